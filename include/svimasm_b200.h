/*
 * svimasm_b200.h -- C ABI of libsvimasm_b200.so: the B200 (sm_100a) drop-in for the
 * alignment-scan + haplotype-pairing hot path of SVIM-asm.
 *
 * The reference (eldariont/svim-asm v1.0.3) is pure Python and has no FFI of its own; the
 * seams this library sits behind are the Python call sites listed in SURVEY.md section 8(b).
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * tree, src/svim_asm/...).  The ctypes binding a maintainer would add is in INTEGRATION.md and
 * in svim_asm_b200/_lib.py.
 *
 * Conventions: plain C, pointers + sizes, no torch types.  Every function returns 0 on success
 * or a negative svb_status; nothing throws across the boundary.  Input host buffers are borrowed
 * for the duration of the call.  The library owns all device memory.  One svb_ctx per device; a
 * ctx is not thread-safe (the reference is single threaded); different ctxs may run concurrently.
 * There is NO CPU fallback: without a CUDA device svb_create() fails with SVB_ERR_CUDA.
 */
#ifndef SVIMASM_B200_H
#define SVIMASM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVB_ABI_VERSION 1

typedef enum {
    SVB_OK = 0,
    SVB_ERR_ARG = -1,       /* bad argument */
    SVB_ERR_CUDA = -2,      /* CUDA runtime error / no device */
    SVB_ERR_NOMEM = -3,
    SVB_ERR_IO = -4,        /* file missing / truncated / not BGZF-BAM */
    SVB_ERR_FORMAT = -5,    /* malformed record / SA tag (reference would raise) */
    SVB_ERR_CAPACITY = -6,  /* internal per-read scratch limit exceeded */
    SVB_ERR_ASSERT = -7     /* a reference `assert end >= start` would have fired (SVCandidate.py:40,83,130,181,266) */
} svb_status;

/* candidate types, in the fixed order of pair_candidates (SVIM_COMBINE.py:180-365) */
enum { SVB_DEL = 0, SVB_INV = 1, SVB_INS = 2, SVB_DUP_TAN = 3, SVB_DUP_INT = 4, SVB_BND = 5 };
/* svb_row.flags */
enum {
    SVB_F_COMPLETE = 1,       /* CandidateInversion.complete            (SVCandidate.py:93)  */
    SVB_F_FULLY_COVERED = 2,  /* CandidateDuplicationTandem.fully_covered (SVCandidate.py:194) */
    SVB_F_CUTPASTE = 4,       /* CandidateDuplicationInterspersed.cutpaste (SVCandidate.py:282) */
    SVB_F_SRC_FWD = 8,        /* CandidateBreakend.source_direction == 'fwd' (SVCandidate.py:357) */
    SVB_F_DST_FWD = 16        /* CandidateBreakend.dest_direction == 'fwd' */
};
/* svb_row.genotype */
enum { SVB_GT_HOM = 0 /* "1/1" */, SVB_GT_HAP1 = 1 /* "1/0" */, SVB_GT_HAP2 = 2 /* "0/1" */ };

/* The nine integer options the hot path reads (SVIM_input_parsing.py:45-95,221,229). */
typedef struct {
    int32_t min_mapq;
    int32_t min_sv_size;
    int32_t max_sv_size;
    int32_t query_gap_tolerance;
    int32_t query_overlap_tolerance;
    int32_t reference_gap_tolerance;
    int32_t reference_overlap_tolerance;
    int32_t partition_max_distance;
    int32_t max_edit_distance;
} svb_params;

/* One BAM record, 32 bytes.  Everything CIGAR-derived is computed on the device. */
typedef struct {
    int32_t tid;           /* AlignedSegment.reference_id */
    int32_t pos;           /* AlignedSegment.reference_start (0-based) */
    uint16_t flag;
    uint8_t mapq;
    uint8_t reserved0;
    uint32_t n_cigar;      /* real op count (after CG:B,I restore), may exceed 65535 */
    uint64_t cigar_off;    /* index of the first op in cigar[]; multiple of 4 (16-byte aligned run) */
    uint32_t l_seq;        /* stored query length */
    uint32_t sa_first;     /* first entry of this record in seg[]; its count is sa_count */
} svb_aln_hdr;

/* One entry of an SA:Z tag after retrieve_other_alignments (SVIM_COLLECT.py:18-55): the pseudo
 * AlignedSegment reduced to the integers analyze_read_segments reads (SVIM_inter.py:68-80). */
typedef struct {
    int32_t tid;           /* bam.get_tid(rname), -1 if unknown */
    int32_t pos;           /* pos - 1 */
    uint8_t is_reverse;    /* strand != '+'  (flag 2064 vs 2048) */
    uint8_t mapq;          /* 0 if the text value is outside 0..255 (SVIM_COLLECT.py:42-45) */
    uint16_t reserved0;
    int32_t ref_end;       /* pysam reference_end: pos + sum(M,D,N,=,X), pos+1 if that sum is 0 */
    int32_t q_astart;      /* pysam query_alignment_start */
    int32_t q_aend;        /* pysam query_alignment_end for l_qseq == 0 */
    int32_t read_len;      /* pysam infer_read_length(): sum(M,I,S,=,X,H) */
    int32_t reserved1;
} svb_segment;

/* One structural-variant candidate (the Candidate* classes of SVCandidate.py), 64 bytes. */
typedef struct {
    uint8_t type;          /* SVB_DEL ... SVB_BND */
    uint8_t flags;         /* SVB_F_* */
    uint8_t genotype;      /* SVB_GT_* */
    uint8_t hap;           /* haplotype of the (first) supporting record: 1 or 2, 0 in haploid runs */
    int32_t src_tid;       /* source_contig  (-1 when the class has none: INS) */
    int32_t src_start;
    int32_t src_end;
    int32_t dst_tid;       /* dest_contig    (-1 when the class has none: DEL/INV/DUP_TAN) */
    int32_t dst_start;
    int32_t dst_end;
    int32_t copies;        /* DUP_TAN */
    uint32_t aln_idx;      /* record that produced it: reads[0] = its query_name; INS sequence source */
    uint32_t seq_pos;      /* INS: start of the inserted bases in that record's query_sequence */
    uint32_t seq_len;      /* INS: number of bases the python slice yields */
    uint32_t mate_aln;     /* paired ("1/1") rows: record index in the OTHER haplotype, else 0xFFFFFFFF */
    uint64_t ordinal;      /* total order = the reference's list-append order within one collect */
    uint64_t reserved0;
} svb_row;


typedef struct svb_ctx svb_ctx;
typedef struct svb_records svb_records;   /* one BAM file (one haplotype) resident in HBM */
typedef struct svb_table svb_table;       /* candidate table resident in HBM */
typedef struct svb_ref svb_ref;           /* reference genome resident in HBM */
typedef struct svb_bam svb_bam;           /* host image of a BAM file produced by the ingest */

/* ---- context ------------------------------------------------------------------------------ */
int svb_abi_version(void);
int svb_create(int device, svb_ctx** out);
void svb_destroy(svb_ctx* ctx);
const char* svb_last_error(const svb_ctx* ctx);          /* valid until the next call on ctx */
int svb_synchronize(svb_ctx* ctx);

/* ---- host ingest: replaces pysam.AlignmentFile / bam.fetch (svim-asm:63,85-86; SVIM_COLLECT.py:62-65)
 * and retrieve_other_alignments (SVIM_COLLECT.py:8-58). Pure host code (zlib inflate, thread pool). */
int svb_bam_open(const char* path, int n_threads, svb_bam** out, char* err, int err_len);
void svb_bam_close(svb_bam* bam);
int64_t svb_bam_n_records(const svb_bam* bam);
int64_t svb_bam_n_ops_padded(const svb_bam* bam);
int64_t svb_bam_n_segments(const svb_bam* bam);
int32_t svb_bam_n_contigs(const svb_bam* bam);
const char* svb_bam_contig_name(const svb_bam* bam, int32_t tid);
const int32_t* svb_bam_contig_lengths(const svb_bam* bam);
const char* svb_bam_sort_order(const svb_bam* bam);      /* @HD SO value or "" (svim-asm:65) */
const svb_aln_hdr* svb_bam_headers(const svb_bam* bam);
const uint32_t* svb_bam_cigar(const svb_bam* bam);
const svb_segment* svb_bam_segments(const svb_bam* bam);
const uint32_t* svb_bam_sa_count(const svb_bam* bam);
const uint8_t* svb_bam_seq4(const svb_bam* bam);         /* 4-bit packed query bases */
const uint64_t* svb_bam_seq_offsets(const svb_bam* bam); /* n_records + 1 byte offsets */
const char* svb_bam_query_name(const svb_bam* bam, int64_t record);
const char* svb_bam_sa_text(const svb_bam* bam, int64_t record);   /* raw SA:Z value or NULL (AlignedSegment.get_tag("SA")) */
/* SA:Z text -> segments for callers that hold records in memory (same rules as the ingest).
 * names/n_contig resolve rname -> tid.  Returns the number of segments written or a negative status. */
int svb_parse_sa(const char* sa_text, const char* const* contig_names, int32_t n_contig,
                 svb_segment* out, int32_t cap);

/* Device ingest: the same BAM file, inflated and split into records ON the GPU.
 * Replaces pysam.AlignmentFile(path) + bam.fetch + AlignedSegment.cigartuples (svim-asm:63,85-86;
 * SVIM_COLLECT.py:65; SVIM_intra.py:37) for the hot path: `rec_out` is the record image svb_load_records would
 * have built from svb_bam_open's arrays (bit-identical), `bam_out` holds the host-side small data (header,
 * svb_aln_hdr array, names, SA texts and segments); its CIGAR / sequence arrays stay empty until
 * svb_bam_materialize_host downloads them (only the per-alignment seams read them).
 * contig_lexrank: rank of every contig name under python string order, or NULL (code-point order of the names). */
int svb_bam_open_device(svb_ctx* ctx, const char* path, int keep_sequences, const int32_t* contig_lexrank,
                        svb_bam** bam_out, svb_records** rec_out, char* err, int err_len);
int svb_bam_materialize_host(svb_ctx* ctx, svb_bam* bam, const svb_records* rec, int what);   /* what: 1 CIGAR ops, 2 query bases, 3 both */

/* ---- device: replaces analyze_alignment_file_coordsorted (SVIM_COLLECT.py:61-83) and below ---- */
int svb_load_records(svb_ctx* ctx, const svb_aln_hdr* hdr, uint32_t n_aln, const uint32_t* cigar,
                     uint64_t n_ops_padded, const svb_segment* seg, const uint32_t* sa_count, uint32_t n_seg,
                     const int32_t* contig_len, const int32_t* contig_lexrank, int32_t n_contig,
                     svb_records** out);
void svb_records_free(svb_records* rec);
/* Attach the 4-bit query sequences (needed only when the table is paired: INS edit distances,
 * SVIM_COMBINE.py:65-76).  seq_off has n_aln + 1 byte offsets. */
int svb_records_set_sequences(svb_ctx* ctx, svb_records* rec, const uint8_t* seq4, const uint64_t* seq_off);

/* The same without the upload: the sequences stay in PINNED host memory of the caller (which must outlive the records) and
 * the device reads the inserted bases it needs in place over PCIe. */
int svb_records_map_sequences_host(svb_ctx* ctx, svb_records* rec, const uint8_t* seq4_pinned, const uint64_t* seq_off_pinned);

/* K2 cigar_scan + K4 segment_walk + K5 ordered merge.  hap = 0 (haploid), 1 or 2.
 * Rows come out in the reference's emission order (SVIM_COLLECT.py:74,79-80). */
int svb_collect(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, svb_table** out);

/* Both haplotypes of a diploid run (svim-asm:97,114: two analyze_alignment_file_coordsorted calls) with ONE host
 * synchronisation: the scans, finalize passes and walk counts of both are enqueued, the host reads all counts at once, the
 * walk rows are written and merged without waiting again.  hap = 1 for rec1, 2 for rec2.  with_pools != 0 also copies the
 * inserted bases of the INS rows next to each table (svb_table_gather_sequences) -- their size is summed on the device
 * before the synchronisation, so this adds no wait either; the records need their sequences (svb_records_set_sequences
 * or svb_records_map_sequences_host). */
int svb_collect2(svb_ctx* ctx, const svb_records* rec1, const svb_records* rec2, const svb_params* p, int with_pools,
                 svb_table** out1, svb_table** out2);

/* analyze_cigar_indel (SVIM_intra.py:8-30) on one op list: rows of (pos_ref, pos_read, length, is_del). */
int svb_cigar_indel(svb_ctx* ctx, const uint32_t* packed_ops, uint32_t n_ops, int32_t min_length,
                    int64_t* out4, uint32_t cap, uint32_t* n_out);

/* Reference genome: upper-cased bases, 1 byte each, contigs concatenated; contig_off has n_contig + 1 entries. */
int svb_ref_load(svb_ctx* ctx, const uint8_t* bases, const uint64_t* contig_off, int32_t n_contig, svb_ref** out);
/* The same from the FASTA file itself (pysam.FastaFile(path), svim-asm:124): the file is copied to the device as it is and
 * one kernel drops the line terminators and upper-cases.  fai: n_contig rows {length, offset, linebases, linewidth} of
 * the .fai index in BAM header order; length 0 marks a contig that the FASTA does not hold. */
int svb_ref_load_fasta(svb_ctx* ctx, const char* path, const uint64_t* fai, int32_t n_contig, svb_ref** out);
void svb_ref_free(svb_ref* ref);

/* pair_candidates (SVIM_COMBINE.py:164-366): form_partitions + compute_distance + pair_haplotypes(_breakends).
 * rec1/rec2 supply the INS sequences of the two haplotypes. Rows come out in the reference's order. */
int svb_pair(svb_ctx* ctx, const svb_table* h1, const svb_table* h2, const svb_records* rec1,
             const svb_records* rec2, const svb_ref* ref, const svb_params* p, svb_table** out);

/* form_partitions (SVIM_COMBINE.py:15-32) for caller-built keys: key = (group << 32) | position with group = rank of
 * (type, contig) under python tuple order and 0 <= position < 2^31.  Stable sort by key (sorted(), :17), then the split
 * of :20-31: a new partition where the group changes or CONSECUTIVE positions are more than max_distance apart.
 * order[i] = input index of the i-th sorted item; part_start[p] = first sorted index of partition p, part_start[n_parts] = n
 * (capacity n + 1). */
int svb_form_partitions(svb_ctx* ctx, const uint64_t* keys, uint32_t n, int64_t max_distance, uint32_t* order,
                        uint32_t* part_start, uint32_t* n_parts);

/* compute_distance for explicit strings (edlib.align(a,b)["editDistance"], SVIM_COMBINE.py:50): one pair of compute_distance (python seam). */
int svb_edit_distance(svb_ctx* ctx, const uint8_t* a, const uint64_t* a_off, const uint8_t* b, const uint64_t* b_off,
                      uint32_t n_pairs, int64_t* out);
/* The same with edlib's k parameter (edlib.align(a, b, k=max_distance)): out[i] = the distance if it is <= max_distance,
 * else -1.  This is the question pair_haplotypes asks of most pairs (fcluster(Z, max_edit_distance, 'distance'),
 * SVIM_COMBINE.py:135): it runs the thresholded wavefront kernel svb_pair uses (O(max_distance^2 + length) per pair). */
int svb_edit_distance_bounded(svb_ctx* ctx, const uint8_t* a, const uint64_t* a_off, const uint8_t* b, const uint64_t* b_off,
                              uint32_t n_pairs, int64_t max_distance, int64_t* out);
/* scipy linkage(method="complete") + fcluster(t, "distance") labels for n <= 32 points given condensed
 * distances (SVIM_COMBINE.py:134-135, SVIM_inter.py:47-48): pair_haplotypes[_breakends] (python seams). */
int svb_cluster_labels(svb_ctx* ctx, const double* condensed, const uint32_t* n_points, uint32_t n_problems,
                       double threshold, int32_t* labels_out /* 32 per problem */);

int64_t svb_table_size(const svb_table* t);
int svb_table_to_host(svb_ctx* ctx, const svb_table* t, svb_row* dst, uint64_t cap, uint64_t* n);
int svb_table_from_host(svb_ctx* ctx, const svb_row* rows, uint64_t n, svb_table** out);
/* device-to-device export / import: the all-gatherv of the candidate table between ranks runs on these. */
int svb_table_export(svb_ctx* ctx, const svb_table* t, void* device_dst, uint64_t cap_rows);
int svb_table_import(svb_ctx* ctx, const void* device_src, uint64_t n_rows, svb_table** out);
void svb_table_free(svb_table* t);

/* ---- Multi-GPU exchange (one process per GPU, records sharded by contig) --------------------
 * The reference is one process with one candidate list per type (SVIM_COLLECT.py:67-91, svim-asm:70-76), so every
 * candidate is visible to pair_candidates (SVIM_COMBINE.py:164).  These calls restore that across ranks without host
 * staging: pack both haplotype tables of a rank into one device buffer, all-gather it (NCCL, by the caller), unpack
 * the gathered buffers into "all ranks' rows in append order, restricted to the key contigs this rank owns". */
void* svb_stream(svb_ctx* ctx);                                   /* cudaStream_t the context works on */
int svb_device_alloc(svb_ctx* ctx, uint64_t bytes, void** out);   /* stream-ordered scratch for the collective */
void svb_device_free(svb_ctx* ctx, void* p);
/* index of every record of a shard in the unsharded batch: aln_idx / ordinal of a table become global with
 * svb_table_remap_records (the reference's append order is the BAM order of the whole file, SVIM_COLLECT.py:65) */
int svb_records_set_global_index(svb_ctx* ctx, svb_records* rec, const uint32_t* global_idx);
int svb_table_remap_records(svb_ctx* ctx, svb_table* t, const svb_records* rec);
/* sizes = {rows of table 1, pool bytes 1, rows 2, pool bytes 2}; svb_exchange_bytes = packed size of such a pair */
int svb_exchange_sizes(const svb_table* t1, const svb_table* t2, uint64_t sizes[4]);
uint64_t svb_exchange_bytes(const uint64_t sizes[4]);
int svb_exchange_pack(svb_ctx* ctx, const svb_table* t1, const svb_table* t2, void* d_buf, uint64_t cap_bytes);
/* d_gathered: world buffers of `stride` bytes each, sizes: world x 4 (host), owner[n_contig]: rank owning each contig
 * (host).  The result keeps the rows whose key contig (Candidate.get_key, SVCandidate.py:17-19) belongs to `rank`. */
int svb_exchange_unpack(svb_ctx* ctx, const void* d_gathered, uint64_t stride, const uint64_t* sizes, int world, int hap,
                        const int32_t* owner, int n_contig, int rank, svb_table** out);
const void* svb_table_device_rows(const svb_table* t);
/* The same exchange over PEER MEMORY, without a collective library (one process per GPU, CUDA IPC): every rank owns a
 * window in its HBM with one slot per rank; svb_exchange_share stores this rank's packed tables straight into its slot of
 * every peer's window (one kernel, NVLink stores), raises a flag there, waits on the flags of its own window (a one-warp
 * kernel on the library's stream) and unpacks; svb_exchange_gather_paired does the same for the paired rows and puts them
 * into pair_candidates' order (type, then contig by python string order) on the device.  Set-up: create, exchange the
 * 64-byte handles between the processes (any transport), open.  Both calls are collective: every rank makes them once
 * per step, in the same order. */
typedef struct svb_exchange svb_exchange;
int svb_exchange_create(svb_ctx* ctx, int world, int rank, uint64_t slot_bytes, uint64_t result_bytes, svb_exchange** out);
int svb_exchange_handle(svb_ctx* ctx, const svb_exchange* x, uint8_t handle_out[64]);
int svb_exchange_open(svb_ctx* ctx, svb_exchange* x, const uint8_t* handles /* world x 64 bytes, rank order */);
void svb_exchange_destroy(svb_ctx* ctx, svb_exchange* x);
int svb_exchange_share(svb_ctx* ctx, svb_exchange* x, const svb_table* t1, const svb_table* t2, const int32_t* owner, int n_contig,
                       svb_table** u1, svb_table** u2);
int svb_exchange_gather_paired(svb_ctx* ctx, svb_exchange* x, const svb_table* paired, const int32_t* contig_lexrank, int n_contig,
                               svb_table** out);            /* device pointer of the rows (all-gather of paired rows) */

/* ---- VCF body: the text of write_final_vcf's record lines (SVIM_COMBINE.py:428-475) and of
 * Candidate*.get_vcf_entry* (SVCandidate.py:53-78,99-125,151-176,203-261,296-347,389-443), assembled on the device: REF /
 * ALT alleles are gathered from the HBM-resident reference and the 4-bit query sequences, one warp per record.
 * The caller has put the entries into the writer's order (sorted_nicely, SVIM_COMBINE.py:369-376) and numbered the IDs. */
enum { SVB_VCF_DEL = 0, SVB_VCF_INV = 1, SVB_VCF_INS = 2, SVB_VCF_TAN_AS_INS = 3, SVB_VCF_TAN_AS_DUP = 4,
       SVB_VCF_INT_AS_INS = 5, SVB_VCF_INT_AS_DUP = 6, SVB_VCF_BND = 7, SVB_VCF_BND_MATE = 8 };
enum { SVB_VCF_SYMBOLIC = 1 };                /* flags: options.symbolic_alleles */
typedef struct {
    uint32_t row;          /* index into the table */
    uint32_t id;           /* number after "svim_asm.<LABEL>." (1-based, per label, in output order) */
    uint32_t mode;         /* SVB_VCF_* : which get_vcf_entry* of the row's class */
} svb_vcf_entry;
/* rec[h] = record image holding the query sequences of haplotype slot h (0 haploid, 1, 2; NULL when unused).
 * names / name_off: contig names concatenated, n_contig + 1 offsets.  *text points into a pinned host buffer owned
 * by ctx, valid until the next svb_vcf_body call or svb_destroy. */
int svb_vcf_body(svb_ctx* ctx, const svb_table* t, const svb_records* const rec[3], const svb_ref* ref, const char* names,
                 const uint32_t* name_off, int32_t n_contig, const svb_vcf_entry* entries, uint64_t n_entries, uint32_t flags,
                 const uint8_t** text, uint64_t* n_bytes);

/* Sequence pools: the inserted bases of the INS rows (candidate.sequence, SVIM_intra.py:42, SVIM_inter.py:117,120)
 * copied next to the table so that it can be paired -- or sent to another rank -- without the record image. */
int svb_table_gather_sequences(svb_ctx* ctx, svb_table* t, const svb_records* rec);          /* device gather */
int svb_table_attach_sequences_host(svb_ctx* ctx, svb_table* t, const uint8_t* seq4, const uint64_t* seq_off);
int svb_table_pool_to_host(svb_ctx* ctx, const svb_table* t, uint8_t* pool_dst, uint64_t cap_bytes, uint64_t* off_dst,
                           uint64_t* pool_bytes);
/* row_off[i] = byte offset of row i's run in `pool` (runs need not be contiguous or in row order: a gathered pool of
 * several ranks is uploaded as is, only the per-row offsets are permuted) */
int svb_table_set_pool_from_host(svb_ctx* ctx, svb_table* t, const uint8_t* pool, uint64_t pool_bytes, const uint64_t* row_off);

#ifdef __cplusplus
}
#endif
#endif /* SVIMASM_B200_H */
