// Shared declarations of libsvimasm_b200 (device layouts, context, error plumbing).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/svimasm_b200_debug.h"

// ---- device image of one BAM file -----------------------------------------------------------
// cigar: BAM-packed ops (len << 4 | op) as uint4 (4 ops = 16 B); every alignment's run starts on
// a uint4 boundary and is padded with op 15.  off4[a] is the run start in uint4 units.
struct svb_records {
    int device = 0;
    cudaStream_t stream = nullptr;    // stream of the owning context (stream-ordered frees)
    uint32_t n_aln = 0;
    uint32_t n_seg = 0;
    int32_t n_contig = 0;
    uint64_t n4 = 0;                  // number of uint4 in cigar
    uint64_t n_ops = 0;               // sum of n_cigar (real ops)
    svb_aln_hdr* d_hdr = nullptr;     // [n_aln]
    uint4* d_cigar = nullptr;         // [cigar_padded_n4(n4)]: the tail is filled with op 15 at load
    uint32_t* d_off4 = nullptr;       // [n_aln + 1]
    uint32_t* d_chunk_first = nullptr;// [ceil(n4 / 256)] alignment that owns the first uint4 of each 1024-op chunk
    svb_segment* d_seg = nullptr;     // [n_seg]
    uint32_t* d_sa_count = nullptr;   // [n_aln]
    int32_t* d_contig_len = nullptr;  // [n_contig]
    int32_t* d_contig_lexrank = nullptr;
    uint4* d_aln_sum = nullptr;       // [n_aln] x: sum(M,D,=,X) y: sum(M,I,S,=,X) z: sum(N) w: sum(H); written by cigar_scan
    uint32_t* d_prim_list = nullptr;  // [n_prim] record indices that carry SA segments
    uint32_t n_prim = 0;
    uint8_t* d_seq4 = nullptr;        // optional 4-bit query sequences
    uint64_t* d_seq_off = nullptr;    // [n_aln + 1]
    uint32_t* d_global_idx = nullptr; // optional [n_aln]: index of each record in the unsharded batch (exchange.cu)
    uint64_t seq_bytes = 0;
    bool seq_borrowed = false;        // d_seq4 / d_seq_off are device views of the caller's pinned host buffers
    std::vector<int32_t> h_contig_len; // host copy of the contig table (svb_pair checks that both haplotypes share it)
};

struct svb_table {
    int device = 0;
    cudaStream_t stream = nullptr;
    svb_row* d_rows = nullptr;
    uint64_t n = 0;
    uint64_t cap = 0;
    uint8_t* d_pool = nullptr;        // optional: 4-bit inserted sequences of the INS rows (seqpool.cu)
    uint64_t* d_pool_off = nullptr;   // [n + 1] byte offset of each row's run in d_pool
    uint64_t pool_bytes = 0;
};
int table_drop_pool(svb_table* t);

struct svb_ref {
    int device = 0;
    uint8_t* d_bases = nullptr;       // upper-cased ASCII, contigs concatenated
    uint64_t* d_contig_off = nullptr; // [n_contig + 1]
    uint8_t* d_class_map = nullptr;   // [256] byte -> symbol class of the edit-distance kernel (255 = none)
    int32_t n_contig = 0;
    uint64_t n_bases = 0;
};

struct TimedSpan {
    int kernel;
    cudaEvent_t start, stop;
};

struct svb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t main_stream = nullptr;   // `stream` outside a SideStream scope
    cudaStream_t side = nullptr;          // second stream: independent halves of a diploid collect overlap (svb_collect2)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    int scan_variant = 0;             // 0 per-warp TMA bulk-copy ring (default: measured faster, profiles/), 1 LDG.128.nc
    int sm_count = 148;
    bool timing_enabled = true;
    std::vector<TimedSpan> spans;     // recorded, not yet folded into `timing`
    std::vector<cudaEvent_t> free_events;
    svb_timing timing;
    cudaEvent_t marks[16] = {};
    uint64_t launches = 0;            // kernels launched by this context (svb_launch_count)
    void* d_scratch = nullptr;        // grow-only device scratch
    size_t scratch_bytes = 0;
    uint32_t* d_status = nullptr;     // device error word (atomicOr of DEV_ERR_*)
    unsigned long long* d_counters = nullptr;   // 64 device counters
    unsigned long long* h_pinned = nullptr;     // 64-word pinned readback area
    uint8_t* h_text = nullptr;        // grow-only pinned buffer the VCF body is returned in (vcf_device.cu)
    size_t h_text_cap = 0;
    void* h_stage = nullptr;          // grow-only pinned staging buffer of svb_table_to_host
    size_t h_stage_cap = 0;
    bool wfa_attr_set = false;        // cudaFuncSetAttribute is per device: one flag per context, not per process
    bool scan_attr_set[8] = {};
    uint64_t ed_stride = 512u * 1024u; // bytes per worker warp for parked bottom-row deltas of the exact edit-distance kernel (grows on demand)
    uint64_t last_pair_stats[4] = {};  // partitions, cross-haplotype pairs, pairs that needed the exact kernel (last svb_pair)
    double ingest_ms[12] = {};        // stage times of the last svb_bam_open_device on this context (svb_bam_device_timings)
    void* uploader = nullptr;         // pinned staging slots + streams of upload_file_range (file_upload.cu)
};

enum : uint32_t {
    DEV_ERR_BAD_TID = 1u,        // a name lookup on tid < 0 or >= n_contig (pysam would raise ValueError)
    DEV_ERR_ASSERT = 2u,         // reference assert end >= start would fire
    DEV_ERR_CAPACITY = 4u,       // per-read scratch exceeded (inversion run > 32, ...)
    DEV_ERR_NOSEQ = 8u,          // an insertion needs query bases that were not uploaded
    DEV_ERR_EXCHANGE = 16u,      // a peer's flag did not arrive within the exchange time-out (exchange.cu)
};

int svb_fail(svb_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess);

#define SVB_CUDA(ctx, call)                                                         \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) return svb_fail((ctx), SVB_ERR_CUDA, #call, e__);   \
    } while (0)

// While one of these is alive everything the library enqueues through ctx->stream goes to the SIDE stream, ordered after what
// the main stream held when the scope began (fork); join() makes the main stream wait for it.  Used where two chains of
// small, latency-bound kernels are independent (the two haplotypes of svb_collect2).
struct SideStream {
    svb_ctx* c;
    explicit SideStream(svb_ctx* ctx) : c(ctx) {
        cudaEventRecord(c->ev_fork, c->main_stream);
        cudaStreamWaitEvent(c->side, c->ev_fork, 0);
        c->stream = c->side;
    }
    ~SideStream() { c->stream = c->main_stream; }
    static void join(svb_ctx* ctx) {
        cudaEventRecord(ctx->ev_join, ctx->side);
        cudaStreamWaitEvent(ctx->main_stream, ctx->ev_join, 0);
    }
};

// scoped kernel timer: records start/stop events on ctx->stream for kernel class `k`
struct KernelTimer {
    svb_ctx* ctx;
    TimedSpan span;
    bool live;
    KernelTimer(svb_ctx* c, int kernel_id);
    ~KernelTimer();
};

// file bytes -> device, parallel pread + pinned staging.  wait_for_stream = false: the caller has synchronised ctx->stream itself
// and calls from a thread of its own (errors are returned, not recorded in the context).  background = true: the upload pauses
// between chunks while a foreground upload of this process is in flight.
int upload_file_range(svb_ctx* ctx, int fd, uint64_t offset, uint64_t n, void* d_dst, bool wait_for_stream = true, bool background = false);
void upload_release(svb_ctx* ctx);
void* svb_scratch(svb_ctx* ctx, size_t bytes);   // grow-only device scratch (nullptr on failure)

// ---- kernels (host launchers) -----------------------------------------------------------------
struct ScanOutput {
    svb_row* rows;            // capacity `cap`
    uint64_t cap;
    unsigned long long* d_count;   // total number of emitted rows (may exceed cap)
    unsigned long long* d_ins_bytes = nullptr;   // optional: bytes of 4-bit inserted sequence the INS rows hold (the sequence pool's size)
};
uint64_t cigar_padded_n4(uint64_t n4);   // uint4 capacity d_cigar must be allocated with (whole 64 KB units)
int launch_build_chunk_index(svb_ctx* ctx, svb_records* rec);
int launch_cigar_scan(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, ScanOutput out);
struct WalkPending;   // a walk whose count pass is in flight (segment_walk.cu)
int walk_count_async(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, unsigned long long* totals, WalkPending** out);   // totals[0] rows, [1] inserted bytes
int walk_write_async(svb_ctx* ctx, WalkPending* w, uint64_t n_rows, svb_row** d_rows_out);                     // consumes w
void walk_discard(svb_ctx* ctx, WalkPending* w);
int gather_pool_known(svb_ctx* ctx, svb_table* t, const uint8_t* seq4, const uint64_t* seq_off, uint64_t pool_bytes);   // seqpool.cu
int launch_scan_u32(svb_ctx* ctx, uint32_t* v, uint32_t n, unsigned long long* d_total);   // exclusive scan in place, v[n] = total (segment_walk.cu)
int launch_merge_tables(svb_ctx* ctx, const svb_row* a, uint64_t na, const svb_row* b, uint64_t nb, svb_row* out);
int run_pairing(svb_ctx* ctx, const svb_table* h1, const svb_table* h2, const svb_records* rec1,
                const svb_records* rec2, const svb_ref* ref, const svb_params* p, svb_table** out);
int run_edit_distance_strings(svb_ctx* ctx, const uint8_t* a, const uint64_t* a_off, const uint8_t* b,
                              const uint64_t* b_off, uint32_t n_pairs, int64_t max_distance, int64_t* out);
int run_form_partitions(svb_ctx* ctx, const uint64_t* keys_host, uint32_t n, int64_t max_distance, uint32_t* order_out,
                        uint32_t* part_start_out, uint32_t* n_parts_out);
int run_cluster_labels(svb_ctx* ctx, const double* condensed, const uint32_t* n_points, uint32_t n_problems,
                       double threshold, int32_t* labels_out);
