// K2 cigar_scan -- the per-alignment CIGAR-op scan of SVIM-asm as one streaming pass over HBM.
//
// Replaces analyze_cigar_indel (reference src/svim_asm/SVIM_intra.py:8-30) and the candidate
// construction of analyze_alignment_indel (SVIM_intra.py:33-44, SVCandidate.py:39-50,129-141),
// applied to every record that passes the filter of analyze_alignment_file_coordsorted
// (SVIM_COLLECT.py:71).
//
// Formulation.  The reference walks each alignment's (op,len) list keeping running pos_ref /
// pos_read and appends an indel when an I or D op has len >= min_sv_size.  Here ALL alignments of
// a BAM file sit in one flat array of BAM-packed ops, and the walk becomes a segmented exclusive
// scan (segments = alignments) fused with an order-preserving stream compaction of the rare
// emitting ops, done in ONE pass over HBM:
//   * a CTA claims a RUN of 8 warps x G chunks x 1024 ops with a ticket; warp w owns G consecutive
//     chunks (4 KB each), so its running "sum since the last alignment head" stays in registers;
//   * a chunk is staged to shared memory by a per-warp ring of TMA bulk copies (cp.async.bulk +
//     mbarrier, two 4 KB stages per warp, the next chunk is in flight while this one is decoded);
//     variant 1 reads the chunk with coalesced 128-bit LDG.nc straight into registers instead;
//   * per op: one 8-byte table entry {multiplier, threshold} from shared memory, one IMAD.WIDE into a
//     packed accumulator (read advance in bits 0..30, reference advance from bit 31 up), one compare
//     that flags the rare ops (I/D with len >= min_sv_size, N/H).  Exact event bits are recomputed
//     only when the flag fires somewhere in the warp (about 1 chunk in 8 for human assemblies);
//   * the run's aggregate and emit count go to a run-status array; ALL 8 warps look back together
//     (256 predecessors per round trip) once per run, so the look-back latency is paid per 128 KB;
//   * only chunks that hold an emitting op are re-read (from L2) to compute the exclusive prefix of
//     that op (two warp reductions per event) and to write the finished 64-byte candidate row at its
//     final, stable position.
// Per-alignment totals (reference span, read span, N and H bases), needed by the split-alignment
// walk for reference_end / infer_read_length (SVIM_inter.py:68-80), fall out as atomics.
//
// Bound: HBM bandwidth.  Algorithmic bytes: 4 B/op + 32 B/alignment + 64 B/emitted row.
#include "common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int ROWS = 8;                    // uint4 per lane per chunk
constexpr int CHUNK4 = 32 * ROWS;          // 256 uint4  = 1024 ops = 4 KB per chunk
constexpr int STAGES = 2;                  // TMA ring depth per warp
// G = chunks per warp per run (template parameter): a run is WARPS * G * CHUNK4 uint4 (G = 16: 512 KB).  Bigger runs
// amortise the run-end barrier + look-back (measured on B200, 2.0e8 ops: G=4 2.5 TB/s, G=8 3.0 TB/s, G=16 3.1 TB/s
// with the LDG variant); smaller ones keep all SMs busy on small inputs.

constexpr uint32_t REF_MASK = (1u << 0) | (1u << 2) | (1u << 7) | (1u << 8);              // M D = X  (SVIM_intra.py:15,23,28)
constexpr uint32_t READ_MASK = (1u << 0) | (1u << 1) | (1u << 4) | (1u << 7) | (1u << 8); // M I S = X (SVIM_intra.py:16,20,26,29)
constexpr uint32_t INDEL_MASK = (1u << 1) | (1u << 2);
constexpr uint32_t NH_MASK = (1u << 3) | (1u << 5);

// run status: three 64-bit words, each tagged with its own state in the top 2 bits so that a reader
// can validate a snapshot without fences (values are self-describing).
constexpr unsigned long long ST_INVALID = 0ull, ST_AGG = 1ull, ST_PREFIX = 2ull;
struct RunStatus {
    unsigned long long w_ref;    // [63:62] state  [32] has_head  [31:0] ref sum since last head
    unsigned long long w_read;   // [63:62] state               [31:0] read sum since last head
    unsigned long long w_cnt;    // [63:62] state  [61:0] emitted rows
    unsigned long long pad;
};

// geometry of one 1024-op chunk, precomputed when the record image is loaded
struct ChunkGeom {
    uint32_t a_lo;        // alignment that owns the first uint4 of the chunk
    uint32_t n_heads;     // alignment runs that START inside the chunk after its first uint4 (a_hi = a_lo + n_heads)
    uint32_t split_rel;   // uint4 offset (from the chunk start) of the first such head, or the chunk's length
    uint32_t head_at_start;   // 1 if a_lo's run starts exactly at the chunk start
};

struct ScanArgs {
    const uint4* cigar;
    uint64_t n4;
    const uint32_t* off4;
    const ChunkGeom* geom;
    const svb_aln_hdr* hdr;
    const int32_t* contig_len;
    uint32_t n_aln;
    int32_t n_contig;
    uint32_t n_runs;
    int32_t min_mapq;
    uint32_t min16;           // min_sv_size << 4 : (x >= min16) <=> (len >= min_sv_size)
    uint32_t hap;
    uint4* aln_sum;
    RunStatus* status;
    unsigned int* ticket;
    svb_row* rows;
    unsigned long long cap;
    unsigned long long* total;
    uint32_t* dev_status;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ready = 0;
    while (!ready) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ready) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_chunk(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// exact classification of one op (rare paths): advance sums, emit bit, N/H bit
__device__ __forceinline__ void decode_op(uint32_t x, uint32_t min16, uint32_t& r, uint32_t& q, uint32_t& ev, uint32_t& nh,
                                          uint32_t bit) {
    const uint32_t op = x & 15u;
    const uint32_t len = x >> 4;
    if ((REF_MASK >> op) & 1u) r += len;
    if ((READ_MASK >> op) & 1u) q += len;
    if (((INDEL_MASK >> op) & 1u) && x >= min16) ev |= bit;
    if (((NH_MASK >> op) & 1u) && len != 0u) nh |= bit;
}

__device__ __forceinline__ bool record_passes(const svb_aln_hdr& h, int32_t min_mapq) {
    // SVIM_COLLECT.py:71  is_unmapped / is_secondary / mapping_quality < min_mapq; records without a contig are
    // never returned by bam.fetch(contig) (:65)
    return h.tid >= 0 && !(h.flag & 0x4) && !(h.flag & 0x100) && static_cast<int32_t>(h.mapq) >= min_mapq;
}

// fast path of one uint4 (4 ops): packed advance sums and the rare flag
__device__ __forceinline__ void fast_row(const uint4 d, const uint2* lut, bool in_head, uint32_t& totR, uint32_t& totQ,
                                         uint32_t& headR, uint32_t& headQ, bool& rare) {
    const uint2 e0 = lut[d.x & 15u], e1 = lut[d.y & 15u], e2 = lut[d.z & 15u], e3 = lut[d.w & 15u];
    unsigned long long acc = static_cast<unsigned long long>(d.x >> 4) * e0.x;
    acc += static_cast<unsigned long long>(d.y >> 4) * e1.x;
    acc += static_cast<unsigned long long>(d.z >> 4) * e2.x;
    acc += static_cast<unsigned long long>(d.w >> 4) * e3.x;
    rare = rare || d.x >= e0.y || d.y >= e1.y || d.z >= e2.y || d.w >= e3.y;
    const uint32_t rr = static_cast<uint32_t>(acc >> 31), qq = static_cast<uint32_t>(acc) & 0x7FFFFFFFu;
    totR += rr;
    totQ += qq;
    if (in_head) {
        headR += rr;
        headQ += qq;
    }
}

// row `r` of the chunk that starts at uint4 index c4, re-read through L2 (rare paths only)
__device__ __forceinline__ uint4 reload_row(const uint4* cigar, uint64_t n4, uint64_t c4, int r, uint32_t lane) {
    const uint64_t g4 = c4 + static_cast<uint64_t>(r) * 32u + lane;
    return g4 < n4 ? cigar[g4] : make_uint4(15u, 15u, 15u, 15u);
}

struct ChunkResult {
    uint32_t tailR, tailQ, head, cnt, evbits;
};

// Phase 1 of one chunk.  `rowsrc(r)` yields this lane's uint4 of row r in the unrolled fast loop (r is a
// compile-time constant there); `raresrc(r)` serves the rare paths, where r is a run-time value.
template <typename RowSrc, typename RareSrc>
__device__ __forceinline__ ChunkResult chunk_phase1(const ScanArgs& a, const uint2* lut, const ChunkGeom g, uint64_t c4, uint32_t lane,
                                                    RowSrc rowsrc, RareSrc raresrc) {
    ChunkResult out;
    out.evbits = 0;
    out.cnt = 0;
    const uint64_t c4end = min(c4 + CHUNK4, a.n4);
    const uint32_t a_lo = g.a_lo, a_hi = g.a_lo + g.n_heads, n_pieces = g.n_heads + 1u;
    // rows of this lane before the first head inside the chunk: 32 r + lane < split_rel
    const int rows_before = (static_cast<int>(g.split_rel) - static_cast<int>(lane) + 31) / 32;
    const int nb = rows_before <= 0 ? 0 : (rows_before >= ROWS ? ROWS : rows_before);
    uint32_t headR = 0, headQ = 0, totR = 0, totQ = 0;
    bool rare = false;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) fast_row(rowsrc(r), lut, r < nb, totR, totQ, headR, headQ, rare);
    const uint32_t first_mask = nb >= 8 ? 0xFFFFFFFFu : ((1u << (4u * static_cast<uint32_t>(nb))) - 1u);
    const bool any_rare = __ballot_sync(0xffffffffu, rare) != 0u;

    uint32_t nhbits = 0, evbits = 0;
    if (any_rare) {                        // exact bits of the rare ops
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r) {
            const uint4 d = raresrc(r);
            uint32_t r0 = 0, q0 = 0;
            decode_op(d.x, a.min16, r0, q0, evbits, nhbits, 1u << (4 * r + 0));
            decode_op(d.y, a.min16, r0, q0, evbits, nhbits, 1u << (4 * r + 1));
            decode_op(d.z, a.min16, r0, q0, evbits, nhbits, 1u << (4 * r + 2));
            decode_op(d.w, a.min16, r0, q0, evbits, nhbits, 1u << (4 * r + 3));
        }
    }
    if (n_pieces <= 2u) {
        headR = __reduce_add_sync(0xffffffffu, headR);
        headQ = __reduce_add_sync(0xffffffffu, headQ);
        if (lane == 0 && (headR | headQ)) {
            atomicAdd(&a.aln_sum[a_lo].x, headR);
            atomicAdd(&a.aln_sum[a_lo].y, headQ);
        }
        out.tailR = headR;
        out.tailQ = headQ;
        if (n_pieces == 2u) {
            const uint32_t restR = __reduce_add_sync(0xffffffffu, totR) - headR;
            const uint32_t restQ = __reduce_add_sync(0xffffffffu, totQ) - headQ;
            if (lane == 0 && (restR | restQ)) {
                atomicAdd(&a.aln_sum[a_hi].x, restR);
                atomicAdd(&a.aln_sum[a_hi].y, restQ);
            }
            out.tailR = restR;
            out.tailQ = restQ;
        }
        if (any_rare) {
            const bool pass0 = record_passes(a.hdr[a_lo], a.min_mapq);
            const bool pass1 = n_pieces == 2u ? record_passes(a.hdr[a_hi], a.min_mapq) : false;
            evbits &= (pass0 ? first_mask : 0u) | (pass1 ? ~first_mask : 0u);      // SVIM_COLLECT.py:71
        }
    } else {
        // generic: several alignment heads inside one 1024-op chunk (short alignments)
        uint32_t keep = 0;
        out.tailR = 0;
        out.tailQ = 0;
        for (uint32_t al = a_lo; al <= a_hi; ++al) {
            const uint64_t lo = max(static_cast<uint64_t>(a.off4[al]), c4);
            const uint64_t hi = min(static_cast<uint64_t>(a.off4[al + 1]), c4end);
            uint32_t sR = 0, sQ = 0, in_mask = 0;
#pragma unroll 1
            for (int r = 0; r < ROWS; ++r) {
                const uint64_t g4 = c4 + static_cast<uint64_t>(r) * 32u + lane;
                if (g4 >= lo && g4 < hi) {
                    const uint4 d = raresrc(r);
                    uint32_t e0 = 0, n0 = 0;
                    decode_op(d.x, a.min16, sR, sQ, e0, n0, 1u);
                    decode_op(d.y, a.min16, sR, sQ, e0, n0, 1u);
                    decode_op(d.z, a.min16, sR, sQ, e0, n0, 1u);
                    decode_op(d.w, a.min16, sR, sQ, e0, n0, 1u);
                    in_mask |= 0xFu << (4 * r);
                }
            }
            sR = __reduce_add_sync(0xffffffffu, sR);
            sQ = __reduce_add_sync(0xffffffffu, sQ);
            if (record_passes(a.hdr[al], a.min_mapq)) keep |= in_mask;
            if (lane == 0 && (sR | sQ)) {
                atomicAdd(&a.aln_sum[al].x, sR);
                atomicAdd(&a.aln_sum[al].y, sQ);
            }
            out.tailR = sR;
            out.tailQ = sQ;
        }
        evbits &= keep;
    }
    if (any_rare) {
        out.cnt = __reduce_add_sync(0xffffffffu, __popc(evbits));
        // rare: N / H ops feed reference_end / infer_read_length of the split-alignment walk
        if (__ballot_sync(0xffffffffu, nhbits != 0u)) {
            uint32_t bits = nhbits;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1u;
                const uint64_t g4 = c4 + static_cast<uint64_t>(b >> 2) * 32u + lane;
                uint32_t al = a_lo;
                while (al < a_hi && static_cast<uint64_t>(a.off4[al + 1]) <= g4) ++al;
                const uint4 d = raresrc(b >> 2);
                const uint32_t x = (b & 3) == 0 ? d.x : (b & 3) == 1 ? d.y : (b & 3) == 2 ? d.z : d.w;
                if ((x & 15u) == 3u) atomicAdd(&a.aln_sum[al].z, x >> 4);
                else atomicAdd(&a.aln_sum[al].w, x >> 4);
            }
        }
    }
    out.evbits = evbits;
    out.head = (g.n_heads > 0u || g.head_at_start) ? 1u : 0u;
    return out;
}

// Phase 3 of one chunk that holds emitting ops: exclusive prefixes and the finished rows.
// carryR/carryQ: advance sums from the start of alignment g.a_lo up to the chunk start (used only if that
// alignment began before the chunk).  `out` is the slot of the chunk's first row.
__device__ __forceinline__ void chunk_emit(const ScanArgs& a, const ChunkGeom g, uint64_t c4, uint32_t lane, uint32_t evbits,
                                        uint32_t carryR, uint32_t carryQ, unsigned long long out) {
    const uint64_t c4end = min(c4 + CHUNK4, a.n4);
    const uint32_t a_lo = g.a_lo, a_hi = g.a_lo + g.n_heads;
    for (uint32_t al = a_lo; al <= a_hi; ++al) {
        const uint64_t lo = max(static_cast<uint64_t>(a.off4[al]), c4);
        const uint64_t hi = min(static_cast<uint64_t>(a.off4[al + 1]), c4end);
        if (hi <= lo) continue;
        uint32_t in_mask = 0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint64_t g4 = c4 + static_cast<uint64_t>(r) * 32u + lane;
            if (g4 >= lo && g4 < hi) in_mask |= 0xFu << (4 * r);
        }
        if (__ballot_sync(0xffffffffu, (evbits & in_mask) != 0u) == 0u) continue;
        const svb_aln_hdr h = a.hdr[al];
        // positions are relative to the alignment start: the carry only applies to a piece that began earlier
        const bool continued = static_cast<uint64_t>(a.off4[al]) < c4;
        const uint32_t baseR = continued ? carryR : 0u;
        const uint32_t baseQ = continued ? carryQ : 0u;
        int32_t clen = 0;
        if (h.tid < 0 || h.tid >= a.n_contig) {
            if (lane == 0) atomicOr(a.dev_status, DEV_ERR_BAD_TID);       // bam.getrname(tid) would raise (SVIM_intra.py:35)
        } else {
            clen = a.contig_len[h.tid];
        }
        uint32_t accR = 0, accQ = 0;                       // this lane's rows of the piece seen so far
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r) {
            const uint64_t g4 = c4 + static_cast<uint64_t>(r) * 32u + lane;
            const bool in = ((in_mask >> (4 * r)) & 1u) != 0u;
            const uint4 d = in ? reload_row(a.cigar, a.n4, c4, r, lane) : make_uint4(15u, 15u, 15u, 15u);
            uint32_t rr = 0, qq = 0, e0 = 0, n0 = 0;
            decode_op(d.x, a.min16, rr, qq, e0, n0, 1u);
            decode_op(d.y, a.min16, rr, qq, e0, n0, 1u);
            decode_op(d.z, a.min16, rr, qq, e0, n0, 1u);
            decode_op(d.w, a.min16, rr, qq, e0, n0, 1u);
            const uint32_t rowbits = in ? ((evbits >> (4 * r)) & 0xFu) : 0u;
            uint32_t bal = __ballot_sync(0xffffffffu, rowbits != 0u);
            while (bal) {
                const int L = __ffs(bal) - 1;
                bal &= bal - 1u;
                const bool earlier = static_cast<int>(lane) < L;
                const uint32_t preR = __reduce_add_sync(0xffffffffu, accR + (earlier ? rr : 0u));
                const uint32_t preQ = __reduce_add_sync(0xffffffffu, accQ + (earlier ? qq : 0u));
                const uint32_t n_emit = __popc(__shfl_sync(0xffffffffu, rowbits, L));
                if (static_cast<int>(lane) == L) {
                    uint32_t pr = baseR + preR, pq = baseQ + preQ;
                    unsigned long long slot = out;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t x = k == 0 ? d.x : k == 1 ? d.y : k == 2 ? d.z : d.w;
                        const uint32_t op = x & 15u, len = x >> 4;
                        if ((rowbits >> k) & 1u) {
                            if (slot < a.cap) {
                                // SVIM_intra.py:38-43 + the clamps of SVCandidate.py:44-46,134-136
                                const long long start = static_cast<long long>(h.pos) + pr;
                                const long long end = start + len;
                                const int32_t cs = static_cast<int32_t>(max(0ll, start));
                                const int32_t ce = static_cast<int32_t>(min(static_cast<long long>(clen), end));
                                const bool del = op == 2u;
                                if (!del && h.l_seq == 0u) atomicOr(a.dev_status, DEV_ERR_ASSERT);   // query_sequence is None
                                const uint32_t seq_len = (del || pq >= h.l_seq) ? 0u : min(len, h.l_seq - pq);
                                const unsigned long long ordinal = (static_cast<unsigned long long>(al) << 32) |
                                                                   static_cast<unsigned long long>((g4 - a.off4[al]) * 4u + k);
                                // svb_row as four 16-byte stores (field order of include/svimasm_b200.h)
                                uint4 w0, w1, w2, w3;
                                w0.x = (del ? SVB_DEL : SVB_INS) | (static_cast<uint32_t>(SVB_GT_HOM) << 16) | (a.hap << 24);
                                w0.y = del ? static_cast<uint32_t>(h.tid) : 0xFFFFFFFFu;     // src_tid
                                w0.z = del ? static_cast<uint32_t>(cs) : 0u;                   // src_start
                                w0.w = del ? static_cast<uint32_t>(ce) : 0u;                   // src_end
                                w1.x = del ? 0xFFFFFFFFu : static_cast<uint32_t>(h.tid);     // dst_tid
                                w1.y = del ? 0u : static_cast<uint32_t>(cs);                   // dst_start
                                w1.z = del ? 0u : static_cast<uint32_t>(ce);                   // dst_end
                                w1.w = 0u;                                                      // copies
                                w2.x = al;                                                      // aln_idx
                                w2.y = pq;                                                      // seq_pos (= pos_read)
                                w2.z = seq_len;
                                w2.w = 0xFFFFFFFFu;                                             // mate_aln
                                w3.x = static_cast<uint32_t>(ordinal);
                                w3.y = static_cast<uint32_t>(ordinal >> 32);
                                w3.z = 0u;
                                w3.w = 0u;
                                uint4* dst = reinterpret_cast<uint4*>(a.rows + slot);
                                dst[0] = w0; dst[1] = w1; dst[2] = w2; dst[3] = w3;
                            }
                            ++slot;
                        }
                        if ((REF_MASK >> op) & 1u) pr += len;
                        if ((READ_MASK >> op) & 1u) pq += len;
                    }
                }
                out += n_emit;
            }
            accR += rr;                                   // rr, qq are 0 for rows outside the piece (pad ops)
            accQ += qq;
        }
    }
}

struct Snap {
    uint32_t R, Q, head, cnt;
};

template <bool USE_TMA, int G>
__global__ void __launch_bounds__(THREADS, USE_TMA ? 3 : 4) cigar_scan_kernel(const ScanArgs a) {
    constexpr int RUN4 = WARPS * G * CHUNK4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* s_ring = reinterpret_cast<uint4*>(smem_raw);       // [WARPS][STAGES][CHUNK4] when USE_TMA
    __shared__ __align__(8) unsigned long long s_mbar[WARPS][STAGES];
    __shared__ uint2 s_lut[16];      // per op code: x = multiplier (bit 0 read advance, bit 31 reference advance), y = rare threshold
    __shared__ uint32_t s_run_id;
    __shared__ uint32_t s_ev[WARPS][G][32];
    __shared__ Snap s_snap[WARPS][G];          // carry of each chunk relative to the warp's first chunk
    __shared__ Snap s_warp[WARPS];             // aggregate of each warp's G chunks
    __shared__ Snap s_wcarry[WARPS];           // carry of each warp relative to the run start
    __shared__ uint32_t s_lbR[WARPS], s_lbQ[WARPS], s_lbFlags[WARPS];
    __shared__ unsigned long long s_lbCnt[WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid >= 32u && tid < 48u) {
        const uint32_t op = tid - 32u;
        uint2 e;
        e.x = (((REF_MASK >> op) & 1u) << 31) | ((READ_MASK >> op) & 1u);
        e.y = ((INDEL_MASK >> op) & 1u) ? a.min16 : (((NH_MASK >> op) & 1u) ? 16u : 0xFFFFFFFFu);
        s_lut[op] = e;
    }
    if (tid == 0) s_run_id = atomicAdd(a.ticket, 1u);       // runs are claimed in scheduling order: look-back cannot deadlock
    if (USE_TMA && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_mbar[warp][s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint32_t run = s_run_id;
    const uint64_t w4 = static_cast<uint64_t>(run) * RUN4 + static_cast<uint64_t>(warp) * (G * CHUNK4);   // this warp's first uint4
    uint4* my_ring = s_ring + static_cast<size_t>(warp) * STAGES * CHUNK4;

    auto chunk_bytes = [&](int j) -> uint32_t {
        const uint64_t c4 = w4 + static_cast<uint64_t>(j) * CHUNK4;
        return c4 >= a.n4 ? 0u : static_cast<uint32_t>(min(static_cast<uint64_t>(CHUNK4), a.n4 - c4)) * 16u;
    };
    if (USE_TMA && lane == 0) {
#pragma unroll
        for (int j = 0; j < STAGES && j < G; ++j) {
            const uint32_t bytes = chunk_bytes(j);
            if (bytes) tma_chunk(my_ring + j * CHUNK4, a.cigar + w4 + static_cast<uint64_t>(j) * CHUNK4, bytes, &s_mbar[warp][j]);
        }
    }

    // ---- phase 1: this warp's G chunks, carry in registers
    uint32_t accR = 0, accQ = 0, accHead = 0, accCnt = 0;
    ChunkGeom g_next;
    g_next.a_lo = 0; g_next.n_heads = 0; g_next.split_rel = 0; g_next.head_at_start = 0;
    if (w4 < a.n4) g_next = a.geom[w4 / CHUNK4];
#pragma unroll 1
    for (int j = 0; j < G; ++j) {
        const uint64_t c4 = w4 + static_cast<uint64_t>(j) * CHUNK4;
        const bool live = c4 < a.n4;
        const ChunkGeom g = g_next;
        if (j + 1 < G && c4 + CHUNK4 < a.n4) g_next = a.geom[c4 / CHUNK4 + 1];        // prefetch next chunk's geometry
        if (lane == 0) {
            Snap s;
            s.R = accR; s.Q = accQ; s.head = accHead; s.cnt = accCnt;
            s_snap[warp][j] = s;
        }
        uint32_t evbits = 0;
        if (live) {
            ChunkResult res;
            if (USE_TMA) {
                mbar_wait(&s_mbar[warp][j % STAGES], static_cast<uint32_t>(j / STAGES) & 1u);
                const uint4* buf = my_ring + (j % STAGES) * CHUNK4;
                const uint32_t here = chunk_bytes(j) / 16u;
                auto from_ring = [&](int r) -> uint4 {
                    const uint32_t i4 = static_cast<uint32_t>(r) * 32u + lane;
                    return i4 < here ? buf[i4] : make_uint4(15u, 15u, 15u, 15u);
                };
                res = chunk_phase1(a, s_lut, g, c4, lane, from_ring, from_ring);
            } else {
                uint4 v[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const uint64_t g4 = c4 + static_cast<uint64_t>(r) * 32u + lane;
                    v[r] = g4 < a.n4 ? ldg_stream(a.cigar + g4) : make_uint4(15u, 15u, 15u, 15u);
                }
                res = chunk_phase1(a, s_lut, g, c4, lane, [&](int r) -> uint4 { return v[r]; },
                                   [&](int r) -> uint4 { return reload_row(a.cigar, a.n4, c4, r, lane); });
            }
            evbits = res.evbits;
            if (res.head) { accR = res.tailR; accQ = res.tailQ; accHead = 1u; }
            else { accR += res.tailR; accQ += res.tailQ; }
            accCnt += res.cnt;
        }
        s_ev[warp][j][lane] = evbits;
        if (USE_TMA) {
            __syncwarp();                                  // every lane is done with this stage
            if (lane == 0 && j + STAGES < G) {
                const uint32_t bytes = chunk_bytes(j + STAGES);
                if (bytes) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    tma_chunk(my_ring + (j % STAGES) * CHUNK4, a.cigar + c4 + static_cast<uint64_t>(STAGES) * CHUNK4, bytes,
                              &s_mbar[warp][j % STAGES]);
                }
            }
        }
    }
    if (lane == 0) {
        Snap s;
        s.R = accR; s.Q = accQ; s.head = accHead; s.cnt = accCnt;
        s_warp[warp] = s;
    }
    __syncthreads();

    // ---- phase 2: run aggregate, publish, look-back by all 8 warps (256 predecessors per round), publish prefix
    uint32_t runR = 0, runQ = 0, runHead = 0, runCnt = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {                      // every thread computes the same serial combine
        if (tid == 0) {
            Snap c;
            c.R = runR; c.Q = runQ; c.head = runHead; c.cnt = runCnt;
            s_wcarry[w] = c;
        }
        const Snap s = s_warp[w];
        if (s.head) { runR = s.R; runQ = s.Q; runHead = 1u; }
        else { runR += s.R; runQ += s.Q; }
        runCnt += s.cnt;
    }
    __syncthreads();                                       // s_wcarry is read by every warp in phase 3
    RunStatus* mine = a.status + run;
    if (run != 0 && tid == 0) {
        st_relaxed(&mine->w_ref, (ST_AGG << 62) | (static_cast<unsigned long long>(runHead) << 32) | runR);
        st_relaxed(&mine->w_read, (ST_AGG << 62) | runQ);
        st_relaxed(&mine->w_cnt, (ST_AGG << 62) | runCnt);
    }
    uint32_t carryR = 0, carryQ = 0;
    unsigned long long excl = 0;
    if (run != 0) {
        bool sums_done = false;
        int64_t look = static_cast<int64_t>(run) - 1;
        while (true) {
            const int64_t t = look - static_cast<int64_t>(warp * 32u + lane);
            unsigned long long wr = (ST_PREFIX << 62), wq = (ST_PREFIX << 62), wc = (ST_PREFIX << 62);   // virtual run -1
            if (t >= 0) {
                const RunStatus* ts = a.status + t;
                while (true) {
                    wr = ld_relaxed(&ts->w_ref);
                    wq = ld_relaxed(&ts->w_read);
                    wc = ld_relaxed(&ts->w_cnt);
                    const unsigned long long s = wr >> 62;
                    if (s != ST_INVALID && s == (wq >> 62) && s == (wc >> 62)) break;
                }
            }
            const bool is_prefix = (wr >> 62) == ST_PREFIX;
            const bool stops_sum = is_prefix || ((wr >> 32) & 1ull);
            const uint32_t pmask = __ballot_sync(0xffffffffu, is_prefix);
            const uint32_t smask = __ballot_sync(0xffffffffu, stops_sum);
            const int k_cnt = pmask ? (__ffs(pmask) - 1) : 31;
            const int k_sum = smask ? (__ffs(smask) - 1) : 31;
            unsigned long long csum = (static_cast<int>(lane) <= k_cnt) ? (wc & ((1ull << 62) - 1ull)) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            const bool take = static_cast<int>(lane) <= k_sum;
            const uint32_t sR = __reduce_add_sync(0xffffffffu, take ? static_cast<uint32_t>(wr) : 0u);
            const uint32_t sQ = __reduce_add_sync(0xffffffffu, take ? static_cast<uint32_t>(wq) : 0u);
            if (lane == 0) {
                s_lbCnt[warp] = csum;
                s_lbR[warp] = sR;
                s_lbQ[warp] = sQ;
                s_lbFlags[warp] = (pmask ? 1u : 0u) | (smask ? 2u : 0u);
            }
            __syncthreads();
            bool cnt_done = false;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {                // nearest window first; every thread computes the same
                if (!cnt_done) {
                    excl += s_lbCnt[w];
                    if (!sums_done) {
                        carryR += s_lbR[w];
                        carryQ += s_lbQ[w];
                        if (s_lbFlags[w] & 2u) sums_done = true;
                    }
                    if (s_lbFlags[w] & 1u) cnt_done = true;
                }
            }
            __syncthreads();                                 // s_lb* are rewritten in the next round
            if (cnt_done) break;
            look -= WARPS * 32;
        }
    }
    if (tid == 0) {
        const uint32_t incR = runHead ? runR : carryR + runR;
        const uint32_t incQ = runHead ? runQ : carryQ + runQ;
        st_relaxed(&mine->w_ref, (ST_PREFIX << 62) | (1ull << 32) | incR);
        st_relaxed(&mine->w_read, (ST_PREFIX << 62) | incQ);
        st_relaxed(&mine->w_cnt, (ST_PREFIX << 62) | (excl + runCnt));
        if (run == a.n_runs - 1u) *a.total = excl + runCnt;
    }

    // ---- phase 3: chunks that hold an emitting op
    const Snap wc = s_wcarry[warp];
#pragma unroll 1
    for (int j = 0; j < G; ++j) {
        const uint32_t evbits = s_ev[warp][j][lane];
        if (__ballot_sync(0xffffffffu, evbits != 0u) == 0u) continue;
        const uint64_t c4 = w4 + static_cast<uint64_t>(j) * CHUNK4;
        const Snap sn = s_snap[warp][j];
        // advance sums since the start of the alignment that spans the chunk start: chunk snapshot, then the
        // warp's carry inside the run, then the run's carry-in -- each level only if no head occurred closer
        uint32_t cR = sn.R, cQ = sn.Q;
        if (!sn.head) {
            cR += wc.R; cQ += wc.Q;
            if (!wc.head) { cR += carryR; cQ += carryQ; }
        }
        chunk_emit(a, a.geom[c4 / CHUNK4], c4, lane, evbits, cR, cQ, excl + wc.cnt + sn.cnt);
    }
}

// geometry of every chunk (one thread per chunk, two binary searches over off4)
__global__ void chunk_index_kernel(const uint32_t* __restrict__ off4, uint32_t n_aln, uint64_t n4, uint64_t n_chunks,
                                   ChunkGeom* __restrict__ geom) {
    const uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const uint64_t x = c * CHUNK4, end = min(x + CHUNK4, n4);
    uint32_t lo = 0, hi = n_aln;                 // upper_bound(off4[0..n_aln), x): first run that starts after x
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        if (static_cast<uint64_t>(off4[mid]) <= x) lo = mid + 1u; else hi = mid;
    }
    const uint32_t a_lo = lo ? lo - 1u : 0u;
    uint32_t lo2 = lo, hi2 = n_aln;              // lower_bound(off4, end): first run that starts at or after `end`
    while (lo2 < hi2) {
        const uint32_t mid = lo2 + (hi2 - lo2) / 2u;
        if (static_cast<uint64_t>(off4[mid]) < end) lo2 = mid + 1u; else hi2 = mid;
    }
    const uint32_t a_hi = lo2 ? lo2 - 1u : 0u;
    ChunkGeom g;
    g.a_lo = a_lo;
    g.n_heads = a_hi > a_lo ? a_hi - a_lo : 0u;
    g.split_rel = g.n_heads ? static_cast<uint32_t>(off4[a_lo + 1u] - x) : static_cast<uint32_t>(end - x);
    g.head_at_start = static_cast<uint64_t>(off4[a_lo]) == x ? 1u : 0u;
    geom[c] = g;
}

}  // namespace

int launch_build_chunk_index(svb_ctx* ctx, svb_records* rec) {
    const uint64_t n_chunks = (rec->n4 + CHUNK4 - 1) / CHUNK4;
    if (n_chunks == 0) return SVB_OK;
    SVB_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&rec->d_chunk_first), n_chunks * sizeof(ChunkGeom), ctx->stream));
    const unsigned blocks = static_cast<unsigned>((n_chunks + 255) / 256);
    chunk_index_kernel<<<blocks, 256, 0, ctx->stream>>>(rec->d_off4, rec->n_aln, rec->n4, n_chunks,
                                                        reinterpret_cast<ChunkGeom*>(rec->d_chunk_first));
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}

template <int G>
static int launch_scan_g(svb_ctx* ctx, const svb_records* rec, ScanArgs a) {
    constexpr int RUN4 = WARPS * G * CHUNK4;
    const uint64_t n_runs64 = (rec->n4 + RUN4 - 1) / RUN4;
    if (n_runs64 > 0x7fffffffull) return svb_fail(ctx, SVB_ERR_ARG, "too many CIGAR ops for one launch");
    const uint32_t n_runs = static_cast<uint32_t>(n_runs64);
    const size_t need = sizeof(RunStatus) * n_runs + 256;
    unsigned char* scratch = static_cast<unsigned char*>(svb_scratch(ctx, need));
    if (!scratch) return svb_fail(ctx, SVB_ERR_NOMEM, "run status scratch");
    SVB_CUDA(ctx, cudaMemsetAsync(scratch, 0, need, ctx->stream));
    a.n_runs = n_runs;
    a.ticket = reinterpret_cast<unsigned int*>(scratch);
    a.status = reinterpret_cast<RunStatus*>(scratch + 256);
    KernelTimer timer(ctx, SVB_K_CIGAR_SCAN);
    if (ctx->scan_variant == 0) {
        const size_t smem = static_cast<size_t>(WARPS) * STAGES * CHUNK4 * sizeof(uint4);
        static bool attr_set = false;
        if (!attr_set) {
            SVB_CUDA(ctx, cudaFuncSetAttribute(cigar_scan_kernel<true, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(smem)));
            attr_set = true;
        }
        cigar_scan_kernel<true, G><<<n_runs, THREADS, smem, ctx->stream>>>(a);
    } else {
        cigar_scan_kernel<false, G><<<n_runs, THREADS, 16, ctx->stream>>>(a);
    }
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}

int launch_cigar_scan(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, ScanOutput out) {
    SVB_CUDA(ctx, cudaMemsetAsync(out.d_count, 0, sizeof(unsigned long long), ctx->stream));
    if (rec->n_aln) SVB_CUDA(ctx, cudaMemsetAsync(rec->d_aln_sum, 0, sizeof(uint4) * rec->n_aln, ctx->stream));
    if (rec->n4 == 0) return SVB_OK;
    ScanArgs a;
    a.cigar = rec->d_cigar;
    a.n4 = rec->n4;
    a.off4 = rec->d_off4;
    a.geom = reinterpret_cast<const ChunkGeom*>(rec->d_chunk_first);
    a.hdr = rec->d_hdr;
    a.contig_len = rec->d_contig_len;
    a.n_aln = rec->n_aln;
    a.n_contig = rec->n_contig;
    a.n_runs = 0;
    a.min_mapq = p->min_mapq;
    const long long m = p->min_sv_size < 0 ? 0 : p->min_sv_size;
    a.min16 = m >= (1ll << 28) ? 0xFFFFFFFFu : static_cast<uint32_t>(m << 4);
    a.hap = static_cast<uint32_t>(hap);
    a.aln_sum = rec->d_aln_sum;
    a.ticket = nullptr;
    a.status = nullptr;
    a.rows = out.rows;
    a.cap = out.cap;
    a.total = out.d_count;
    a.dev_status = ctx->d_status;
    // run length: as long as possible while at least two full waves of CTAs (4 per SM) remain
    const uint64_t two_waves = 2ull * 4ull * static_cast<uint64_t>(ctx->sm_count);
    const uint64_t chunks = (rec->n4 + CHUNK4 - 1) / CHUNK4;
    if (chunks / (WARPS * 16) >= two_waves) return launch_scan_g<16>(ctx, rec, a);
    if (chunks / (WARPS * 8) >= two_waves) return launch_scan_g<8>(ctx, rec, a);
    if (chunks / (WARPS * 4) >= two_waves) return launch_scan_g<4>(ctx, rec, a);
    return launch_scan_g<2>(ctx, rec, a);
}
