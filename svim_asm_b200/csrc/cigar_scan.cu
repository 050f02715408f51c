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
//   * the array is cut into UNITS of G chunks x 1024 ops (G = 16: 64 KB); one warp owns one unit, so its
//     running "sum since the last alignment head" stays in registers and warps never wait for each other:
//     no block barrier, no inter-block look-back in the streaming kernel;
//   * a chunk (4 KB) is staged to shared memory by a per-warp ring of TMA bulk copies (cp.async.bulk +
//     mbarrier, two stages: the next chunk is in flight while this one is decoded; default), or, variant 1,
//     read with coalesced 128-bit LDG.nc into registers;
//   * per op: one 8-byte table entry {multiplier, threshold} from shared memory, one IMAD.WIDE into a
//     packed accumulator (read advance in bits 0..30, reference advance from bit 31 up), one compare
//     that flags the rare ops (I/D with len >= min_sv_size, N/H).  Exact event bits are recomputed
//     only for the rows where the flag fired (about 1 chunk in 8 holds one for human assemblies);
//   * a chunk that holds an emitting op writes its finished 64-byte rows at once, into a staging area at an
//     atomically reserved slot, tagged with (unit, index inside the unit).  Rows of the alignment that was
//     already running when the unit began (about 15 % of the rows) cannot know their offset yet: they are
//     flagged and carry their raw partial sums;
//   * two tiny kernels finish the job: a segmented scan over the per-unit aggregates (carry-in and row
//     base of every unit) and a per-row pass that moves each row to its final, stable position and
//     completes the flagged ones (clamps of SVCandidate.py:44-46,134-136 applied then).
// Per-alignment totals (reference span, read span, N and H bases), needed by the split-alignment
// walk for reference_end / infer_read_length (SVIM_inter.py:68-80), fall out as atomics.
//
// Bound: HBM bandwidth.  Algorithmic bytes: 4 B/op + 32 B/alignment + 64 B/emitted row.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
#ifndef K2_ROWS
#define K2_ROWS 8
#endif
constexpr int ROWS = K2_ROWS;              // uint4 per lane per chunk (8: 4 KB chunks, 4: 2 KB chunks)
constexpr int CHUNK4 = 32 * ROWS;          // 256 uint4  = 1024 ops = 4 KB per chunk
constexpr int STAGES = 2;                  // TMA ring depth per warp
#ifndef K2_TMA_CTAS
#define K2_TMA_CTAS 3                      // resident CTAs per SM the TMA variant is compiled for (64 KB ring each)
#endif
#ifndef K2_LDG_CTAS
#define K2_LDG_CTAS 4                      // resident CTAs per SM the LDG variant is compiled for (64 registers)
#endif
// G = chunks per unit (template parameter, picked at launch): one warp walks G chunks of 4 KB with its carry in
// registers.  Long units amortise the per-unit work; short ones keep every SM busy on small inputs.

constexpr uint32_t REF_MASK = (1u << 0) | (1u << 2) | (1u << 7) | (1u << 8);              // M D = X  (SVIM_intra.py:15,23,28)
constexpr uint32_t READ_MASK = (1u << 0) | (1u << 1) | (1u << 4) | (1u << 7) | (1u << 8); // M I S = X (SVIM_intra.py:16,20,26,29)
constexpr uint32_t INDEL_MASK = (1u << 1) | (1u << 2);
constexpr uint32_t NH_MASK = (1u << 3) | (1u << 5);

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
    uint32_t n_units;
    int32_t min_mapq;
    uint32_t min16;           // min_sv_size << 4 : (x >= min16) <=> (len >= min_sv_size)
    uint32_t mul28;           // 1 << 28 (see op_len)
    uint32_t hap;
    uint4* aln_sum;
    uint4* unit_agg;          // [n_units] x: ref sum since last head, y: read sum, z: head seen, w: rows emitted
    svb_row* rows;            // staging rows (slots reserved with atomics)
    unsigned long long cap;
    unsigned long long* total;     // number of staged rows (may exceed cap)
    unsigned int* unit_counter;    // next unit to hand out (persistent warps claim units)
    uint32_t* dev_status;
};

constexpr uint8_t ROW_NEEDS_CARRY = 0x80;     // staging flag: offsets are relative to the unit start

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// mbarrier wait / TMA bulk copy on shared-window addresses computed once per warp
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
    uint32_t ready = 0;
    while (!ready) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ready) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_chunk_addr(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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

// Decode table, one entry per op code: x = multiplier (bit 0: the op advances the read, bit 31: it advances the
// reference), y = threshold of the rare test on the PACKED op (I/D: min_sv_size << 4, N/H: 16 i.e. len != 0, else never).
// File scope so that every lookup is LDS.64 [offset + table] (a pointer parameter costs an extra add per op).
__shared__ uint2 s_lut[16];

// fast path of one uint4 (4 ops): packed advance sums (read advance in bits 0..30, reference advance from bit 31 up)
// added to `acc`, and the rare flag, accumulated over the whole chunk.  Lengths have 28 bits, so EIGHT ops (two uint4)
// cannot overflow the 31-bit field: 8 x (2^28 - 1) < 2^31.
#ifndef K2_LEN_BY_MULHI
#define K2_LEN_BY_MULHI 0      // measured on B200: IMAD.HI costs more than it saves (0.164 ms vs 0.155 ms per whole-genome launch)
#endif
// op length = x >> 4.  The integer ALU pipe (LOP3 / SHF / ISETP) is the busiest unit of this kernel, the FMA pipe is
// half idle, so the shift is done there: mulhi(x, 2^28) == x >> 4, with 2^28 a kernel argument the compiler cannot fold.
__device__ __forceinline__ uint32_t op_len(uint32_t x, uint32_t mul28) {
#if K2_LEN_BY_MULHI
    return __umulhi(x, mul28);
#else
    return x >> 4;
#endif
}

__device__ __forceinline__ unsigned long long fast_row(const uint4 d, unsigned long long acc, bool& rare, uint32_t mul28) {
    const uint2 e0 = s_lut[d.x & 15u], e1 = s_lut[d.y & 15u], e2 = s_lut[d.z & 15u], e3 = s_lut[d.w & 15u];
    acc += static_cast<unsigned long long>(op_len(d.x, mul28)) * e0.x;
    acc += static_cast<unsigned long long>(op_len(d.y, mul28)) * e1.x;
    acc += static_cast<unsigned long long>(op_len(d.z, mul28)) * e2.x;
    acc += static_cast<unsigned long long>(op_len(d.w, mul28)) * e3.x;
    rare |= (d.x >= e0.y) | (d.y >= e1.y) | (d.z >= e2.y) | (d.w >= e3.y);
    return acc;
}

// row `r` of the chunk that starts at uint4 index c4, re-read through L2 (rare paths only)
__device__ __forceinline__ uint4 reload_row(const uint4* cigar, uint64_t c4, int r, uint32_t lane) {
    return cigar[c4 + static_cast<uint64_t>(r) * 32u + lane];        // the buffer is padded to whole units (op 15)
}

// What a warp carries from chunk to chunk inside its unit.  The advance sums of the alignment piece that is running
// stay LANE-LOCAL (each lane adds up its own rows); they are reduced over the warp only when something needs the
// total: an alignment head, an emitting op, the end of the unit.
struct WarpState {
    uint32_t laneR, laneQ;     // this lane's share of the advance sums since the last head inside the unit / the last flush
    uint32_t cur_aln;          // alignment those sums belong to
    uint32_t head;             // 1 once an alignment head was seen inside the unit
    uint32_t cnt;              // rows emitted by the unit so far
};

// add warp totals to the alignment the running piece belongs to
__device__ __forceinline__ void add_to_alignment(const ScanArgs& a, uint32_t aln, uint32_t R, uint32_t Q, uint32_t lane) {
    if (lane == 0 && (R | Q)) {
        atomicAdd(&a.aln_sum[aln].x, R);
        atomicAdd(&a.aln_sum[aln].y, Q);
    }
}

__device__ __forceinline__ void chunk_emit(const ScanArgs& a, const ChunkGeom g, uint64_t c4, uint32_t lane,
                                           uint32_t evbits, const uint4* rows4, uint32_t carryR, uint32_t carryQ,
                                           bool resolved, uint32_t unit, uint32_t local0, unsigned long long base);

// Everything a chunk may need beyond the plain sums: alignment heads inside the chunk, ops that pass their class
// threshold (emitting I/D, N, H).  About one chunk in five comes here.  totR/totQ: this lane's sums over the whole
// chunk (from the fast loop); `rare`: this lane saw an op over its threshold; `rowsrc(r)`: the lane's uint4 of row r
// (r is a run-time value here); `stage_rows()` makes the chunk available in shared memory and returns it.
template <typename RowSrc, typename StageRows>
__device__ __forceinline__ void chunk_slow(const ScanArgs& a, WarpState& st, const ChunkGeom g, uint64_t c4, uint32_t lane, uint32_t unit,
                                           uint32_t totR, uint32_t totQ, bool rare, bool any_rare, RowSrc rowsrc, StageRows stage_rows) {
    const uint64_t c4end = min(c4 + CHUNK4, a.n4);
    const uint32_t a_lo = g.a_lo, a_hi = g.a_lo + g.n_heads, n_pieces = g.n_heads + 1u;
    // rows of this lane before the first head inside the chunk: 32 r + lane < split_rel
    const int rows_before = (static_cast<int>(g.split_rel) - static_cast<int>(lane) + 31) / 32;
    const int nb = rows_before <= 0 ? 0 : (rows_before >= ROWS ? ROWS : rows_before);
    const uint32_t first_mask = nb >= 8 ? 0xFFFFFFFFu : ((1u << (4u * static_cast<uint32_t>(nb))) - 1u);   // ROWS <= 8

    uint32_t nhbits = 0, evbits = 0;
    if (any_rare && rare) {               // exact bits of this lane's 32 ops
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r) {
            const uint4 d = rowsrc(r);
            uint32_t r0 = 0, q0 = 0;
            decode_op(d.x, a.min16, r0, q0, evbits, nhbits, 1u << (4 * r + 0));
            decode_op(d.y, a.min16, r0, q0, evbits, nhbits, 1u << (4 * r + 1));
            decode_op(d.z, a.min16, r0, q0, evbits, nhbits, 1u << (4 * r + 2));
            decode_op(d.w, a.min16, r0, q0, evbits, nhbits, 1u << (4 * r + 3));
        }
    }
    __syncwarp();
    // advance sums since the last head, up to the start of this chunk (what an emitted row of the running piece adds)
    const uint32_t carryR = __reduce_add_sync(0xffffffffu, st.laneR), carryQ = __reduce_add_sync(0xffffffffu, st.laneQ);
    const bool resolved = st.head != 0u;

    // per-piece bookkeeping; afterwards st.lane* hold the LAST piece's sums inside this chunk
    uint32_t headR = 0, headQ = 0;        // n_pieces == 2: this lane's sums over the rows of the first piece
    uint32_t lastR = 0, lastQ = 0;        // n_pieces > 2: warp totals of the last piece
    if (n_pieces == 1u) {
        if (any_rare && !record_passes(a.hdr[a_lo], a.min_mapq)) evbits = 0u;          // SVIM_COLLECT.py:71
    } else if (n_pieces == 2u) {
#pragma unroll 1
        for (int r = 0; r < nb; ++r) {
            bool ignore = false;
            const unsigned long long acc = fast_row(rowsrc(r), 0ull, ignore, a.mul28);
            headR += static_cast<uint32_t>(acc >> 31);
            headQ += static_cast<uint32_t>(acc) & 0x7FFFFFFFu;
        }
        if (any_rare) {
            const bool pass0 = record_passes(a.hdr[a_lo], a.min_mapq), pass1 = record_passes(a.hdr[a_hi], a.min_mapq);
            evbits &= (pass0 ? first_mask : 0u) | (pass1 ? ~first_mask : 0u);
        }
    } else {
        // generic: several alignment heads inside one 1024-op chunk (short alignments)
        uint32_t keep = 0;
        for (uint32_t al = a_lo; al <= a_hi; ++al) {
            const uint64_t lo = max(static_cast<uint64_t>(a.off4[al]), c4);
            const uint64_t hi = min(static_cast<uint64_t>(a.off4[al + 1]), c4end);
            uint32_t sR = 0, sQ = 0, in_mask = 0;
#pragma unroll 1
            for (int r = 0; r < ROWS; ++r) {
                const uint64_t g4 = c4 + static_cast<uint64_t>(r) * 32u + lane;
                if (g4 >= lo && g4 < hi) {
                    const uint4 d = rowsrc(r);
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
            if (al == a_lo) {              // the running piece ends here: its carry goes with it
                sR += carryR;
                sQ += carryQ;
            }
            if (al == a_hi) {              // the last piece keeps running
                lastR = sR;
                lastQ = sQ;
            } else if (lane == 0 && (sR | sQ)) {
                atomicAdd(&a.aln_sum[al].x, sR);
                atomicAdd(&a.aln_sum[al].y, sQ);
            }
        }
        evbits &= keep;
    }
    if (any_rare) {
        const uint32_t cnt = __reduce_add_sync(0xffffffffu, __popc(evbits));
        // rare: N / H ops feed reference_end / infer_read_length of the split-alignment walk
        if (__ballot_sync(0xffffffffu, nhbits != 0u)) {
            uint32_t bits = nhbits;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1u;
                const uint64_t g4 = c4 + static_cast<uint64_t>(b >> 2) * 32u + lane;
                uint32_t al = a_lo;
                while (al < a_hi && static_cast<uint64_t>(a.off4[al + 1]) <= g4) ++al;
                const uint4 d = rowsrc(b >> 2);
                const uint32_t x = (b & 3) == 0 ? d.x : (b & 3) == 1 ? d.y : (b & 3) == 2 ? d.z : d.w;
                if ((x & 15u) == 3u) atomicAdd(&a.aln_sum[al].z, x >> 4);
                else atomicAdd(&a.aln_sum[al].w, x >> 4);
            }
        }
        if (cnt) {
            const uint4* rows4 = stage_rows();
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.total, static_cast<unsigned long long>(cnt));
            base = __shfl_sync(0xffffffffu, base, 0);
            chunk_emit(a, g, c4, lane, evbits, rows4, carryR, carryQ, resolved, unit, st.cnt, base);
            st.cnt += cnt;
        }
    }
    // hand the sums on
    if (n_pieces == 1u) {
        st.laneR += totR;
        st.laneQ += totQ;
    } else if (n_pieces == 2u) {
        add_to_alignment(a, a_lo, carryR + __reduce_add_sync(0xffffffffu, headR), carryQ + __reduce_add_sync(0xffffffffu, headQ), lane);
        st.laneR = totR - headR;
        st.laneQ = totQ - headQ;
        st.cur_aln = a_hi;
        st.head = 1u;
    } else {
        st.laneR = lane == 0 ? lastR : 0u;      // (the first piece, carry included, was added to its alignment above)
        st.laneQ = lane == 0 ? lastQ : 0u;
        st.cur_aln = a_hi;
        st.head = 1u;
    }
}

// Rows of one chunk that holds emitting ops.  carryR/carryQ: advance sums since the last alignment head inside
// the unit (or since the unit start when `resolved` is false: then rows of the alignment that spans the chunk
// start are flagged ROW_NEEDS_CARRY and finished by finalize_rows_kernel).  Rows go to staging slots
// base, base+1, ...; `local0` is the index of the chunk's first row inside its unit.  `rows4` = the chunk in
// shared memory (ring stage or spill buffer), 256 uint4.
__device__ __forceinline__ void chunk_emit(const ScanArgs& a, const ChunkGeom g, uint64_t c4, uint32_t lane,
                                           uint32_t evbits, const uint4* rows4, uint32_t carryR, uint32_t carryQ,
                                           bool resolved, uint32_t unit, uint32_t local0, unsigned long long base) {
    const uint64_t c4end = min(c4 + CHUNK4, a.n4);
    const uint32_t a_lo = g.a_lo, a_hi = g.a_lo + g.n_heads;
    uint32_t emitted = 0;
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
        const bool needs_carry = continued && !resolved;
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
            const uint32_t i4 = static_cast<uint32_t>(r) * 32u + lane;
            const bool in = ((in_mask >> (4 * r)) & 1u) != 0u;
            const uint4 d = in ? rows4[i4] : make_uint4(15u, 15u, 15u, 15u);
            bool rare0 = false;
            const unsigned long long packed = fast_row(d, 0ull, rare0, a.mul28);
            const uint32_t rr = static_cast<uint32_t>(packed >> 31), qq = static_cast<uint32_t>(packed) & 0x7FFFFFFFu;
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
                    uint32_t idx = emitted;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t x = k == 0 ? d.x : k == 1 ? d.y : k == 2 ? d.z : d.w;
                        const uint32_t op = x & 15u, len = x >> 4;
                        if ((rowbits >> k) & 1u) {
                            const unsigned long long slot = base + idx;
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
                                // svb_row as four 16-byte stores (field order of include/svimasm_b200.h); staging uses
                                // copies = index inside the unit, mate_aln = unit, reserved0 = (len << 32) | raw pos_ref
                                uint4 w0, w1, w2, w3;
                                w0.x = (del ? SVB_DEL : SVB_INS) | (needs_carry ? (static_cast<uint32_t>(ROW_NEEDS_CARRY) << 8) : 0u) |
                                       (static_cast<uint32_t>(SVB_GT_HOM) << 16) | (a.hap << 24);
                                w0.y = del ? static_cast<uint32_t>(h.tid) : 0xFFFFFFFFu;     // src_tid
                                w0.z = del ? static_cast<uint32_t>(cs) : 0u;                   // src_start
                                w0.w = del ? static_cast<uint32_t>(ce) : 0u;                   // src_end
                                w1.x = del ? 0xFFFFFFFFu : static_cast<uint32_t>(h.tid);     // dst_tid
                                w1.y = del ? 0u : static_cast<uint32_t>(cs);                   // dst_start
                                w1.z = del ? 0u : static_cast<uint32_t>(ce);                   // dst_end
                                w1.w = local0 + idx;                                            // (staging) index inside the unit
                                w2.x = al;                                                      // aln_idx
                                // seq_pos = pos_read; an insertion's is clamped like the python slice start (SVIM_intra.py:42)
                                w2.y = (del || needs_carry) ? pq : min(pq, h.l_seq);
                                w2.z = seq_len;
                                w2.w = unit;                                                    // (staging) unit
                                w3.x = static_cast<uint32_t>(ordinal);
                                w3.y = static_cast<uint32_t>(ordinal >> 32);
                                w3.z = pr;                                                      // (staging) raw pos_ref
                                w3.w = len;                                                     // (staging) op length
                                uint4* dst = reinterpret_cast<uint4*>(a.rows + slot);
                                dst[0] = w0; dst[1] = w1; dst[2] = w2; dst[3] = w3;
                            }
                            ++idx;
                        }
                        if ((REF_MASK >> op) & 1u) pr += len;
                        if ((READ_MASK >> op) & 1u) pq += len;
                    }
                }
                emitted += n_emit;
            }
            accR += rr;                                   // rr, qq are 0 for rows outside the piece (pad ops)
            accQ += qq;
        }
    }
}

template <bool USE_TMA, int G>
__global__ void __launch_bounds__(THREADS, USE_TMA ? K2_TMA_CTAS : K2_LDG_CTAS) cigar_scan_kernel(const ScanArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // USE_TMA: [WARPS][STAGES][CHUNK4] ring.  LDG: [WARPS][CHUNK4] spill buffer, written only for chunks with events
    uint4* s_buf = reinterpret_cast<uint4*>(smem_raw);
    __shared__ __align__(8) unsigned long long s_mbar[WARPS][STAGES];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid >= 32u && tid < 48u) {
        const uint32_t op = tid - 32u;
        uint2 e;
        e.x = (((REF_MASK >> op) & 1u) << 31) | ((READ_MASK >> op) & 1u);
        e.y = ((INDEL_MASK >> op) & 1u) ? a.min16 : (((NH_MASK >> op) & 1u) ? 16u : 0xFFFFFFFFu);
        s_lut[op] = e;
    }
    if (USE_TMA && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_mbar[warp][s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();                 // the only block-wide barrier: the decode table
    uint4* my_buf = s_buf + static_cast<size_t>(warp) * (USE_TMA ? STAGES : 1) * CHUNK4;
    constexpr uint64_t UNIT4 = static_cast<uint64_t>(G) * CHUNK4;
    auto load_geom = [&](uint64_t chunk) -> ChunkGeom {                  // one 16-byte load
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.geom) + chunk);
        ChunkGeom g;
        g.a_lo = v.x; g.n_heads = v.y; g.split_rel = v.z; g.head_at_start = v.w;
        return g;
    };
    // Persistent warps: units are claimed from a counter, so every warp keeps streaming until the input is used up
    // (no tail of half-empty SMs, one table/barrier set-up per CTA).  The next unit is claimed while the current one
    // is being scanned, which lets the TMA ring run across the unit boundary without a bubble.
    auto claim = [&]() -> uint32_t {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(a.unit_counter, 1u);
        return __shfl_sync(0xffffffffu, u, 0);
    };
    // the CIGAR buffer is padded (op 15) to whole units of 16 chunks: a chunk that starts inside the data is complete
    auto chunk_exists = [&](uint32_t u, uint32_t j) -> bool {
        return u < a.n_units && static_cast<uint64_t>(u) * UNIT4 + static_cast<uint64_t>(j) * CHUNK4 < a.n4;
    };
    const uint32_t bar_addr = smem_u32(&s_mbar[warp][0]), buf_addr = smem_u32(my_buf);
    auto issue = [&](uint32_t u, uint32_t j, uint32_t stage) {           // lane 0 only
        tma_chunk_addr(buf_addr + stage * (CHUNK4 * 16u), a.cigar + static_cast<uint64_t>(u) * UNIT4 + static_cast<uint64_t>(j) * CHUNK4,
                       CHUNK4 * 16u, bar_addr + stage * 8u);
    };

    uint32_t unit = claim();
    if (unit >= a.n_units) return;
    if (USE_TMA && lane == 0) {
#pragma unroll
        for (uint32_t j = 0; j < static_cast<uint32_t>(STAGES); ++j)
            if (j < static_cast<uint32_t>(G) && chunk_exists(unit, j)) issue(unit, j, j);
    }
    uint32_t next_unit = claim();
    uint32_t k = 0;                       // chunks this warp has consumed: ring stage k % STAGES, parity (k / STAGES) & 1
    static_assert(G >= STAGES, "the ring is primed with the first STAGES chunks of a unit");

    while (true) {
        const uint64_t w4 = static_cast<uint64_t>(unit) * UNIT4;         // this unit's first uint4
        WarpState st;
        st.laneR = 0; st.laneQ = 0; st.cur_aln = 0; st.head = 0; st.cnt = 0;
        ChunkGeom g_next = load_geom(w4 / CHUNK4);
#pragma unroll 1
        for (int j = 0; j < G; ++j) {
            const uint64_t c4 = w4 + static_cast<uint64_t>(j) * CHUNK4;
            if (c4 >= a.n4) break;
            const ChunkGeom g = g_next;
            if (j + 1 < G && c4 + CHUNK4 < a.n4) g_next = load_geom(c4 / CHUNK4 + 1);  // prefetch next chunk's geometry
            if (g.head_at_start) {        // an alignment starts exactly here: the running piece is complete
                add_to_alignment(a, st.cur_aln, __reduce_add_sync(0xffffffffu, st.laneR), __reduce_add_sync(0xffffffffu, st.laneQ), lane);
                st.laneR = 0;
                st.laneQ = 0;
                st.head = 1u;
            }
            st.cur_aln = g.a_lo;
            uint32_t totR = 0, totQ = 0;
            bool rare = false;            // some op of this lane passes its class threshold (I/D >= min_sv_size, N, H)
            if (USE_TMA) {
                const uint32_t stage = k % STAGES;
                mbar_wait_addr(bar_addr + stage * 8u, (k / STAGES) & 1u);
                const uint4* buf = my_buf + stage * CHUNK4;
#pragma unroll
                for (int r = 0; r < ROWS; r += 2) {
                    unsigned long long acc = fast_row(buf[static_cast<uint32_t>(r) * 32u + lane], 0ull, rare, a.mul28);
                    if (r + 1 < ROWS) acc = fast_row(buf[static_cast<uint32_t>(r + 1) * 32u + lane], acc, rare, a.mul28);
                    totR += static_cast<uint32_t>(acc >> 31);
                    totQ += static_cast<uint32_t>(acc) & 0x7FFFFFFFu;
                }
                const bool any_rare = __any_sync(0xffffffffu, rare);
                if (g.n_heads == 0u && !any_rare) {
                    st.laneR += totR;
                    st.laneQ += totQ;
                } else {
                    chunk_slow(a, st, g, c4, lane, unit, totR, totQ, rare, any_rare,
                               [&](int r) -> uint4 { return buf[static_cast<uint32_t>(r) * 32u + lane]; },
                               [&]() -> const uint4* { return buf; });
                }
                __syncwarp();                              // every lane is done with this stage
                if (lane == 0) {                           // refill it with the chunk STAGES ahead (maybe of the next unit)
                    const uint32_t ja = static_cast<uint32_t>(j) + STAGES;
                    const uint32_t au = ja < static_cast<uint32_t>(G) ? unit : next_unit;
                    const uint32_t aj = ja < static_cast<uint32_t>(G) ? ja : ja - static_cast<uint32_t>(G);
                    if (chunk_exists(au, aj)) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue(au, aj, stage);
                    }
                }
                ++k;
            } else {
                uint4 v[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) v[r] = ldg_stream(a.cigar + c4 + static_cast<uint64_t>(r) * 32u + lane);
#pragma unroll
                for (int r = 0; r < ROWS; r += 2) {
                    unsigned long long acc = fast_row(v[r], 0ull, rare, a.mul28);
                    if (r + 1 < ROWS) acc = fast_row(v[r + 1 < ROWS ? r + 1 : r], acc, rare, a.mul28);
                    totR += static_cast<uint32_t>(acc >> 31);
                    totQ += static_cast<uint32_t>(acc) & 0x7FFFFFFFu;
                }
                const bool any_rare = __any_sync(0xffffffffu, rare);
                if (g.n_heads == 0u && !any_rare) {
                    st.laneR += totR;
                    st.laneQ += totQ;
                } else {
                    chunk_slow(a, st, g, c4, lane, unit, totR, totQ, rare, any_rare,
                               [&](int r) -> uint4 { return reload_row(a.cigar, c4, r, lane); },
                               [&]() -> const uint4* {          // park the chunk in shared memory for the row writer
#pragma unroll
                                   for (int r = 0; r < ROWS; ++r) my_buf[r * 32 + lane] = v[r];
                                   __syncwarp();
                                   return my_buf;
                               });
                    __syncwarp();                          // the spill buffer is reused by the next chunk with events
                }
            }
        }
        {
            const uint32_t R = __reduce_add_sync(0xffffffffu, st.laneR), Q = __reduce_add_sync(0xffffffffu, st.laneQ);
            add_to_alignment(a, st.cur_aln, R, Q, lane);
            if (lane == 0) a.unit_agg[unit] = make_uint4(R, Q, st.head, st.cnt);
        }
        unit = next_unit;
        if (unit >= a.n_units) break;
        next_unit = claim();
    }
}

// ---- finalize 1: segmented exclusive scan of the unit aggregates (chained scan over CTAs) ------------------
// carry[u] = advance sums since the last alignment head before unit u (x: ref, y: read); base[u] = rows before u.
// One unit per thread: warp-level segmented scan by shuffles, the 32 warp totals scanned by warp 0, and the CTA's
// exclusive prefix obtained by looking back over the aggregates of ALL earlier CTAs (n_units / 1024 of them: 25 for
// a whole genome), which are published with one 16-byte store each {R, Q, rows, READY | head}.
struct UnitPrefix {
    uint32_t R, Q;
    unsigned long long base;
};

struct SegSum {              // monoid of the segmented scan: x then y
    uint32_t R, Q, H, C;
};
__device__ __forceinline__ SegSum seg_combine(const SegSum x, const SegSum y) {
    SegSum o;
    o.R = y.H ? y.R : x.R + y.R;
    o.Q = y.H ? y.Q : x.Q + y.Q;
    o.H = x.H | y.H;
    o.C = x.C + y.C;
    return o;
}
__device__ __forceinline__ SegSum seg_shfl_up(const SegSum v, uint32_t d) {
    SegSum o;
    o.R = __shfl_up_sync(0xffffffffu, v.R, d);
    o.Q = __shfl_up_sync(0xffffffffu, v.Q, d);
    o.H = __shfl_up_sync(0xffffffffu, v.H, d);
    o.C = __shfl_up_sync(0xffffffffu, v.C, d);
    return o;
}
__device__ __forceinline__ SegSum seg_warp_inclusive(SegSum v, uint32_t lane) {
#pragma unroll
    for (uint32_t d = 1; d < 32u; d <<= 1) {
        const SegSum up = seg_shfl_up(v, d);
        if (lane >= d) v = seg_combine(up, v);
    }
    return v;
}

constexpr uint32_t US_THREADS = 1024;

// cta_agg[b] = {R, Q, rows (32 bit: a CTA covers at most 1024 units), head seen}; cta_ready[b] is raised after a
// __threadfence() (the PTX memory model gives no single-copy atomicity to 16-byte vectors, so the flag is a word of its
// own).  ticket: dynamic CTA index, so that a CTA only ever waits for CTAs that already run.  The block size is a
// launch parameter (tests run the look-back with 32-thread CTAs on small inputs).
__global__ void __launch_bounds__(US_THREADS) unit_scan_kernel(const uint4* __restrict__ agg, uint32_t n_units,
                                                               UnitPrefix* __restrict__ prefix, uint4* cta_agg,
                                                               unsigned int* cta_ready, unsigned int* ticket) {
    __shared__ SegSum s_warp[32];
    __shared__ uint32_t s_cta;
    __shared__ uint32_t s_preR, s_preQ, s_preH;
    __shared__ unsigned long long s_preC;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5, n_warps = blockDim.x >> 5;
    if (t == 0) s_cta = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t cta = s_cta;
    const uint32_t u = cta * blockDim.x + t;
    SegSum mine;
    mine.R = mine.Q = mine.H = mine.C = 0u;
    if (u < n_units) {
        const uint4 x = __ldcg(agg + u);
        mine.R = x.x; mine.Q = x.y; mine.H = x.z ? 1u : 0u; mine.C = x.w;
    }
    const SegSum incl = seg_warp_inclusive(mine, lane);
    if (lane == 31u) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        SegSum wt;
        wt.R = wt.Q = wt.H = wt.C = 0u;
        if (lane < n_warps) wt = s_warp[lane];
        const SegSum winc = seg_warp_inclusive(wt, lane);
        if (lane == 31u) {                                  // publish the CTA aggregate before looking back
            uint4 st;
            st.x = winc.R; st.y = winc.Q; st.z = winc.C; st.w = winc.H;
            __stcg(cta_agg + cta, st);
            __threadfence();
            *reinterpret_cast<volatile unsigned int*>(cta_ready + cta) = 1u;
        }
        SegSum wex = seg_shfl_up(winc, 1);                  // exclusive prefix of each warp inside the CTA
        if (lane == 0u) wex.R = wex.Q = wex.H = wex.C = 0u;
        s_warp[lane] = wex;
        // look back: lanes take the predecessors cta-1-lane, cta-33-lane, ...; nearest first
        uint32_t R = 0, Q = 0, H = 0;                       // fold of the CTAs seen so far (they FOLLOW the ones still to come)
        unsigned long long C = 0;
        for (int first = static_cast<int>(cta) - 1; first >= 0; first -= 32) {
            const int p = first - static_cast<int>(lane);
            SegSum v;
            v.R = v.Q = v.H = v.C = 0u;
            if (p >= 0) {
                while (*reinterpret_cast<volatile unsigned int*>(cta_ready + p) == 0u) {
                }
                __threadfence();
                const uint4 st = __ldcg(cta_agg + p);
                v.R = st.x; v.Q = st.y; v.C = st.z; v.H = st.w;
            }
            // lane 0 holds the NEAREST predecessor: combine in order far -> near, i.e. reversed lanes
            SegSum rv;
            rv.R = __shfl_sync(0xffffffffu, v.R, 31u - lane);
            rv.Q = __shfl_sync(0xffffffffu, v.Q, 31u - lane);
            rv.H = __shfl_sync(0xffffffffu, v.H, 31u - lane);
            rv.C = __shfl_sync(0xffffffffu, v.C, 31u - lane);
            const SegSum winsum = seg_warp_inclusive(rv, lane);     // lane 31: fold of this window in CTA order
            const uint32_t wR = __shfl_sync(0xffffffffu, winsum.R, 31), wQ = __shfl_sync(0xffffffffu, winsum.Q, 31);
            const uint32_t wH = __shfl_sync(0xffffffffu, winsum.H, 31);
            const unsigned long long wC = __reduce_add_sync(0xffffffffu, v.C);
            if (!H) { R += wR; Q += wQ; }                   // window precedes what was folded so far
            H |= wH;
            C += wC;
            // (rows need every predecessor; the carries stop mattering once a head was seen)
        }
        if (lane == 0u) { s_preR = R; s_preQ = Q; s_preH = H; s_preC = C; }
    }
    __syncthreads();
    SegSum ex = seg_shfl_up(incl, 1);                       // the lanes before this one
    if (lane == 0u) ex.R = ex.Q = ex.H = ex.C = 0u;
    if (u < n_units) {
        // exclusive prefix of this unit: CTA prefix, then warp prefix, then the lanes before it
        SegSum pre;
        pre.R = s_preR; pre.Q = s_preQ; pre.H = s_preH; pre.C = 0u;
        const SegSum wex = s_warp[warp];
        const SegSum tot = seg_combine(seg_combine(pre, wex), ex);
        UnitPrefix o;
        o.R = tot.R; o.Q = tot.Q;
        o.base = s_preC + wex.C + ex.C;
        prefix[u] = o;
    }
}

// ---- finalize 2: every staged row moves to its final slot; flagged rows get their carry and their clamps --
__global__ void finalize_rows_kernel(const svb_row* __restrict__ staged, const unsigned long long* __restrict__ n_staged,
                                     unsigned long long cap, const UnitPrefix* __restrict__ prefix,
                                     const svb_aln_hdr* __restrict__ hdr, const int32_t* __restrict__ contig_len, int32_t n_contig,
                                     svb_row* __restrict__ out, unsigned long long* __restrict__ ins_bytes) {
    const unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned long long n = min(*n_staged, cap);
    unsigned long long mine = 0;                                  // bytes of inserted sequence of this row (the pool's size is their sum)
    if (i < n) {
        svb_row r = staged[i];
        const UnitPrefix p = prefix[r.mate_aln];
        const unsigned long long dst = p.base + static_cast<uint32_t>(r.copies);
        if (r.flags & ROW_NEEDS_CARRY) {
            const svb_aln_hdr h = hdr[r.aln_idx];
            const uint32_t pr = static_cast<uint32_t>(r.reserved0) + p.R, len = static_cast<uint32_t>(r.reserved0 >> 32);
            const uint32_t pq = r.seq_pos + p.Q;
            const int32_t clen = (h.tid >= 0 && h.tid < n_contig) ? contig_len[h.tid] : 0;
            const long long start = static_cast<long long>(h.pos) + pr, end = start + len;
            const int32_t cs = static_cast<int32_t>(max(0ll, start));
            const int32_t ce = static_cast<int32_t>(min(static_cast<long long>(clen), end));
            if (r.type == SVB_DEL) { r.src_start = cs; r.src_end = ce; r.seq_len = 0; }
            else { r.dst_start = cs; r.dst_end = ce; r.seq_len = pq >= h.l_seq ? 0u : min(len, h.l_seq - pq); }
            r.seq_pos = r.type == SVB_DEL ? pq : min(pq, h.l_seq);
        }
        r.flags = 0;
        r.copies = 0;
        r.mate_aln = 0xFFFFFFFFu;
        r.reserved0 = 0;
        if (dst < cap) {
            out[dst] = r;
            if (r.type == SVB_INS) mine = (r.seq_len + 1u) / 2u;
        }
    }
    if (ins_bytes) {                                              // (warp-uniform)
        for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
        if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(ins_bytes, mine);
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

__global__ void pad_fill_kernel(uint4* cigar, uint64_t from4, uint64_t to4) {
    const uint64_t i = from4 + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < to4) cigar[i] = make_uint4(15u, 15u, 15u, 15u);
}

}  // namespace

uint64_t cigar_padded_n4(uint64_t n4) {
    const uint64_t unit4 = 16ull * CHUNK4;          // the longest unit a launch may pick (G = 16 chunks of 4 KB)
    return (n4 + unit4 - 1) / unit4 * unit4;
}

int launch_build_chunk_index(svb_ctx* ctx, svb_records* rec) {
    // pad the tail of the buffer (allocated with cigar_padded_n4) so that no load of the scan needs a bounds check
    const uint64_t pad_to = cigar_padded_n4(rec->n4);
    if (pad_to > rec->n4) {
        pad_fill_kernel<<<static_cast<unsigned>((pad_to - rec->n4 + 255) / 256), 256, 0, ctx->stream>>>(rec->d_cigar, rec->n4, pad_to);
        ctx->launches += 1;
    }
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
static int launch_scan_g(svb_ctx* ctx, const svb_records* rec, ScanArgs a, svb_row* final_rows, unsigned long long* ins_bytes) {
    const uint64_t unit4 = static_cast<uint64_t>(G) * CHUNK4;
    const uint64_t n_units64 = (rec->n4 + unit4 - 1) / unit4;
    if (n_units64 > 0x7fffffffull) return svb_fail(ctx, SVB_ERR_ARG, "too many CIGAR ops for one launch");
    const uint32_t n_units = static_cast<uint32_t>(n_units64);
    // scratch: unit aggregates, unit prefixes, CTA status words of the unit scan (+ its ticket), staging rows
    const size_t agg_bytes = (sizeof(uint4) * n_units + 255) & ~static_cast<size_t>(255);
    const size_t pre_bytes = (sizeof(UnitPrefix) * n_units + 255) & ~static_cast<size_t>(255);
    uint32_t us_threads = US_THREADS;
    if (const char* env = getenv("SVB_UNIT_SCAN_THREADS")) {            // test hook: small CTAs exercise the look-back
        const int v = atoi(env);
        if (v >= 32 && v <= static_cast<int>(US_THREADS) && v % 32 == 0) us_threads = static_cast<uint32_t>(v);
    }
    const uint32_t us_ctas = (n_units + us_threads - 1) / us_threads;
    const size_t st_bytes = ((sizeof(uint4) + sizeof(unsigned int)) * (us_ctas + 2) + 255) & ~static_cast<size_t>(255);
    unsigned char* scratch = static_cast<unsigned char*>(svb_scratch(ctx, agg_bytes + pre_bytes + st_bytes + sizeof(svb_row) * a.cap));
    if (!scratch) return svb_fail(ctx, SVB_ERR_NOMEM, "cigar_scan scratch");
    a.n_units = n_units;
    a.unit_agg = reinterpret_cast<uint4*>(scratch);
    UnitPrefix* prefix = reinterpret_cast<UnitPrefix*>(scratch + agg_bytes);
    uint4* cta_status = reinterpret_cast<uint4*>(scratch + agg_bytes + pre_bytes);
    unsigned int* cta_ready = reinterpret_cast<unsigned int*>(cta_status + us_ctas);
    unsigned int* ticket = cta_ready + us_ctas;
    a.unit_counter = ticket + 1;
    a.rows = reinterpret_cast<svb_row*>(scratch + agg_bytes + pre_bytes + st_bytes);
    SVB_CUDA(ctx, cudaMemsetAsync(cta_status, 0, st_bytes, ctx->stream));
    // persistent: as many CTAs as stay resident, each warp claims units until none is left
    const unsigned resident = static_cast<unsigned>(ctx->sm_count) * (ctx->scan_variant == 0 ? K2_TMA_CTAS : K2_LDG_CTAS);
    const unsigned blocks = std::min<unsigned>((n_units + WARPS - 1) / WARPS, resident);
    {
        KernelTimer timer(ctx, SVB_K_CIGAR_SCAN);           // the streaming kernel alone (the roofline's launch duration)
        if (ctx->scan_variant == 0) {
            const size_t smem = static_cast<size_t>(WARPS) * STAGES * CHUNK4 * sizeof(uint4);
            // function attributes are per DEVICE: remembered per context (one context per device), not per process
            constexpr int slot = G == 2 ? 0 : G == 4 ? 1 : G == 8 ? 2 : G == 16 ? 3 : G == 32 ? 4 : G == 64 ? 5 : 6;
            if (!ctx->scan_attr_set[slot]) {
                SVB_CUDA(ctx, cudaFuncSetAttribute(cigar_scan_kernel<true, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(smem)));
                ctx->scan_attr_set[slot] = true;
            }
            cigar_scan_kernel<true, G><<<blocks, THREADS, smem, ctx->stream>>>(a);
        } else {
            const size_t smem = static_cast<size_t>(WARPS) * CHUNK4 * sizeof(uint4);
            cigar_scan_kernel<false, G><<<blocks, THREADS, smem, ctx->stream>>>(a);
        }
    }
    {
        KernelTimer timer(ctx, SVB_K_SCAN_FINALIZE);        // unit scan + per-row finalize
        unit_scan_kernel<<<us_ctas, us_threads, 0, ctx->stream>>>(a.unit_agg, n_units, prefix, cta_status, cta_ready, ticket);
        const unsigned long long fin_blocks = (a.cap + 255) / 256;
        finalize_rows_kernel<<<static_cast<unsigned>(std::min<unsigned long long>(fin_blocks, 0x7fffffffull)), 256, 0, ctx->stream>>>(
            a.rows, a.total, a.cap, prefix, rec->d_hdr, rec->d_contig_len, rec->n_contig, final_rows, ins_bytes);
    }
    ctx->launches += 3;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}

int launch_cigar_scan(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, ScanOutput out) {
    SVB_CUDA(ctx, cudaMemsetAsync(out.d_count, 0, sizeof(unsigned long long), ctx->stream));
    if (out.d_ins_bytes) SVB_CUDA(ctx, cudaMemsetAsync(out.d_ins_bytes, 0, sizeof(unsigned long long), ctx->stream));
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
    a.n_units = 0;
    a.min_mapq = p->min_mapq;
    const long long m = p->min_sv_size < 0 ? 0 : p->min_sv_size;
    a.min16 = m >= (1ll << 28) ? 0xFFFFFFFFu : static_cast<uint32_t>(m << 4);
    a.hap = static_cast<uint32_t>(hap);
    a.mul28 = 1u << 28;
    a.aln_sum = rec->d_aln_sum;
    a.unit_agg = nullptr;
    a.rows = nullptr;
    a.unit_counter = nullptr;
    a.cap = out.cap;
    a.total = out.d_count;
    a.dev_status = ctx->d_status;
    // unit length: as long as possible while at least four full waves of warps (32 per SM) remain
    const uint64_t four_waves = 4ull * 32ull * static_cast<uint64_t>(ctx->sm_count);
    const uint64_t chunks = (rec->n4 + CHUNK4 - 1) / CHUNK4;
    constexpr int S = 8 / ROWS;                      // keep the unit length in ops when chunks are smaller
    if (chunks / (16 * S) >= four_waves) return launch_scan_g<16 * S>(ctx, rec, a, out.rows, out.d_ins_bytes);
    if (chunks / (8 * S) >= four_waves) return launch_scan_g<8 * S>(ctx, rec, a, out.rows, out.d_ins_bytes);
    if (chunks / (4 * S) >= four_waves) return launch_scan_g<4 * S>(ctx, rec, a, out.rows, out.d_ins_bytes);
    return launch_scan_g<2 * S>(ctx, rec, a, out.rows, out.d_ins_bytes);
}
