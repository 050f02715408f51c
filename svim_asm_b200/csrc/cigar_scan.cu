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
// emitting ops.  It is done in ONE pass with a decoupled look-back over 8192-op tiles:
//   * a tile (32 KB) is staged to shared memory with one TMA bulk copy (cp.async.bulk + mbarrier),
//     or, variant 1, read with coalesced 128-bit LDG.nc straight into registers;
//   * each lane owns one uint4 (4 ops) per 128-op row; it only accumulates its advance sums and
//     a bit mask of emitting ops -- no per-op prefix is formed in the common case;
//   * per-warp (1024 ops) and per-tile aggregates "sum since the last alignment head" and the
//     emit count are published to a tile-status array; predecessors are combined by a warp-wide
//     look-back exactly once per tile, so every op is read from HBM exactly once;
//   * only warps that contain an emitting op recompute the exclusive prefix of that op (two
//     warp reductions per event) and write the finished 64-byte candidate row at its final,
//     stable position.
// Per-alignment totals (reference span, read span, N and H bases), needed by the split-alignment
// walk for reference_end / infer_read_length (SVIM_inter.py:68-80), fall out as atomics.
//
// Bound: HBM bandwidth.  Algorithmic bytes: 4 B/op + 32 B/alignment + 64 B/emitted row.
#include "common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int ROWS = 8;                    // uint4 per lane
constexpr int CHUNK4 = 32 * ROWS;          // 256 uint4  = 1024 ops per warp
constexpr int TILE4 = WARPS * CHUNK4;      // 2048 uint4 = 8192 ops = 32 KB per CTA

constexpr uint32_t REF_MASK = (1u << 0) | (1u << 2) | (1u << 7) | (1u << 8);              // M D = X  (SVIM_intra.py:15,23,28)
constexpr uint32_t READ_MASK = (1u << 0) | (1u << 1) | (1u << 4) | (1u << 7) | (1u << 8); // M I S = X (SVIM_intra.py:16,20,26,29)
constexpr uint32_t INDEL_MASK = (1u << 1) | (1u << 2);
constexpr uint32_t NH_MASK = (1u << 3) | (1u << 5);

// tile status: three 64-bit words, each tagged with its own state in the top 2 bits so that a
// reader can validate a snapshot without fences (values are self-describing).
constexpr unsigned long long ST_INVALID = 0ull, ST_AGG = 1ull, ST_PREFIX = 2ull;
struct TileStatus {
    unsigned long long w_ref;    // [63:62] state  [32] has_head  [31:0] ref sum since last head
    unsigned long long w_read;   // [63:62] state               [31:0] read sum since last head
    unsigned long long w_cnt;    // [63:62] state  [61:0] emitted rows
    unsigned long long pad;
};

struct ScanArgs {
    const uint4* cigar;
    uint64_t n4;
    const uint32_t* off4;
    const uint32_t* chunk_first;
    const svb_aln_hdr* hdr;
    const int32_t* contig_len;
    uint32_t n_aln;
    int32_t n_contig;
    uint32_t n_tiles;
    int32_t min_mapq;
    uint32_t min16;           // min_sv_size << 4 : (x >= min16) <=> (len >= min_sv_size)
    uint32_t hap;
    uint4* aln_sum;
    TileStatus* status;
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

// One op: advance sums, emit bit, N/H bit.
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
    // SVIM_COLLECT.py:71  is_unmapped / is_secondary / mapping_quality < min_mapq
    return !(h.flag & 0x4) && !(h.flag & 0x100) && static_cast<int32_t>(h.mapq) >= min_mapq;
}

template <bool USE_TMA>
__device__ __forceinline__ uint4 reload_row_fn(const uint4* s_tile, const uint4* cigar, uint64_t tile4, uint32_t here4, uint32_t i4) {
    if (i4 >= here4) return make_uint4(15u, 15u, 15u, 15u);
    if (USE_TMA) return s_tile[i4];
    return cigar[tile4 + i4];
}

// fast path of one uint4 (4 ops): packed advance sums and the rare flag
__device__ __forceinline__ void fast_row_fn(const uint4 d, const uint2* lut, bool in_head, uint32_t& totR, uint32_t& totQ,
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

template <bool USE_TMA>
__global__ void __launch_bounds__(THREADS, 5) cigar_scan_kernel(const ScanArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* s_tile = reinterpret_cast<uint4*>(smem_raw);       // TILE4 uint4 when USE_TMA
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ uint32_t s_tile_id;
    __shared__ uint2 s_lut[16];      // per op code: x = multiplier (bit 0 read advance, bit 31 reference advance), y = rare threshold
    __shared__ uint32_t s_wR[WARPS], s_wQ[WARPS], s_wCnt[WARPS], s_wHead[WARPS];
    __shared__ uint32_t s_cR[WARPS], s_cQ[WARPS], s_cResolved[WARPS], s_cBase[WARPS];
    __shared__ uint32_t s_tileR, s_tileQ;
    __shared__ unsigned long long s_tileBase;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    if (tid >= 32u && tid < 48u) {
        const uint32_t op = tid - 32u;
        uint2 e;
        e.x = (((REF_MASK >> op) & 1u) << 31) | ((READ_MASK >> op) & 1u);
        e.y = ((INDEL_MASK >> op) & 1u) ? a.min16 : (((NH_MASK >> op) & 1u) ? 16u : 0xFFFFFFFFu);
        s_lut[op] = e;
    }
    if (tid == 0) {
        s_tile_id = atomicAdd(a.ticket, 1u);      // tiles are claimed in scheduling order: look-back cannot deadlock
        if (USE_TMA) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads();
    const uint32_t tile = s_tile_id;
    const uint64_t tile4 = static_cast<uint64_t>(tile) * TILE4;
    const uint32_t here4 = static_cast<uint32_t>(min(static_cast<uint64_t>(TILE4), a.n4 - tile4));

    if (USE_TMA) {
        if (tid == 0) {
            const uint32_t bytes = here4 * 16u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         :: "r"(smem_u32(&s_mbar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(s_tile)), "l"(a.cigar + tile4), "r"(bytes), "r"(smem_u32(&s_mbar)) : "memory");
        }
    }

    // ---- geometry of this warp's 1024-op chunk
    const uint64_t g4base = tile4 + static_cast<uint64_t>(warp) * CHUNK4;
    const uint64_t g4end = min(g4base + CHUNK4, a.n4);
    const bool chunk_live = g4base < a.n4;
    uint32_t a_lo = 0, a_hi = 0;
    if (chunk_live) {
        a_lo = a.chunk_first[g4base / CHUNK4];
        a_hi = a_lo;
        while (true) {                                    // heads inside the chunk (off4 is non-decreasing)
            const uint32_t idx = a_hi + 1u + lane;
            const bool in = idx < a.n_aln && static_cast<uint64_t>(a.off4[idx]) < g4end;
            const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, in));
            a_hi += cnt;
            if (cnt < 32u) break;
        }
    }

    // ---- load + per-lane decode, phase 1: totals and emit counts of the chunk's alignment pieces.
    // Fast path per op (no branches, three pipes): one 8-byte table entry {multiplier, threshold} from
    // shared memory; IMAD.WIDE adds len * multiplier to a packed accumulator (read sum in bits 0..30,
    // reference sum from bit 31 up); one compare-OR flags the rare ops (I/D with len >= min_sv_size,
    // N/H with len > 0).  Exact event bits are only recomputed when the flag fires somewhere in the warp.
    if (USE_TMA) {
        uint32_t ready = 0;
        while (!ready) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ready) : "r"(smem_u32(&s_mbar)) : "memory");
        }
    }
    // re-read of a row for the rare paths (shared memory, or L2 for the LDG variant)
    const uint32_t lane4 = warp * CHUNK4 + lane;            // this lane's uint4 index inside the tile, row 0
#define reload_row(r) reload_row_fn<USE_TMA>(s_tile, a.cigar, tile4, here4, lane4 + static_cast<uint32_t>(r) * 32u)

    uint32_t evbits = 0;
    uint32_t tailR = 0, tailQ = 0, warp_cnt = 0;
    if (chunk_live) {
        const uint32_t n_pieces = a_hi - a_lo + 1u;
        // rows of this lane that lie before the first head inside the chunk: g4base + 32 r + lane < split4
        const uint64_t split4 = (n_pieces >= 2u) ? static_cast<uint64_t>(a.off4[a_lo + 1u]) : g4end;
        const long long rows_before = (static_cast<long long>(split4) - static_cast<long long>(g4base) - lane + 31) / 32;
        const int nb = rows_before <= 0 ? 0 : (rows_before >= ROWS ? ROWS : static_cast<int>(rows_before));
        uint32_t headR = 0, headQ = 0, totR = 0, totQ = 0;
        bool rare = false;
        if (USE_TMA) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const uint32_t i4 = warp * CHUNK4 + static_cast<uint32_t>(r) * 32u + lane;
                fast_row_fn((i4 < here4) ? s_tile[i4] : make_uint4(15u, 15u, 15u, 15u), s_lut, r < nb, totR, totQ, headR, headQ, rare);
            }
        } else {
#pragma unroll
            for (int half = 0; half < 2; ++half) {         // two batches of four 128-bit loads in flight per lane
                uint4 v[ROWS / 2];
#pragma unroll
                for (int k = 0; k < ROWS / 2; ++k) {
                    const uint64_t g4 = g4base + static_cast<uint64_t>(half * (ROWS / 2) + k) * 32u + lane;
                    v[k] = (g4 < a.n4) ? ldg_stream(a.cigar + g4) : make_uint4(15u, 15u, 15u, 15u);
                }
#pragma unroll
                for (int k = 0; k < ROWS / 2; ++k) fast_row_fn(v[k], s_lut, half * (ROWS / 2) + k < nb, totR, totQ, headR, headQ, rare);
            }
        }
        const uint32_t first_mask = nb >= 8 ? 0xFFFFFFFFu : ((1u << (4u * static_cast<uint32_t>(nb))) - 1u);
        const bool any_rare = __ballot_sync(0xffffffffu, rare) != 0u;

        // exact bits of the rare ops, only when the warp saw one (about 1 chunk in 8 for human assemblies)
        uint32_t nhbits = 0;
        if (any_rare) {
#pragma unroll 1
            for (int r = 0; r < ROWS; ++r) {
                const uint4 d = reload_row(r);
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
            tailR = headR;
            tailQ = headQ;
            if (n_pieces == 2u) {
                const uint32_t restR = __reduce_add_sync(0xffffffffu, totR) - headR;
                const uint32_t restQ = __reduce_add_sync(0xffffffffu, totQ) - headQ;
                if (lane == 0 && (restR | restQ)) {
                    atomicAdd(&a.aln_sum[a_hi].x, restR);
                    atomicAdd(&a.aln_sum[a_hi].y, restQ);
                }
                tailR = restR;
                tailQ = restQ;
            }
            if (any_rare) {
                const bool pass0 = record_passes(a.hdr[a_lo], a.min_mapq);
                const bool pass1 = n_pieces == 2u ? record_passes(a.hdr[a_hi], a.min_mapq) : false;
                evbits &= (pass0 ? first_mask : 0u) | (pass1 ? ~first_mask : 0u);      // SVIM_COLLECT.py:71
                warp_cnt = __reduce_add_sync(0xffffffffu, __popc(evbits));
            }
        } else {
            // generic: several alignment heads inside one 1024-op chunk (short alignments)
            uint32_t keep = 0;
            for (uint32_t al = a_lo; al <= a_hi; ++al) {
                const uint64_t lo = max(static_cast<uint64_t>(a.off4[al]), g4base);
                const uint64_t hi = min(static_cast<uint64_t>(a.off4[al + 1]), g4end);
                uint32_t sR = 0, sQ = 0, in_mask = 0;
#pragma unroll 1
                for (int r = 0; r < ROWS; ++r) {
                    const uint64_t g4 = g4base + static_cast<uint64_t>(r) * 32u + lane;
                    if (g4 >= lo && g4 < hi) {
                        const uint4 d = reload_row(r);
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
                tailR = sR;
                tailQ = sQ;
            }
            evbits &= keep;
            if (any_rare) warp_cnt = __reduce_add_sync(0xffffffffu, __popc(evbits));
        }
        // rare: N / H ops feed reference_end / infer_read_length of the split-alignment walk
        if (any_rare && __ballot_sync(0xffffffffu, nhbits != 0u)) {
            uint32_t bits = nhbits;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1u;
                const uint64_t g4 = g4base + static_cast<uint64_t>(b >> 2) * 32u + lane;
                uint32_t al = a_lo;
                while (al < a_hi && static_cast<uint64_t>(a.off4[al + 1]) <= g4) ++al;
                const uint4 d = reload_row(b >> 2);
                const uint32_t x = (b & 3) == 0 ? d.x : (b & 3) == 1 ? d.y : (b & 3) == 2 ? d.z : d.w;
                if ((x & 15u) == 3u) atomicAdd(&a.aln_sum[al].z, x >> 4);
                else atomicAdd(&a.aln_sum[al].w, x >> 4);
            }
        }
    }
    if (lane == 0) {
        s_wR[warp] = tailR;
        s_wQ[warp] = tailQ;
        s_wCnt[warp] = warp_cnt;
        s_wHead[warp] = chunk_live && (a_hi > a_lo || static_cast<uint64_t>(a.off4[a_lo]) >= g4base) ? 1u : 0u;
    }
    __syncthreads();

    // ---- phase 2 (warp 0): combine the 8 warp aggregates, publish, look back, publish prefix
    if (warp == 0) {
        uint32_t runR = 0, runQ = 0, runHead = 0, runCnt = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            if (lane == 0) {
                s_cR[w] = runR;
                s_cQ[w] = runQ;
                s_cResolved[w] = runHead;
                s_cBase[w] = runCnt;
            }
            if (s_wHead[w]) {
                runR = s_wR[w];
                runQ = s_wQ[w];
                runHead = 1u;
            } else {
                runR += s_wR[w];
                runQ += s_wQ[w];
            }
            runCnt += s_wCnt[w];
        }
        TileStatus* mine = a.status + tile;
        if (tile != 0 && lane == 0) {
            st_relaxed(&mine->w_ref, (ST_AGG << 62) | (static_cast<unsigned long long>(runHead) << 32) | runR);
            st_relaxed(&mine->w_read, (ST_AGG << 62) | runQ);
            st_relaxed(&mine->w_cnt, (ST_AGG << 62) | runCnt);
        }
        // look-back: 32 predecessors per round, nearest = lane 0
        uint32_t carryR = 0, carryQ = 0;
        bool sums_done = false;
        unsigned long long excl = 0;
        bool cnt_done = (tile == 0);
        if (tile == 0) sums_done = true;
        int64_t look = static_cast<int64_t>(tile) - 1;
        while (!cnt_done) {
            const int64_t t = look - lane;
            unsigned long long wr = (ST_PREFIX << 62), wq = (ST_PREFIX << 62), wc = (ST_PREFIX << 62);   // virtual tile -1: empty prefix
            if (t >= 0) {
                const TileStatus* ts = a.status + t;
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
            const int k_cnt = pmask ? (__ffs(pmask) - 1) : 31;          // count: lanes 0..k_cnt
            const int k_sum = smask ? (__ffs(smask) - 1) : 31;          // sums: lanes 0..k_sum (nearest stop)
            const unsigned long long c = (static_cast<int>(lane) <= k_cnt) ? (wc & ((1ull << 62) - 1ull)) : 0ull;
            unsigned long long csum = c;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            excl += csum;
            if (!sums_done) {
                const bool take = static_cast<int>(lane) <= k_sum;
                carryR += __reduce_add_sync(0xffffffffu, take ? static_cast<uint32_t>(wr) : 0u);
                carryQ += __reduce_add_sync(0xffffffffu, take ? static_cast<uint32_t>(wq) : 0u);
                if (smask) sums_done = true;
            }
            if (pmask) cnt_done = true;
            look -= 32;
        }
        if (lane == 0) {
            const uint32_t incR = runHead ? runR : carryR + runR;
            const uint32_t incQ = runHead ? runQ : carryQ + runQ;
            st_relaxed(&mine->w_ref, (ST_PREFIX << 62) | (1ull << 32) | incR);
            st_relaxed(&mine->w_read, (ST_PREFIX << 62) | incQ);
            st_relaxed(&mine->w_cnt, (ST_PREFIX << 62) | (excl + runCnt));
            s_tileR = carryR;
            s_tileQ = carryQ;
            s_tileBase = excl;
            if (tile == a.n_tiles - 1u) *a.total = excl + runCnt;
        }
    }
    __syncthreads();

    // ---- phase 3: only warps that hold an emitting op (about 1 in 8 for human assemblies)
    if (warp_cnt == 0u) return;
    uint32_t carryR = s_cR[warp], carryQ = s_cQ[warp];
    if (!s_cResolved[warp]) {
        carryR += s_tileR;
        carryQ += s_tileQ;
    }
    unsigned long long out = s_tileBase + s_cBase[warp];

    for (uint32_t al = a_lo; al <= a_hi; ++al) {
        const uint64_t lo = max(static_cast<uint64_t>(a.off4[al]), g4base);
        const uint64_t hi = min(static_cast<uint64_t>(a.off4[al + 1]), g4end);
        if (hi <= lo) continue;
        // rows of this lane inside the piece; pieces without a surviving event are skipped
        uint32_t in_mask = 0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint64_t g4 = g4base + static_cast<uint64_t>(r) * 32u + lane;
            if (g4 >= lo && g4 < hi) in_mask |= 0xFu << (4 * r);
        }
        if (__ballot_sync(0xffffffffu, (evbits & in_mask) != 0u) == 0u) continue;
        const svb_aln_hdr h = a.hdr[al];
        // positions are relative to the alignment start: the carry only applies to a piece that began earlier
        const bool continued = static_cast<uint64_t>(a.off4[al]) < g4base;
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
            const uint64_t g4 = g4base + static_cast<uint64_t>(r) * 32u + lane;
            const bool in = ((in_mask >> (4 * r)) & 1u) != 0u;
            const uint4 d = in ? reload_row(r) : make_uint4(15u, 15u, 15u, 15u);
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

#undef reload_row

// chunk_first[c] = last alignment whose run starts at or before uint4 index c * CHUNK4
__global__ void chunk_index_kernel(const uint32_t* __restrict__ off4, uint32_t n_aln, uint64_t n_chunks,
                                   uint32_t* __restrict__ chunk_first) {
    const uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const uint64_t x = c * CHUNK4;
    uint32_t lo = 0, hi = n_aln;                 // upper_bound over off4[0..n_aln)
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        if (static_cast<uint64_t>(off4[mid]) <= x) lo = mid + 1u; else hi = mid;
    }
    chunk_first[c] = lo ? lo - 1u : 0u;
}

}  // namespace

int launch_build_chunk_index(svb_ctx* ctx, svb_records* rec) {
    const uint64_t n_chunks = (rec->n4 + CHUNK4 - 1) / CHUNK4;
    if (n_chunks == 0) return SVB_OK;
    SVB_CUDA(ctx, cudaMalloc(&rec->d_chunk_first, n_chunks * sizeof(uint32_t)));
    const unsigned blocks = static_cast<unsigned>((n_chunks + 255) / 256);
    chunk_index_kernel<<<blocks, 256, 0, ctx->stream>>>(rec->d_off4, rec->n_aln, n_chunks, rec->d_chunk_first);
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}

int launch_cigar_scan(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, ScanOutput out) {
    SVB_CUDA(ctx, cudaMemsetAsync(out.d_count, 0, sizeof(unsigned long long), ctx->stream));
    if (rec->n_aln) SVB_CUDA(ctx, cudaMemsetAsync(rec->d_aln_sum, 0, sizeof(uint4) * rec->n_aln, ctx->stream));
    if (rec->n4 == 0) return SVB_OK;
    const uint64_t n_tiles64 = (rec->n4 + TILE4 - 1) / TILE4;
    if (n_tiles64 > 0x7fffffffull) return svb_fail(ctx, SVB_ERR_ARG, "too many CIGAR ops for one launch");
    const uint32_t n_tiles = static_cast<uint32_t>(n_tiles64);
    const size_t need = sizeof(TileStatus) * n_tiles + 256;
    unsigned char* scratch = static_cast<unsigned char*>(svb_scratch(ctx, need));
    if (!scratch) return svb_fail(ctx, SVB_ERR_NOMEM, "tile status scratch");
    SVB_CUDA(ctx, cudaMemsetAsync(scratch, 0, need, ctx->stream));

    ScanArgs a;
    a.cigar = rec->d_cigar;
    a.n4 = rec->n4;
    a.off4 = rec->d_off4;
    a.chunk_first = rec->d_chunk_first;
    a.hdr = rec->d_hdr;
    a.contig_len = rec->d_contig_len;
    a.n_aln = rec->n_aln;
    a.n_contig = rec->n_contig;
    a.n_tiles = n_tiles;
    a.min_mapq = p->min_mapq;
    const long long m = p->min_sv_size < 0 ? 0 : p->min_sv_size;
    a.min16 = m >= (1ll << 28) ? 0xFFFFFFFFu : static_cast<uint32_t>(m << 4);
    a.hap = static_cast<uint32_t>(hap);
    a.aln_sum = rec->d_aln_sum;
    a.ticket = reinterpret_cast<unsigned int*>(scratch);
    a.status = reinterpret_cast<TileStatus*>(scratch + 256);
    a.rows = out.rows;
    a.cap = out.cap;
    a.total = out.d_count;
    a.dev_status = ctx->d_status;

    KernelTimer timer(ctx, SVB_K_CIGAR_SCAN);
    if (ctx->scan_variant == 0) {
        const size_t smem = static_cast<size_t>(TILE4) * sizeof(uint4);
        static bool attr_set = false;
        if (!attr_set) {
            SVB_CUDA(ctx, cudaFuncSetAttribute(cigar_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(smem)));
            attr_set = true;
        }
        cigar_scan_kernel<true><<<n_tiles, THREADS, smem, ctx->stream>>>(a);
    } else {
        cigar_scan_kernel<false><<<n_tiles, THREADS, 16, ctx->stream>>>(a);
    }
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}
