// K6-K9 diploid haplotype pairing on the device.
//
// Replaces pair_candidates (reference src/svim_asm/SVIM_COMBINE.py:164-366):
//   K6  form_partitions' sorted(key=get_key)  (:17; keys SVCandidate.py:17-19,147-148,292-293,386-387)
//       -> stable LSD radix sort (8-bit digits, warp match_any ranking) of both haplotypes' rows on
//          (type, contig rank under python string order, position); hap 1 precedes hap 2 on ties.
//   K7  the linear split into partitions (:20-31) -> head flags + exclusive scan.
//   K8  compute_distance (:35-102) for every cross-haplotype pair of a partition with 2..10 members
//       -> edit_distance.cu (job list built here); span_position_distance_breakends (:105-117) inline.
//   K9  pair_haplotypes / pair_haplotypes_breakends (:120-161): complete linkage, cut, clusters in scipy's
//       label order -> linkage.cuh, one thread per partition; then the genotype / merge rules of :184-365.
// Output rows appear in the reference's order: type by type, partition by partition, label by label.
#include <algorithm>
#include <cstdlib>

#include "pairing.cuh"
#include "walk.cuh"
#include "wfa_core.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 2;          // 512 keys per tile: a whole-genome diploid sample (18,000 keys) spreads over 37 CTAs
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;

struct PairArgs {
    const svb_row* rows;            // concatenated h1 ++ h2 (hap set)
    const uint32_t* order;          // sorted position -> index into rows
    const uint32_t* part_start;     // [n_parts + 1]
    const unsigned long long* n_parts_dev;   // number of partitions, on the device (the host does not wait for it)
    uint32_t n_rows;                // h1 + h2: upper bound of the number of partitions
    const int32_t* contig_len;      // BAM header lengths (constructor clamps)
    const int32_t* contig_lexrank;
    int32_t n_contig;
    const uint64_t* ref_off;        // reference contig offsets (get_reference_length of the FASTA)
    int32_t ref_n_contig;
    double max_edit_distance;
    double* dist;                   // per job: the distance, or the lower end of what is known about it (pairing.cuh)
    double* dist_hi;                // per job: the upper end (== dist when exact)
    uint32_t* job_slot;             // [partition * PAIR_DIST_STRIDE + condensed index] -> job
    EditJob* jobs;
    unsigned long long* job_count;
    uint32_t* exact_list;           // jobs the exact kernel has to run
    unsigned long long* exact_count;
    unsigned long long* cell_count;  // sum of la x lb over the jobs: the full tables an exhaustive aligner would fill (bench.py)
    int all_exact;                  // the wavefront kernel did not run (threshold too large): every job is exact
    uint32_t* counts;               // [n_rows + 1]
    svb_row* out;
    uint32_t* dev_status;
};

__device__ __forceinline__ void key_of(const svb_row& r, int32_t& tid, int32_t& pos) {
    switch (r.type) {
        case SVB_DEL: case SVB_INV: case SVB_DUP_TAN:
            tid = r.src_tid;
            pos = static_cast<int32_t>((static_cast<long long>(r.src_start) + r.src_end) >> 1);   // (start + end) // 2
            break;
        case SVB_INS: case SVB_DUP_INT:
            tid = r.dst_tid; pos = r.dst_start; break;
        default:
            tid = r.src_tid; pos = r.src_start; break;
    }
}

__global__ void concat_keys_kernel(const svb_row* __restrict__ h1, uint32_t n1, const svb_row* __restrict__ h2, uint32_t n2,
                                   const uint64_t* __restrict__ pool_off1, const uint64_t* __restrict__ pool_off2,
                                   const uint64_t* __restrict__ seq_off1, const uint64_t* __restrict__ seq_off2,
                                   const int32_t* __restrict__ lexrank, int32_t n_contig, uint32_t rank_bits,
                                   svb_row* __restrict__ rows, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                                   uint32_t* dev_status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1 + n2) return;
    svb_row r = i < n1 ? h1[i] : h2[i - n1];
    r.hap = i < n1 ? 1 : 2;
    // where this row's inserted sequence starts, as a nibble index: in the table's pool if it has one, else in
    // the haplotype's full 4-bit query sequences
    r.reserved0 = ~0ull;
    if (r.type == SVB_INS) {
        const uint64_t* pool_off = i < n1 ? pool_off1 : pool_off2;
        const uint64_t* seq_off = i < n1 ? seq_off1 : seq_off2;
        if (pool_off) r.reserved0 = pool_off[i < n1 ? i : i - n1] * 2ull;
        else if (seq_off) r.reserved0 = seq_off[r.aln_idx] * 2ull + r.seq_pos;
        else atomicOr(dev_status, DEV_ERR_NOSEQ);
    }
    rows[i] = r;
    int32_t tid, pos;
    key_of(r, tid, pos);
    uint32_t rank = 0;
    if (tid < 0 || tid >= n_contig) atomicOr(dev_status, DEV_ERR_BAD_TID);
    else rank = static_cast<uint32_t>(lexrank[tid]);
    keys[i] = (static_cast<unsigned long long>(r.type) << (32u + rank_bits)) | (static_cast<unsigned long long>(rank) << 32) |
              static_cast<uint32_t>(pos);
    vals[i] = i;
}

__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t shift,
                                                                uint32_t* __restrict__ hist, uint32_t n_blocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const unsigned long long* __restrict__ keys_in,
                                                                   const uint32_t* __restrict__ vals_in,
                                                                   unsigned long long* __restrict__ keys_out,
                                                                   uint32_t* __restrict__ vals_out, uint32_t n, uint32_t shift,
                                                                   const uint32_t* __restrict__ hist_scanned, uint32_t n_blocks) {
    __shared__ uint32_t s_base[256], s_running[256];
    __shared__ uint32_t s_wc[RS_THREADS / 32][256];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    s_base[tid] = hist_scanned[tid * n_blocks + blockIdx.x];
    s_running[tid] = 0;
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) s_wc[w][tid] = 0;
        __syncthreads();
        const uint32_t i = base + r * RS_THREADS + tid;
        const bool valid = i < n;
        unsigned long long key = 0;
        uint32_t val = 0, d = 0x100u | lane;                 // invalid lanes match nobody
        if (valid) {
            key = keys_in[i];
            val = vals_in[i];
            d = static_cast<uint32_t>(key >> shift) & 255u;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0u) s_wc[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t before = 0;
            for (uint32_t w = 0; w < warp; ++w) before += s_wc[w][d];
            const uint32_t dst = s_base[d] + s_running[d] + before + rank_in_warp;     // stable: round, warp, lane order
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        __syncthreads();
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) add += s_wc[w][tid];
        s_running[tid] += add;
        __syncthreads();
    }
}

// form_partitions' split (SVIM_COMBINE.py:20-31): a new partition starts when type or contig differ or the
// key positions of CONSECUTIVE items are more than max_distance apart
__global__ void heads_kernel(const unsigned long long* __restrict__ keys, uint32_t n, long long max_distance,
                             uint32_t* __restrict__ head) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = 1;
    if (i > 0) {
        const unsigned long long a = keys[i - 1], b = keys[i];
        const long long pa = static_cast<int32_t>(static_cast<uint32_t>(a)), pb = static_cast<int32_t>(static_cast<uint32_t>(b));
        const long long gap = pa > pb ? pa - pb : pb - pa;
        h = ((a >> 32) != (b >> 32) || gap > max_distance) ? 1u : 0u;
    }
    head[i] = h;
}

__global__ void part_start_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ excl, uint32_t n,
                                  long long max_distance, uint32_t* __restrict__ part_start,
                                  const unsigned long long* __restrict__ n_parts_dev) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) part_start[static_cast<uint32_t>(*n_parts_dev)] = n;
    if (i >= n) return;
    bool h = true;
    if (i > 0) {
        const unsigned long long a = keys[i - 1], b = keys[i];
        const long long pa = static_cast<int32_t>(static_cast<uint32_t>(a)), pb = static_cast<int32_t>(static_cast<uint32_t>(b));
        const long long gap = pa > pb ? pa - pb : pb - pa;
        h = (a >> 32) != (b >> 32) || gap > max_distance;
    }
    if (h) part_start[excl[i]] = i;
}

// Coordinates come from BAM records, the bases from the FASTA: a FASTA contig shorter than the BAM header says (or absent:
// length 0) must not be read past its end.  FastaFile.fetch truncates at the contig end, so every piece is clamped to the
// FASTA length of its contig (`lo` / `hi` already are, SVIM_COMBINE.py:45-46,68-69).
__device__ void make_desc(const svb_row& r, long long lo, long long hi, const PairArgs& a, HapDesc& d) {
    d.m_kind = HAP_MID_NONE; d.m_len = 0; d.m_base = 0; d.m_unit = 1; d.seq_sel = 0;
    if (r.type == SVB_DEL || r.type == SVB_INV || r.type == SVB_DUP_TAN) {
        const uint64_t base = a.ref_off[r.src_tid];
        const long long clen = static_cast<long long>(a.ref_off[r.src_tid + 1] - base);
        const long long s = min(static_cast<long long>(r.src_start), clen), e = min(static_cast<long long>(r.src_end), clen);
        d.l_base = base + static_cast<uint64_t>(lo);
        d.l_len = static_cast<uint32_t>(s > lo ? s - lo : 0);
        d.r_base = base + static_cast<uint64_t>(e);
        d.r_len = static_cast<uint32_t>(hi > e ? hi - e : 0);
        if (r.type == SVB_INV) {
            d.m_kind = HAP_MID_REVCOMP; d.m_base = base + static_cast<uint64_t>(s); d.m_len = static_cast<uint32_t>(e > s ? e - s : 0);
        } else if (r.type == SVB_DUP_TAN) {
            const uint32_t unit = static_cast<uint32_t>(e > s ? e - s : 0);
            d.m_kind = HAP_MID_REPEAT; d.m_base = base + static_cast<uint64_t>(s); d.m_unit = unit ? unit : 1u;
            d.m_len = unit * static_cast<uint32_t>(r.copies + 1 > 0 ? r.copies + 1 : 0);
        }
    } else {   // INS, DUP_INT: the window is anchored on dest_start (SVIM_COMBINE.py:68-69,92-93)
        const uint64_t base = a.ref_off[r.dst_tid];
        const long long clen = static_cast<long long>(a.ref_off[r.dst_tid + 1] - base);
        const long long s = min(static_cast<long long>(r.dst_start), clen);
        d.l_base = base + static_cast<uint64_t>(lo);
        d.l_len = static_cast<uint32_t>(s > lo ? s - lo : 0);
        d.r_base = base + static_cast<uint64_t>(s);
        d.r_len = static_cast<uint32_t>(hi > s ? hi - s : 0);
        if (r.type == SVB_INS) {
            d.m_kind = HAP_MID_SEQ4; d.seq_sel = r.hap == 2 ? 1u : 0u;
            if (r.reserved0 != ~0ull) {
                d.m_base = r.reserved0;
                d.m_len = r.seq_len;
            }
        } else {
            const bool src_ok = r.src_tid >= 0 && r.src_tid < a.ref_n_contig;
            const uint64_t sbase = src_ok ? a.ref_off[r.src_tid] : 0ull;
            const long long slen = src_ok ? static_cast<long long>(a.ref_off[r.src_tid + 1] - sbase) : 0ll;
            const long long ss = min(static_cast<long long>(r.src_start), slen), se = min(static_cast<long long>(r.src_end), slen);
            if (!src_ok) atomicOr(a.dev_status, DEV_ERR_BAD_TID);
            d.m_kind = HAP_MID_REF; d.m_base = sbase + static_cast<uint64_t>(ss);
            d.m_len = static_cast<uint32_t>(se > ss ? se - ss : 0);
        }
    }
}

// one thread per partition: the cross-haplotype pairs that need an edit distance
// (launched before the host knows the number of partitions: the bound is read from the device)
__device__ void enumerate_partition(const PairArgs& a, uint32_t p) {
    const uint32_t first = __ldcg(a.part_start + p), n = __ldcg(a.part_start + p + 1) - first;
    if (n < 2u || n > static_cast<uint32_t>(PAIR_MAX)) return;
    const svb_row r0 = a.rows[__ldcg(a.order + first)];
    if (r0.type == SVB_BND) return;
    for (uint32_t i = 0; i + 1 < n; ++i) {
        const svb_row ri = a.rows[__ldcg(a.order + first + i)];
        for (uint32_t j = i + 1; j < n; ++j) {
            const svb_row rj = a.rows[__ldcg(a.order + first + j)];
            if (ri.hap == rj.hap) continue;
            const bool by_source = ri.type == SVB_DEL || ri.type == SVB_INV || ri.type == SVB_DUP_TAN;
            const int32_t tid = by_source ? ri.src_tid : ri.dst_tid;
            if (tid < 0 || tid >= a.ref_n_contig) { atomicOr(a.dev_status, DEV_ERR_BAD_TID); continue; }
            const long long clen = static_cast<long long>(a.ref_off[tid + 1] - a.ref_off[tid]);   // reference.get_reference_length
            long long lo, hi;
            if (by_source) {
                lo = max(0ll, static_cast<long long>(min(ri.src_start, rj.src_start)) - 100);
                hi = min(clen, static_cast<long long>(max(ri.src_end, rj.src_end)) + 100);
            } else {
                lo = max(0ll, static_cast<long long>(min(ri.dst_start, rj.dst_start)) - 100);
                hi = min(clen, static_cast<long long>(max(ri.dst_start, rj.dst_start)) + 100);
            }
            EditJob job;
            make_desc(ri, lo, hi, a, job.a);
            make_desc(rj, lo, hi, a, job.b);
            const unsigned long long slot = atomicAdd(a.job_count, 1ull);
            job.out_index = static_cast<uint32_t>(slot);
            job.pad = 0;
            a.jobs[slot] = job;
            atomicAdd(a.cell_count, static_cast<unsigned long long>(hap_length(job.a)) * hap_length(job.b));
            a.job_slot[static_cast<size_t>(p) * PAIR_DIST_STRIDE + static_cast<uint32_t>(link_cidx(static_cast<int>(n), static_cast<int>(i), static_cast<int>(j)))] =
                static_cast<uint32_t>(slot);
        }
    }
}

__global__ void enumerate_jobs_kernel(const PairArgs a, const unsigned long long* __restrict__ n_parts_dev) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= static_cast<uint32_t>(*n_parts_dev)) return;
    enumerate_partition(a, p);
}

// ---- K6 + K7 + job enumeration in ONE launch ---------------------------------------------------------------------------
// For a diploid human sample the pairing front is latency, not work: 18,000 keys, 40 key bits, and until round 2 a chain of
// 22 launches (keys, 5 x (histogram, scan, scatter), heads, scan, partition starts, jobs: 0.14 ms of launch gaps and 6 us
// kernels).  Here a handful of co-resident CTAs (cooperative launch; one per 512-key tile, at most FRONT_MAX_TILES) run the
// same phases back to back, separated by a grid barrier: a counter in global memory that every CTA bumps once per phase.
// The per-tile histograms live in global memory as before (digit-major); instead of a separate scan launch every CTA
// rebuilds the offsets of its own tile from them (256 x tiles words, read from L2).
constexpr uint32_t FRONT_MAX_TILES = 148;

struct FrontArgs {
    const svb_row* h1; const svb_row* h2; uint32_t n1, n2;
    const uint64_t* pool_off1; const uint64_t* pool_off2; const uint64_t* seq_off1; const uint64_t* seq_off2;
    uint32_t rank_bits, key_bits, n_tiles;
    long long max_distance;
    svb_row* rows;
    unsigned long long* keys[2];
    uint32_t* vals[2];
    uint32_t* hist;                  // [256][n_tiles]
    uint32_t* tile_heads;            // [n_tiles]
    uint32_t* part_start;
    unsigned long long* n_parts_dev;
    unsigned int* barrier;           // zeroed by the host
    PairArgs pa;                     // pa.order must point at the buffer the LAST pass writes
};

__device__ __forceinline__ void front_barrier(unsigned int* counter, unsigned int& phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++phase;
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned int target = phase * gridDim.x;
        while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RS_THREADS) pair_front_kernel(const __grid_constant__ FrontArgs f) {
    __shared__ uint32_t s_hist[256], s_base[256], s_running[256], s_scan[256];
    __shared__ uint32_t s_wc[RS_THREADS / 32][256];
    __shared__ uint32_t s_tile_off[FRONT_MAX_TILES + 1];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n = f.n1 + f.n2;
    unsigned int phase = 0;

    // ---- keys (concat_keys_kernel)
    for (uint32_t i = blockIdx.x * RS_THREADS + tid; i < n; i += gridDim.x * RS_THREADS) {
        svb_row r = i < f.n1 ? f.h1[i] : f.h2[i - f.n1];
        r.hap = i < f.n1 ? 1 : 2;
        r.reserved0 = ~0ull;
        if (r.type == SVB_INS) {
            const uint64_t* pool_off = i < f.n1 ? f.pool_off1 : f.pool_off2;
            const uint64_t* seq_off = i < f.n1 ? f.seq_off1 : f.seq_off2;
            if (pool_off) r.reserved0 = pool_off[i < f.n1 ? i : i - f.n1] * 2ull;
            else if (seq_off) r.reserved0 = seq_off[r.aln_idx] * 2ull + r.seq_pos;
            else atomicOr(f.pa.dev_status, DEV_ERR_NOSEQ);
        }
        f.rows[i] = r;
        int32_t ktid, pos;
        key_of(r, ktid, pos);
        uint32_t rank = 0;
        if (ktid < 0 || ktid >= f.pa.n_contig) atomicOr(f.pa.dev_status, DEV_ERR_BAD_TID);
        else rank = static_cast<uint32_t>(f.pa.contig_lexrank[ktid]);
        f.keys[0][i] = (static_cast<unsigned long long>(r.type) << (32u + f.rank_bits)) | (static_cast<unsigned long long>(rank) << 32) |
                       static_cast<uint32_t>(pos);
        f.vals[0][i] = i;
    }
    front_barrier(f.barrier, phase);

    // ---- stable LSD radix sort, 8 bits per pass
    int cur = 0;
    for (uint32_t shift = 0; shift < f.key_bits; shift += 8) {
        const unsigned long long* kin = f.keys[cur];
        const uint32_t* vin = f.vals[cur];
        unsigned long long* kout = f.keys[cur ^ 1];
        uint32_t* vout = f.vals[cur ^ 1];
        for (uint32_t t = blockIdx.x; t < f.n_tiles; t += gridDim.x) {          // histogram of every tile of this CTA
            s_hist[tid] = 0;
            __syncthreads();
            const uint32_t base = t * RS_TILE;
            for (int r = 0; r < RS_ROUNDS; ++r) {
                const uint32_t i = base + r * RS_THREADS + tid;
                if (i < n) atomicAdd(&s_hist[(__ldcg(kin + i) >> shift) & 255u], 1u);
            }
            __syncthreads();
            f.hist[tid * f.n_tiles + t] = s_hist[tid];
            __syncthreads();
        }
        front_barrier(f.barrier, phase);
        for (uint32_t t = blockIdx.x; t < f.n_tiles; t += gridDim.x) {
            // offsets of tile t: keys with a smaller digit (all tiles) + keys with this digit in earlier tiles
            uint32_t total = 0, before = 0;
            for (uint32_t q = 0; q < f.n_tiles; ++q) {
                const uint32_t c = __ldcg(f.hist + tid * f.n_tiles + q);
                total += c;
                if (q < t) before += c;
            }
            // exclusive scan of `total` over the 256 digits (one per thread)
            uint32_t inc = total;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
                if (static_cast<int>(lane) >= d) inc += y;
            }
            if (lane == 31u) s_scan[warp] = inc;
            __syncthreads();
            uint32_t warp_before = 0;
            for (uint32_t w = 0; w < warp; ++w) warp_before += s_scan[w];
            s_base[tid] = warp_before + inc - total + before;
            s_running[tid] = 0;
            __syncthreads();
            const uint32_t base = t * RS_TILE;
            for (int r = 0; r < RS_ROUNDS; ++r) {                                 // radix_scatter_kernel's rounds
#pragma unroll
                for (int w = 0; w < RS_THREADS / 32; ++w) s_wc[w][tid] = 0;
                __syncthreads();
                const uint32_t i = base + r * RS_THREADS + tid;
                const bool valid = i < n;
                unsigned long long key = 0;
                uint32_t val = 0, d = 0x100u | lane;                              // invalid lanes match nobody
                if (valid) {
                    key = __ldcg(kin + i);
                    val = __ldcg(vin + i);
                    d = static_cast<uint32_t>(key >> shift) & 255u;
                }
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
                if (valid && rank_in_warp == 0u) s_wc[warp][d] = __popc(peers);
                __syncthreads();
                if (valid) {
                    uint32_t ahead = 0;
                    for (uint32_t w = 0; w < warp; ++w) ahead += s_wc[w][d];
                    const uint32_t dst = s_base[d] + s_running[d] + ahead + rank_in_warp;      // stable: round, warp, lane order
                    kout[dst] = key;
                    vout[dst] = val;
                }
                __syncthreads();
                uint32_t add = 0;
#pragma unroll
                for (int w = 0; w < RS_THREADS / 32; ++w) add += s_wc[w][tid];
                s_running[tid] += add;
                __syncthreads();
            }
        }
        front_barrier(f.barrier, phase);
        cur ^= 1;
    }
    const unsigned long long* keys = f.keys[cur];

    // ---- partition heads (heads_kernel): count per tile, offsets, starts
    auto is_head = [&](uint32_t i) -> uint32_t {
        if (i == 0) return 1u;
        const unsigned long long a = __ldcg(keys + i - 1), b = __ldcg(keys + i);
        const long long pa = static_cast<int32_t>(static_cast<uint32_t>(a)), pb = static_cast<int32_t>(static_cast<uint32_t>(b));
        const long long gap = pa > pb ? pa - pb : pb - pa;
        return ((a >> 32) != (b >> 32) || gap > f.max_distance) ? 1u : 0u;
    };
    for (uint32_t t = blockIdx.x; t < f.n_tiles; t += gridDim.x) {
        uint32_t mine = 0;
        for (int r = 0; r < RS_ROUNDS; ++r) {
            const uint32_t i = t * RS_TILE + r * RS_THREADS + tid;
            if (i < n) mine += is_head(i);
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if (lane == 0) s_scan[warp] = mine;
        __syncthreads();
        if (tid == 0) {
            uint32_t sum = 0;
            for (int w = 0; w < RS_THREADS / 32; ++w) sum += s_scan[w];
            f.tile_heads[t] = sum;
        }
        __syncthreads();
    }
    front_barrier(f.barrier, phase);
    if (tid == 0) {
        uint32_t sum = 0;
        for (uint32_t q = 0; q < f.n_tiles; ++q) {
            s_tile_off[q] = sum;
            sum += __ldcg(f.tile_heads + q);
        }
        s_tile_off[f.n_tiles] = sum;
        if (blockIdx.x == 0) {
            *f.n_parts_dev = sum;
            f.part_start[sum] = n;
        }
    }
    __syncthreads();
    for (uint32_t t = blockIdx.x; t < f.n_tiles; t += gridDim.x) {
        uint32_t running = s_tile_off[t];
        for (int r = 0; r < RS_ROUNDS; ++r) {                 // rows of a tile in order: round, then thread
            const uint32_t i = t * RS_TILE + r * RS_THREADS + tid;
            const uint32_t h = i < n ? is_head(i) : 0u;
            const uint32_t bal = __ballot_sync(0xffffffffu, h != 0u);
            if (lane == 0) s_scan[warp] = __popc(bal);
            __syncthreads();
            uint32_t ahead = 0, round_total = 0;
            for (uint32_t w = 0; w < RS_THREADS / 32; ++w) {
                if (w < warp) ahead += s_scan[w];
                round_total += s_scan[w];
            }
            if (h) f.part_start[running + ahead + __popc(bal & ((1u << lane) - 1u))] = i;
            running += round_total;
            __syncthreads();
        }
    }
    front_barrier(f.barrier, phase);

    // ---- the cross-haplotype pairs of every partition (enumerate_jobs_kernel)
    const uint32_t n_parts = static_cast<uint32_t>(*reinterpret_cast<volatile unsigned long long*>(f.n_parts_dev));
    for (uint32_t p = blockIdx.x * RS_THREADS + tid; p < n_parts; p += gridDim.x * RS_THREADS) enumerate_partition(f.pa, p);
}

// Which pairs need their EXACT distance?  The wavefront kernel (wfa.cu) left, for every cross-haplotype pair, either the
// distance (it is <= t), or an interval [lo, hi] above t, or "unknown".  Complete linkage + fcluster only COMPARE distances
// (linkage.cuh), so the clustering is run once on the intervals: if no comparison's outcome depended on where inside their
// intervals the far values lie, any representative (the cluster kernel uses `lo`) gives the labels of the true values and
// nothing has to be computed.  A partition of two has a single pair above t: never.  A shared variant next to a private
// one, or two shared variants close together: the far values only ever meet exact values below t or the
// same-haplotype constant above every interval: never.  Two far pairs that are compared with each other (App. D's
// A2, B, A, A3): their order decides scipy's labels, all far pairs of that partition go to the exact kernel.
// One thread per partition.
__global__ void resolve_kernel(const PairArgs a) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= static_cast<uint32_t>(*a.n_parts_dev)) return;
    const uint32_t first = a.part_start[p], n = a.part_start[p + 1] - first;
    if (n < 2u || n > static_cast<uint32_t>(PAIR_MAX)) return;
    if (a.rows[a.order[first]].type == SVB_BND) return;
    uint8_t hap[PAIR_MAX];
    for (uint32_t i = 0; i < n; ++i) hap[i] = a.rows[a.order[first + i]].hap;
    LinkInterval D[PAIR_DIST_STRIDE];
    bool unknown = a.all_exact != 0, any_far = false;
    int m = 0;
    for (uint32_t i = 0; i + 1 < n; ++i)
        for (uint32_t j = i + 1; j < n; ++j, ++m) {
            LinkInterval v;
            v.id = 0xFFFFFFFFu;
            if (hap[i] == hap[j]) {
                v.lo = v.hi = 1000000000.0;            // SVIM_COMBINE.py:40-41
            } else {
                v.id = a.job_slot[static_cast<size_t>(p) * PAIR_DIST_STRIDE + m];
                if (a.all_exact) {
                    v.lo = 0.0; v.hi = 1.0 / 0.0;
                } else {
                    v.lo = a.dist[v.id];
                    v.hi = a.dist_hi[v.id];
                }
                if (v.lo < 0.0) unknown = true;
                if (v.lo < v.hi) any_far = true;
            }
            D[m] = v;
        }
    if (!unknown && !any_far) return;
    bool need = unknown;
    if (!need && n > 2u) {
        LinkInterval work[PAIR_DIST_STRIDE];
        int labels[PAIR_MAX];
        for (int x = 0; x < m; ++x) work[x] = D[x];
        need = !link_labels_determined(static_cast<int>(n), work, a.max_edit_distance, labels);
    }
    if (!need) return;
    for (int x = 0; x < m; ++x)
        if (D[x].id != 0xFFFFFFFFu && (a.all_exact || D[x].lo < D[x].hi || D[x].lo < 0.0))
            a.exact_list[atomicAdd(a.exact_count, 1ull)] = D[x].id;
}

// pair_candidates' per-cluster rules (SVIM_COMBINE.py:184-365) for one cluster of 1 or 2 members
__device__ void emit_cluster(const svb_row& first, const svb_row* second, const PairArgs& a, svb_row* out, uint32_t slot) {
    svb_row r = first;
    r.mate_aln = 0xFFFFFFFFu;
    if (!second) {
        r.genotype = first.hap == 1 ? SVB_GT_HAP1 : SVB_GT_HAP2;                          // "1/0" / "0/1"
    } else {
        r.genotype = SVB_GT_HOM;
        r.mate_aln = second->aln_idx;                                                      // reads = c0.reads + c1.reads
        if (r.type != SVB_BND) r.flags = first.flags | (second->flags & (SVB_F_COMPLETE | SVB_F_FULLY_COVERED | SVB_F_CUTPASTE));
        if (r.type == SVB_DUP_TAN) {                                                       // round(mean([c0, c1])): half to even (:290)
            const long long s = static_cast<long long>(first.copies) + second->copies;
            long long k = s >> 1;
            if (s & 1) k += (k & 1);
            r.copies = static_cast<int32_t>(k);
        }
    }
    if (r.type == SVB_BND)       // the constructor runs again on the stored fields (:342-363)
        wk_fill_bnd(r, a.contig_len, a.contig_lexrank, first.src_tid, first.src_start, (first.flags & SVB_F_SRC_FWD) != 0,
                    first.dst_tid, first.dst_start, (first.flags & SVB_F_DST_FWD) != 0);
    r.ordinal = slot;
    r.reserved0 = 0;
    out[slot] = r;
}

// pair_haplotypes / pair_haplotypes_breakends for one partition (SVIM_COMBINE.py:120-161): flat-cluster labels in scipy's
// numbering.  Returns the number of clusters; labels[i] in 1 .. clusters.
__device__ int partition_labels(const PairArgs& a, uint32_t p, uint32_t first, uint32_t n, int* labels) {
    const bool bnd = a.rows[a.order[first]].type == SVB_BND;
    const double threshold = bnd ? 0.3 : a.max_edit_distance;
    double D[PAIR_DIST_STRIDE];
    int m = 0;
    for (uint32_t i = 0; i + 1 < n; ++i) {
        const svb_row& ri = a.rows[a.order[first + i]];
        for (uint32_t j = i + 1; j < n; ++j, ++m) {
            const svb_row& rj = a.rows[a.order[first + j]];
            double d;
            if (bnd) {                                                                 // :105-117 (dest contig never compared)
                const bool same_dirs = ((ri.flags ^ rj.flags) & (SVB_F_SRC_FWD | SVB_F_DST_FWD)) == 0;
                if (ri.hap != rj.hap && same_dirs) {
                    const long long d1 = static_cast<long long>(ri.src_start) - rj.src_start;
                    const long long d2 = static_cast<long long>(ri.dst_start) - rj.dst_start;
                    d = static_cast<double>((d1 < 0 ? -d1 : d1) + (d2 < 0 ? -d2 : d2)) / 3000.0;
                } else {
                    d = 99999.0;
                }
            } else if (ri.hap == rj.hap) {
                d = 1000000000.0;                                                      // :40-41
            } else {
                d = a.dist[a.job_slot[static_cast<size_t>(p) * PAIR_DIST_STRIDE + m]];
            }
            D[m] = d;
        }
    }
    if (n == 2u) {          // one merge at D[0]: a single cluster when it is within the cut, else the two leaves left to right
        labels[0] = 1;
        labels[1] = D[0] <= threshold ? 1 : 2;
        return labels[1];
    }
    return link_complete_fcluster(static_cast<int>(n), D, threshold, labels);
}

// clusters of 1 or 2 members become rows (other sizes: logged, skipped, :204-205); WRITE = false only counts them
template <bool WRITE>
__device__ uint32_t partition_rows(const PairArgs& a, uint32_t first, uint32_t n, const int* labels, int n_clusters, uint32_t slot) {
    uint32_t produced = 0;
    for (int c = 1; c <= n_clusters; ++c) {
        int members = 0, i0 = -1, i1 = -1;
        for (uint32_t i = 0; i < n; ++i)
            if (labels[i] == c) {
                if (members == 0) i0 = static_cast<int>(i);
                else if (members == 1) i1 = static_cast<int>(i);
                ++members;
            }
        if (members == 1 || members == 2) {
            if (WRITE) {
                const svb_row f = a.rows[a.order[first + i0]];
                if (members == 2) {
                    const svb_row s = a.rows[a.order[first + i1]];
                    emit_cluster(f, &s, a, a.out, slot + produced);
                } else {
                    emit_cluster(f, nullptr, a, a.out, slot + produced);
                }
            }
            ++produced;
        }
    }
    return produced;
}

template <bool WRITE>
__global__ void __launch_bounds__(64) cluster_kernel(const PairArgs a) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= static_cast<uint32_t>(*a.n_parts_dev)) {
        if (!WRITE && p <= a.n_rows) a.counts[p] = 0u;          // the scan runs over the upper bound
        return;
    }
    const uint32_t first = a.part_start[p], n = a.part_start[p + 1] - first;
    const uint32_t slot = WRITE ? a.counts[p] : 0u;
    uint32_t produced = 0;
    if (n == 1u) {
        if (WRITE) emit_cluster(a.rows[a.order[first]], nullptr, a, a.out, slot);
        produced = 1;
    } else if (n <= static_cast<uint32_t>(PAIR_MAX)) {
        int labels[PAIR_MAX];
        const int n_clusters = partition_labels(a, p, first, n, labels);
        produced = partition_rows<WRITE>(a, first, n, labels, n_clusters, slot);
    }
    if (!WRITE) a.counts[p] = produced;
}

// ---- K9 in ONE launch: labels, row counts, their prefix over the partitions, rows ---------------------------------------
// (count pass + scan launch + write pass until round 2: the linkage of every partition ran twice and the row offsets made
// a round trip through a single-CTA scan).  Cooperative launch: tile = 64 partitions; phase 1 computes the labels of a
// partition once and keeps them packed (4 bits per member) next to its row count; one grid barrier; phase 2 derives the
// tile's first row from the tile sums and writes the rows.
struct ClusterFusedArgs {
    PairArgs pa;
    unsigned long long* packed;      // [n_rows] labels (4 bits each), clusters << 40, rows << 48
    uint32_t* tile_sum;              // [ceil(n_rows / 64)]
    unsigned int* barrier;
    unsigned long long* total;       // rows written
};

__global__ void __launch_bounds__(64) cluster_fused_kernel(const __grid_constant__ ClusterFusedArgs c) {
    const PairArgs& a = c.pa;
    __shared__ uint32_t s_w[2];
    __shared__ uint32_t s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_parts = static_cast<uint32_t>(*a.n_parts_dev), n_tiles = (n_parts + 63u) / 64u;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint32_t p = t * 64u + tid;
        uint32_t produced = 0;
        if (p < n_parts) {
            const uint32_t first = a.part_start[p], n = a.part_start[p + 1] - first;
            unsigned long long pk = 0;
            if (n == 1u) {
                produced = 1;
            } else if (n <= static_cast<uint32_t>(PAIR_MAX)) {
                int labels[PAIR_MAX];
                const int n_clusters = partition_labels(a, p, first, n, labels);
                produced = partition_rows<false>(a, first, n, labels, n_clusters, 0u);
                for (uint32_t i = 0; i < n; ++i) pk |= static_cast<unsigned long long>(labels[i]) << (4u * i);
                pk |= static_cast<unsigned long long>(n_clusters) << 40;
            }
            c.packed[p] = pk | (static_cast<unsigned long long>(produced) << 48);
        }
        const uint32_t wsum = __reduce_add_sync(0xffffffffu, produced);
        if (lane == 0) s_w[warp] = wsum;
        __syncthreads();
        if (tid == 0) c.tile_sum[t] = s_w[0] + s_w[1];
        __syncthreads();
    }
    unsigned int phase = 0;
    front_barrier(c.barrier, phase);
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        uint32_t before = 0;
        for (uint32_t q = tid; q < t; q += 64u) before += __ldcg(c.tile_sum + q);
        before = __reduce_add_sync(0xffffffffu, before);
        if (lane == 0) s_w[warp] = before;
        __syncthreads();
        if (tid == 0) s_base = s_w[0] + s_w[1];
        __syncthreads();
        const uint32_t base = s_base;
        const uint32_t p = t * 64u + tid;
        const unsigned long long pk = p < n_parts ? c.packed[p] : 0ull;
        const uint32_t produced = static_cast<uint32_t>(pk >> 48);
        uint32_t inc = produced;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
            if (static_cast<int>(lane) >= d) inc += y;
        }
        if (lane == 31u) s_w[warp] = inc;
        __syncthreads();
        const uint32_t slot = base + (warp ? s_w[0] : 0u) + inc - produced;
        if (p < n_parts) {
            const uint32_t first = a.part_start[p], n = a.part_start[p + 1] - first;
            if (n == 1u) {
                emit_cluster(a.rows[a.order[first]], nullptr, a, a.out, slot);
            } else if (n <= static_cast<uint32_t>(PAIR_MAX)) {
                int labels[PAIR_MAX];
                for (uint32_t i = 0; i < n; ++i) labels[i] = static_cast<int>((pk >> (4u * i)) & 15ull);
                partition_rows<true>(a, first, n, labels, static_cast<int>((pk >> 40) & 255ull), slot);
            }
        }
        if (t == n_tiles - 1u && tid == 63u) *c.total = static_cast<unsigned long long>(slot) + produced;
        __syncthreads();
    }
    if (n_tiles == 0u && blockIdx.x == 0 && tid == 0) *c.total = 0ull;
}

}  // namespace

// One stream-ordered chain, ONE synchronisation at the end: every count the next kernel needs (partitions, jobs, pairs for
// the exact kernel, output rows) stays on the device; grids cover the upper bound h1 + h2.
static int run_pairing_once(svb_ctx* ctx, const svb_table* h1, const svb_table* h2, const svb_records* rec1, const svb_records* rec2,
                            const svb_ref* ref, const svb_params* p, svb_table** out, unsigned long long* need_out) {
    *out = nullptr;
    *need_out = 0;
    const uint64_t n64 = h1->n + h2->n;
    if (n64 >= 0x7fffffffull) return svb_fail(ctx, SVB_ERR_ARG, "svb_pair: too many candidates");
    const uint32_t n1 = static_cast<uint32_t>(h1->n), n2 = static_cast<uint32_t>(h2->n), n = n1 + n2;
    svb_table* result = new (std::nothrow) svb_table();
    if (!result) return svb_fail(ctx, SVB_ERR_NOMEM, "svb_pair");
    result->device = ctx->device;
    result->stream = ctx->stream;
    if (n == 0) {
        *reinterpret_cast<uint32_t*>(ctx->h_pinned + 12) = 0u;       // nothing ran: no device status to report
        SVB_CUDA(ctx, cudaMallocAsync(&result->d_rows, sizeof(svb_row), ctx->stream));
        result->cap = 1;
        *out = result;
        return SVB_OK;
    }
    if (!rec1 || !rec2) { delete result; return svb_fail(ctx, SVB_ERR_ARG, "svb_pair: records of both haplotypes are required"); }
    // rows of both tables are keyed, clamped and named by tid: the two record images must share one contig table
    // (pair_candidates takes every name and length from the first BAM, SVIM_COMBINE.py:164-366)
    // (a table that brings its own sequence pool does not consult its record image: it may have been renumbered already)
    if (!h2->d_pool_off && rec1->h_contig_len != rec2->h_contig_len) {
        delete result;
        return svb_fail(ctx, SVB_ERR_ARG, "svb_pair: the two record images list different contigs; renumber the second table onto the first header");
    }
    const svb_records* rec = rec1;
    uint32_t rank_bits = 1;
    while ((1u << rank_bits) < static_cast<uint32_t>(std::max(rec->n_contig, 1))) ++rank_bits;
    const uint32_t key_bits = 32u + rank_bits + 3u;
    const uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
    const uint32_t max_jobs = static_cast<uint32_t>(std::min<uint64_t>(static_cast<uint64_t>(n) * 5 / 2 + 1, 0x7fffffffull));

    // one slab for everything that lives only during this call
    auto align = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
    const size_t sz_rows = align(sizeof(svb_row) * n), sz_keys = align(sizeof(unsigned long long) * n), sz_vals = align(sizeof(uint32_t) * n);
    const size_t sz_hist = align(sizeof(uint32_t) * (256ull * n_blocks + 1)), sz_head = align(sizeof(uint32_t) * (static_cast<size_t>(n) + 2));
    const size_t sz_slot = align(sizeof(uint32_t) * PAIR_DIST_STRIDE * n), sz_jobs = align(sizeof(EditJob) * max_jobs);
    const size_t sz_dist = align(sizeof(double) * max_jobs), sz_list = align(sizeof(uint32_t) * max_jobs), sz_big = align(sizeof(uint4) * max_jobs);
    const size_t total = sz_rows + 2 * sz_keys + 2 * sz_vals + sz_hist + 3 * sz_head + sz_slot + sz_jobs + 2 * sz_dist + sz_list + sz_big;
    unsigned char* slab = nullptr;
    if (cudaMallocAsync(&slab, total, ctx->stream) != cudaSuccess) { delete result; return svb_fail(ctx, SVB_ERR_NOMEM, "svb_pair: scratch"); }
    unsigned char* cur = slab;
    auto carve = [&](size_t bytes) { unsigned char* q = cur; cur += bytes; return q; };
    svb_row* rows = reinterpret_cast<svb_row*>(carve(sz_rows));
    unsigned long long* keys[2] = {reinterpret_cast<unsigned long long*>(carve(sz_keys)), reinterpret_cast<unsigned long long*>(carve(sz_keys))};
    uint32_t* vals[2] = {reinterpret_cast<uint32_t*>(carve(sz_vals)), reinterpret_cast<uint32_t*>(carve(sz_vals))};
    uint32_t* hist = reinterpret_cast<uint32_t*>(carve(sz_hist));
    uint32_t* head = reinterpret_cast<uint32_t*>(carve(sz_head));
    uint32_t* part_start = reinterpret_cast<uint32_t*>(carve(sz_head));
    uint32_t* counts = reinterpret_cast<uint32_t*>(carve(sz_head));
    uint32_t* job_slot = reinterpret_cast<uint32_t*>(carve(sz_slot));
    EditJob* jobs = reinterpret_cast<EditJob*>(carve(sz_jobs));
    double* dist = reinterpret_cast<double*>(carve(sz_dist));
    double* dist_hi = reinterpret_cast<double*>(carve(sz_dist));
    uint32_t* exact_list = reinterpret_cast<uint32_t*>(carve(sz_list));
    uint4* big = reinterpret_cast<uint4*>(carve(sz_big));

    auto fail = [&](int rc) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeAsync(slab, ctx->stream);
        if (result->d_rows) cudaFreeAsync(result->d_rows, ctx->stream);
        delete result;
        return rc;
    };
#define PAIR_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) return fail(svb_fail(ctx, SVB_ERR_CUDA, #call, e__));     \
    } while (0)

    // device counters of this call: [2] partitions, [3] jobs, [4] pairs for the exact kernel, [5] output rows,
    // [6] longest text the exact kernel could not park, [7..8] the wavefront kernel's three 32-bit counters
    unsigned long long* cnt = ctx->d_counters;
    PAIR_CUDA(cudaMemsetAsync(cnt + 2, 0, 9 * sizeof(unsigned long long), ctx->stream));      // [9] radix scan total, [10] table cells
    const uint32_t n_passes = (key_bits + 7u) / 8u;
    const bool fused_front = n_blocks <= FRONT_MAX_TILES && !getenv("SVB_PAIR_UNFUSED");
    int cur_buf = fused_front ? static_cast<int>(n_passes & 1u) : 0;

    PairArgs a;
    a.rows = rows;
    a.order = vals[cur_buf];          // (the unfused path sets it again once its passes are enqueued)
    a.part_start = part_start;
    a.n_parts_dev = cnt + 2;
    a.n_rows = n;
    a.contig_len = rec->d_contig_len;
    a.contig_lexrank = rec->d_contig_lexrank;
    a.n_contig = rec->n_contig;
    a.ref_off = ref->d_contig_off;
    a.ref_n_contig = ref->n_contig;
    a.max_edit_distance = static_cast<double>(p->max_edit_distance);
    a.dist = dist;
    a.dist_hi = dist_hi;
    a.job_slot = job_slot;
    a.jobs = jobs;
    a.job_count = cnt + 3;
    a.exact_list = exact_list;
    a.exact_count = cnt + 4;
    a.cell_count = cnt + 10;
    a.all_exact = (p->max_edit_distance < 0 || static_cast<uint32_t>(p->max_edit_distance) > WFA_MAX_T) ? 1 : 0;
    a.counts = counts;
    a.out = nullptr;
    a.dev_status = ctx->d_status;
    if (fused_front) {
        // keys, radix passes, partition starts and the job list in ONE cooperative launch (pair_front_kernel)
        FrontArgs f;
        f.h1 = h1->d_rows; f.h2 = h2->d_rows; f.n1 = n1; f.n2 = n2;
        f.pool_off1 = h1->d_pool_off; f.pool_off2 = h2->d_pool_off;
        f.seq_off1 = h1->d_pool_off ? nullptr : rec1->d_seq_off;
        f.seq_off2 = h2->d_pool_off ? nullptr : rec2->d_seq_off;
        f.rank_bits = rank_bits; f.key_bits = key_bits; f.n_tiles = n_blocks;
        f.max_distance = p->partition_max_distance;
        f.rows = rows;
        f.keys[0] = keys[0]; f.keys[1] = keys[1]; f.vals[0] = vals[0]; f.vals[1] = vals[1];
        f.hist = hist;
        f.tile_heads = head;
        f.part_start = part_start;
        f.n_parts_dev = cnt + 2;
        f.barrier = reinterpret_cast<unsigned int*>(cnt + 9);          // zeroed with the other counters above
        f.pa = a;
        void* kargs[] = {&f};
        KernelTimer timer(ctx, SVB_K_SORT);
        PAIR_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(pair_front_kernel), dim3(n_blocks), dim3(RS_THREADS), kargs, 0, ctx->stream));
        ctx->launches += 1;
    } else {
        {
            KernelTimer timer(ctx, SVB_K_SORT);
            concat_keys_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(h1->d_rows, n1, h2->d_rows, n2, h1->d_pool_off, h2->d_pool_off,
                                                                         h1->d_pool_off ? nullptr : rec1->d_seq_off,
                                                                         h2->d_pool_off ? nullptr : rec2->d_seq_off, rec->d_contig_lexrank, rec->n_contig,
                                                                         rank_bits, rows, keys[0], vals[0], ctx->d_status);
            ctx->launches += 1;
            for (uint32_t shift = 0; shift < key_bits; shift += 8) {
                radix_hist_kernel<<<n_blocks, RS_THREADS, 0, ctx->stream>>>(keys[cur_buf], n, shift, hist, n_blocks);
                int rc = launch_scan_u32(ctx, hist, 256u * n_blocks, cnt + 9);
                if (rc != SVB_OK) return fail(rc);
                radix_scatter_kernel<<<n_blocks, RS_THREADS, 0, ctx->stream>>>(keys[cur_buf], vals[cur_buf], keys[cur_buf ^ 1], vals[cur_buf ^ 1], n,
                                                                               shift, hist, n_blocks);
                cur_buf ^= 1;
                ctx->launches += 2;
            }
            heads_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(keys[cur_buf], n, p->partition_max_distance, head);
            ctx->launches += 1;
            int rc = launch_scan_u32(ctx, head, n, cnt + 2);
            if (rc != SVB_OK) return fail(rc);
        }
        PAIR_CUDA(cudaGetLastError());
        a.order = vals[cur_buf];
        {
            KernelTimer timer(ctx, SVB_K_SORT);
            part_start_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(keys[cur_buf], head, n, p->partition_max_distance, part_start, cnt + 2);
            enumerate_jobs_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(a, cnt + 2);
            ctx->launches += 2;
        }
    }
    PAIR_CUDA(cudaGetLastError());
    // K8: thresholded wavefronts for every pair, then the exact kernel for the few pairs whose exact value can matter
    const uint8_t* seq_a = h1->d_pool_off ? h1->d_pool : rec1->d_seq4;
    const uint8_t* seq_b = h2->d_pool_off ? h2->d_pool : rec2->d_seq4;
    if (!a.all_exact) {
        WfaArgs w;
        w.jobs = jobs; w.n_jobs_dev = cnt + 3; w.counters = reinterpret_cast<unsigned int*>(cnt + 7);
        w.big = big; w.big_cap = max_jobs; w.ref = ref->d_bases; w.seq4_a = seq_a; w.seq4_b = seq_b; w.class_map = ref->d_class_map;
        w.dist = dist; w.dist_hi = dist_hi; w.t = static_cast<uint32_t>(p->max_edit_distance); w.cap_chars = 0; w.stage = 0;
        int rc = launch_wfa(ctx, w);
        if (rc != SVB_OK) return fail(rc);
    }
    {
        KernelTimer timer(ctx, SVB_K_EDIT_DISTANCE);
        resolve_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(a);
        ctx->launches += 1;
    }
    PAIR_CUDA(cudaGetLastError());
    {
        int rc = launch_edit_distance(ctx, jobs, exact_list, cnt + 4, max_jobs, ref->d_bases, seq_a, seq_b, ref->d_class_map, dist, dist_hi, cnt + 6);
        if (rc != SVB_OK) return fail(rc);
    }
    // pairing never makes rows (every output row is one input row or the merge of two): the table is allocated for n
    // rows and written without waiting for the exact count, which comes back with the final synchronisation
    result->cap = std::max<uint64_t>(n, 1);
    PAIR_CUDA(cudaMallocAsync(&result->d_rows, sizeof(svb_row) * result->cap, ctx->stream));
    a.out = result->d_rows;
    const unsigned cluster_tiles = (n + 63u) / 64u;
    if (cluster_tiles <= static_cast<unsigned>(ctx->sm_count) * 8u && !getenv("SVB_PAIR_UNFUSED")) {
        ClusterFusedArgs cf;
        cf.pa = a;
        cf.packed = reinterpret_cast<unsigned long long*>(keys[cur_buf ^ 1]);        // the sort's spare key buffer: n words, free by now
        cf.tile_sum = head;                                                       // free since the partition starts exist
        cf.barrier = reinterpret_cast<unsigned int*>(cnt + 9) + 1;                 // second word of the barrier slot (zeroed above)
        cf.total = cnt + 5;
        void* kargs[] = {&cf};
        KernelTimer timer(ctx, SVB_K_CLUSTER);
        PAIR_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(cluster_fused_kernel), dim3(cluster_tiles), dim3(64), kargs, 0, ctx->stream));
        ctx->launches += 1;
    } else {
        {
            KernelTimer timer(ctx, SVB_K_CLUSTER);
            cluster_kernel<false><<<(n + 64) / 64, 64, 0, ctx->stream>>>(a);
            ctx->launches += 1;
            int rc = launch_scan_u32(ctx, counts, n, cnt + 5);
            if (rc != SVB_OK) return fail(rc);
        }
        PAIR_CUDA(cudaGetLastError());
        {
            KernelTimer timer(ctx, SVB_K_CLUSTER);
            cluster_kernel<true><<<(n + 64) / 64, 64, 0, ctx->stream>>>(a);
            ctx->launches += 1;
        }
    }
    PAIR_CUDA(cudaGetLastError());
    PAIR_CUDA(cudaMemcpyAsync(ctx->h_pinned + 2, cnt + 2, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    uint32_t* h_status = reinterpret_cast<uint32_t*>(ctx->h_pinned + 12);      // the device error word rides along
    PAIR_CUDA(cudaMemcpyAsync(h_status, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PAIR_CUDA(cudaFreeAsync(slab, ctx->stream));
    PAIR_CUDA(cudaStreamSynchronize(ctx->stream));
    result->n = ctx->h_pinned[5];
    ctx->last_pair_stats[0] = ctx->h_pinned[2];
    ctx->last_pair_stats[1] = ctx->h_pinned[3];
    ctx->last_pair_stats[2] = ctx->h_pinned[4];
    ctx->last_pair_stats[3] = ctx->h_pinned[10];
    *need_out = ctx->h_pinned[6];
    if (*h_status) {                       // svb_pair turns it into the reference's error (check_device_status)
        *out = result;
        return SVB_OK;
    }
    if (result->n > result->cap) return fail(svb_fail(ctx, SVB_ERR_CAPACITY, "svb_pair: more rows out than in"));
#undef PAIR_CUDA
    *out = result;
    return SVB_OK;
}

int run_pairing(svb_ctx* ctx, const svb_table* h1, const svb_table* h2, const svb_records* rec1, const svb_records* rec2,
                const svb_ref* ref, const svb_params* p, svb_table** out) {
    for (int attempt = 0;; ++attempt) {
        unsigned long long need = 0;
        int rc = run_pairing_once(ctx, h1, h2, rec1, rec2, ref, p, out, &need);
        if (rc != SVB_OK || need == 0) return rc;
        // a multi-stripe table of the exact kernel had a text longer than a worker's slice of parked deltas (Mb-scale pair
        // that the wavefront kernel could not settle): grow the slices and run the call again
        if (*out) { svb_table_free(*out); *out = nullptr; }
        if (attempt == 1) return svb_fail(ctx, SVB_ERR_CAPACITY, "svb_pair: parked-delta stride");
        ctx->ed_stride = need + 4096;
    }
}

// ---- form_partitions for explicit keys (svb_form_partitions): the K6 sort + K7 split on caller-supplied keys ------
namespace {
__global__ void iota_kernel(uint32_t* v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
}  // namespace

int run_form_partitions(svb_ctx* ctx, const uint64_t* keys_host, uint32_t n, int64_t max_distance, uint32_t* order_out,
                        uint32_t* part_start_out, uint32_t* n_parts_out) {
    *n_parts_out = 0;
    if (!n) return SVB_OK;
    const uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
    auto align = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
    const size_t sz_keys = align(sizeof(unsigned long long) * n), sz_vals = align(sizeof(uint32_t) * (static_cast<size_t>(n) + 1));
    const size_t sz_hist = align(sizeof(uint32_t) * (256ull * n_blocks + 1));
    unsigned char* slab = nullptr;
    SVB_CUDA(ctx, cudaMallocAsync(&slab, 2 * sz_keys + 4 * sz_vals + sz_hist, ctx->stream));
    unsigned char* cur = slab;
    auto carve = [&](size_t bytes) { unsigned char* q = cur; cur += bytes; return q; };
    unsigned long long* keys[2] = {reinterpret_cast<unsigned long long*>(carve(sz_keys)), reinterpret_cast<unsigned long long*>(carve(sz_keys))};
    uint32_t* vals[2] = {reinterpret_cast<uint32_t*>(carve(sz_vals)), reinterpret_cast<uint32_t*>(carve(sz_vals))};
    uint32_t* head = reinterpret_cast<uint32_t*>(carve(sz_vals));
    uint32_t* part_start = reinterpret_cast<uint32_t*>(carve(sz_vals));
    uint32_t* hist = reinterpret_cast<uint32_t*>(carve(sz_hist));
    auto fail = [&](int rc) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeAsync(slab, ctx->stream);
        return rc;
    };
    cudaError_t e = cudaMemcpyAsync(keys[0], keys_host, sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return fail(svb_fail(ctx, SVB_ERR_CUDA, "svb_form_partitions upload", e));
    int cur_buf = 0;
    {
        KernelTimer timer(ctx, SVB_K_SORT);
        iota_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(vals[0], n);
        ctx->launches += 1;
        for (uint32_t shift = 0; shift < 64; shift += 8) {
            radix_hist_kernel<<<n_blocks, RS_THREADS, 0, ctx->stream>>>(keys[cur_buf], n, shift, hist, n_blocks);
            int rc = launch_scan_u32(ctx, hist, 256u * n_blocks, ctx->d_counters + 9);
            if (rc != SVB_OK) return fail(rc);
            radix_scatter_kernel<<<n_blocks, RS_THREADS, 0, ctx->stream>>>(keys[cur_buf], vals[cur_buf], keys[cur_buf ^ 1], vals[cur_buf ^ 1], n,
                                                                           shift, hist, n_blocks);
            cur_buf ^= 1;
            ctx->launches += 2;
        }
        heads_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(keys[cur_buf], n, max_distance, head);
        ctx->launches += 1;
        int rc = launch_scan_u32(ctx, head, n, ctx->d_counters + 2);
        if (rc != SVB_OK) return fail(rc);
        part_start_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(keys[cur_buf], head, n, max_distance, part_start, ctx->d_counters + 2);
        ctx->launches += 1;
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_pinned + 2, ctx->d_counters + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(order_out, vals[cur_buf], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(part_start_out, part_start, sizeof(uint32_t) * (static_cast<size_t>(n) + 1), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(svb_fail(ctx, SVB_ERR_CUDA, "svb_form_partitions", e));
    *n_parts_out = static_cast<uint32_t>(ctx->h_pinned[2]);
    cudaFreeAsync(slab, ctx->stream);
    return SVB_OK;
}

// ---- test hook: scipy-compatible flat cluster labels (svb_cluster_labels) ------------------------------
namespace {
__global__ void cluster_labels_kernel(const double* __restrict__ condensed, const uint32_t* __restrict__ n_points,
                                      const uint64_t* __restrict__ offsets, uint32_t n_problems, double threshold,
                                      int32_t* __restrict__ labels_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_problems) return;
    const int n = static_cast<int>(n_points[i]);
    double work[LINK_MAXN * (LINK_MAXN - 1) / 2];
    int labels[LINK_MAXN];
    const double* src = condensed + offsets[i];
    for (int k = 0; k < n * (n - 1) / 2; ++k) work[k] = src[k];
    link_complete_fcluster(n, work, threshold, labels);
    for (int k = 0; k < LINK_MAXN; ++k) labels_out[static_cast<size_t>(i) * LINK_MAXN + k] = k < n ? labels[k] : 0;
}
}  // namespace

int run_cluster_labels(svb_ctx* ctx, const double* condensed, const uint32_t* n_points, uint32_t n_problems,
                       double threshold, int32_t* labels_out) {
    if (!n_problems) return SVB_OK;
    std::vector<uint64_t> off(n_problems + 1, 0);
    for (uint32_t i = 0; i < n_problems; ++i) {
        if (n_points[i] > LINK_MAXN) return svb_fail(ctx, SVB_ERR_ARG, "svb_cluster_labels: more than 32 points");
        off[i + 1] = off[i] + static_cast<uint64_t>(n_points[i]) * (n_points[i] - 1) / 2;
    }
    double* d_c = nullptr;
    uint32_t* d_n = nullptr;
    uint64_t* d_o = nullptr;
    int32_t* d_l = nullptr;
    SVB_CUDA(ctx, cudaMallocAsync(&d_c, sizeof(double) * std::max<uint64_t>(off[n_problems], 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_n, sizeof(uint32_t) * n_problems, ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_o, sizeof(uint64_t) * (n_problems + 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_l, sizeof(int32_t) * LINK_MAXN * n_problems, ctx->stream));
    if (off[n_problems])
        SVB_CUDA(ctx, cudaMemcpyAsync(d_c, condensed, sizeof(double) * off[n_problems], cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_n, n_points, sizeof(uint32_t) * n_problems, cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_o, off.data(), sizeof(uint64_t) * (n_problems + 1), cudaMemcpyHostToDevice, ctx->stream));
    {
        KernelTimer timer(ctx, SVB_K_CLUSTER);
        cluster_labels_kernel<<<(n_problems + 63) / 64, 64, 0, ctx->stream>>>(d_c, d_n, d_o, n_problems, threshold, d_l);
        ctx->launches += 1;
    }
    SVB_CUDA(ctx, cudaGetLastError());
    SVB_CUDA(ctx, cudaMemcpyAsync(labels_out, d_l, sizeof(int32_t) * LINK_MAXN * n_problems, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(d_c, ctx->stream);
    cudaFreeAsync(d_n, ctx->stream);
    cudaFreeAsync(d_o, ctx->stream);
    cudaFreeAsync(d_l, ctx->stream);
    return SVB_OK;
}
