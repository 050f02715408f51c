// K8 edit_distance -- exact unit-cost global edit distance of haplotype-string pairs, one warp per pair.
//
// Replaces compute_distance's edlib.align(h1, h2)["editDistance"] (reference SVIM_COMBINE.py:35-102).
// The shorter string is the pattern: its rows are cut into 64-row blocks (edit_core.cuh), block b of a
// 2048-row stripe lives in lane b, and the columns of the text stream through the lanes as a software
// pipeline -- lane b works on column t-b at step t and hands (text class, horizontal delta) to lane b+1
// with one shuffle.  Patterns longer than 2048 rows take several stripes; the horizontal deltas of a
// stripe's last row are parked in a per-warp byte buffer in HBM.  Strings are never materialised:
// bases are read through HapDesc (reference bytes / 4-bit query bases resident in HBM).
// Bound: integer issue rate (about 60 instructions per column step per warp), not bandwidth.
#include "common.cuh"
#include "edit_core.cuh"
#include "pairing.cuh"

namespace {

constexpr int ED_WARPS = 4;

__global__ void __launch_bounds__(ED_WARPS * 32) edit_distance_kernel(const EditJob* __restrict__ jobs, uint32_t n_jobs,
                                                                       unsigned int* next_job,
                                                                       const uint8_t* __restrict__ ref,
                                                                       const uint8_t* __restrict__ seq4_a,
                                                                       const uint8_t* __restrict__ seq4_b,
                                                                       const uint8_t* __restrict__ class_map,
                                                                       signed char* hbuf_pool, uint64_t hbuf_stride,
                                                                       double* __restrict__ out) {
    __shared__ unsigned long long s_peq[ED_WARPS][ED_NCLASS][32];
    __shared__ uint8_t s_class[256];
    __shared__ uint8_t s_tcls[ED_WARPS][64];        // ring: symbol class of the band's columns
    __shared__ signed char s_th[ED_WARPS][64];       // ring: horizontal delta entering the stripe's top row
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) s_class[i] = class_map[i];
    __syncthreads();
    signed char* hbuf = hbuf_pool ? hbuf_pool + (static_cast<uint64_t>(blockIdx.x) * ED_WARPS + warp) * hbuf_stride : nullptr;

    // Two sweeps over the job list: patterns longer than one 2048-row stripe first (they are the long poles),
    // then the single-stripe jobs.
    for (int sweep = 0; sweep < 2; ++sweep) {
        while (true) {
            uint32_t job_id = 0;
            if (lane == 0) job_id = atomicAdd(next_job + sweep, 1u);
            job_id = __shfl_sync(0xffffffffu, job_id, 0);
            if (job_id >= n_jobs) break;
            const EditJob job = jobs[job_id];
            const uint32_t la0 = hap_length(job.a), lb0 = hap_length(job.b);
            if ((min(la0, lb0) > 2048u) != (sweep == 0)) continue;
            // The edit distance does not change when the common prefix and suffix are removed; the two haplotypes of
            // a shared variant are mostly identical, so this alone finishes most pairs (warp-parallel compare).
            const uint32_t lim = min(la0, lb0);
            uint32_t pre = 0;
            while (pre < lim) {
                const uint32_t i = pre + lane;
                const bool same = i < lim && hap_char(job.a, i, ref, seq4_a, seq4_b) == hap_char(job.b, i, ref, seq4_a, seq4_b);
                const uint32_t mask = __ballot_sync(0xffffffffu, same);
                if (mask != 0xffffffffu) { pre += static_cast<uint32_t>(__ffs(~mask) - 1); break; }
                pre += 32u;
            }
            pre = min(pre, lim);
            const uint32_t lim2 = lim - pre;
            uint32_t suf = 0;
            while (suf < lim2) {
                const uint32_t i = suf + lane;
                const bool same = i < lim2 && hap_char(job.a, la0 - 1u - i, ref, seq4_a, seq4_b) == hap_char(job.b, lb0 - 1u - i, ref, seq4_a, seq4_b);
                const uint32_t mask = __ballot_sync(0xffffffffu, same);
                if (mask != 0xffffffffu) { suf += static_cast<uint32_t>(__ffs(~mask) - 1); break; }
                suf += 32u;
            }
            suf = min(suf, lim2);
            const uint32_t la = la0 - pre - suf, lb = lb0 - pre - suf;
            const bool a_is_pattern = la <= lb;                // pattern = the shorter string
            const uint32_t m = a_is_pattern ? la : lb, n = a_is_pattern ? lb : la;
            const HapDesc P = a_is_pattern ? job.a : job.b;
            const HapDesc T = a_is_pattern ? job.b : job.a;
            long long dist = n;
            if (m != 0u) {
                // Ukkonen cut-off: an alignment of cost d stays on the diagonals [-d, (n - m) + d], so a band of
                // half-width K gives the exact distance whenever the result is <= K; otherwise widen and repeat.
                // Single-stripe patterns are computed in full at once.
                const uint32_t bands[3] = {max(128u, m / 16u), max(1024u, m / 4u), 0xFFFFFFFFu};
                for (int attempt = m > 2048u ? 0 : 2; attempt < 3; ++attempt) {
                    const unsigned long long K = bands[attempt];
                    long long anchor = 0;                      // D'[last row of the previous stripe, jlo - 1]
                    uint32_t prev_jhi = 0;
                    for (uint32_t row0 = 0; row0 < m; row0 += 2048u) {
                        const uint32_t rows = min(2048u, m - row0);
                        const bool final_stripe = row0 + rows == m;
                        const uint32_t nblk = (rows + 63u) / 64u, last_lane = nblk - 1u;
                        const uint32_t jlo = static_cast<unsigned long long>(row0) > K ? static_cast<uint32_t>(row0 - K) : 0u;
                        const uint32_t jhi = static_cast<uint32_t>(min(static_cast<unsigned long long>(n),
                                                                       static_cast<unsigned long long>(row0) + rows + (n - m) + K));
                        // first column of the NEXT stripe's band: the running anchor stops there
                        const uint32_t next_row0 = row0 + rows;
                        const uint32_t next_jlo = static_cast<unsigned long long>(next_row0) > K ? static_cast<uint32_t>(next_row0 - K) : 0u;
                        // per-class match masks of this lane's 64 rows
#pragma unroll 4
                        for (int c = 0; c < ED_NCLASS; ++c) s_peq[warp][c][lane] = 0ull;
                        for (uint32_t r = 0; r < 64u; ++r) {
                            const uint32_t idx = lane * 64u + r;
                            if (idx < rows) {
                                const uint8_t cls = s_class[hap_char(P, pre + row0 + idx, ref, seq4_a, seq4_b)];
                                if (cls < ED_NCLASS) s_peq[warp][cls][lane] |= 1ull << r;
                            }
                        }
                        __syncwarp();
                        uint64_t pv = ~0ull, mv = 0ull;
                        const uint64_t hibit = (final_stripe && lane == last_lane) ? (1ull << ((rows - 1u) & 63u)) : (1ull << 63);
                        int sum_all = 0, sum_anchor = 0;        // bottom-row deltas over [jlo, jhi) and over [jlo, next_jlo)
                        const uint32_t width = jhi - jlo, steps = width + nblk - 1u;
                        // Text classes and top-row deltas of the band's columns live in a 64-entry ring in shared memory,
                        // refilled 32 columns at a time one block ahead of their use.  Every lane reads the class of ITS
                        // column (t - lane) from the ring, so the match mask of step t+1 is fetched while step t computes
                        // and only the 2-bit horizontal delta travels lane to lane: the loop-carried path is one shuffle
                        // plus the block recurrence.
                        auto fetch = [&](uint32_t rel) -> uint32_t {           // (class, top delta + 1) of band column `rel`
                            const uint32_t j = jlo + rel;
                            const uint32_t cls = (rel < width) ? s_class[hap_char(T, pre + j, ref, seq4_a, seq4_b)] : 255u;
                            const int h = (rel < width && row0 != 0u && j < prev_jhi) ? hbuf[j] : 1;
                            return cls | (static_cast<uint32_t>(h + 1) << 8);
                        };
                        {
                            const uint32_t first = fetch(lane);
                            s_tcls[warp][lane] = static_cast<uint8_t>(first & 0xFFu);
                            s_th[warp][lane] = static_cast<signed char>(static_cast<int>(first >> 8) - 1);
                        }
                        uint32_t nxt = fetch(32u + lane);
                        __syncwarp();
                        auto lookup = [&](uint32_t t) -> uint64_t {            // match mask of this lane's block at step t
                            const long long rel = static_cast<long long>(t) - lane;
                            if (rel < 0 || rel >= static_cast<long long>(width)) return 0ull;
                            const uint32_t cls = s_tcls[warp][static_cast<uint32_t>(rel) & 63u];
                            return cls < ED_NCLASS ? s_peq[warp][cls][lane] : 0ull;
                        };
                        uint64_t eq = lookup(0u);
                        int hout = 0;
                        for (uint32_t t = 0; t < steps; ++t) {
                            int hin = __shfl_up_sync(0xffffffffu, hout, 1);
                            if (lane == 0u) hin = s_th[warp][t & 63u];
                            const long long rel = static_cast<long long>(t) - lane;
                            hout = 0;
                            if (lane < nblk && rel >= 0 && rel < static_cast<long long>(width)) {
                                hout = myers_block(pv, mv, eq, hin, hibit);
                                if (lane == last_lane) {
                                    const uint32_t j = jlo + static_cast<uint32_t>(rel);
                                    sum_all += hout;
                                    if (j < next_jlo) sum_anchor += hout;
                                    if (!final_stripe) hbuf[j] = static_cast<signed char>(hout);
                                }
                            }
                            if (((t + 1u) & 31u) == 0u) {              // columns t+1 .. t+32 enter the ring, t+33 .. t+64 are fetched
                                const uint32_t slot = (t + 1u + lane) & 63u;
                                s_tcls[warp][slot] = static_cast<uint8_t>(nxt & 0xFFu);
                                s_th[warp][slot] = static_cast<signed char>(static_cast<int>(nxt >> 8) - 1);
                                nxt = fetch(t + 33u + lane);
                                __syncwarp();
                            }
                            eq = lookup(t + 1u);
                        }
                        __syncwarp();
                        sum_all = __shfl_sync(0xffffffffu, sum_all, last_lane);
                        sum_anchor = __shfl_sync(0xffffffffu, sum_anchor, last_lane);
                        if (final_stripe) dist = anchor + rows + sum_all;
                        else anchor += static_cast<long long>(rows) + sum_anchor;
                        prev_jhi = jhi;
                    }
                    if (static_cast<unsigned long long>(dist) <= K) break;      // exact
                }
            }
            if (lane == 0) out[job.out_index] = static_cast<double>(dist);
        }
    }
}

// strings given explicitly (test hook svb_edit_distance): bytes are the reference, HapDesc = one left piece
__global__ void make_string_jobs(const uint64_t* __restrict__ a_off, const uint64_t* __restrict__ b_off, uint64_t b_shift,
                                 uint32_t n_pairs, EditJob* __restrict__ jobs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    EditJob j;
    j.a.l_base = a_off[i]; j.a.l_len = static_cast<uint32_t>(a_off[i + 1] - a_off[i]);
    j.a.r_base = 0; j.a.r_len = 0; j.a.m_base = 0; j.a.m_len = 0; j.a.m_kind = HAP_MID_NONE; j.a.m_unit = 1; j.a.seq_sel = 0;
    j.b = j.a;
    j.b.l_base = b_shift + b_off[i]; j.b.l_len = static_cast<uint32_t>(b_off[i + 1] - b_off[i]);
    j.out_index = i;
    j.pad = 0;
    jobs[i] = j;
}

}  // namespace

int build_class_map(const uint8_t* bases, uint64_t n, uint8_t* map256, std::string* why) {
    // classes 0..15: the nt16 letters a 4-bit query base can decode to; further classes: any other byte of
    // the (upper-cased) reference.  Equality of bytes == equality of classes, so distances stay exact.
    memset(map256, 255, 256);
    const char* nt16 = "=ACMGRSVTWYHKDBN";
    int next = 0;
    for (int i = 0; i < 16; ++i) map256[static_cast<uint8_t>(nt16[i])] = static_cast<uint8_t>(next++);
    bool seen[256] = {false};
    for (uint64_t i = 0; i < n; ++i) seen[bases[i]] = true;
    for (int c = 0; c < 256; ++c) {
        if (!seen[c] || map256[c] != 255) continue;
        if (next >= ED_NCLASS) {
            if (why) *why = "reference uses more than 32 distinct symbols";
            return SVB_ERR_FORMAT;
        }
        map256[c] = static_cast<uint8_t>(next++);
    }
    return SVB_OK;
}

int launch_edit_distance(svb_ctx* ctx, const EditJob* d_jobs, uint32_t n_jobs, uint64_t max_text_multi_stripe,
                         const uint8_t* d_ref, const uint8_t* d_seq4_a, const uint8_t* d_seq4_b, const uint8_t* d_class_map,
                         double* d_out) {
    if (!n_jobs) return SVB_OK;
    const unsigned blocks = static_cast<unsigned>(std::min<uint64_t>((n_jobs + ED_WARPS - 1) / ED_WARPS,
                                                                     static_cast<uint64_t>(ctx->sm_count) * 8));
    signed char* hbuf = nullptr;
    const uint64_t stride = (max_text_multi_stripe + 127) & ~127ull;
    if (stride) SVB_CUDA(ctx, cudaMallocAsync(&hbuf, stride * blocks * ED_WARPS, ctx->stream));
    SVB_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 8, 0, sizeof(unsigned long long), ctx->stream));   // two 32-bit job counters
    {
        KernelTimer timer(ctx, SVB_K_EDIT_DISTANCE);
        edit_distance_kernel<<<blocks, ED_WARPS * 32, 0, ctx->stream>>>(d_jobs, n_jobs, reinterpret_cast<unsigned int*>(ctx->d_counters + 8),
                                                                       d_ref, d_seq4_a, d_seq4_b, d_class_map, hbuf, stride, d_out);
        ctx->launches += 1;
    }
    SVB_CUDA(ctx, cudaGetLastError());
    if (hbuf) SVB_CUDA(ctx, cudaFreeAsync(hbuf, ctx->stream));
    return SVB_OK;
}

int run_edit_distance_strings(svb_ctx* ctx, const uint8_t* a, const uint64_t* a_off, const uint8_t* b, const uint64_t* b_off,
                              uint32_t n_pairs, int64_t* out) {
    if (!n_pairs) return SVB_OK;
    const uint64_t na = a_off[n_pairs], nb = b_off[n_pairs];
    std::vector<uint8_t> both(na + nb);
    if (na) memcpy(both.data(), a, na);
    if (nb) memcpy(both.data() + na, b, nb);
    uint8_t map[256];
    {
        // explicit strings may use any byte: give every distinct byte its own class (up to 32)
        memset(map, 255, 256);
        int next = 0;
        bool seen[256] = {false};
        for (uint8_t c : both) seen[c] = true;
        for (int c = 0; c < 256; ++c)
            if (seen[c]) {
                if (next >= ED_NCLASS) return svb_fail(ctx, SVB_ERR_FORMAT, "svb_edit_distance: more than 32 distinct symbols");
                map[c] = static_cast<uint8_t>(next++);
            }
    }
    uint64_t max_multi = 0;
    for (uint32_t i = 0; i < n_pairs; ++i) {
        const uint64_t la = a_off[i + 1] - a_off[i], lb = b_off[i + 1] - b_off[i];
        if (std::min(la, lb) > 2048) max_multi = std::max(max_multi, std::max(la, lb));
    }
    uint8_t *d_bytes = nullptr, *d_map = nullptr;
    uint64_t *d_aoff = nullptr, *d_boff = nullptr;
    EditJob* d_jobs = nullptr;
    double* d_out = nullptr;
    SVB_CUDA(ctx, cudaMallocAsync(&d_bytes, std::max<uint64_t>(both.size(), 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_map, 256, ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_aoff, sizeof(uint64_t) * (n_pairs + 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_boff, sizeof(uint64_t) * (n_pairs + 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_jobs, sizeof(EditJob) * n_pairs, ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_out, sizeof(double) * n_pairs, ctx->stream));
    if (!both.empty()) SVB_CUDA(ctx, cudaMemcpyAsync(d_bytes, both.data(), both.size(), cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_map, map, 256, cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_aoff, a_off, sizeof(uint64_t) * (n_pairs + 1), cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_boff, b_off, sizeof(uint64_t) * (n_pairs + 1), cudaMemcpyHostToDevice, ctx->stream));
    make_string_jobs<<<(n_pairs + 127) / 128, 128, 0, ctx->stream>>>(d_aoff, d_boff, na, n_pairs, d_jobs);
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    int rc = launch_edit_distance(ctx, d_jobs, n_pairs, max_multi, d_bytes, nullptr, nullptr, d_map, d_out);
    if (rc != SVB_OK) return rc;
    std::vector<double> h(n_pairs);
    SVB_CUDA(ctx, cudaMemcpyAsync(h.data(), d_out, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < n_pairs; ++i) out[i] = static_cast<int64_t>(h[i]);
    cudaFreeAsync(d_bytes, ctx->stream);
    cudaFreeAsync(d_map, ctx->stream);
    cudaFreeAsync(d_aoff, ctx->stream);
    cudaFreeAsync(d_boff, ctx->stream);
    cudaFreeAsync(d_jobs, ctx->stream);
    cudaFreeAsync(d_out, ctx->stream);
    return SVB_OK;
}
