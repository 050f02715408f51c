// K8a wfa -- THRESHOLDED edit distance of haplotype-string pairs: "is edlib.align(h1, h2)['editDistance'] <= t, and
// if so what is it?", one CTA per pair.
//
// pair_haplotypes (reference SVIM_COMBINE.py:120-140) feeds compute_distance's values (:35-102) into complete linkage cut
// at t = max_edit_distance.  Complete linkage is ordinal (nearest-neighbour chain, max, stable sort, `<= t`): a distance
// above t is only ever COMPARED, so for a partition of two candidates -- 99 % of a diploid sample -- "more than t" is all
// the clustering needs, and in larger partitions the exact value of a far pair matters only when it could tie or swap
// with another far pair of the same partition (pair.cu: resolve_kernel sends exactly those to the exact kernel,
// edit_distance.cu).  This kernel therefore computes furthest-reaching wavefronts (wfa_core.cuh) up to wave t:
//   * common prefix / suffix stripped by two warps from both ends (most shared variants end here),
//   * |la - lb| > t: far, nothing to compute,
//   * the trimmed strings become symbol-class bytes in shared memory (sentinels behind both: no bounds checks),
//   * wave s: one thread per diagonal (at most t + 1 after pruning), three shared-memory reads, a word-wise match
//     extension of up to 32 symbols; a diagonal that is still matching after that (the optimal path of two similar
//     haplotypes) is extended by the whole warp, 128 symbols per round; one barrier per wave doubles as the "done" vote.
// Work per pair is O(t^2 + length) instead of O(length^2 / 64): the 10,000-base insertion pairs that were the critical
// path of the exact kernel (0.5 ms for one warp) take tens of microseconds here.
// Two launches of the same kernel: `stage 0` takes every job, trims it and runs the ones whose strings fit a small
// shared-memory window (many CTAs per SM); the few that do not are queued, with their trim, for `stage 1` (one CTA per SM,
// the whole 227 KB).  What fits neither is reported as unknown and goes to the exact kernel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "edit_strings.cuh"
#include "wfa_core.cuh"

namespace {
using namespace edstr;

constexpr int WFA_THREADS = 256;         // 8 warps: 4 per direction of the bidirectional wavefront
constexpr int WFA_SIDE = WFA_THREADS / 2;
#ifndef WFA_PRUNE_SLACK
#define WFA_PRUNE_SLACK 1          // (0 rebuilds the one-sided bound that loses a distance of exactly t: tests/test_pair_gpu.py must catch it)
#endif
constexpr uint32_t WFA_SLACK = 160;               // readable bytes behind the second string (warp-wide extension reads ahead)

struct WfaControl {
    uint32_t job;
    uint32_t trim[2];
    uint32_t hit;              // bit 0: the waves met at total 2 r - 1, bit 1: at 2 r
};

// One diagonal of a wave, one thread.  `prev` is this side's previous wave, `other` the OTHER side's previous wave (both
// complete since the last barrier), `cur` this side's new wave.  Everything is branch-free sentinel arithmetic: a diagonal
// that was not reached holds WFA_NEG, and WFA_NEG + anything stays far below every real row, so
//     v = min(max(prev[k] + 1, prev[k - 1], prev[k + 1] + 1), la, lb - k)          (furthest row with distance <= r on k)
// needs no validity tests (the clamp keeps the point inside the table: a cell next to a reached cell of the last row or
// column is within one more edit).  Then eight symbols of match extension with every load issued at once; a diagonal still
// matching after that is taken on by the whole warp, 128 symbols per step.
// The forward side also tests for an overlap with the backward wave g = other[kd - k] of the previous round:
//     prev[k] + g >= la : the waves of round r - 1 had met                      -> total 2 r - 2   (bit 0)
//     v + g >= la       : this wave meets the backward wave of round r - 1      -> total 2 r - 1   (bit 1)
// (wfa_core.cuh: F[sf][k] + G[sb][kd - k] >= la  <=>  distance <= sf + sb; every diagonal on which the waves of round r - 1
// can meet within t edits lies in the range of round r, see the loop below.)
template <bool BACKWARD>
__device__ __forceinline__ uint32_t wave_diagonal(const int* prev, const int* other, int* cur, const uint32_t* Aw, const uint32_t* Bw, int k, bool active,
                                                  int klo, int plo, int phi, int mid, int ila, int ilb, int kd, uint32_t lane) {
    const int kk = active ? k : klo;                              // a safe diagonal for the threads beyond the range
    const int fm1 = prev[mid + kk - 1], f0 = prev[mid + kk], fp1 = prev[mid + kk + 1];
    const int kb = kd - kk;
    const bool partner = !BACKWARD && active && kb >= plo && kb <= phi;
    const int g = partner ? other[mid + kb] : WFA_NEG;
    const int x = wfa_next_clamped(fm1, f0, fp1, kk, ila, ilb);
    const bool valid = active && x > WFA_NEG / 2;
    const int xs = valid ? x : 0;                                 // safe offsets for the loads below
    const int jb = xs + (valid ? kk : 0);
    uint32_t x0, x1, n0, n1;
    if (!BACKWARD) {
        const uint32_t oa = WFA_FRONT + static_cast<uint32_t>(xs), ob = WFA_FRONT + static_cast<uint32_t>(jb);
        x0 = wfa_load4(Aw, oa) ^ wfa_load4(Bw, ob);
        x1 = wfa_load4(Aw, oa + 4u) ^ wfa_load4(Bw, ob + 4u);
        n0 = (static_cast<uint32_t>(__ffs(static_cast<int>(x0 | 0x80000000u))) - 1u) >> 3;
        n1 = (static_cast<uint32_t>(__ffs(static_cast<int>(x1 | 0x80000000u))) - 1u) >> 3;
    } else {
        const uint32_t oa = WFA_FRONT + static_cast<uint32_t>(ila - 1 - xs), ob = WFA_FRONT + static_cast<uint32_t>(ilb - 1 - jb);
        x0 = wfa_load4(Aw, oa - 3u) ^ wfa_load4(Bw, ob - 3u);
        x1 = wfa_load4(Aw, oa - 7u) ^ wfa_load4(Bw, ob - 7u);
        n0 = static_cast<uint32_t>(__clz(static_cast<int>(x0 | 1u))) >> 3;
        n1 = static_cast<uint32_t>(__clz(static_cast<int>(x1 | 1u))) >> 3;
    }
    const uint32_t run8 = x0 ? n0 : 4u + (x1 ? n1 : 4u);
    int v = valid ? x + static_cast<int>(run8) : WFA_NEG;
    // a diagonal that is still matching: the whole warp goes on, 128 symbols per step
    uint32_t pending = __ballot_sync(FULL, valid && (x0 | x1) == 0u);
    while (pending) {
        const int src = __ffs(static_cast<int>(pending)) - 1;
        pending &= pending - 1u;
        const int v0 = __shfl_sync(FULL, v, src), k0 = __shfl_sync(FULL, kk, src);
        uint32_t run = 0;
        while (true) {
            uint32_t y;
            if (!BACKWARD) {
                const uint32_t ia = WFA_FRONT + static_cast<uint32_t>(v0) + run + 4u * lane;
                y = wfa_load4(Aw, ia) ^ wfa_load4(Bw, ia + static_cast<uint32_t>(k0));
            } else {
                const int pa = static_cast<int>(WFA_FRONT) + ila - 1 - v0 - 3 - static_cast<int>(run) - 4 * static_cast<int>(lane);
                y = wfa_load4s(Aw, pa) ^ wfa_load4s(Bw, pa + kd - k0);     // pb = FRONT + lb - 1 - (v0 + k0) = pa + kd - k0
            }
            const uint32_t bal = __ballot_sync(FULL, y != 0u);
            if (bal) {
                const int f = __ffs(static_cast<int>(bal)) - 1;
                const uint32_t yf = __shfl_sync(FULL, y, f);
                run += 4u * static_cast<uint32_t>(f) + (BACKWARD ? wfa_last_diff(yf) : wfa_first_diff(yf));
                break;
            }
            run += 128u;
        }
        if (static_cast<int>(lane) == src) v = v0 + static_cast<int>(run);
    }
    if (active) cur[mid + k] = v;
    return (f0 + g >= ila ? 1u : 0u) | (v + g >= ila ? 2u : 0u);   // (g == WFA_NEG: neither)
}

__global__ void __launch_bounds__(WFA_THREADS) wfa_kernel(const WfaArgs a) {
    extern __shared__ uint32_t smem_w[];
    const int t = static_cast<int>(a.t), W = 2 * t + 7, mid = t + 3;
    int* const Fw = reinterpret_cast<int*>(smem_w);          // [side][wave parity][W]: forward 0 / 1, backward 0 / 1
    uint8_t* cls2 = reinterpret_cast<uint8_t*>(Fw + 4 * W);
    WfaControl* ctl = reinterpret_cast<WfaControl*>(cls2 + 512);
    uint32_t* Aw = reinterpret_cast<uint32_t*>(ctl + 1);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t i = tid; i < 256u; i += WFA_THREADS) {
        cls2[i] = a.class_map[i];
        cls2[256u + i] = a.class_map[hap_complement(static_cast<uint8_t>(i))];
    }
    const uint32_t n_jobs = a.stage == 0 ? static_cast<uint32_t>(*a.n_jobs_dev) : min(a.counters[1], a.big_cap);
    // stage 0 goes over the job list twice: the long pairs first (a 200-wave pair claimed last would be the tail of the launch)
    int pass = a.stage == 0 ? 0 : 1;
    while (true) {
        __syncthreads();
        if (tid == 0) ctl->job = atomicAdd(a.counters + (a.stage != 0 ? 2 : (pass == 0 ? 0 : 3)), 1u);
        __syncthreads();
        const uint32_t claim = ctl->job;
        if (claim >= n_jobs) {
            if (a.stage == 0 && pass == 0) { pass = 1; continue; }
            break;
        }
        uint32_t job_id = claim, pre = 0, suf = 0;
        if (a.stage != 0) {
            const uint4 e = a.big[claim];
            job_id = e.x; pre = e.y; suf = e.z;
        }
        const EditJob job = a.jobs[job_id];
        const long long clock_begin = a.profile ? clock64() : 0ll;
        const uint32_t la0 = hap_length(job.a), lb0 = hap_length(job.b);
        if (a.stage == 0 && (min(la0, lb0) >= 2048u) != (pass == 0)) continue;
        if (a.stage == 0) {
            // common prefix (warp 0) and suffix (warp 1) at the same time, each at most half of the shorter string; the side
            // that ran into its half-way mark goes on if the other one stopped early
            const uint32_t lim = min(la0, lb0), half = (lim + 1u) / 2u;
            if (warp < 2u) {
                const uint32_t r = warp == 0 ? common_run<false>(job.a, job.b, la0, lb0, half, a.ref, a.seq4_a, a.seq4_b, cls2, lane)
                                             : common_run<true>(job.a, job.b, la0, lb0, lim - half, a.ref, a.seq4_a, a.seq4_b, cls2, lane);
                if (lane == 0) ctl->trim[warp] = r;
            }
            __syncthreads();
            pre = ctl->trim[0];
            suf = ctl->trim[1];
            __syncthreads();
            if (pre == half && suf < lim - half) {
                if (warp == 0) {
                    const uint32_t r = common_run<false>(job.a, job.b, la0, lb0, lim - suf, a.ref, a.seq4_a, a.seq4_b, cls2, lane, half);
                    if (lane == 0) ctl->trim[0] = r;
                }
                __syncthreads();
                pre = ctl->trim[0];
            } else if (suf == lim - half && pre < half) {
                if (warp == 0) {
                    const uint32_t r = common_run<true>(job.a, job.b, la0, lb0, lim - pre, a.ref, a.seq4_a, a.seq4_b, cls2, lane, lim - half);
                    if (lane == 0) ctl->trim[1] = r;
                }
                __syncthreads();
                suf = ctl->trim[1];
            }
        }
        const uint32_t la = la0 - pre - suf, lb = lb0 - pre - suf;
        const uint32_t longer = max(la, lb), gap = la > lb ? la - lb : lb - la;
        double lo, hi;
        bool settled = true;
        if (min(la, lb) == 0u) {                       // one string is a prefix + suffix of the other
            lo = hi = static_cast<double>(longer);
        } else if (gap > a.t) {                        // the distance is at least the length difference
            lo = static_cast<double>(gap);
            hi = static_cast<double>(longer);
        } else {
            settled = false;
            lo = hi = 0.0;
        }
        if (settled) {
            if (tid == 0) {
                a.dist[job.out_index] = lo; a.dist_hi[job.out_index] = hi;
                if (a.profile) a.profile[job_id] = make_uint4(la, lb, 0xFFFFu, static_cast<uint32_t>(clock64() - clock_begin));
            }
            continue;
        }
        const uint32_t a_bytes = (WFA_FRONT + la + WFA_PAD + 3u) & ~3u, b_bytes = (WFA_FRONT + lb + WFA_PAD + 3u) & ~3u;
        if (a_bytes + b_bytes + WFA_SLACK > a.cap_chars) {
            if (tid == 0) {
                bool queued = false;
                if (a.stage == 0) {
                    const uint32_t slot = atomicAdd(a.counters + 1, 1u);
                    if (slot < a.big_cap) { a.big[slot] = make_uint4(job_id, pre, suf, 0u); queued = true; }
                }
                if (!queued) { a.dist[job.out_index] = -1.0; a.dist_hi[job.out_index] = static_cast<double>(longer); }      // unknown: exact kernel
            }
            continue;
        }
        // ---- the trimmed strings as class bytes in shared memory
        uint32_t* Bw = Aw + a_bytes / 4u;
        uint8_t* As = reinterpret_cast<uint8_t*>(Aw);
        uint8_t* Bs = reinterpret_cast<uint8_t*>(Bw);
        for (uint32_t i0 = 0; i0 < longer; i0 += 4u * WFA_THREADS) {
            uint32_t ba[4], ma[4], bb[4], mb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * WFA_THREADS + tid;
                ma[u] = mb[u] = TOK_NONE;
                ba[u] = bb[u] = 0u;
                if (i < la) ba[u] = hap_fetch(job.a, pre + i, a.ref, a.seq4_a, a.seq4_b, ma[u]);
                if (i < lb) bb[u] = hap_fetch(job.b, pre + i, a.ref, a.seq4_a, a.seq4_b, mb[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * WFA_THREADS + tid;
                if (i < la) {
                    const uint32_t c = tok_class(cls2, ba[u], ma[u]);
                    As[WFA_FRONT + i] = c < ED_NOCLASS ? static_cast<uint8_t>(c) : WFA_NOCLASS_A;
                }
                if (i < lb) {
                    const uint32_t c = tok_class(cls2, bb[u], mb[u]);
                    Bs[WFA_FRONT + i] = c < ED_NOCLASS ? static_cast<uint8_t>(c) : WFA_NOCLASS_B;
                }
            }
        }
        if (tid < WFA_FRONT) { As[tid] = WFA_FRONT_A; Bs[tid] = WFA_FRONT_B; }
        for (uint32_t i = WFA_FRONT + la + tid; i < a_bytes; i += WFA_THREADS) As[i] = WFA_END_A;
        for (uint32_t i = WFA_FRONT + lb + tid; i < b_bytes + WFA_SLACK; i += WFA_THREADS) Bs[i] = WFA_END_B;
        for (int x = static_cast<int>(tid); x < 4 * W; x += WFA_THREADS) Fw[x] = (x == W + mid || x == 3 * W + mid) ? -1 : WFA_NEG;      // "wave -1": row -1 on diagonal 0, so that wave 0 starts at row 0
        if (tid == 0) ctl->hit = 0u;
        __syncthreads();

        // ---- waves, from both ends at once (wfa_core.cuh: bidirectional wavefronts).  Warps 0-3 advance the forward wave, warps
        // 4-7 the backward one (the same recurrence on the reversed strings).  A pair that differs by more than t is settled
        // after t / 2 rounds instead of t waves.  A round of a long far pair is a chain of dependent instructions of ONE warp
        // (measured: a 9,748-symbol pair alone on the GPU took 1,800 cycles per round with 220-330 instructions per warp and
        // round), so the round is kept short: sentinel arithmetic instead of validity branches, the overlap test folded into
        // the wave (one more shared-memory read), one barrier.
        const int ila = static_cast<int>(la), ilb = static_cast<int>(lb), kd = ilb - ila;
        const bool backward = tid >= static_cast<uint32_t>(WFA_SIDE);
        const int st = static_cast<int>(tid & static_cast<uint32_t>(WFA_SIDE - 1));       // thread within its side
        int* const mine = Fw + (backward ? 2 * W : 0);          // this side's two arrays (wave r lives in array r & 1)
        const int* const theirs = Fw + (backward ? 0 : 2 * W);
        int result = -1, waves = 0;
        // ONE barrier per round.  Round r: both sides compute wave r from wave r - 1; the forward threads test, on the way, the
        // totals 2 r - 2 (waves r - 1 of both sides) and 2 r - 1 (their new wave against the backward wave r - 1), so that the
        // totals 0, 1, 2, ... are tested in order, two per round, and the answer is read after the round's barrier.
        // Coverage: the waves of round r - 1 can meet within t edits only on a diagonal with |kd - k| <= sb = r - 1, and round r
        // visits the diagonals with |kd - k| <= (t + 1) - r: the range is pruned as if t + 1 edits were allowed, because
        // r - 1 <= t + 1 - r is 2 r - 2 <= t, which holds in every round that runs (with the one-sided bound t - r the last round
        // of an even t lost the outermost diagonal, and with it a distance of exactly t: found by the host statement of these
        // rounds, wfa_rounds_serial, against a plain DP).
        int plo = 0, phi = -1;                                  // the range of the previous round's waves (forward and backward alike)
        for (int r = 0;; ++r) {
            waves = r + 1;
            const int* prev = mine + ((r & 1) ^ 1) * W;
            const int* other = theirs + ((r & 1) ^ 1) * W;
            int* cur = mine + (r & 1) * W;
            int klo, khi;
            wfa_range(r, t + WFA_PRUNE_SLACK, kd, ila, ilb, klo, khi);        // pruned as if one more edit were allowed: see above
            uint32_t hit = 0;
            for (int k = klo + st; k <= ((khi - klo) | (WFA_SIDE - 1)) + klo; k += WFA_SIDE) {      // whole warps go round together
                if (!backward) hit |= wave_diagonal<false>(prev, other, cur, Aw, Bw, k, k <= khi, klo, plo, phi, mid, ila, ilb, kd, lane);
                else wave_diagonal<true>(prev, other, cur, Aw, Bw, k, k <= khi, klo, plo, phi, mid, ila, ilb, kd, lane);
            }
            if (st == 0) {             // the next wave reads one diagonal beyond this range on either side
                cur[mid + klo - 1] = cur[mid + klo - 2] = WFA_NEG;
                cur[mid + khi + 1] = cur[mid + khi + 2] = WFA_NEG;
            }
            if (hit) atomicOr(&ctl->hit, hit);
            __syncthreads();
            // (every warp reads the word right behind the barrier; the next write to it is a whole wave away, and a hit ends the
            // pair, so the word is zero whenever a round starts)
            const uint32_t h = ctl->hit;
            if (h & 1u) { result = 2 * r - 2; break; }
            if (h & 2u) { result = 2 * r - 1 <= t ? 2 * r - 1 : -1; break; }
            if (2 * r - 1 >= t) break;                          // every total up to t has been tested
            plo = klo;
            phi = khi;
        }
        if (tid == 0) {
            if (result >= 0) {
                a.dist[job.out_index] = a.dist_hi[job.out_index] = static_cast<double>(result);
            } else {                                   // more than t edits: only bounds are known
                a.dist[job.out_index] = static_cast<double>(a.t + 1u);
                a.dist_hi[job.out_index] = static_cast<double>(longer);
            }
            if (a.profile)
                a.profile[job_id] = make_uint4(la, lb, static_cast<uint32_t>(waves) | (static_cast<uint32_t>(a.stage) << 16), static_cast<uint32_t>(clock64() - clock_begin));
        }
    }
}

}  // namespace

// counters: [0] next job of stage 0, [1] size of the queue for stage 1, [2] next job of stage 1 (zeroed by the caller)
int launch_wfa(svb_ctx* ctx, WfaArgs a) {
    if (a.t > WFA_MAX_T) return svb_fail(ctx, SVB_ERR_ARG, "launch_wfa: threshold too large");
    const size_t fixed = sizeof(int) * 4u * (2u * a.t + 7u) + 512u + sizeof(WfaControl);
    const size_t small = fixed + 24u * 1024u;          // pairs of up to about 12,000 trimmed symbols each: 7-8 CTAs per SM
    const size_t big = 227u * 1024u - 1024u;           // the rest, one CTA per SM
    if (!ctx->wfa_attr_set) {                          // per context = per device (function attributes are per device)
        SVB_CUDA(ctx, cudaFuncSetAttribute(wfa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(big)));
        ctx->wfa_attr_set = true;
    }
    // profiling aid: SVB_WFA_PROFILE=<file> dumps one uint4 per job of the launch (the count must be known: only tools use it)
    const char* profile_path = getenv("SVB_WFA_PROFILE");
    a.profile = nullptr;
    if (profile_path && a.big_cap) {
        SVB_CUDA(ctx, cudaMallocAsync(&a.profile, sizeof(uint4) * a.big_cap, ctx->stream));
        SVB_CUDA(ctx, cudaMemsetAsync(a.profile, 0, sizeof(uint4) * a.big_cap, ctx->stream));
    }
    {
    KernelTimer timer(ctx, SVB_K_EDIT_DISTANCE);
    a.stage = 0;
    a.cap_chars = static_cast<uint32_t>(small - fixed);
    wfa_kernel<<<static_cast<unsigned>(ctx->sm_count) * 4u, WFA_THREADS, small, ctx->stream>>>(a);
    a.stage = 1;
    a.cap_chars = static_cast<uint32_t>(big - fixed);
    wfa_kernel<<<static_cast<unsigned>(ctx->sm_count), WFA_THREADS, big, ctx->stream>>>(a);
    ctx->launches += 2;
    }
    SVB_CUDA(ctx, cudaGetLastError());
    if (a.profile) {
        std::vector<uint4> h(a.big_cap);
        SVB_CUDA(ctx, cudaMemcpyAsync(h.data(), a.profile, sizeof(uint4) * a.big_cap, cudaMemcpyDeviceToHost, ctx->stream));
        SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (FILE* f = fopen(profile_path, "wb")) {
            fwrite(h.data(), sizeof(uint4), a.big_cap, f);
            fclose(f);
        }
        cudaFreeAsync(a.profile, ctx->stream);
    }
    return SVB_OK;
}
