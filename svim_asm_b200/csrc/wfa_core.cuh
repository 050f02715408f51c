// K8 core of the THRESHOLDED edit distance: furthest-reaching wavefronts (Myers 1986 O(ND) / Ukkonen 1985).
//
// pair_haplotypes clusters a partition by complete linkage cut at max_edit_distance (reference SVIM_COMBINE.py:120-140,
// fcluster(Z, t, 'distance')).  Complete linkage only compares distances, so for most pairs the only question is
// "d <= t, and if so which value?": wave s holds, for every diagonal k = j - i, the furthest row i reachable with s
// edits; the distance is the first s whose wave reaches (la, lb), and after wave t the answer is "more than t" without
// ever touching the rest of the table.  Cost O(t^2 + la) instead of O(la * lb / 64): a 10,000-base pair of haplotypes
// that differ by 500 edits stops after 200 waves of at most 201 diagonals.
//
// Building blocks shared by the device kernel (wfa.cu: one CTA per pair, diagonals across the threads) and the host
// build (tests/hostcheck: serial driver checked against a plain DP).  Strings are symbol-class bytes followed by
// WFA_PAD sentinel bytes (different for the two strings), so the match extension needs no bounds checks.
#pragma once
#include <stdint.h>

#include "linkage.cuh"   // SVB_HD

constexpr int WFA_NEG = -(1 << 29);              // "diagonal not reached"
constexpr uint32_t WFA_PAD = 12;                 // sentinel bytes behind each string (word-wise extension reads ahead)
constexpr uint8_t WFA_END_A = 0xFD, WFA_END_B = 0xFC, WFA_NOCLASS_A = 0xFB, WFA_NOCLASS_B = 0xFA;
constexpr uint32_t WFA_MAX_T = 1024;             // thresholds above this take the exact kernel (t^2 work stops paying)

SVB_HD int wfa_imax(int a, int b) { return a > b ? a : b; }
SVB_HD int wfa_imin(int a, int b) { return a < b ? a : b; }

// Diagonals wave s has to compute: reachable with s edits (|k| <= s), still able to reach the end diagonal kd = lb - la with
// the t - s edits that remain, inside the table.  A diagonal outside this range is never read by a later wave's range.
SVB_HD void wfa_range(int s, int t, int kd, int la, int lb, int& klo, int& khi) {
    klo = wfa_imax(wfa_imax(-s, kd - (t - s)), -la);
    khi = wfa_imin(wfa_imin(s, kd + (t - s)), lb);
}

// Furthest row on diagonal k after one more edit, from the previous wave's rows on k - 1, k, k + 1:
//   substitution (k): (i, j) -> (i + 1, j + 1);  insertion (from k - 1): (i, j) -> (i, j + 1);  deletion (from k + 1): (i, j) -> (i + 1, j)
// A move that leaves the table is not made.
SVB_HD int wfa_next(int fm1, int f0, int fp1, int k, int la, int lb) {
    int v = WFA_NEG;
    if (f0 > WFA_NEG / 2 && f0 + 1 <= la && f0 + 1 + k <= lb) v = f0 + 1;
    if (fm1 > WFA_NEG / 2 && fm1 + k <= lb) v = wfa_imax(v, fm1);
    if (fp1 > WFA_NEG / 2 && fp1 + 1 <= la) v = wfa_imax(v, fp1 + 1);
    return v;
}

// four bytes starting at byte offset `off` of a word-aligned buffer (little endian)
SVB_HD uint32_t wfa_load4(const uint32_t* w, uint32_t off) {
    const uint32_t idx = off >> 2, sh = (off & 3u) * 8u;
    const uint32_t lo = w[idx], hi = w[idx + 1u];
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}

SVB_HD uint32_t wfa_first_diff(uint32_t x) {     // index of the lowest non-zero byte of x (x != 0)
#ifdef __CUDA_ARCH__
    return (static_cast<uint32_t>(__ffs(static_cast<int>(x))) - 1u) >> 3;
#else
    uint32_t n = 0;
    while (!(x & 0xFFu)) { x >>= 8; ++n; }
    return n;
#endif
}

// Slide down the diagonal while the strings agree, at most `limit` symbols (a multiple of 4).  Returns the number of
// matching symbols; `*more` is set when the limit was reached with every symbol matching (the caller goes on, in the
// kernel with the whole warp).  The sentinels behind the strings end every run.
SVB_HD uint32_t wfa_extend(const uint32_t* A, const uint32_t* B, uint32_t i, uint32_t j, uint32_t limit, bool* more) {
    uint32_t run = 0;
    *more = false;
    while (run < limit) {
        const uint32_t x = wfa_load4(A, i + run) ^ wfa_load4(B, j + run);
        if (x) return run + wfa_first_diff(x);
        run += 4u;
    }
    *more = true;
    return run;
}

// The same for exactly 8 symbols with every load issued at once (one shared-memory latency instead of two dependent
// rounds): what a thread of the kernel does before the whole warp takes over a diagonal that keeps matching.
SVB_HD uint32_t wfa_extend8(const uint32_t* A, const uint32_t* B, uint32_t i, uint32_t j, bool* more) {
    const uint32_t ia = i >> 2, sa = (i & 3u) * 8u, ib = j >> 2, sb = (j & 3u) * 8u;
    const uint32_t a0 = A[ia], a1 = A[ia + 1u], a2 = A[ia + 2u], b0 = B[ib], b1 = B[ib + 1u], b2 = B[ib + 2u];
#ifdef __CUDA_ARCH__
    const uint32_t x0 = __funnelshift_r(a0, a1, sa) ^ __funnelshift_r(b0, b1, sb);
    const uint32_t x1 = __funnelshift_r(a1, a2, sa) ^ __funnelshift_r(b1, b2, sb);
#else
    const uint32_t x0 = (sa ? (a0 >> sa) | (a1 << (32u - sa)) : a0) ^ (sb ? (b0 >> sb) | (b1 << (32u - sb)) : b0);
    const uint32_t x1 = (sa ? (a1 >> sa) | (a2 << (32u - sa)) : a1) ^ (sb ? (b1 >> sb) | (b2 << (32u - sb)) : b1);
#endif
    *more = false;
    if (x0) return wfa_first_diff(x0);
    if (x1) return 4u + wfa_first_diff(x1);
    *more = true;
    return 8u;
}

// Serial driver (host check, and the specification of what the kernel computes): the distance if it is <= t, else -1.
// A, B: word-aligned class bytes with sentinels; F0, F1: 2 t + 7 ints each.
SVB_HD long long wfa_distance_serial(const uint32_t* A, int la, const uint32_t* B, int lb, int t, int* F0, int* F1) {
    const int kd = lb - la, W = 2 * t + 7, mid = t + 3;
    if ((kd < 0 ? -kd : kd) > t) return -1;
    for (int x = 0; x < W; ++x) F0[x] = F1[x] = WFA_NEG;
    int* prev = F0;
    int* cur = F1;
    for (int s = 0; s <= t; ++s) {
        int klo, khi;
        wfa_range(s, t, kd, la, lb, klo, khi);
        for (int k = klo; k <= khi; ++k) {
            int v = s == 0 ? 0 : wfa_next(prev[mid + k - 1], prev[mid + k], prev[mid + k + 1], k, la, lb);
            if (v > WFA_NEG / 2) {
                bool more = true;
                v += static_cast<int>(wfa_extend8(A, B, static_cast<uint32_t>(v), static_cast<uint32_t>(v + k), &more));
                while (more) v += static_cast<int>(wfa_extend(A, B, static_cast<uint32_t>(v), static_cast<uint32_t>(v + k), 32u, &more));
            }
            cur[mid + k] = v;
            if (k == kd && v >= la) return s;
        }
        // the next wave reads one diagonal beyond this range on either side
        cur[mid + klo - 1] = cur[mid + klo - 2] = WFA_NEG;
        cur[mid + khi + 1] = cur[mid + khi + 2] = WFA_NEG;
        int* tmp = prev; prev = cur; cur = tmp;
    }
    return -1;
}

// ---- bidirectional wavefronts ------------------------------------------------------------------------------------------
// A forward wavefront from (0, 0) and a backward one from (la, lb) (the same recurrence on the reversed strings) meet in the
// middle: with F[s][k] = the furthest row on diagonal k whose prefix distance is <= s and G[s][k'] the same for the reversed
// strings (k' = kd - k in forward terms), d <= sf + sb exactly when some diagonal has F[sf][k] + G[sb][kd - k] >= la
// (prefix distances do not decrease along a diagonal, so the backward point is reachable forward with <= sf edits once the
// forward reach is at or beyond it).  Testing the totals 0, 1, 2, ... in order gives d, and "more than t" after t: the same
// answer as the one-sided run in half as many DEPENDENT waves, because the two sides advance at the same time.
// Strings carry WFA_FRONT sentinel bytes in front as well (the backward extension runs off the beginning).
constexpr uint32_t WFA_FRONT = 12;
constexpr uint8_t WFA_FRONT_A = 0xF9, WFA_FRONT_B = 0xF8;

SVB_HD uint32_t wfa_last_diff(uint32_t x) {      // number of equal bytes from the TOP of the word down (x != 0)
#ifdef __CUDA_ARCH__
    return static_cast<uint32_t>(__clz(static_cast<int>(x))) >> 3;
#else
    uint32_t n = 0;
    while (!(x & 0xFF000000u)) { x <<= 8; ++n; }
    return n;
#endif
}

// Backward match extension: the bytes at offsets pa, pa - 1, ... of A against pb, pb - 1, ... of B (byte offsets into the
// buffers, sentinels included), at most `limit` (a multiple of 4).
SVB_HD uint32_t wfa_rextend(const uint32_t* A, const uint32_t* B, uint32_t pa, uint32_t pb, uint32_t limit, bool* more) {
    uint32_t run = 0;
    *more = false;
    while (run < limit) {
        const uint32_t x = wfa_load4(A, pa - 3u - run) ^ wfa_load4(B, pb - 3u - run);
        if (x) return run + wfa_last_diff(x);
        run += 4u;
    }
    *more = true;
    return run;
}

// four bytes at a byte offset that may be NEGATIVE (the warp-wide backward extension looks up to 128 bytes before the front
// sentinels: the caller guarantees that much readable memory in front of the buffer)
SVB_HD uint32_t wfa_load4s(const uint32_t* w, int off) {
    const int idx = off >> 2;
    const uint32_t sh = (static_cast<uint32_t>(off) & 3u) * 8u;
    const uint32_t lo = w[idx], hi = w[idx + 1];
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}

// the backward twin of wfa_extend8: 8 symbols ending at pa / pb, every load issued at once
SVB_HD uint32_t wfa_rextend8(const uint32_t* A, const uint32_t* B, uint32_t pa, uint32_t pb, bool* more) {
    const uint32_t x0 = wfa_load4(A, pa - 3u) ^ wfa_load4(B, pb - 3u);
    const uint32_t x1 = wfa_load4(A, pa - 7u) ^ wfa_load4(B, pb - 7u);
    *more = false;
    if (x0) return wfa_last_diff(x0);
    if (x1) return 4u + wfa_last_diff(x1);
    *more = true;
    return 8u;
}

// monotone closure of wfa_next: a diagonal keeps what it had reached with fewer edits (F[s][k] = max row with distance <= s)
SVB_HD int wfa_next_closed(int fm1, int f0, int fp1, int k, int la, int lb) {
    return wfa_imax(wfa_next(fm1, f0, fp1, k, la, lb), f0);
}

// Serial driver of the bidirectional run: the distance if it is <= t, else -1.  A, B: word-aligned buffers, WFA_FRONT front
// sentinels + symbols + WFA_PAD back sentinels; Ff / Fb: two arrays of 2 t + 7 ints each.
SVB_HD long long wfa_bidir_serial(const uint32_t* A, int la, const uint32_t* B, int lb, int t, int* Ff0, int* Ff1, int* Fb0, int* Fb1) {
    const int kd = lb - la, W = 2 * t + 7, mid = t + 3;
    if ((kd < 0 ? -kd : kd) > t) return -1;
    for (int x = 0; x < W; ++x) Ff0[x] = Ff1[x] = Fb0[x] = Fb1[x] = WFA_NEG;
    int* fprev = Ff0; int* fcur = Ff1; int* bprev = Fb0; int* bcur = Fb1;
    auto wave = [&](bool backward, int s, const int* prev, int* cur) {
        int klo, khi;
        wfa_range(s, t, kd, la, lb, klo, khi);
        for (int k = klo; k <= khi; ++k) {
            int v = s == 0 ? 0 : wfa_next_closed(prev[mid + k - 1], prev[mid + k], prev[mid + k + 1], k, la, lb);
            if (v > WFA_NEG / 2) {
                bool more = true;
                while (more) {
                    if (!backward) v += static_cast<int>(wfa_extend(A, B, WFA_FRONT + static_cast<uint32_t>(v), WFA_FRONT + static_cast<uint32_t>(v + k), 32u, &more));
                    else v += static_cast<int>(wfa_rextend8(A, B, WFA_FRONT + static_cast<uint32_t>(la - 1 - v), WFA_FRONT + static_cast<uint32_t>(lb - 1 - (v + k)), &more));
                }
            }
            cur[mid + k] = v;
        }
        cur[mid + klo - 1] = cur[mid + klo - 2] = WFA_NEG;
        cur[mid + khi + 1] = cur[mid + khi + 2] = WFA_NEG;
    };
    auto overlap = [&](int sf, const int* F, int sb, const int* G) {
        int klo, khi, glo, ghi;
        wfa_range(sf, t, kd, la, lb, klo, khi);
        wfa_range(sb, t, kd, la, lb, glo, ghi);
        for (int k = klo; k <= khi; ++k) {
            const int kb = kd - k;
            if (kb < glo || kb > ghi) continue;
            const int f = F[mid + k], g = G[mid + kb];
            if (f > WFA_NEG / 2 && g > WFA_NEG / 2 && f + g >= la) return true;
        }
        return false;
    };
    wave(false, 0, fprev, fcur);
    wave(true, 0, bprev, bcur);
    if (overlap(0, fcur, 0, bcur)) return 0;
    int sf = 0, sb = 0;
    while (sf + sb < t) {
        { int* x = fprev; fprev = fcur; fcur = x; }
        ++sf;
        wave(false, sf, fprev, fcur);
        if (overlap(sf, fcur, sb, bcur)) return sf + sb;
        if (sf + sb >= t) break;
        { int* x = bprev; bprev = bcur; bcur = x; }
        ++sb;
        wave(true, sb, bprev, bcur);
        if (overlap(sf, fcur, sb, bcur)) return sf + sb;
    }
    return -1;
}

// ---- the round structure of the kernel (wfa.cu), stated serially --------------------------------------------------------------
// Branch-free form of the recurrence: an unreached diagonal holds WFA_NEG, which survives the additions, and the clamp keeps the
// point inside the table (a cell next to a reached cell of the last row or column is within one more edit, so the clamped value
// is still "the furthest row with distance <= s").
SVB_HD int wfa_next_clamped(int fm1, int f0, int fp1, int k, int la, int lb) {
    return wfa_imin(wfa_imax(wfa_imax(f0 + 1, fm1), fp1 + 1), wfa_imin(la, lb - k));
}

// What one CTA of wfa_kernel computes for a pair with 0 < min(la, lb) and |lb - la| <= t, round by round: both waves advance
// from "wave -1" (row -1 on diagonal 0); in round r the forward diagonals also test the backward wave g of round r - 1 on their
// partner diagonal: prev + g >= la (total 2 r - 2), new + g >= la (total 2 r - 1).  Returns the distance if it is <= t, else -1.
// F: four arrays of 2 t + 7 ints (forward parity 0 / 1, backward parity 0 / 1).
SVB_HD long long wfa_rounds_serial(const uint32_t* A, int la, const uint32_t* B, int lb, int t, int* F) {
    const int kd = lb - la, W = 2 * t + 7, mid = t + 3;
    for (int x = 0; x < 4 * W; ++x) F[x] = (x == W + mid || x == 3 * W + mid) ? -1 : WFA_NEG;
    int plo = 0, phi = -1;
    for (int r = 0;; ++r) {
        int klo, khi;
        wfa_range(r, t + 1, kd, la, lb, klo, khi);       // one diagonal more than the one-sided pruning: see wfa.cu
        unsigned hit = 0;
        for (int side = 0; side < 2; ++side) {
            const bool backward = side == 1;
            const int* prev = F + (2 * side + ((r & 1) ^ 1)) * W;
            const int* other = F + (2 * (1 - side) + ((r & 1) ^ 1)) * W;
            int* cur = F + (2 * side + (r & 1)) * W;
            for (int k = klo; k <= khi; ++k) {
                const int fm1 = prev[mid + k - 1], f0 = prev[mid + k], fp1 = prev[mid + k + 1];
                const int kb = kd - k;
                const int g = (!backward && kb >= plo && kb <= phi) ? other[mid + kb] : WFA_NEG;
                const int x = wfa_next_clamped(fm1, f0, fp1, k, la, lb);
                int v = WFA_NEG;
                if (x > WFA_NEG / 2) {
                    v = x;
                    bool more = true;
                    while (more) {
                        if (!backward) v += static_cast<int>(wfa_extend8(A, B, WFA_FRONT + static_cast<uint32_t>(v), WFA_FRONT + static_cast<uint32_t>(v + k), &more));
                        else v += static_cast<int>(wfa_rextend8(A, B, WFA_FRONT + static_cast<uint32_t>(la - 1 - v), WFA_FRONT + static_cast<uint32_t>(lb - 1 - (v + k)), &more));
                    }
                }
                cur[mid + k] = v;
                if (f0 + g >= la) hit |= 1u;
                if (v + g >= la) hit |= 2u;
            }
            cur[mid + klo - 1] = cur[mid + klo - 2] = WFA_NEG;
            cur[mid + khi + 1] = cur[mid + khi + 2] = WFA_NEG;
        }
        if (hit & 1u) return 2 * r - 2;
        if (hit & 2u) return 2 * r - 1 <= t ? 2 * r - 1 : -1;
        if (2 * r - 1 >= t) return -1;
        plo = klo;
        phi = khi;
    }
}
