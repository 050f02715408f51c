// K8 core -- bit-parallel global edit distance (Myers 1999 / Hyyro 2003 block recurrence).
//
// The reference computes edlib.align(h1, h2)["editDistance"] (SVIM_COMBINE.py:50,64,76,88,100):
// edlib's default is mode="NW", task="distance", i.e. the unit-cost global Levenshtein distance
// [ext: edlib is an un-vendored dependency, setup.py:38].  That number is unique, so any exact
// algorithm reproduces it.  This header holds the 64-row block step and the virtual haplotype
// strings of compute_distance (SVIM_COMBINE.py:35-102); edit_distance.cu pipelines the blocks of
// one pattern across the 32 lanes of a warp.  The block step also compiles for the host
// (tests/hostcheck) where it is checked against a plain DP.
#pragma once
#include <stdint.h>

#include "linkage.cuh"   // SVB_HD

// One column step of a 64-row block.  pv/mv: vertical +1/-1 delta bit vectors (in/out).  eq: rows whose
// pattern char equals the text char.  Horizontal deltas travel as 2-bit codes (bit 0: +1, bit 1: -1): `hin`
// enters at the top row, the code of row `hshift` is returned.  Branch-free; the path hin -> return value is
// the loop-carried dependency of the lane pipeline (or, and, 64-bit add, xor, or, one LOP3, shift).
SVB_HD uint32_t myers_step(uint64_t& pv, uint64_t& mv, uint64_t eq, uint32_t hin, uint32_t hshift) {
    const uint64_t hin_neg = hin >> 1, hin_pos = hin & 1u;
    const uint64_t xv = eq | mv;
    eq |= hin_neg;
    const uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
    uint64_t ph = mv | ~(xh | pv);
    uint64_t mh = pv & xh;
    const uint32_t hout = static_cast<uint32_t>((ph >> hshift) & 1ull) | (static_cast<uint32_t>((mh >> hshift) & 1ull) << 1);
    ph = (ph << 1) | hin_pos;
    mh = (mh << 1) | hin_neg;
    pv = mh | ~(xv | ph);
    mv = ph & xv;
    return hout;
}

// Same step with signed deltas (-1, 0, +1) and a one-hot output row; used by the host check against a plain DP.
SVB_HD int myers_block(uint64_t& pv, uint64_t& mv, uint64_t eq, int hin, uint64_t hibit) {
    uint32_t shift = 0;
    while (!((hibit >> shift) & 1ull)) ++shift;
    const uint32_t out = myers_step(pv, mv, eq, hin > 0 ? 1u : (hin < 0 ? 2u : 0u), shift);
    return static_cast<int>(out & 1u) - static_cast<int>(out >> 1);
}

// ---- sliding window over the Ukkonen band -----------------------------------------------------------
// A global alignment of cost d never leaves the diagonals [-d, (n - m) + d] (n >= m), so with half-width K
// only the 64-row blocks b = 0 .. (m-1)/64 on the text columns [win_jlo(b), win_jhi(b)) matter, and the
// result is exact whenever it is <= K.  One warp holds 32 consecutive blocks of that band at a time: lane
// b % 32 works on block b, and once block b has run out of columns the lane moves on to block b + 32.
// That needs block b + 32 to start after block b ends, start(b) = win_jlo(b) + b (lanes are skewed by one
// column per block): 64 * 32 + 32 - K >= 64 + (n - m) + K (with 64-row blocks; 32-row blocks halve the band and the
// work per step).  A block that enters the band starts from
// "everything above and to the left is one more per row" (vertical deltas +1, horizontal input +1), which
// can only overestimate cells whose optimal path leaves the band.  With S_b = sum of the block's bottom-row
// horizontal deltas over the columns [win_jlo(b), win_jlo(b + 1)) (last block: up to n),
//     D[m][n] = m + sum_b S_b.
struct WinGeom {
    uint32_t m, n, K, last_block;
    uint32_t bw;           // rows per block: 64 (one 64-bit word per lane) or 32 (half the work per step, half the band)
    uint32_t dlt;          // upper band offset: n - m for a whole alignment; a HALF of a split alignment (first m rows of a
                           // longer pattern, first n columns of its text) keeps the whole problem's value
};
constexpr uint32_t WIN_SLACK = 8;     // idle steps guaranteed between two blocks of one lane (prefetch priming)

SVB_HD WinGeom win_geom(uint32_t m, uint32_t n, uint32_t K, uint32_t bw) {
    WinGeom g;
    g.m = m; g.n = n; g.K = K; g.bw = bw; g.last_block = (m - 1u) / bw; g.dlt = n - m;
    return g;
}
SVB_HD uint32_t win_kmax(uint32_t m, uint32_t n, uint32_t bw) {   // widest half-width one warp can slide over (0: none)
    const uint32_t dlt = n - m, room = bw * 32u + 32u - bw - WIN_SLACK;
    return dlt >= room ? 0u : (room - dlt) / 2u;
}
SVB_HD uint32_t win_jlo(const WinGeom& g, uint32_t b) {
    const uint64_t x = static_cast<uint64_t>(g.bw) * b;
    return x > g.K ? static_cast<uint32_t>(x - g.K) : 0u;
}
SVB_HD uint32_t win_jhi(const WinGeom& g, uint32_t b) {
    const uint64_t x = static_cast<uint64_t>(g.bw) * b + g.bw + g.dlt + g.K;
    return x < g.n ? static_cast<uint32_t>(x) : g.n;
}
// ---- split alignment (two warps per pair) ------------------------------------------------------------------------------
// D[m][n] = min over j of F(j) + B(n - j), F(j) = D[mh][j] from the first mh pattern rows, B(j') the same quantity of the
// REVERSED strings from their first m - mh rows.  Inside the band the optimal path crosses row mh at a column where both
// halves are exact, everywhere else both are upper bounds, so the minimum is D[m][n] whenever it is <= K.  Each half only
// needs the text columns its band reaches: about n / 2 + K steps instead of n, and the two halves run concurrently.
// m and n are multiples of the block height (the caller pads with the sentinel symbol).
struct WinSplit {
    uint32_t rows_f, rows_b;   // pattern rows of the forward / backward half
    uint32_t cols_f, cols_b;   // text columns each half processes
};
SVB_HD WinSplit win_split(uint32_t m, uint32_t n, uint32_t K, uint32_t bw) {
    WinSplit s;
    const uint32_t blocks = m / bw, dlt = n - m;
    s.rows_f = (blocks / 2u) * bw;
    s.rows_b = m - s.rows_f;
    const uint64_t cf = static_cast<uint64_t>(s.rows_f) + dlt + K, cb = static_cast<uint64_t>(s.rows_b) + dlt + K;
    s.cols_f = cf < n ? static_cast<uint32_t>(cf) : n;
    s.cols_b = cb < n ? static_cast<uint32_t>(cb) : n;
    return s;
}
SVB_HD WinGeom win_geom_half(uint32_t rows, uint32_t cols, uint32_t dlt, uint32_t K, uint32_t bw) {
    WinGeom g;
    g.m = rows; g.n = cols; g.K = K; g.bw = bw; g.last_block = (rows - 1u) / bw; g.dlt = dlt;
    return g;
}

// per-block step limits, all relative to the block's first column (rel = column - win_jlo(b)):
//   rel < width: the block is inside the band;  rel < hin_lim: the block above is too (else +1 enters);
//   rel < cnt_lim: the bottom-row delta counts towards D[m][n]
struct WinBlock {
    uint32_t start;        // global step of its first column: win_jlo(b) + b
    uint32_t width, hin_lim, cnt_lim, hshift;
};
SVB_HD WinBlock win_block(const WinGeom& g, uint32_t b) {
    WinBlock w;
    const uint32_t lo = win_jlo(g, b), hi = win_jhi(g, b);
    w.start = lo + b;
    w.width = hi - lo;
    w.hin_lim = b == 0u ? 0u : win_jhi(g, b - 1u) - lo;
    w.cnt_lim = b == g.last_block ? w.width : win_jlo(g, b + 1u) - lo;
    w.hshift = b == g.last_block ? ((g.m - 1u) % g.bw) : g.bw - 1u;
    return w;
}

// The block step on a 32-row block (same recurrence as myers_step on one 32-bit word).
SVB_HD uint32_t myers_step32(uint32_t& pv, uint32_t& mv, uint32_t eq, uint32_t hin, uint32_t hshift) {
    const uint32_t hin_neg = hin >> 1, hin_pos = hin & 1u;
    const uint32_t xv = eq | mv;
    eq |= hin_neg;
    const uint32_t xh = (((eq & pv) + pv) ^ pv) | eq;
    uint32_t ph = mv | ~(xh | pv);
    uint32_t mh = pv & xh;
    const uint32_t hout = ((ph >> hshift) & 1u) | (((mh >> hshift) & 1u) << 1);
    ph = (ph << 1) | hin_pos;
    mh = (mh << 1) | hin_neg;
    pv = mh | ~(xv | ph);
    mv = ph & xv;
    return hout;
}

// ---- virtual haplotype strings --------------------------------------------------------------------
// compute_distance builds  ref[lo:s] + MIDDLE + ref[e:hi]  for both candidates; MIDDLE depends on the
// type.  Nothing is materialised: a descriptor addresses the bases in HBM.
enum : uint32_t {
    HAP_MID_NONE = 0,      // DEL  (SVIM_COMBINE.py:48-49)
    HAP_MID_REVCOMP = 1,   // INV: reverse complement of ref[s:e], non-ACGT unchanged (:56-63)
    HAP_MID_REPEAT = 2,    // DUP_TAN: ref[s:e] * (copies + 1) (:82-87)
    HAP_MID_SEQ4 = 3,      // INS: the candidate's own sequence, 4-bit packed query bases (:70-75)
    HAP_MID_REF = 4        // DUP_INT: ref[source_start:source_end] of the source contig (:94-99)
};

struct HapDesc {
    uint64_t l_base;       // absolute index into the concatenated reference of ref[lo]
    uint64_t r_base;       // ... of ref[e]
    uint64_t m_base;       // reference index of the unit start, or NIBBLE index into seq4
    uint32_t l_len, r_len, m_len;
    uint32_t m_kind;
    uint32_t m_unit;       // HAP_MID_REPEAT: length of one copy
    uint32_t seq_sel;      // HAP_MID_SEQ4: 0 = haplotype-1 sequences, 1 = haplotype-2 sequences
};

SVB_HD uint32_t hap_length(const HapDesc& d) { return d.l_len + d.m_len + d.r_len; }

SVB_HD uint8_t hap_complement(uint8_t c) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        default: return c;
    }
}

SVB_HD uint8_t hap_char(const HapDesc& d, uint32_t i, const uint8_t* ref, const uint8_t* seq4_a, const uint8_t* seq4_b) {
    if (i < d.l_len) return ref[d.l_base + i];
    i -= d.l_len;
    if (i < d.m_len) {
        switch (d.m_kind) {
            case HAP_MID_REVCOMP: return hap_complement(ref[d.m_base + (d.m_len - 1u - i)]);
            case HAP_MID_REPEAT: return ref[d.m_base + (i % d.m_unit)];
            case HAP_MID_SEQ4: {
                const uint8_t* s = d.seq_sel ? seq4_b : seq4_a;
                const uint64_t nib = d.m_base + i;
                const uint8_t byte = s[nib >> 1];
                const uint32_t code = (nib & 1ull) ? (byte & 15u) : (byte >> 4);
                return static_cast<uint8_t>("=ACMGRSVTWYHKDBN"[code]);
            }
            default: return ref[d.m_base + i];
        }
    }
    i -= d.m_len;
    return ref[d.r_base + i];
}

// Split form of hap_char for latency hiding: hap_fetch computes the address and issues ONE load whose result is not
// touched; `mode` says how to turn the byte into a symbol later (0: reference byte, 1: complemented reference byte,
// 2 / 3: high / low nibble of a packed query byte).
enum : uint32_t { TOK_REF = 0, TOK_REF_COMP = 1, TOK_NIB_HI = 2, TOK_NIB_LO = 3, TOK_NONE = 4 };

SVB_HD uint32_t hap_fetch(const HapDesc& d, uint32_t i, const uint8_t* ref, const uint8_t* seq4_a, const uint8_t* seq4_b,
                          uint32_t& mode) {
    const uint8_t* p;
    mode = TOK_REF;
    if (i < d.l_len) {
        p = ref + d.l_base + i;
    } else {
        i -= d.l_len;
        if (i < d.m_len) {
            if (d.m_kind == HAP_MID_SEQ4) {
                const uint64_t nib = d.m_base + i;
                p = (d.seq_sel ? seq4_b : seq4_a) + (nib >> 1);
                mode = TOK_NIB_HI + static_cast<uint32_t>(nib & 1ull);
            } else if (d.m_kind == HAP_MID_REVCOMP) {
                p = ref + d.m_base + (d.m_len - 1u - i);
                mode = TOK_REF_COMP;
            } else if (d.m_kind == HAP_MID_REPEAT) {
                p = ref + d.m_base + (i % d.m_unit);
            } else {
                p = ref + d.m_base + i;
            }
        } else {
            p = ref + d.r_base + (i - d.m_len);
        }
    }
    return *p;
}
