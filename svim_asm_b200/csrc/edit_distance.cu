// K8 edit_distance -- exact unit-cost global edit distance of haplotype-string pairs, one warp per pair.
//
// Replaces compute_distance's edlib.align(h1, h2)["editDistance"] (reference SVIM_COMBINE.py:35-102).
// The shorter string is the pattern: its rows are cut into 64-row blocks (edit_core.cuh), block b of a
// 2048-row stripe lives in lane b, and the columns of the text stream through the lanes as a software
// pipeline -- lane b works on column t-b at step t and hands its 2-bit horizontal delta to lane b+1 with one
// shuffle.  Patterns longer than 2048 rows take several stripes; the horizontal deltas of a stripe's last row
// are parked in a per-warp byte buffer in HBM.  Strings are never materialised: bases are read through HapDesc
// (reference bytes / 4-bit query bases resident in HBM).
//
// Bound: the latency of ONE warp's dependency chain (the longest pair is the critical path of the launch), so
// the step loop is written for latency: shuffle + block recurrence are the only loop-carried work; the match
// mask of step t+1 and the symbol class of step t+2 are loaded while step t computes, the steady state runs
// without masks or divergent branches, and the global loads that feed the column ring are issued 32..64 steps
// before their first use and only decoded (class lookup) when they enter the ring.
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "edit_core.cuh"
#include "pairing.cuh"
#include "edit_strings.cuh"
#include "wfa_core.cuh"

namespace {
using namespace edstr;

constexpr int ED_WARPS = 2;
constexpr uint32_t ED_SENTINEL = ED_NCLASS + 1;   // padding symbol of window_pass32: matches only itself
constexpr uint32_t ED_SKIP = ED_NCLASS + 2;       // "no row here" while match masks are built
constexpr uint32_t PEQ_ROWS = ED_NCLASS + 2;

struct EdShared {
    // [buffer][class][lane]: match mask of the lane's 64 rows.  The striped path uses buffer 0; the sliding window
    // alternates (a lane's next block is built while it still works on the current one).
    unsigned long long peq[ED_WARPS][2][PEQ_ROWS][32];
    uint8_t cls2[512];                                     // byte -> class, complemented byte -> class
    uint8_t tcls[ED_WARPS][128];                           // ring: symbol class of the band's columns
    uint8_t th[ED_WARPS][64];                              // ring: delta code entering the stripe's top row
    // split alignment of a long pair (both warps of the CTA on one job): bottom-row delta codes of each half's last block,
    // the value they start from, the job and the verdict
    uint8_t rec[2][1280];
    long long half_base[2];
    long long split_result;
    uint32_t cta_job, cta_worker;
    uint32_t trim[2];
};

// ---- sliding window (edit_core.cuh): ONE pass of n + (m-1)/64 steps over a band of half-width K ------------------
// Lane l works on the blocks l, l + 32, l + 64, ... one after the other; at global step t the lane that holds block b
// is on text column t - b.  Returns the window's value, which is D[m][n] whenever it is <= K.
// Every 32 steps ("grid point") the warp (a) moves up to 32 further text columns from registers into the 128-slot
// class ring and requests the next 32, (b) builds the match masks of the next block that will enter the band into
// the idle mask buffer of its lane; the pattern bytes for that were requested one grid point earlier.
template <int BW>
__device__ __forceinline__ long long window_pass(const HapDesc& P, const HapDesc& T, uint32_t pre, uint32_t m, uint32_t n,
                                              uint32_t K, const uint8_t* ref, const uint8_t* sa, const uint8_t* sb,
                                              const uint8_t* cls2tab, unsigned long long* peq_mem, uint8_t* tcls, uint32_t lane) {
    static_assert(BW == 64 || BW == 32, "block = one 64-bit or one 32-bit word per lane");
    using Word = typename std::conditional<BW == 64, uint64_t, uint32_t>::type;
    constexpr uint32_t WPB = BW / 32;                                // 32-bit words per block
    const WinGeom g = win_geom(m, n, K, BW);
    Word* peqw = reinterpret_cast<Word*>(peq_mem);                   // [buffer][class][lane]
    uint32_t* peq32 = reinterpret_cast<uint32_t*>(peq_mem);          // [buffer][class][lane][word of the block]

    // ---- masks of the blocks 0 .. 31 (buffer 0), 32 rows per round like the striped path
    __syncwarp();
    for (uint32_t i = lane; i < 2u * PEQ_ROWS * 32u * WPB; i += 32u) peq32[i] = 0u;
    for (uint32_t i = lane; i < 128u; i += 32u) tcls[i] = static_cast<uint8_t>(ED_NOCLASS);
    __syncwarp();
    const uint32_t rows0 = min(m, 32u * BW), ngroups = (rows0 + 31u) / 32u;
    for (uint32_t g0 = 0; g0 < ngroups; g0 += 4u) {
        uint32_t byte[4], mode[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t idx = (g0 + k) * 32u + lane;
            mode[k] = TOK_NONE;
            byte[k] = 0u;
            if (idx < rows0) byte[k] = hap_fetch(P, pre + idx, ref, sa, sb, mode[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t cls = mode[k] == TOK_NONE ? ED_NOCLASS + 1u : tok_class(cls2tab, byte[k], mode[k]);
            const uint32_t peers = __match_any_sync(FULL, cls);
            if (cls < ED_NOCLASS && static_cast<uint32_t>(__ffs(peers) - 1) == lane) peq32[cls * (32u * WPB) + (g0 + k)] = peers;
        }
    }
    // ---- later blocks: bytes requested ahead (pat_*), masks built at a grid point
    uint32_t next_build = 32u;                                       // next block whose masks are to be built
    uint32_t pat_byte[WPB], pat_mode[WPB];
    auto pattern_fetch = [&]() {
#pragma unroll
        for (int k = 0; k < static_cast<int>(WPB); ++k) {
            const uint64_t idx = static_cast<uint64_t>(BW) * next_build + 32u * k + lane;
            pat_mode[k] = TOK_NONE;
            pat_byte[k] = 0u;
            if (next_build <= g.last_block && idx < m) pat_byte[k] = hap_fetch(P, pre + static_cast<uint32_t>(idx), ref, sa, sb, pat_mode[k]);
        }
    };
    auto pattern_build = [&]() {                                     // block next_build -> buffer (b / 32) & 1, column b % 32
        const uint32_t buf = (next_build >> 5) & 1u, col = next_build & 31u;
        for (uint32_t i = lane; i < PEQ_ROWS * WPB; i += 32u) peq32[((buf * PEQ_ROWS + i / WPB) * 32u + col) * WPB + (i % WPB)] = 0u;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < static_cast<int>(WPB); ++k) {
            const uint32_t cls = pat_mode[k] == TOK_NONE ? ED_NOCLASS + 1u : tok_class(cls2tab, pat_byte[k], pat_mode[k]);
            const uint32_t peers = __match_any_sync(FULL, cls);
            if (cls < ED_NOCLASS && static_cast<uint32_t>(__ffs(peers) - 1) == lane)
                peq32[((buf * PEQ_ROWS + cls) * 32u + col) * WPB + static_cast<uint32_t>(k)] = peers;
        }
        __syncwarp();
    };
    pattern_fetch();

    // ---- column ring: columns [0, c_ins) are in; the batch [c_ins, c_ins + 32) waits in registers
    uint32_t c_ins = 0, nxt_byte = 0, nxt_mode = TOK_NONE;
    auto ring_fetch = [&]() {
        const uint32_t col = c_ins + lane;
        nxt_mode = TOK_NONE;
        nxt_byte = 0u;
        if (col < n) nxt_byte = hap_fetch(T, pre + col, ref, sa, sb, nxt_mode);
    };
    auto ring_insert = [&]() {
        tcls[(c_ins + lane) & 127u] = static_cast<uint8_t>(nxt_mode == TOK_NONE ? ED_NOCLASS : tok_class(cls2tab, nxt_byte, nxt_mode));
        c_ins += 32u;
    };
    for (int k = 0; k < 3; ++k) {
        ring_fetch();
        ring_insert();
    }
    ring_fetch();
    __syncwarp();

    // ---- per-lane block state.  What a step does with its block depends on where it is inside the block's columns
    // (rel = t - start): rel < width: inside the band (commit the new vertical deltas); rel < hin_lim: the block above
    // is inside too (take its delta, else +1); rel < cnt_lim: the bottom-row delta counts.
    uint32_t blk = lane, buf = 0;
    WinBlock w;
    w.start = 0; w.width = 0; w.hin_lim = 0; w.cnt_lim = 0; w.hshift = BW - 1u;
    if (blk <= g.last_block) w = win_block(g, blk);
    int colbase = -static_cast<int>(blk);                            // text column of step t: t + colbase
    Word pv = static_cast<Word>(~0ull), mv = 0;
    uint32_t acc_codes = 0, acc_minus = 0;                           // sum of counted delta codes (0, 1, 2) / number of -1 among them
    uint32_t hout = 0;
    auto ring_class = [&](int col) -> uint32_t { return (col >= 0 && static_cast<uint32_t>(col) < n) ? tcls[static_cast<uint32_t>(col) & 127u] : ED_NOCLASS; };
    const Word* peq_lane = peqw + lane;                              // + (buffer * PEQ_ROWS + class) * 32
    Word eq0 = peq_lane[ring_class(colbase) * 32u];
    uint32_t cls1 = ring_class(colbase + 1);
    uint32_t b_lo = 0;                                               // smallest block that still has columns to do
    const uint32_t t_end = n + g.last_block;
    const uint32_t src_lane = (lane + 31u) & 31u;
    // A lane changes block when its block has run out of columns: that is the only EVENT of the step loop (warp-wide
    // minimum of the lanes' next switch step, one per bw + 1 steps); the three per-step conditions above are plain
    // compares on rel, off the dependency chain of the recurrence (the loop is latency bound, they cost nothing).
    constexpr uint32_t NEVER = 0xFFFFFFFFu;
    auto switch_step = [&]() -> uint32_t { return w.width != 0u ? w.start + w.width : NEVER; };
    auto lane_switch = [&](uint32_t t) {
        if (t == switch_step()) {                                    // on to block blk + 32 (WIN_SLACK idle steps follow)
            blk += 32u;
            buf ^= 1u;
            colbase -= 32;
            pv = static_cast<Word>(~0ull);
            mv = 0;
            peq_lane = peqw + buf * (PEQ_ROWS * 32u) + lane;
            w.start = 0; w.width = 0; w.hin_lim = 0; w.cnt_lim = 0; w.hshift = BW - 1u;
            if (blk <= g.last_block) w = win_block(g, blk);
        }
    };
    uint32_t next_event = __reduce_min_sync(FULL, switch_step());

    for (uint32_t t0 = 0; t0 < t_end; t0 += 32u) {
        if (t0) {                                                    // grid point
            while (b_lo <= g.last_block && win_jhi(g, b_lo) + b_lo <= t0) ++b_lo;
            const int oldest = static_cast<int>(t0) - static_cast<int>(b_lo) - 31;     // oldest column any lane still reads
            if (static_cast<int>(c_ins) <= oldest + 94) {
                ring_insert();
                ring_fetch();
            }
            if (next_build <= g.last_block && t0 > win_jlo(g, next_build - 32u) + (next_build - 32u)) {
                pattern_build();                                     // its lane has moved on to block next_build - 32
                ++next_build;
                pattern_fetch();
            }
            __syncwarp();
        }
        const uint32_t t1 = min(t_end, t0 + 32u);
        uint32_t t = t0;
        while (t < t1) {
            if (t == next_event) {
                lane_switch(t);
                next_event = __reduce_min_sync(FULL, switch_step());
            }
            const uint32_t t_stop = min(t1, next_event);             // next_event > t here
            uint32_t ridx = static_cast<uint32_t>(static_cast<int>(t) + 2 + colbase);
            uint32_t rel = t - w.start;                              // wraps to a huge value before the block starts
            const uint32_t width = w.width, hin_lim = w.hin_lim, cnt_lim = w.cnt_lim, hshift = w.hshift;
#pragma unroll 4
            for (; t < t_stop; ++t) {
                const uint32_t shin = __shfl_sync(FULL, hout, src_lane);
                const uint32_t hin = rel < hin_lim ? shin : 1u;
                const Word eq1 = peq_lane[cls1 * 32u];
                const uint32_t cls2 = tcls[ridx & 127u];
                ++ridx;
                Word npv = pv, nmv = mv;
                uint32_t ho;
                if constexpr (BW == 64) ho = myers_step(npv, nmv, eq0, hin, hshift);
                else ho = myers_step32(npv, nmv, eq0, hin, hshift);
                const bool active = rel < width;
                pv = active ? npv : pv;
                mv = active ? nmv : mv;
                const uint32_t counted = rel < cnt_lim ? ho : 0u;
                acc_codes += counted;
                acc_minus += counted >> 1;
                hout = ho;
                eq0 = eq1;
                cls1 = cls2;
                ++rel;
            }
        }
    }
    // code 1 is +1, code 2 is -1: sum of deltas = (#1) - (#2) = acc_codes - 3 * acc_minus
    const int partial = static_cast<int>(acc_codes) - 3 * static_cast<int>(acc_minus);
    __syncwarp();
    return static_cast<long long>(m) + __reduce_add_sync(FULL, partial);
}

// ---- the same window with 32-row blocks, written for the shortest dependency chain per step -------------------------
// (the step loop is bound by ONE chain: shuffle -> block recurrence -> shuffle)
//  * both strings are padded with a sentinel symbol to a multiple of 32 rows: D(P + S^k, T + S^k) = D(P, T), and the
//    bottom row of EVERY block is bit 31 (a constant shift instead of a variable one);
//  * the horizontal delta travels as two registers (is -1 / is +1) in two shuffles, so the -1 flag ORs straight into
//    the match mask and nothing has to be unpacked on the chain:
//    select, or-and, add, xor-or, and-or-not, shift = 6 dependent instructions between the shuffles.
// m0, n0: the real lengths.  Returns the window's value, which is D[m0][n0] whenever it is <= K.
// HALF = false: the whole alignment.  HALF = true: one half of a split alignment (edit_core.cuh, win_split): the first
// `rows` rows of the padded pattern against the first `cols` columns of the padded text, both read from the far end when
// `rev` is set; the last block's bottom-row deltas are not summed but recorded per column in `rec` (2-bit codes), and the
// return value is D[rows][first column of that block], the base the caller's prefix sums start from.
template <bool HALF>
__device__ __forceinline__ long long window_core32(const HapDesc& P, const HapDesc& T, uint32_t pre, uint32_t m0, uint32_t n0,
                                                   uint32_t K, const uint8_t* ref, const uint8_t* sa, const uint8_t* sb,
                                                   const uint8_t* cls2tab, unsigned long long* peq_mem, uint8_t* tcls, uint32_t lane,
                                                   uint32_t rows, uint32_t cols, bool rev, uint8_t* rec) {
    constexpr int BW = 32;
    using Word = uint32_t;
    constexpr uint32_t WPB = 1;
    const uint32_t padk = (32u - (m0 & 31u)) & 31u, mp = m0 + padk, np_ = n0 + padk;    // padded lengths of the whole strings
    const uint32_t m = HALF ? rows : mp, n = HALF ? cols : np_;                          // what this pass covers
    const WinGeom g = HALF ? win_geom_half(m, n, np_ - mp, K, BW) : win_geom(m, n, K, BW);
    // row i / column j of this pass -> index in the padded strings (sentinel symbols sit at the far end of both)
    auto prow = [&](uint32_t i) -> uint32_t { return (HALF && rev) ? mp - 1u - i : i; };
    auto tcol = [&](uint32_t j) -> uint32_t { return (HALF && rev) ? np_ - 1u - j : j; };
    Word* peqw = reinterpret_cast<Word*>(peq_mem);                   // [buffer][class][lane]
    uint32_t* peq32 = reinterpret_cast<uint32_t*>(peq_mem);          // [buffer][class][lane][word of the block]

    // ---- masks of the blocks 0 .. 31 (buffer 0), 32 rows per round like the striped path
    __syncwarp();
    for (uint32_t i = lane; i < 2u * PEQ_ROWS * 32u * WPB; i += 32u) peq32[i] = 0u;
    for (uint32_t i = lane; i < 128u; i += 32u) tcls[i] = static_cast<uint8_t>(ED_NOCLASS);
    __syncwarp();
    const uint32_t rows0 = min(m, 32u * BW), ngroups = (rows0 + 31u) / 32u;
    for (uint32_t g0 = 0; g0 < ngroups; g0 += 4u) {
        uint32_t byte[4], mode[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t idx = (g0 + k) * 32u + lane;
            mode[k] = TOK_NONE;
            byte[k] = 0u;
            if (idx < rows0 && prow(idx) < m0) byte[k] = hap_fetch(P, pre + prow(idx), ref, sa, sb, mode[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t idx = (g0 + k) * 32u + lane;
            uint32_t cls = mode[k] == TOK_NONE ? ED_SKIP : tok_class(cls2tab, byte[k], mode[k]);
            if (idx < rows0 && prow(idx) >= m0) cls = ED_SENTINEL;
            const uint32_t peers = __match_any_sync(FULL, cls);
            if (cls != ED_NOCLASS && cls < ED_SKIP && static_cast<uint32_t>(__ffs(peers) - 1) == lane) peq32[cls * (32u * WPB) + (g0 + k)] = peers;
        }
    }
    // ---- later blocks: bytes requested ahead (pat_*), masks built at a grid point
    uint32_t next_build = 32u;                                       // next block whose masks are to be built
    uint32_t pat_byte[WPB], pat_mode[WPB];
    auto pattern_fetch = [&]() {
#pragma unroll
        for (int k = 0; k < static_cast<int>(WPB); ++k) {
            const uint64_t idx = static_cast<uint64_t>(BW) * next_build + 32u * k + lane;
            pat_mode[k] = TOK_NONE;
            pat_byte[k] = 0u;
            if (next_build <= g.last_block && idx < m && prow(static_cast<uint32_t>(idx)) < m0)
                pat_byte[k] = hap_fetch(P, pre + prow(static_cast<uint32_t>(idx)), ref, sa, sb, pat_mode[k]);
        }
    };
    auto pattern_build = [&]() {                                     // block next_build -> buffer (b / 32) & 1, column b % 32
        const uint32_t buf = (next_build >> 5) & 1u, col = next_build & 31u;
        for (uint32_t i = lane; i < PEQ_ROWS * WPB; i += 32u) peq32[((buf * PEQ_ROWS + i / WPB) * 32u + col) * WPB + (i % WPB)] = 0u;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < static_cast<int>(WPB); ++k) {
            const uint64_t idx = static_cast<uint64_t>(BW) * next_build + 32u * k + lane;
            uint32_t cls = pat_mode[k] == TOK_NONE ? ED_SKIP : tok_class(cls2tab, pat_byte[k], pat_mode[k]);
            if (idx < m && prow(static_cast<uint32_t>(idx)) >= m0) cls = ED_SENTINEL;
            const uint32_t peers = __match_any_sync(FULL, cls);
            if (cls != ED_NOCLASS && cls < ED_SKIP && static_cast<uint32_t>(__ffs(peers) - 1) == lane)
                peq32[((buf * PEQ_ROWS + cls) * 32u + col) * WPB + static_cast<uint32_t>(k)] = peers;
        }
        __syncwarp();
    };
    pattern_fetch();

    // ---- column ring: columns [0, c_ins) are in; the batch [c_ins, c_ins + 32) waits in registers
    uint32_t c_ins = 0, nxt_byte = 0, nxt_mode = TOK_NONE;
    auto ring_fetch = [&]() {
        const uint32_t col = c_ins + lane;
        nxt_mode = TOK_NONE;
        nxt_byte = 0u;
        if (col < n && tcol(col) < n0) nxt_byte = hap_fetch(T, pre + tcol(col), ref, sa, sb, nxt_mode);
    };
    auto ring_insert = [&]() {
        const uint32_t col = c_ins + lane;
        uint32_t cls = nxt_mode == TOK_NONE ? ED_NOCLASS : tok_class(cls2tab, nxt_byte, nxt_mode);
        if (col < n && tcol(col) >= n0) cls = ED_SENTINEL;
        tcls[col & 127u] = static_cast<uint8_t>(cls);
        c_ins += 32u;
    };
    for (int k = 0; k < 3; ++k) {
        ring_fetch();
        ring_insert();
    }
    ring_fetch();
    __syncwarp();

    // ---- per-lane block state.  What a step does with its block depends on where it is inside the block's columns
    // (rel = t - start): rel < width: inside the band (commit the new vertical deltas); rel < hin_lim: the block above
    // is inside too (take its delta, else +1); rel < cnt_lim: the bottom-row delta counts.
    uint32_t blk = lane, buf = 0;
    WinBlock w;
    w.start = 0; w.width = 0; w.hin_lim = 0; w.cnt_lim = 0; w.hshift = BW - 1u;
    if (blk <= g.last_block) w = win_block(g, blk);
    bool is_last = HALF && blk == g.last_block;
    if (is_last) w.cnt_lim = 0u;
    int colbase = -static_cast<int>(blk);                            // text column of step t: t + colbase
    Word pv = static_cast<Word>(~0ull), mv = 0;
    uint32_t acc_pos = 0, acc_neg = 0;                               // counted +1 / -1 deltas of the bottom rows
    uint32_t hneg_out = 0, hpos_out = 0;
    auto ring_class = [&](int col) -> uint32_t { return (col >= 0 && static_cast<uint32_t>(col) < n) ? tcls[static_cast<uint32_t>(col) & 127u] : ED_NOCLASS; };
    const Word* peq_lane = peqw + lane;                              // + (buffer * PEQ_ROWS + class) * 32
    Word eq0 = peq_lane[ring_class(colbase) * 32u];
    uint32_t cls1 = ring_class(colbase + 1);
    uint32_t b_lo = 0;                                               // smallest block that still has columns to do
    const uint32_t t_end = n + g.last_block;
    const uint32_t src_lane = (lane + 31u) & 31u;
    // A lane changes block when its block has run out of columns: that is the only EVENT of the step loop (warp-wide
    // minimum of the lanes' next switch step, one per bw + 1 steps); the three per-step conditions above are plain
    // compares on rel, off the dependency chain of the recurrence (the loop is latency bound, they cost nothing).
    constexpr uint32_t NEVER = 0xFFFFFFFFu;
    auto switch_step = [&]() -> uint32_t { return w.width != 0u ? w.start + w.width : NEVER; };
    auto lane_switch = [&](uint32_t t) {
        if (t == switch_step()) {                                    // on to block blk + 32 (WIN_SLACK idle steps follow)
            blk += 32u;
            buf ^= 1u;
            colbase -= 32;
            pv = static_cast<Word>(~0ull);
            mv = 0;
            peq_lane = peqw + buf * (PEQ_ROWS * 32u) + lane;
            w.start = 0; w.width = 0; w.hin_lim = 0; w.cnt_lim = 0; w.hshift = BW - 1u;
            if (blk <= g.last_block) w = win_block(g, blk);
            is_last = HALF && blk == g.last_block;
            if (is_last) w.cnt_lim = 0u;
        }
    };
    uint32_t next_event = __reduce_min_sync(FULL, switch_step());

    for (uint32_t t0 = 0; t0 < t_end; t0 += 32u) {
        if (t0) {                                                    // grid point
            while (b_lo <= g.last_block && win_jhi(g, b_lo) + b_lo <= t0) ++b_lo;
            const int oldest = static_cast<int>(t0) - static_cast<int>(b_lo) - 31;     // oldest column any lane still reads
            if (static_cast<int>(c_ins) <= oldest + 94) {
                ring_insert();
                ring_fetch();
            }
            if (next_build <= g.last_block && t0 > win_jlo(g, next_build - 32u) + (next_build - 32u)) {
                pattern_build();                                     // its lane has moved on to block next_build - 32
                ++next_build;
                pattern_fetch();
            }
            __syncwarp();
        }
        const uint32_t t1 = min(t_end, t0 + 32u);
        uint32_t t = t0;
        while (t < t1) {
            if (t == next_event) {
                lane_switch(t);
                next_event = __reduce_min_sync(FULL, switch_step());
            }
            const uint32_t t_stop = min(t1, next_event);             // next_event > t here
            uint32_t ridx = static_cast<uint32_t>(static_cast<int>(t) + 2 + colbase);
            uint32_t rel = t - w.start;                              // wraps to a huge value before the block starts
            const uint32_t width = w.width, hin_lim = w.hin_lim, cnt_lim = w.cnt_lim;
#pragma unroll 4
            for (; t < t_stop; ++t) {
                const uint32_t sneg = __shfl_sync(FULL, hneg_out, src_lane), spos = __shfl_sync(FULL, hpos_out, src_lane);
                const bool take = rel < hin_lim;
                const uint32_t hneg = take ? sneg : 0u, hpos = take ? spos : 1u;
                const Word eq1 = peq_lane[cls1 * 32u];
                const uint32_t cls2 = tcls[ridx & 127u];
                ++ridx;
                const uint32_t xv = eq0 | mv;
                const uint32_t eqn = eq0 | hneg;
                const uint32_t xh = (((eqn & pv) + pv) ^ pv) | eqn;
                uint32_t ph = mv | ~(xh | pv);
                uint32_t mh = pv & xh;
                const uint32_t pos_o = ph >> 31, neg_o = mh >> 31;
                ph = (ph << 1) | hpos;
                mh = (mh << 1) | hneg;
                const bool active = rel < width;
                pv = active ? (mh | ~(xv | ph)) : pv;
                mv = active ? (ph & xv) : mv;
                const bool cnt = rel < cnt_lim;
                acc_pos += cnt ? pos_o : 0u;
                acc_neg += cnt ? neg_o : 0u;
                if (HALF && is_last && active) rec[rel] = static_cast<uint8_t>(pos_o | (neg_o << 1));
                hpos_out = pos_o;
                hneg_out = neg_o;
                eq0 = eq1;
                cls1 = cls2;
                ++rel;
            }
        }
    }
    const int partial = static_cast<int>(acc_pos) - static_cast<int>(acc_neg);
    __syncwarp();
    return static_cast<long long>(m) + __reduce_add_sync(FULL, partial);
}

__device__ __forceinline__ long long window_pass32(const HapDesc& P, const HapDesc& T, uint32_t pre, uint32_t m0, uint32_t n0,
                                                   uint32_t K, const uint8_t* ref, const uint8_t* sa, const uint8_t* sb,
                                                   const uint8_t* cls2tab, unsigned long long* peq_mem, uint8_t* tcls, uint32_t lane) {
    return window_core32<false>(P, T, pre, m0, n0, K, ref, sa, sb, cls2tab, peq_mem, tcls, lane, 0u, 0u, false, nullptr);
}

__global__ void __launch_bounds__(ED_WARPS * 32) edit_distance_kernel(const EditJob* __restrict__ jobs,
                                                                       const uint32_t* __restrict__ job_list,
                                                                       const unsigned long long* __restrict__ n_list_dev,
                                                                       unsigned int* next_job,
                                                                       const uint8_t* __restrict__ ref,
                                                                       const uint8_t* __restrict__ seq4_a,
                                                                       const uint8_t* __restrict__ seq4_b,
                                                                       const uint8_t* __restrict__ class_map,
                                                                       uint8_t* hbuf_pool, uint64_t hbuf_stride,
                                                                       double* __restrict__ out, double* __restrict__ out_hi,
                                                                       uint4* __restrict__ profile,
                                                                       unsigned int* sm_tokens, unsigned long long* need) {
    __shared__ EdShared sh;
    const uint32_t n_jobs = static_cast<uint32_t>(*n_list_dev);
    if (n_jobs == 0u) return;                         // the usual case: the wavefront kernel settled every pair
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) {
        sh.cls2[i] = class_map[i];
        sh.cls2[256u + i] = class_map[hap_complement(static_cast<uint8_t>(i))];
    }
    // the maskless steady state lets lanes without a block read ring slots no refill has written yet: any class <= 32 is fine
    for (uint32_t i = threadIdx.x; i < ED_WARPS * 128u; i += blockDim.x) (&sh.tcls[0][0])[i] = static_cast<uint8_t>(ED_NOCLASS);
    __syncthreads();
    unsigned long long (*peq)[32] = sh.peq[warp][0];                          // the striped path works in buffer 0
    uint32_t* peq32 = reinterpret_cast<uint32_t*>(&sh.peq[warp][0][0][0]);    // [class][lane][half]
    uint8_t* tcls = sh.tcls[warp];
    uint8_t* th = sh.th[warp];

    // Two sweeps over the job list: pairs with a string longer than one 2048-row stripe first (they are the long poles),
    // then the short ones.
    // Sweep 0 is the critical path of the launch (a 10,000-row pair is one warp's dependency chain for most of a
    // millisecond), so at most ONE warp per SM sub-partition works on it: the first warp that claims the token of its
    // (SM, scheduler) pair.  The others start with the short jobs right away.
    // (CTA level: both warps of a worker CTA take the SAME long job, one half each, see below)
    // Only the workers run tables of several stripes, so only they own a slice of the parked-delta pool (slot = order of
    // winning a token; at most two per SM).
    if (threadIdx.x == 0) {
        uint32_t smid, warpid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
        const bool won = atomicExch(sm_tokens + ((smid & 1023u) * 4u + ((warpid >> 1) & 1u)), 1u) == 0u;
        sh.cta_worker = won ? 1u + atomicAdd(next_job + 4, 1u) : 0u;
    }
    __syncthreads();
    const bool long_worker = sh.cta_worker != 0u;
    uint8_t* hbuf = long_worker ? hbuf_pool + (static_cast<uint64_t>(sh.cta_worker - 1u) * ED_WARPS + warp) * hbuf_stride : nullptr;
    // Long pairs are handed out longest first (three size classes of the longer string, one pass over the job list each):
    // with 2 worker CTAs per SM the makespan is max(longest pair, total / workers) only if the big ones start early.
    for (int pass = long_worker ? 0 : 3; pass < 4; ++pass) {
        const int sweep = pass < 3 ? 0 : 1;
        while (true) {
            uint32_t job_id = 0;
            if (sweep == 0) {              // the whole CTA takes one long job (two barriers per round, both warps always)
                if (threadIdx.x == 0) sh.cta_job = atomicAdd(next_job + pass, 1u);
                __syncthreads();
                job_id = sh.cta_job;
                __syncthreads();
            } else {
                if (lane == 0) job_id = atomicAdd(next_job + 3, 1u);
                job_id = __shfl_sync(FULL, job_id, 0);
            }
            if (job_id >= n_jobs) break;
            const EditJob job = jobs[job_list[job_id]];
            const uint32_t la0 = hap_length(job.a), lb0 = hap_length(job.b);
            const uint32_t longer = max(la0, lb0);      // classes: >= 8192, >= 4096, > 2048 (one stripe), the rest
            if ((longer >= 8192u ? 0 : longer >= 4096u ? 1 : longer > 2048u ? 2 : 3) != pass) continue;
            const long long clock_begin = profile ? clock64() : 0ll;
            uint32_t steps_total = 0;
            // The edit distance does not change when the common prefix and suffix are removed; the two haplotypes of
            // a shared variant are mostly identical, so this alone finishes most pairs (warp-parallel compare).
            const uint32_t lim = min(la0, lb0);
            uint32_t pre, suf;
            if (sweep == 0) {
                // The CTA's two warps scan from the two ends at the same time, each at most half of the shorter string (an
                // identical pair, the common case, is settled in half the time).  If one side ran into its half-way mark
                // and the other did not, the first continues into the other half (both warps do, to stay in step).
                const uint32_t half = (lim + 1u) / 2u;
                const uint32_t r = warp == 0 ? common_run<false>(job.a, job.b, la0, lb0, half, ref, seq4_a, seq4_b, sh.cls2, lane)
                                             : common_run<true>(job.a, job.b, la0, lb0, lim - half, ref, seq4_a, seq4_b, sh.cls2, lane);
                if (lane == 0) sh.trim[warp] = r;
                __syncthreads();           // (the next write to trim[] comes after the next round's claim barriers)
                pre = sh.trim[0];
                suf = sh.trim[1];
                if (pre == half && suf < lim - half)
                    pre = common_run<false>(job.a, job.b, la0, lb0, lim - suf, ref, seq4_a, seq4_b, sh.cls2, lane, half);
                else if (suf == lim - half && pre < half)
                    suf = common_run<true>(job.a, job.b, la0, lb0, lim - pre, ref, seq4_a, seq4_b, sh.cls2, lane, lim - half);
            } else {
                pre = common_run<false>(job.a, job.b, la0, lb0, lim, ref, seq4_a, seq4_b, sh.cls2, lane);
                suf = common_run<true>(job.a, job.b, la0, lb0, lim - pre, ref, seq4_a, seq4_b, sh.cls2, lane);
            }
            const uint32_t la = la0 - pre - suf, lb = lb0 - pre - suf;
            const bool a_is_pattern = la <= lb;                // pattern = the shorter string
            const uint32_t m = a_is_pattern ? la : lb, n = a_is_pattern ? lb : la;
            const HapDesc P = a_is_pattern ? job.a : job.b;
            const HapDesc T = a_is_pattern ? job.b : job.a;
            long long dist = n;
            bool done = m == 0u;
            int first_attempt = 0;
            // Long patterns: one sliding-window pass over the Ukkonen band settles every pair whose distance is within
            // the band in about n steps: first with 32-row blocks (band of about 500 either side, half the work per
            // step), then with 64-row blocks (about 1000).  What is left goes to the striped band attempts.
            unsigned long long known_above = 0;           // the distance is known to exceed this
            bool col_split = false;
            if (sweep == 0) {
                // Both warps of the CTA on this pair: D[m][n] = min_j F(j) + B(n - j) with F from the top half of the rows
                // (warp 0) and B from the bottom half of the REVERSED strings (warp 1); each half needs about n / 2 + K
                // steps and they run at the same time.  Exact whenever the minimum is inside the band.
                const uint32_t padk = (32u - (m & 31u)) & 31u, mp = m + padk, np_ = n + padk;
                const uint32_t kw = win_kmax(mp, np_, 32u);
                const bool split = !done && kw >= 128u && mp >= 2048u;
                if (split) {
                    const WinSplit sp = win_split(mp, np_, kw, 32u);
                    const uint32_t rows = warp == 0 ? sp.rows_f : sp.rows_b, cols = warp == 0 ? sp.cols_f : sp.cols_b;
                    const long long base = window_core32<true>(P, T, pre, m, n, kw, ref, seq4_a, seq4_b, sh.cls2, &sh.peq[warp][0][0][0],
                                                               tcls, lane, rows, cols, warp != 0, sh.rec[warp]);
                    steps_total += cols + (rows - 1u) / 32u;
                    // values along the meeting row, F(j) for j = jlo .. jhi of this half's last block: inclusive scan of the
                    // recorded deltas, kept in this warp's (now idle) mask buffer
                    const WinGeom gh = win_geom_half(rows, cols, np_ - mp, kw, 32u);
                    const uint32_t jlo = win_jlo(gh, gh.last_block), len = win_jhi(gh, gh.last_block) - jlo;
                    int* val = reinterpret_cast<int*>(&sh.peq[warp][0][0][0]);
                    __syncwarp();
                    int running = static_cast<int>(base);
                    if (lane == 0) val[0] = running;
                    for (uint32_t c0 = 0; c0 < len; c0 += 32u) {
                        const uint32_t c = c0 + lane;
                        const uint32_t code = c < len ? sh.rec[warp][c] : 0u;
                        int d = static_cast<int>(code & 1u) - static_cast<int>(code >> 1);
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int up = __shfl_up_sync(FULL, d, o);
                            if (static_cast<int>(lane) >= o) d += up;
                        }
                        if (c < len) val[c + 1u] = running + d;
                        running += __shfl_sync(FULL, d, 31);
                    }
                    if (lane == 0) sh.half_base[warp] = (static_cast<long long>(jlo) << 32) | len;     // geometry for the other warp
                    __syncthreads();
                    if (warp == 0) {
                        const int* vf = val;
                        const int* vb = reinterpret_cast<const int*>(&sh.peq[1][0][0][0]);
                        const uint32_t jlo_b = static_cast<uint32_t>(sh.half_base[1] >> 32), len_b = static_cast<uint32_t>(sh.half_base[1]);
                        int best = 0x7fffffff;
                        for (uint32_t i = lane; i <= len; i += 32u) {
                            const long long j = static_cast<long long>(jlo) + i;
                            const long long ib = static_cast<long long>(np_) - j - static_cast<long long>(jlo_b);
                            if (ib >= 0 && ib <= static_cast<long long>(len_b)) best = min(best, vf[i] + vb[ib]);
                        }
                        best = __reduce_min_sync(FULL, best);
                        if (lane == 0) sh.split_result = best == 0x7fffffff ? -1ll : static_cast<long long>(best);
                    }
                    __syncthreads();
                    const long long d = sh.split_result;
                    if (d >= 0 && d <= static_cast<long long>(kw)) {
                        dist = d;
                        done = true;
                    } else {
                        known_above = kw;
                    }
                }
                // A short pattern against a long text (two unrelated insertions side by side: n - m far beyond any band) needs
                // the full table, n steps for one warp.  Two warps split the TEXT: every alignment path crosses the line
                // between text columns nh - 1 | nh at some row i, so D[m][n] = min_i Df[i][nh] + Db[m - i][n - nh] with Df the
                // table of P against T[0:nh] and Db that of the reversed strings; the column values are the vertical deltas
                // left in the lanes' registers after the last column.  n / 2 steps each, at the same time.
                col_split = !done && m != 0u && m <= 2048u && n >= 4096u &&
                            static_cast<unsigned long long>(n) / 2u + 64u < static_cast<unsigned long long>((n + 2047u) / 2048u) * (m + 31u);
                if (warp != 0 && !col_split) continue;       // what is left of this pair (fallbacks, the result) is warp 0's
            }
            if (!done && m > 1024u && known_above == 0) {
                const uint32_t kw = win_kmax(m, n, 32u);
                if (kw >= 128u) {
                    const long long d = window_pass32(P, T, pre, m, n, kw, ref, seq4_a, seq4_b, sh.cls2, &sh.peq[warp][0][0][0], tcls, lane);
                    steps_total += n + (m - 1u) / 32u;
                    if (d <= static_cast<long long>(kw)) {
                        dist = d;
                        done = true;
                    } else {
                        known_above = kw;
                    }
                }
            }
            if (!done && m > 2048u) {
                const uint32_t kw = win_kmax(m, n, 64u);
                // (after a failed 32-row pass the wider window rarely succeeds: only worth it against an expensive full table)
                const unsigned long long full_cost = static_cast<unsigned long long>((m + 2047u) / 2048u) * (n + 31u);
                if (kw >= 128u && kw > known_above && (known_above == 0 || full_cost > 3ull * n)) {
                    const long long d = window_pass<64>(P, T, pre, m, n, kw, ref, seq4_a, seq4_b, sh.cls2, &sh.peq[warp][0][0][0], tcls, lane);
                    steps_total += n + (m - 1u) / 64u;
                    if (d <= static_cast<long long>(kw)) {
                        dist = d;
                        done = true;
                    } else {
                        known_above = kw;
                    }
                }
            }
            if (!done && m > 1024u) {
                // beyond the windows: these are mostly unrelated alleles, so the stripes start four times wider
                while ((256ull << (2 * first_attempt)) < 4ull * known_above) ++first_attempt;
                // the maskless steady state of the striped path reads ring slots no refill has written yet
                __syncwarp();
                for (uint32_t i = lane; i < 128u; i += 32u) tcls[i] = static_cast<uint8_t>(ED_NOCLASS);
                __syncwarp();
            }
            // The stripes may take either string as the pattern: a full table costs ceil(rows / 2048) x (columns + 31) steps,
            // and for a short pattern against a long text (two unrelated insertions next to each other) the transposed table
            // is several times cheaper.  The band attempts need n >= m, so the swap implies the full table.
            uint32_t m_s = m, n_s = n;
            HapDesc P_s = P, T_s = T;
            bool force_full = false;
            const bool rev_s = col_split && warp != 0;                   // this warp reads both strings from the far end
            const uint32_t n_view = !col_split ? 0u : (warp == 0 ? n / 2u : n - n / 2u);     // text columns of this warp's half
            if (!done && !col_split) {
                // the distance is at least n - m: attempts with a narrower band cannot succeed
                while ((256ull << (2 * first_attempt)) < static_cast<unsigned long long>(n - m)) ++first_attempt;
                const unsigned long long K0 = 256ull << (2 * first_attempt);
                if (m <= 2048u || K0 >= m) {
                    const unsigned long long as_is = static_cast<unsigned long long>((m + 2047u) / 2048u) * (n + 31u);
                    const unsigned long long swapped = static_cast<unsigned long long>((n + 2047u) / 2048u) * (m + 31u);
                    if (swapped < as_is) {
                        m_s = n; n_s = m;
                        P_s = T; T_s = P;
                        force_full = true;
                    }
                }
            }
            if (!done && m_s > 2048u && static_cast<uint64_t>(n_s) > hbuf_stride) {
                // several stripes park one byte per text column: this pair needs a longer slice than the pool has.  It stays
                // unknown (negative), the longest such text is reported and the host repeats the call with a larger stride.
                if (lane == 0) {
                    atomicMax(need, static_cast<unsigned long long>(n_s));
                    out[job.out_index] = -2.0;
                }
                continue;
            }
            if (!done) {
                // Ukkonen cut-off: an alignment of cost d stays on the diagonals [-d, (n - m) + d], so a band of
                // half-width K gives the exact distance whenever the result is <= K; otherwise widen (x4) and repeat.
                // Single-stripe patterns are computed in full at once.
                for (int attempt = first_attempt;; ++attempt) {
                    unsigned long long K = 256ull << (2 * attempt);
                    const bool full_table = force_full || m_s <= 2048u || K >= m_s;
                    if (full_table) K = ~0ull >> 1;
                    long long anchor = 0;                      // D'[last row of the previous stripe, jlo - 1]
                    uint32_t prev_jhi = 0;
                    for (uint32_t row0 = 0; row0 < m_s; row0 += 2048u) {
                        const uint32_t rows = min(2048u, m_s - row0);
                        const bool final_stripe = row0 + rows == m_s;
                        const uint32_t nblk = (rows + 63u) / 64u, last_lane = nblk - 1u;
                        const uint32_t jlo = static_cast<unsigned long long>(row0) > K ? static_cast<uint32_t>(row0 - K) : 0u;
                        const uint32_t jhi = static_cast<uint32_t>(min(static_cast<unsigned long long>(col_split ? n_view : n_s),
                                                                       static_cast<unsigned long long>(row0) + rows + (n_s - m_s) + K));
                        // first column of the NEXT stripe's band: the running anchor stops there
                        const uint32_t next_row0 = row0 + rows;
                        const uint32_t next_jlo = static_cast<unsigned long long>(next_row0) > K ? static_cast<uint32_t>(next_row0 - K) : 0u;
                        const int width = static_cast<int>(jhi - jlo);
                        const int steps = width + static_cast<int>(nblk) - 1;
                        const int anchor_rel = static_cast<int>(min(next_jlo, jhi) - jlo);

                        // ---- match masks: 32 rows per round, the lanes holding the same class form one 32-bit mask
                        __syncwarp();
                        for (uint32_t i = lane; i < (ED_NCLASS + 1) * 64u; i += 32u) peq32[i] = 0u;
                        __syncwarp();
                        const uint32_t ngroups = (rows + 31u) / 32u;
                        for (uint32_t g0 = 0; g0 < ngroups; g0 += 4u) {
                            uint32_t byte[4], mode[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t idx = (g0 + k) * 32u + lane;
                                mode[k] = TOK_NONE;
                                byte[k] = 0u;
                                if (idx < rows) byte[k] = hap_fetch(P_s, pre + (rev_s ? m_s - 1u - (row0 + idx) : row0 + idx), ref, seq4_a, seq4_b, mode[k]);
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t cls = mode[k] == TOK_NONE ? ED_NOCLASS + 1u : tok_class(sh.cls2, byte[k], mode[k]);
                                const uint32_t peers = __match_any_sync(FULL, cls);
                                if (cls < ED_NOCLASS && static_cast<uint32_t>(__ffs(peers) - 1) == lane)
                                    peq32[cls * 64u + (g0 + k)] = peers;            // word (g0+k): block (g0+k)/2, half (g0+k)&1
                            }
                        }

                        // ---- column ring: classes and top-row delta codes, 32 columns per refill
                        uint32_t nxt_byte = 0, nxt_mode = TOK_NONE, nxt_h = 1;
                        auto fetch = [&](int rel) {                                 // issue the loads of band column `rel`
                            nxt_mode = TOK_NONE;
                            nxt_byte = 0u;
                            nxt_h = 1u;                                             // outside the previous band: +1
                            if (rel < width) {
                                const uint32_t j = jlo + static_cast<uint32_t>(rel);
                                nxt_byte = hap_fetch(T_s, pre + (rev_s ? n_s - 1u - j : j), ref, seq4_a, seq4_b, nxt_mode);
                                if (row0 != 0u && j < prev_jhi) nxt_h = __ldcg(hbuf + j);
                            }
                        };
                        auto insert = [&](int first_rel) {                          // decode the pending loads into the ring
                            const uint32_t slot = static_cast<uint32_t>(first_rel + static_cast<int>(lane)) & 63u;
                            tcls[slot] = static_cast<uint8_t>(nxt_mode == TOK_NONE ? ED_NOCLASS : tok_class(sh.cls2, nxt_byte, nxt_mode));
                            th[slot] = static_cast<uint8_t>(nxt_h);
                        };
                        fetch(static_cast<int>(lane));
                        insert(0);
                        fetch(32 + static_cast<int>(lane));
                        __syncwarp();

                        uint64_t pv = ~0ull, mv = 0ull;
                        const uint32_t hshift = (final_stripe && lane == last_lane) ? ((rows - 1u) & 63u) : 63u;
                        const bool park = !final_stripe && lane == last_lane;
                        int sum_all = 0, sum_anchor = 0;        // bottom-row deltas over [jlo, jhi) and over [jlo, next_jlo)
                        const int ilane = static_cast<int>(lane);
                        auto ring_class = [&](int rel) -> uint32_t {
                            return (rel >= 0 && rel < width) ? tcls[static_cast<uint32_t>(rel) & 63u] : ED_NOCLASS;
                        };
                        // pipeline registers: eq0 = match mask of step t, cls1 = class of step t+1, th0 = top code of step t
                        uint64_t eq0 = peq[ring_class(-ilane)][lane];
                        uint32_t cls1 = ring_class(1 - ilane);
                        uint32_t th0 = th[0];
                        uint32_t hout = 0;

                        auto run = [&](auto masked_tag, int t, const int t_end) {
                            constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll 2
                            for (; t < t_end; ++t) {
                                uint32_t hin = __shfl_up_sync(FULL, hout, 1);
                                hin = lane == 0u ? th0 : hin;
                                const int rel = t - ilane;
                                // loads for the next steps (independent of the recurrence)
                                const uint64_t eq1 = peq[cls1][lane];
                                uint32_t cls2;
                                if (MASKED) cls2 = ring_class(rel + 2);
                                else cls2 = tcls[static_cast<uint32_t>(rel + 2) & 63u];
                                const uint32_t th1 = th[static_cast<uint32_t>(t + 1) & 63u];
                                uint64_t npv = pv, nmv = mv;
                                uint32_t ho = myers_step(npv, nmv, eq0, hin, hshift);
                                if (MASKED) {
                                    const bool active = lane < nblk && rel >= 0 && rel < width;
                                    pv = active ? npv : pv;
                                    mv = active ? nmv : mv;
                                    ho = active ? ho : 0u;
                                    if (park && active) hbuf[jlo + static_cast<uint32_t>(rel)] = static_cast<uint8_t>(ho);
                                } else {
                                    pv = npv;
                                    mv = nmv;
                                    if (park) hbuf[jlo + static_cast<uint32_t>(rel)] = static_cast<uint8_t>(ho);
                                }
                                if (rel == anchor_rel) sum_anchor = sum_all;
                                sum_all += static_cast<int>(ho & 1u) - static_cast<int>(ho >> 1);
                                hout = ho;
                                eq0 = eq1;
                                cls1 = cls2;
                                th0 = th1;
                            }
                        };
                        // The ring is refilled before the steps t = 30 (mod 32): columns t+2 .. t+33 enter, t+34 .. t+65 are
                        // requested.  Chunks that lie inside [nblk-1, width) need no masks: every block lane is on a column.
                        int t = 0, next_refill = 30;
                        steps_total += static_cast<uint32_t>(steps);
                        while (t < steps) {
                            if (t == next_refill) {
                                insert(t + 2);
                                fetch(t + 34 + ilane);
                                __syncwarp();
                                next_refill += 32;
                            }
                            const int t_end = min(steps, next_refill);
                            if (t >= static_cast<int>(nblk) - 1 && t_end <= width) run(std::false_type(), t, t_end);
                            else run(std::true_type(), t, t_end);
                            t = t_end;
                        }
                        __syncwarp();
                        if (col_split) {       // D[i][n_view] for i = 0 .. m_s (one stripe): prefix sums of the vertical deltas
                            int* val = reinterpret_cast<int*>(&sh.peq[warp][1][0][0]);      // the second mask buffer is idle here
                            const uint32_t valid = lane < nblk ? min(64u, rows - 64u * lane) : 0u;
                            const uint64_t vmask = valid == 64u ? ~0ull : ((1ull << valid) - 1ull);
                            int incl = __popcll(pv & vmask) - __popcll(mv & vmask);
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int up = __shfl_up_sync(FULL, incl, o);
                                if (static_cast<int>(lane) >= o) incl += up;
                            }
                            const int excl = incl - (__popcll(pv & vmask) - __popcll(mv & vmask));
                            if (lane == 0) val[0] = static_cast<int>(n_view);
                            for (uint32_t r = 0; r < valid; ++r) {
                                const uint64_t bits = r == 63u ? ~0ull : ((2ull << r) - 1ull);
                                val[64u * lane + r + 1u] = static_cast<int>(n_view) + excl + __popcll(pv & bits) - __popcll(mv & bits);
                            }
                        }
                        if (anchor_rel >= width) sum_anchor = sum_all;
                        sum_all = __shfl_sync(FULL, sum_all, last_lane);
                        sum_anchor = __shfl_sync(FULL, sum_anchor, last_lane);
                        if (final_stripe) dist = anchor + rows + sum_all;
                        else anchor += static_cast<long long>(rows) + sum_anchor;
                        prev_jhi = jhi;
                    }
                    if (full_table || static_cast<unsigned long long>(dist) <= K) break;      // exact
                }
            }
            if (col_split) {
                __syncthreads();
                if (warp == 0) {
                    const int* vf = reinterpret_cast<const int*>(&sh.peq[0][1][0][0]);
                    const int* vb = reinterpret_cast<const int*>(&sh.peq[1][1][0][0]);
                    int best = 0x7fffffff;
                    for (uint32_t i = lane; i <= m_s; i += 32u) best = min(best, vf[i] + vb[m_s - i]);
                    dist = __reduce_min_sync(FULL, best);
                }
                __syncthreads();
                if (warp != 0) continue;
            }
            if (lane == 0) {
                out[job.out_index] = static_cast<double>(dist);
                out_hi[job.out_index] = static_cast<double>(dist);
            }
            if (profile && lane == 0)      // SVB_ED_PROFILE: {trimmed pattern rows, trimmed text columns, column steps, cycles}
                profile[job_id] = make_uint4(m, n, steps_total, static_cast<uint32_t>(clock64() - clock_begin));
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

int build_class_map_from_seen(const bool* seen, uint8_t* map256, std::string* why) {
    // classes 0..15: the nt16 letters a 4-bit query base can decode to; further classes: any other byte of
    // the (upper-cased) reference.  Equality of bytes == equality of classes, so distances stay exact.
    memset(map256, 255, 256);
    const char* nt16 = "=ACMGRSVTWYHKDBN";
    int next = 0;
    for (int i = 0; i < 16; ++i) map256[static_cast<uint8_t>(nt16[i])] = static_cast<uint8_t>(next++);
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

int build_class_map(const uint8_t* bases, uint64_t n, uint8_t* map256, std::string* why) {
    bool seen[256] = {false};
    for (uint64_t i = 0; i < n; ++i) seen[bases[i]] = true;
    return build_class_map_from_seen(seen, map256, why);
}

int launch_edit_distance(svb_ctx* ctx, const EditJob* d_jobs, const uint32_t* d_list, const unsigned long long* d_n_list,
                         uint32_t max_jobs, const uint8_t* d_ref, const uint8_t* d_seq4_a, const uint8_t* d_seq4_b,
                         const uint8_t* d_class_map, double* d_dist, double* d_dist_hi, unsigned long long* d_need) {
    if (!max_jobs) return SVB_OK;
    // persistent grid: two worker CTAs per SM for the long pairs plus two for the short ones; every CTA leaves at once when
    // the list is empty (the count is read on the device: the host does not wait for it)
    const unsigned blocks = static_cast<unsigned>(std::min<uint64_t>((static_cast<uint64_t>(max_jobs) + ED_WARPS - 1) / ED_WARPS,
                                                                     static_cast<uint64_t>(ctx->sm_count) * 4));
    // parked bottom-row deltas of multi-stripe tables, one byte per text column, one slice per worker warp
    const uint64_t stride = (std::max<uint64_t>(ctx->ed_stride, 2048) + 127) & ~127ull;
    const uint64_t workers = std::min<uint64_t>(blocks, static_cast<uint64_t>(ctx->sm_count) * 2);
    uint8_t* hbuf = nullptr;
    uint4* d_profile = nullptr;
    unsigned int* sm_tokens = nullptr;                 // one word per (SM, scheduler): who runs the long jobs; + counters
    auto release = [&]() {
        if (hbuf) cudaFreeAsync(hbuf, ctx->stream);
        if (d_profile) cudaFreeAsync(d_profile, ctx->stream);
        if (sm_tokens) cudaFreeAsync(sm_tokens, ctx->stream);
    };
#define ED_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            release();                                                             \
            return svb_fail(ctx, SVB_ERR_CUDA, #call, e__);                        \
        }                                                                          \
    } while (0)
    ED_CUDA(cudaMallocAsync(&hbuf, stride * workers * ED_WARPS, ctx->stream));
    // profiling aid: SVB_ED_PROFILE=<file> dumps one uint4 per listed job {rows, columns, column steps, SM cycles}
    const char* profile_path = getenv("SVB_ED_PROFILE");
    if (profile_path) {
        ED_CUDA(cudaMallocAsync(&d_profile, sizeof(uint4) * max_jobs, ctx->stream));
        ED_CUDA(cudaMemsetAsync(d_profile, 0, sizeof(uint4) * max_jobs, ctx->stream));
    }
    ED_CUDA(cudaMallocAsync(&sm_tokens, sizeof(unsigned int) * (4096 + 8), ctx->stream));      // + four pass counters, the worker counter
    ED_CUDA(cudaMemsetAsync(sm_tokens, 0, sizeof(unsigned int) * (4096 + 8), ctx->stream));
    {
        KernelTimer timer(ctx, SVB_K_EDIT_DISTANCE);
        edit_distance_kernel<<<blocks, ED_WARPS * 32, 0, ctx->stream>>>(d_jobs, d_list, d_n_list, sm_tokens + 4096,
                                                                       d_ref, d_seq4_a, d_seq4_b, d_class_map, hbuf, stride, d_dist, d_dist_hi,
                                                                       d_profile, sm_tokens, d_need);
        ctx->launches += 1;
    }
    ED_CUDA(cudaGetLastError());
    if (d_profile) {
        std::vector<uint4> h(max_jobs);
        ED_CUDA(cudaMemcpyAsync(h.data(), d_profile, sizeof(uint4) * max_jobs, cudaMemcpyDeviceToHost, ctx->stream));
        ED_CUDA(cudaStreamSynchronize(ctx->stream));
        if (FILE* f = fopen(profile_path, "wb")) {
            fwrite(h.data(), sizeof(uint4), max_jobs, f);
            fclose(f);
        }
    }
#undef ED_CUDA
    release();
    return SVB_OK;
}

// jobs whose wavefront pass did not settle them (dist < 0), or all jobs (bounded == 0) -> list for the exact kernel
namespace {
__global__ void list_jobs_kernel(const double* __restrict__ dist, uint32_t n, int only_unknown, uint32_t* __restrict__ list,
                                 unsigned long long* __restrict__ count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!only_unknown || dist[i] < 0.0) list[atomicAdd(count, 1ull)] = i;
}
}  // namespace

// max_distance < 0: exact distances (edlib.align(a, b)["editDistance"]).  max_distance >= 0: edlib's k parameter -- the
// distance if it is <= max_distance, else -1 -- through the wavefront kernel (wfa.cu), the way svb_pair computes them.
int run_edit_distance_strings(svb_ctx* ctx, const uint8_t* a, const uint64_t* a_off, const uint8_t* b, const uint64_t* b_off,
                              uint32_t n_pairs, int64_t max_distance, int64_t* out) {
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
    const bool bounded = max_distance >= 0 && max_distance <= static_cast<int64_t>(WFA_MAX_T);
    unsigned char* slab = nullptr;
    auto align = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
    const size_t sz_bytes = align(std::max<uint64_t>(both.size(), 1)), sz_off = align(sizeof(uint64_t) * (n_pairs + 1));
    const size_t sz_jobs = align(sizeof(EditJob) * n_pairs), sz_d = align(sizeof(double) * n_pairs), sz_list = align(sizeof(uint32_t) * n_pairs);
    const size_t sz_big = align(sizeof(uint4) * n_pairs), sz_cnt = 256;
    SVB_CUDA(ctx, cudaMallocAsync(&slab, sz_bytes + 256 + 2 * sz_off + sz_jobs + 2 * sz_d + sz_list + sz_big + sz_cnt, ctx->stream));
    unsigned char* cur = slab;
    auto carve = [&](size_t bytes) { unsigned char* q = cur; cur += bytes; return q; };
    uint8_t* d_bytes = carve(sz_bytes);
    uint8_t* d_map = carve(256);
    uint64_t* d_aoff = reinterpret_cast<uint64_t*>(carve(sz_off));
    uint64_t* d_boff = reinterpret_cast<uint64_t*>(carve(sz_off));
    EditJob* d_jobs = reinterpret_cast<EditJob*>(carve(sz_jobs));
    double* d_dist = reinterpret_cast<double*>(carve(sz_d));
    double* d_hi = reinterpret_cast<double*>(carve(sz_d));
    uint32_t* d_list = reinterpret_cast<uint32_t*>(carve(sz_list));
    uint4* d_big = reinterpret_cast<uint4*>(carve(sz_big));
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(carve(sz_cnt));     // [0] n jobs, [1] list size, [2] need, [4..] wfa counters
    auto fail = [&](int rc) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeAsync(slab, ctx->stream);
        return rc;
    };
#define STR_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) return fail(svb_fail(ctx, SVB_ERR_CUDA, #call, e__));    \
    } while (0)
    std::vector<double> h(n_pairs), h_hi(n_pairs);
    for (int attempt = 0; attempt < 2; ++attempt) {
        unsigned long long h_cnt[32] = {0};
        h_cnt[0] = n_pairs;
        STR_CUDA(cudaMemcpyAsync(d_cnt, h_cnt, sz_cnt, cudaMemcpyHostToDevice, ctx->stream));
        if (!both.empty()) STR_CUDA(cudaMemcpyAsync(d_bytes, both.data(), both.size(), cudaMemcpyHostToDevice, ctx->stream));
        STR_CUDA(cudaMemcpyAsync(d_map, map, 256, cudaMemcpyHostToDevice, ctx->stream));
        STR_CUDA(cudaMemcpyAsync(d_aoff, a_off, sizeof(uint64_t) * (n_pairs + 1), cudaMemcpyHostToDevice, ctx->stream));
        STR_CUDA(cudaMemcpyAsync(d_boff, b_off, sizeof(uint64_t) * (n_pairs + 1), cudaMemcpyHostToDevice, ctx->stream));
        make_string_jobs<<<(n_pairs + 127) / 128, 128, 0, ctx->stream>>>(d_aoff, d_boff, na, n_pairs, d_jobs);
        ctx->launches += 1;
        if (bounded) {
            WfaArgs w;
            w.jobs = d_jobs; w.n_jobs_dev = d_cnt; w.counters = reinterpret_cast<unsigned int*>(d_cnt + 4);
            w.big = d_big; w.big_cap = n_pairs; w.ref = d_bytes; w.seq4_a = nullptr; w.seq4_b = nullptr; w.class_map = d_map;
            w.dist = d_dist; w.dist_hi = d_hi; w.t = static_cast<uint32_t>(max_distance); w.cap_chars = 0; w.stage = 0;
            int rc = launch_wfa(ctx, w);
            if (rc != SVB_OK) return fail(rc);
        }
        list_jobs_kernel<<<(n_pairs + 127) / 128, 128, 0, ctx->stream>>>(d_dist, n_pairs, bounded ? 1 : 0, d_list, d_cnt + 1);
        ctx->launches += 1;
        STR_CUDA(cudaGetLastError());
        int rc = launch_edit_distance(ctx, d_jobs, d_list, d_cnt + 1, n_pairs, d_bytes, nullptr, nullptr, d_map, d_dist, d_hi, d_cnt + 2);
        if (rc != SVB_OK) return fail(rc);
        STR_CUDA(cudaMemcpyAsync(h.data(), d_dist, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
        STR_CUDA(cudaMemcpyAsync(h_hi.data(), d_hi, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
        STR_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        STR_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h_cnt[2] == 0) break;
        if (attempt == 1) return fail(svb_fail(ctx, SVB_ERR_CAPACITY, "svb_edit_distance: parked-delta stride"));
        ctx->ed_stride = h_cnt[2] + 4096;               // a multi-stripe table with a longer text than the slices: once more
    }
#undef STR_CUDA
    for (uint32_t i = 0; i < n_pairs; ++i) {
        const bool exact = h[i] == h_hi[i];
        if (max_distance < 0) out[i] = static_cast<int64_t>(h[i]);
        else out[i] = (exact && h[i] <= static_cast<double>(max_distance)) ? static_cast<int64_t>(h[i]) : -1;
    }
    cudaFreeAsync(slab, ctx->stream);
    return SVB_OK;
}
