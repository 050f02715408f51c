// Device helpers shared by the edit-distance kernels (edit_distance.cu: exact bit-parallel tables; wfa.cu: thresholded
// wavefronts): symbol classes of the virtual haplotype strings and the warp-parallel common prefix / suffix scan.
#pragma once
#include "common.cuh"
#include "edit_core.cuh"
#include "pairing.cuh"

namespace edstr {

constexpr uint32_t ED_NOCLASS = ED_NCLASS;      // "matches nothing" (extra all-zero row of the match-mask tables)
constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t tok_class(const uint8_t* cls2, uint32_t byte, uint32_t mode) {
    uint32_t c;
    if (mode >= TOK_NIB_HI) c = (mode == TOK_NIB_HI) ? (byte >> 4) : (byte & 15u);      // class i == nt16 code i
    else c = cls2[mode * 256u + byte];
    return min(c, ED_NOCLASS);
}

// A run of `count` consecutive positions starting at `first` (ascending) that lies inside ONE linearly stored piece of the
// virtual string: the left / right reference slice, a reference middle (DUP_INT) or the 4-bit inserted bases.  Returns
// false when the run touches a piece boundary or a reverse-complemented / repeated middle (the caller then takes the
// general per-position path).  On success position first + k is byte `base[(nib0 + k) >> shift]`, nibble parity
// (nib0 + k) & 1 when shift == 1.
struct LinearRun {
    const uint8_t* base;
    uint64_t nib0;          // shift == 1: nibble index of the first position; shift == 0: unused (0)
    uint32_t shift;         // 0: one byte per position (reference), 1: two positions per byte (4-bit query bases)
};
__device__ __forceinline__ bool linear_run(const HapDesc& d, uint32_t first, uint32_t count, const uint8_t* ref, const uint8_t* sa,
                                           const uint8_t* sb, LinearRun& out) {
    out.nib0 = 0;
    out.shift = 0;
    if (first + count <= d.l_len) {
        out.base = ref + d.l_base + first;
        return true;
    }
    if (first < d.l_len) return false;
    const uint32_t im = first - d.l_len;
    if (im + count <= d.m_len) {
        if (d.m_kind == HAP_MID_SEQ4) {
            out.base = d.seq_sel ? sb : sa;
            out.nib0 = d.m_base + im;
            out.shift = 1;
            return true;
        }
        if (d.m_kind == HAP_MID_REF) {
            out.base = ref + d.m_base + im;
            return true;
        }
        return false;
    }
    if (im < d.m_len) return false;
    out.base = ref + d.r_base + (im - d.m_len);
    return true;
}

// ---- word-wise rounds of the common prefix / suffix scan ---------------------------------------------------------------
// When both strings are inside a linearly stored piece of the same kind -- reference bytes against reference bytes, or 4-bit
// query bases against 4-bit query bases, which is what two haplotypes of a shared variant look like almost everywhere --
// symbols are equal exactly when their stored bits are (distinct bytes have distinct classes, class i is nt16 code i), so a
// lane compares one 32-bit word per load: 4 reference bases or 8 query bases, at any alignment of either string (two
// aligned loads and a funnel shift each).  WORDS_PER_LANE loads per string are in flight: one round covers 512 reference
// bases or 1024 query bases instead of 128.
constexpr int CR_WORDS = 4;

// 32 bits starting `bit` bits into the little-endian word stream at `w` (the second word is only read when it is needed,
// so nothing behind the last byte of a string is touched)
template <bool NIBBLES>
__device__ __forceinline__ uint32_t cr_load_bits(const uint32_t* w, uint32_t bit) {
    uint32_t lo = w[0], hi = bit ? w[1] : 0u;
    if (NIBBLES) {      // the first base of a byte is its HIGH nibble: swap so that the stream is little-endian nibble by nibble
        lo = ((lo & 0x0F0F0F0Fu) << 4) | ((lo >> 4) & 0x0F0F0F0Fu);
        hi = ((hi & 0x0F0F0F0Fu) << 4) | ((hi >> 4) & 0x0F0F0F0Fu);
    }
    return __funnelshift_r(lo, hi, bit);
}

// One round over `R = 32 * CR_WORDS * (NIBBLES ? 8 : 4)` symbols that both runs hold linearly (ascending symbol s of run r is
// byte r.base[s] / nibble r.nib0 + s).  Scan order is ascending (REVERSED = false) or descending.  Returns the number of
// symbols that agree before the first difference in scan order (R if none differ).
template <bool REVERSED, bool NIBBLES>
__device__ __forceinline__ uint32_t cr_word_round(const LinearRun& ra, const LinearRun& rb, uint32_t lane) {
    constexpr uint32_t SYM = NIBBLES ? 8u : 4u, R = 32u * CR_WORDS * SYM;
    uint32_t x[CR_WORDS];
#pragma unroll
    for (int k = 0; k < CR_WORDS; ++k) {
        const uint32_t j = 32u * k + lane;                                   // word number in scan order
        const uint32_t s0 = REVERSED ? R - SYM - SYM * j : SYM * j;           // its first symbol, ascending
        uint32_t va, vb;
        if (NIBBLES) {
            const uint64_t na = ra.nib0 + s0, nb = rb.nib0 + s0;
            va = cr_load_bits<true>(reinterpret_cast<const uint32_t*>(ra.base) + (na >> 3), static_cast<uint32_t>(na & 7ull) * 4u);
            vb = cr_load_bits<true>(reinterpret_cast<const uint32_t*>(rb.base) + (nb >> 3), static_cast<uint32_t>(nb & 7ull) * 4u);
        } else {
            const uintptr_t qa = reinterpret_cast<uintptr_t>(ra.base) + s0, qb = reinterpret_cast<uintptr_t>(rb.base) + s0;
            va = cr_load_bits<false>(reinterpret_cast<const uint32_t*>(qa & ~static_cast<uintptr_t>(3)), static_cast<uint32_t>(qa & 3u) * 8u);
            vb = cr_load_bits<false>(reinterpret_cast<const uint32_t*>(qb & ~static_cast<uintptr_t>(3)), static_cast<uint32_t>(qb & 3u) * 8u);
        }
        x[k] = va ^ vb;
    }
    uint32_t agree = R;
#pragma unroll
    for (int k = CR_WORDS - 1; k >= 0; --k) {
        const uint32_t mask = __ballot_sync(FULL, x[k] != 0u);
        if (mask) {
            const int f = __ffs(static_cast<int>(mask)) - 1;                 // first lane (scan order) whose word differs
            const uint32_t xf = __shfl_sync(FULL, x[k], f);
            const uint32_t inside = REVERSED ? static_cast<uint32_t>(__clz(static_cast<int>(xf))) / (32u / SYM)
                                             : (static_cast<uint32_t>(__ffs(static_cast<int>(xf))) - 1u) / (32u / SYM);
            agree = SYM * (32u * k + static_cast<uint32_t>(f)) + inside;
        }
    }
    return agree;
}

// Length of the common prefix (REVERSED = false) or suffix of the two strings, at most `lim`, given that the first
// `start` positions are known to agree; 128 positions per round so that four loads per string are in flight (eight per
// string were measured slower).  Rounds that stay inside one linearly stored piece of both strings (nearly all of them)
// address their bytes directly instead of walking the piece table per position.
template <bool REVERSED>
__device__ __forceinline__ uint32_t common_run(const HapDesc& A, const HapDesc& B, uint32_t la, uint32_t lb, uint32_t lim,
                                               const uint8_t* ref, const uint8_t* sa, const uint8_t* sb, const uint8_t* cls2,
                                               uint32_t lane, uint32_t start = 0u) {
    constexpr int U = 4;
    uint32_t run = start;
    bool done = false;
    while (run < lim && !done) {
        uint32_t ba[U], ma[U], bb[U], mb[U];
        LinearRun ra, rb;
        // word-wise rounds while both strings stay inside one linear piece of the same kind (see cr_word_round)
        {
            constexpr uint32_t RB = 32u * CR_WORDS * 4u, RN = 32u * CR_WORDS * 8u;
            if (run + RN <= lim && linear_run(A, REVERSED ? la - run - RN : run, RN, ref, sa, sb, ra) && ra.shift == 1u &&
                linear_run(B, REVERSED ? lb - run - RN : run, RN, ref, sa, sb, rb) && rb.shift == 1u) {
                const uint32_t agree = cr_word_round<REVERSED, true>(ra, rb, lane);
                run += agree;
                if (agree < RN) done = true;
                continue;
            }
            if (run + RB <= lim && linear_run(A, REVERSED ? la - run - RB : run, RB, ref, sa, sb, ra) && ra.shift == 0u &&
                linear_run(B, REVERSED ? lb - run - RB : run, RB, ref, sa, sb, rb) && rb.shift == 0u) {
                const uint32_t agree = cr_word_round<REVERSED, false>(ra, rb, lane);
                run += agree;
                if (agree < RB) done = true;
                continue;
            }
        }
        const bool whole = run + 32u * U <= lim;         // a full round: positions run .. run + 127 of both strings
        // ascending position of the round's LAST element when scanning backwards, of its first otherwise
        const uint32_t fa = REVERSED ? la - run - 32u * U : run, fb = REVERSED ? lb - run - 32u * U : run;
        if (whole && linear_run(A, fa, 32u * U, ref, sa, sb, ra) && linear_run(B, fb, 32u * U, ref, sa, sb, rb)) {
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const uint32_t off = REVERSED ? 32u * U - 1u - (32u * k + lane) : 32u * k + lane;     // offset inside the run
                const uint64_t na = ra.nib0 + off, nb = rb.nib0 + off;
                ba[k] = ra.shift ? ra.base[na >> 1] : ra.base[off];
                bb[k] = rb.shift ? rb.base[nb >> 1] : rb.base[off];
                ma[k] = ra.shift ? TOK_NIB_HI + static_cast<uint32_t>(na & 1ull) : TOK_REF;
                mb[k] = rb.shift ? TOK_NIB_HI + static_cast<uint32_t>(nb & 1ull) : TOK_REF;
            }
        } else {
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const uint32_t i = run + 32u * k + lane;
                ma[k] = mb[k] = TOK_NONE;
                ba[k] = bb[k] = 0u;
                if (i < lim) {
                    ba[k] = hap_fetch(A, REVERSED ? la - 1u - i : i, ref, sa, sb, ma[k]);
                    bb[k] = hap_fetch(B, REVERSED ? lb - 1u - i : i, ref, sa, sb, mb[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            bool same = false;
            if (ma[k] != TOK_NONE) {
                const uint32_t ca = tok_class(cls2, ba[k], ma[k]);
                same = ca < ED_NOCLASS && ca == tok_class(cls2, bb[k], mb[k]);
            }
            const uint32_t mask = __ballot_sync(FULL, same);
            if (!done && mask != FULL) {
                run += 32u * k + static_cast<uint32_t>(__ffs(~mask) - 1);
                done = true;
            }
        }
        if (!done) run += 32u * U;
    }
    return min(run, lim);
}

}  // namespace edstr
