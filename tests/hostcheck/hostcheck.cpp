// TEST INFRASTRUCTURE: compiles the __host__ __device__ cores of the kernels (linkage, segment walk)
// with g++ so that their logic is checked on a CPU-only box.  Never shipped, never loaded by the
// product: the product path is the CUDA library only.
#include <stdint.h>
#include <vector>
#include <algorithm>
#include "../../svim_asm_b200/csrc/linkage.cuh"
#include "../../svim_asm_b200/csrc/walk.cuh"

extern "C" int hc_cluster_labels(const double* condensed, int n, double threshold, int* labels) {
    double work[LINK_MAXN * (LINK_MAXN - 1) / 2];
    for (int i = 0; i < n * (n - 1) / 2; ++i) work[i] = condensed[i];
    return link_complete_fcluster(n, work, threshold, labels);
}

// The interval run of the linkage (pair.cu resolve_kernel): 1 when the labels are the same for every choice of values inside
// [lo[i], hi[i]] (then written to `labels`), else 0.  Every non-degenerate interval is its own unknown (id = index).
extern "C" int hc_labels_determined(const double* lo, const double* hi, int n, double threshold, int* labels) {
    LinkInterval work[LINK_MAXN * (LINK_MAXN - 1) / 2];
    for (int i = 0; i < n * (n - 1) / 2; ++i) { work[i].lo = lo[i]; work[i].hi = hi[i]; work[i].id = static_cast<uint32_t>(i); }
    return link_labels_determined(n, work, threshold, labels) ? 1 : 0;
}

// segs: k rows of (q_start, q_end, tid, ref_start, ref_end, rev), primary first.  params: min_mapq,
// min_sv, max_sv, qgt, qot, rgt, rot.  Returns the number of rows, or -(error bits) on a reference abort.
extern "C" int hc_walk(const int32_t* segs, int k, int32_t read_len, uint32_t l_seq, const int32_t* params,
                       const int32_t* contig_len, const int32_t* lexrank, int32_t n_contig, uint32_t aln_idx,
                       uint32_t hap, svb_row* rows_out, int cap) {
    std::vector<WalkScratch> sc(static_cast<size_t>(k) + 1);
    for (int i = 0; i < k; ++i) {
        WalkSeg s;
        s.q_start = segs[6 * i]; s.q_end = segs[6 * i + 1]; s.tid = segs[6 * i + 2];
        s.ref_start = segs[6 * i + 3]; s.ref_end = segs[6 * i + 4]; s.rev = segs[6 * i + 5];
        sc[i].seg = s;
    }
    WalkParams p{params[0], params[1], params[2], params[3], params[4], params[5], params[6]};
    WalkRead rd{aln_idx, hap, read_len, l_seq, contig_len, lexrank, n_contig};
    std::vector<svb_row> rows(static_cast<size_t>(k) * k + 8);
    WalkOut o{rows.data(), static_cast<uint32_t>(rows.size()), 0, 0, 0};
    walk_read(rd, p, sc.data(), static_cast<uint32_t>(k), o);
    if (o.err) return -static_cast<int>(o.err);
    for (uint32_t i = 0; i < o.n && static_cast<int>(i) < cap; ++i) rows_out[i] = rows[i];
    return static_cast<int>(o.n);
}

#include "../../svim_asm_b200/csrc/edit_core.cuh"
#include <string.h>
// Sequential use of the 64-row block step (what the GPU pipelines across the lanes of a warp):
// pattern = a (rows), text = b (columns); stripes of `blocks_per_stripe` blocks with the horizontal
// deltas of the stripe's last row carried in hbuf, exactly like edit_distance.cu.
extern "C" long long hc_myers(const uint8_t* a, long long m, const uint8_t* b, long long n, int blocks_per_stripe) {
    if (m == 0) return n;
    if (n == 0) return m;
    std::vector<signed char> hbuf(static_cast<size_t>(n), 1);
    long long total = 0;
    const long long stripe_rows = 64ll * blocks_per_stripe;
    for (long long row0 = 0; row0 < m; row0 += stripe_rows) {
        const long long rows = (m - row0 < stripe_rows) ? (m - row0) : stripe_rows;
        const bool final_stripe = row0 + rows == m;
        const int nblk = static_cast<int>((rows + 63) / 64);
        std::vector<uint64_t> pv(nblk, ~0ull), mv(nblk, 0ull);
        std::vector<uint64_t> peq(static_cast<size_t>(nblk) * 256, 0ull);
        for (long long r = 0; r < rows; ++r) peq[static_cast<size_t>(r / 64) * 256 + a[row0 + r]] |= 1ull << (r & 63);
        for (long long j = 0; j < n; ++j) {
            int h = hbuf[static_cast<size_t>(j)];
            for (int k = 0; k < nblk; ++k) {
                const bool last = k == nblk - 1;
                const uint64_t hibit = (last && final_stripe) ? (1ull << ((rows - 1) & 63)) : (1ull << 63);
                h = myers_block(pv[k], mv[k], peq[static_cast<size_t>(k) * 256 + b[j]], h, hibit);
            }
            if (final_stripe) total += h; else hbuf[static_cast<size_t>(j)] = static_cast<signed char>(h);
        }
    }
    return m + total;
}

// The sliding-window pass of edit_distance.cu, emulated with the SAME step order: global step t, lane l = b % 32
// works on column t - b of block b, horizontal deltas reach the lane below one step later, lanes move on to
// block b + 32 when theirs has run out of columns.  bw = rows per block (64 or 32).  Returns the window's value
// (exact iff <= K), -1 if K is too wide for one warp.
extern "C" long long hc_myers_window(const uint8_t* a, long long m64, const uint8_t* b, long long n64, unsigned K, unsigned bw) {
    const WinGeom g = win_geom(static_cast<uint32_t>(m64), static_cast<uint32_t>(n64), K, bw);
    struct Lane { uint32_t blk; bool has; WinBlock w; uint64_t pv, mv; uint32_t hout; long long partial; };
    Lane L[32];
    auto masks = [&](uint32_t blk, uint8_t c) {
        uint64_t eq = 0;
        for (uint32_t r = 0; r < bw && static_cast<uint64_t>(bw) * blk + r < g.m; ++r) if (a[static_cast<uint64_t>(bw) * blk + r] == c) eq |= 1ull << r;
        return eq;
    };
    const uint64_t fresh = bw == 64u ? ~0ull : 0xFFFFFFFFull;
    for (uint32_t l = 0; l < 32u; ++l) {
        L[l].blk = l; L[l].has = l <= g.last_block; L[l].pv = fresh; L[l].mv = 0; L[l].hout = 0; L[l].partial = 0;
        if (L[l].has) L[l].w = win_block(g, l);
    }
    const uint32_t t_end = g.n + g.last_block;
    for (uint32_t t = 0; t < t_end; ++t) {
        uint32_t prev_hout[32];
        for (uint32_t l = 0; l < 32u; ++l) prev_hout[l] = L[l].hout;
        for (uint32_t l = 0; l < 32u; ++l) {
            Lane& x = L[l];
            if (!x.has) continue;
            const int rel = static_cast<int>(t) - static_cast<int>(x.w.start);
            if (rel >= 0 && static_cast<uint32_t>(rel) < x.w.width) {
                const uint32_t j = win_jlo(g, x.blk) + static_cast<uint32_t>(rel);
                const uint32_t hin = static_cast<uint32_t>(rel) < x.w.hin_lim ? prev_hout[(l + 31u) & 31u] : 1u;
                uint32_t ho;
                if (bw == 64u) {
                    ho = myers_step(x.pv, x.mv, masks(x.blk, b[j]), hin, x.w.hshift);
                } else {
                    uint32_t pv = static_cast<uint32_t>(x.pv), mv = static_cast<uint32_t>(x.mv);
                    ho = myers_step32(pv, mv, static_cast<uint32_t>(masks(x.blk, b[j])), hin, x.w.hshift);
                    x.pv = pv; x.mv = mv;
                }
                if (static_cast<uint32_t>(rel) < x.w.cnt_lim) x.partial += static_cast<int>(ho & 1u) - static_cast<int>(ho >> 1);
                x.hout = ho;
            }
            if (rel + 1 == static_cast<int>(x.w.width)) {          // out of columns: move on to block blk + 32
                x.blk += 32u;
                x.has = x.blk <= g.last_block;
                x.pv = fresh; x.mv = 0;
                if (x.has) {
                    const WinBlock nw = win_block(g, x.blk);
                    if (nw.start < t + 1u + WIN_SLACK) return -1;  // the lane is not free in time: K too wide for one warp
                    x.w = nw;
                }
            }
        }
    }
    long long total = g.m;
    for (uint32_t l = 0; l < 32u; ++l) total += L[l].partial;
    return total;
}
#include "../../svim_asm_b200/csrc/wfa_core.cuh"
// The thresholded wavefront distance of wfa.cu through its serial driver: strings as bytes (any value below 0xFA), padded
// with the sentinels here.  Returns the distance if it is <= t, else -1 (what edlib.align(a, b, k=t) answers).
extern "C" long long hc_wfa(const uint8_t* a, long long la, const uint8_t* b, long long lb, int t) {
    std::vector<uint32_t> A((static_cast<size_t>(la) + WFA_PAD + 7) / 4 + 1, 0u), B((static_cast<size_t>(lb) + WFA_PAD + 7) / 4 + 1, 0u);
    uint8_t* pa = reinterpret_cast<uint8_t*>(A.data());
    uint8_t* pb = reinterpret_cast<uint8_t*>(B.data());
    for (long long i = 0; i < la; ++i) pa[i] = a[i];
    for (long long i = 0; i < lb; ++i) pb[i] = b[i];
    for (uint32_t i = 0; i < WFA_PAD; ++i) { pa[la + i] = WFA_END_A; pb[lb + i] = WFA_END_B; }
    std::vector<int> F0(2 * static_cast<size_t>(t) + 7), F1(2 * static_cast<size_t>(t) + 7);
    return wfa_distance_serial(A.data(), static_cast<int>(la), B.data(), static_cast<int>(lb), t, F0.data(), F1.data());
}

// the bidirectional run (two wavefronts meeting in the middle), strings with front and back sentinels
extern "C" long long hc_wfa_bidir(const uint8_t* a, long long la, const uint8_t* b, long long lb, int t) {
    std::vector<uint32_t> A((static_cast<size_t>(la) + WFA_FRONT + WFA_PAD + 7) / 4 + 1, 0u), B((static_cast<size_t>(lb) + WFA_FRONT + WFA_PAD + 7) / 4 + 1, 0u);
    uint8_t* pa = reinterpret_cast<uint8_t*>(A.data());
    uint8_t* pb = reinterpret_cast<uint8_t*>(B.data());
    for (uint32_t i = 0; i < WFA_FRONT; ++i) { pa[i] = WFA_FRONT_A; pb[i] = WFA_FRONT_B; }
    for (long long i = 0; i < la; ++i) pa[WFA_FRONT + i] = a[i];
    for (long long i = 0; i < lb; ++i) pb[WFA_FRONT + i] = b[i];
    for (uint32_t i = 0; i < WFA_PAD; ++i) { pa[WFA_FRONT + la + i] = WFA_END_A; pb[WFA_FRONT + lb + i] = WFA_END_B; }
    const size_t W = 2 * static_cast<size_t>(t) + 7;
    std::vector<int> F0(W), F1(W), G0(W), G1(W);
    return wfa_bidir_serial(A.data(), static_cast<int>(la), B.data(), static_cast<int>(lb), t, F0.data(), F1.data(), G0.data(), G1.data());
}

// the kernel's own round structure (sentinel recurrence, overlap test folded into the forward wave), stated serially; the
// cases the kernel settles before its rounds (an empty string, a length difference beyond t) are settled the same way here
extern "C" long long hc_wfa_rounds(const uint8_t* a, long long la, const uint8_t* b, long long lb, int t) {
    const long long longer = la > lb ? la : lb, gap = la > lb ? la - lb : lb - la;
    if (la == 0 || lb == 0) return longer <= t ? longer : -1;
    if (gap > t) return -1;
    std::vector<uint32_t> A((WFA_FRONT + static_cast<size_t>(la) + WFA_PAD + 7) / 4 + 1, 0u), B((WFA_FRONT + static_cast<size_t>(lb) + WFA_PAD + 7) / 4 + 1, 0u);
    uint8_t* pa = reinterpret_cast<uint8_t*>(A.data());
    uint8_t* pb = reinterpret_cast<uint8_t*>(B.data());
    for (uint32_t i = 0; i < WFA_FRONT; ++i) { pa[i] = WFA_FRONT_A; pb[i] = WFA_FRONT_B; }
    for (long long i = 0; i < la; ++i) pa[WFA_FRONT + i] = a[i];
    for (long long i = 0; i < lb; ++i) pb[WFA_FRONT + i] = b[i];
    for (uint32_t i = 0; i < WFA_PAD; ++i) { pa[WFA_FRONT + la + i] = WFA_END_A; pb[WFA_FRONT + lb + i] = WFA_END_B; }
    std::vector<int> F(4 * (2 * static_cast<size_t>(t) + 7));
    return wfa_rounds_serial(A.data(), static_cast<int>(la), B.data(), static_cast<int>(lb), t, F.data());
}

extern "C" unsigned hc_win_kmax(unsigned m, unsigned n, unsigned bw) { return win_kmax(m, n, bw); }

#include "../../svim_asm_b200/csrc/inflate_core.cuh"
// The per-member DEFLATE decoder of bgzf_inflate.cu, on the host: checked against zlib's output.
extern "C" int hc_inflate(const uint8_t* src, unsigned src_len, uint8_t* dst, unsigned out_len) {
    return inflate_member(src, src_len, dst, out_len);
}

// The device decoder's flow (bam_device.cu: block header, direct tables, the symbol loop every lane runs in lock step, matches
// copied by the lanes side by side) with one host thread standing in for the warp.  `misalign` shifts the compressed
// bytes inside their buffer: the word refill must cope with any alignment of a member.
extern "C" int hc_inflate_fast(const uint8_t* src_in, unsigned src_len, uint8_t* dst, unsigned out_len, unsigned misalign) {
    std::vector<uint32_t> backing((src_len + 16) / 4 + 4);
    uint8_t* src = reinterpret_cast<uint8_t*>(backing.data()) + (misalign & 3u);
    memcpy(src, src_in, src_len);
    uint16_t lencnt[16], lensym[288], distcnt[16], distsym[32];
    uint8_t lengths[320];
    std::vector<inf_len_t> tlen(1u << INF_LEN_BITS);
    std::vector<uint32_t> tdist(1u << INF_DIST_BITS);
    std::vector<uint8_t> ring(INF_RING, 0xEE);
    InfHuff lencode{lencnt, lensym}, distcode{distcnt, distsym};
    InfBits b{src, src + src_len, 0ull, 0, 0};
    InfOut o = inf_out(ring.data(), dst, out_len);
    auto flush = [&]() {
        for (uint32_t lane = 0; lane < 32u; ++lane) inf_flush(o, o.pos, lane, 32u);
        o.flushed = o.pos;
    };
    int last = 0;
    do {
        uint32_t type = 0, st_off = 0, st_len = 0;
        int err = inf_block_header(b, src, o.pos, out_len, lencode, distcode, lengths, &last, &type, &st_off, &st_len);
        if (err) return err;
        if (type == 0u) {                                   // stored: through the ring in pieces, later matches may point into it
            for (uint32_t done = 0; done < st_len;) {
                const uint32_t chunk = std::min(st_len - done, INF_FLUSH_AT);
                for (uint32_t i = 0; i < chunk; ++i) ring[(o.rbase + o.pos + i) & INF_RMASK] = src[st_off + done + i];
                o.pos += chunk;
                done += chunk;
                flush();
            }
            continue;
        }
        std::fill(tlen.begin(), tlen.end(), static_cast<inf_len_t>(0));
        std::fill(tdist.begin(), tdist.end(), 0u);
        for (uint32_t lane = 0; lane < 32u; ++lane) {
            inf_fill_table(lencnt, lensym, tlen.data(), INF_LEN_BITS, false, lane, 32u);
            inf_fill_table(distcnt, distsym, tdist.data(), INF_DIST_BITS, true, lane, 32u);
        }
        err = inf_run_lanes(b, lencode, distcode, tlen.data(), tdist.data(), o, 0u, 1u);      // one host thread stands in for the warp
        if (err) return err;
    } while (!last);
    if (o.pos > out_len) return INF_ERR_OUTPUT;
    flush();
    return o.pos == out_len ? INF_OK : INF_ERR_SIZE;
}

// ---- split window (two halves meeting at row mh), the algorithm of window_pass_split in edit_distance.cu ---------------
namespace {
struct HalfResult { long long base; uint32_t jlo; std::vector<int> delta; bool ok; };

// window over the first g.m rows of P and the first g.n columns of T (32-row blocks); the LAST block does not count
// towards the sum: its bottom-row deltas are recorded per column instead
template <typename PF, typename TF>
HalfResult emulate_half(const WinGeom& g, PF P, TF T) {
    struct Lane { uint32_t blk; bool has; WinBlock w; uint32_t pv, mv, hout; long long partial; };
    Lane L[32];
    HalfResult out;
    out.ok = true;
    out.jlo = win_jlo(g, g.last_block);
    out.delta.assign(win_jhi(g, g.last_block) - out.jlo, 0);
    for (uint32_t l = 0; l < 32u; ++l) {
        L[l].blk = l; L[l].has = l <= g.last_block; L[l].pv = 0xFFFFFFFFu; L[l].mv = 0; L[l].hout = 0; L[l].partial = 0;
        if (L[l].has) L[l].w = win_block(g, l);
    }
    const uint32_t t_end = g.n + g.last_block;
    for (uint32_t t = 0; t < t_end; ++t) {
        uint32_t prev_hout[32];
        for (uint32_t l = 0; l < 32u; ++l) prev_hout[l] = L[l].hout;
        for (uint32_t l = 0; l < 32u; ++l) {
            Lane& x = L[l];
            if (!x.has) continue;
            const int rel = static_cast<int>(t) - static_cast<int>(x.w.start);
            if (rel >= 0 && static_cast<uint32_t>(rel) < x.w.width) {
                const uint32_t j = win_jlo(g, x.blk) + static_cast<uint32_t>(rel);
                uint32_t eq = 0;
                for (uint32_t r = 0; r < 32u; ++r) if (P(32u * x.blk + r) == T(j)) eq |= 1u << r;
                const uint32_t hin = static_cast<uint32_t>(rel) < x.w.hin_lim ? prev_hout[(l + 31u) & 31u] : 1u;
                const uint32_t ho = myers_step32(x.pv, x.mv, eq, hin, 31u);
                const int d = static_cast<int>(ho & 1u) - static_cast<int>(ho >> 1);
                if (x.blk == g.last_block) out.delta[j - out.jlo] = d;
                else if (static_cast<uint32_t>(rel) < x.w.cnt_lim) x.partial += d;
                x.hout = ho;
            }
            if (rel + 1 == static_cast<int>(x.w.width)) {
                x.blk += 32u;
                x.has = x.blk <= g.last_block;
                x.pv = 0xFFFFFFFFu; x.mv = 0;
                if (x.has) {
                    const WinBlock nw = win_block(g, x.blk);
                    if (nw.start < t + 1u + WIN_SLACK) out.ok = false;
                    x.w = nw;
                }
            }
        }
    }
    out.base = g.m;
    for (uint32_t l = 0; l < 32u; ++l) out.base += L[l].partial;
    return out;
}
}  // namespace

extern "C" long long hc_myers_window_split(const uint8_t* a, long long m0, const uint8_t* b, long long n0, unsigned K) {
    const uint32_t pad = (32u - (static_cast<uint32_t>(m0) & 31u)) & 31u;
    const uint32_t m = static_cast<uint32_t>(m0) + pad, n = static_cast<uint32_t>(n0) + pad, dlt = n - m;
    auto P = [&](uint32_t i) -> int { return i < static_cast<uint32_t>(m0) ? a[i] : (i < m ? 256 : -1); };      // 256: the sentinel
    auto T = [&](uint32_t j) -> int { return j < static_cast<uint32_t>(n0) ? b[j] : (j < n ? 256 : -2); };
    auto PR = [&](uint32_t i) -> int { return i < m ? P(m - 1u - i) : -1; };
    auto TR = [&](uint32_t j) -> int { return j < n ? T(n - 1u - j) : -2; };
    const WinSplit sp = win_split(m, n, K, 32u);
    if (sp.rows_f == 0 || sp.rows_b == 0) return -1;
    const HalfResult f = emulate_half(win_geom_half(sp.rows_f, sp.cols_f, dlt, K, 32u), P, T);
    const HalfResult r = emulate_half(win_geom_half(sp.rows_b, sp.cols_b, dlt, K, 32u), PR, TR);
    if (!f.ok || !r.ok) return -1;
    // F(j), j in [f.jlo, f.jlo + len]; B(j'), j' in [r.jlo, r.jlo + len]; D = min F(j) + B(n - j)
    std::vector<long long> F(f.delta.size() + 1), B(r.delta.size() + 1);
    F[0] = f.base;
    for (size_t i = 0; i < f.delta.size(); ++i) F[i + 1] = F[i] + f.delta[i];
    B[0] = r.base;
    for (size_t i = 0; i < r.delta.size(); ++i) B[i + 1] = B[i] + r.delta[i];
    long long best = -2;
    for (size_t i = 0; i < F.size(); ++i) {
        const long long j = static_cast<long long>(f.jlo) + static_cast<long long>(i);
        const long long jb = static_cast<long long>(n) - j - static_cast<long long>(r.jlo);
        if (jb < 0 || jb >= static_cast<long long>(B.size())) continue;
        const long long v = F[i] + B[static_cast<size_t>(jb)];
        if (best < 0 || v < best) best = v;
    }
    return best;
}

#include "../../svim_asm_b200/csrc/vcf_core.cuh"

// VCF body lines of `n_entries` entries on the host: the plan of vcf_core.cuh executed byte by byte.
// Returns the number of bytes (written up to `cap`).
extern "C" long long hc_vcf_body(const svb_row* rows, const svb_vcf_entry* entries, long long n_entries, const uint8_t* bases,
                                 const uint64_t* contig_off, int n_contig, const uint8_t* names, const uint32_t* name_off,
                                 const uint8_t* seq4_0, const uint64_t* seq_off_0, const uint8_t* seq4_1, const uint64_t* seq_off_1,
                                 const uint8_t* seq4_2, const uint64_t* seq_off_2, unsigned flags, uint8_t* dst, long long cap) {
    VcfEnv env = {};
    env.bases = bases;
    env.contig_off = contig_off;
    env.n_contig = n_contig;
    env.names = names;
    env.name_off = name_off;
    env.seq4[0] = seq4_0; env.seq_off[0] = seq_off_0;
    env.seq4[1] = seq4_1; env.seq_off[1] = seq_off_1;
    env.seq4[2] = seq4_2; env.seq_off[2] = seq_off_2;
    env.flags = flags;
    long long pos = 0;
    VcfPlan plan;
    for (long long e = 0; e < n_entries; ++e) {
        vcf_plan(rows[entries[e].row], entries[e], env, plan);
        for (uint32_t k = 0; k < plan.n_pieces; ++k)
            for (uint64_t i = 0; i < plan.piece[k].len; ++i, ++pos)
                if (pos < cap) dst[pos] = vcf_piece_byte(plan, plan.piece[k], env, i);
    }
    return pos;
}
