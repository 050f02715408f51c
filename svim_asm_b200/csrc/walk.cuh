// K4 core -- the split-alignment inter-segment walk of one read, as a __host__ __device__ routine.
//
// Replaces analyze_read_segments (reference src/svim_asm/SVIM_inter.py:62-340) with its helpers
// is_similar (:12-16), reciprocal_overlap_distance (:19-39), process_overlapping_inversions
// (:42-60), and the constructors of SVCandidate.py that it calls (clamps, asserts, the breakend
// normalisation by contig-name string order, :351-376).
//
// Everything is integer arithmetic on (q_start, q_end, tid, ref_start, ref_end, is_reverse) tuples
// except the inversion clustering (fp64, linkage.cuh).  The quirks of the reference are kept on
// purpose and marked QUIRK.
#pragma once
#include "linkage.cuh"
#include "../../include/svimasm_b200.h"

struct WalkSeg {
    int32_t q_start, q_end, tid, ref_start, ref_end, rev;
};
struct WalkTandem {
    int32_t tid, start, end, fully, dir_fwd;
};
struct WalkTrans {
    int32_t dir1_fwd, dir2_fwd, tid1, pos1, tid2, pos2;
};
struct WalkInv {
    int32_t tid, start, end, side;      // side 0 = "left_*", 1 = "right_*"
};
// one scratch slot per (primary + SA segment); the four lists of a read share the slot range
struct WalkScratch {
    WalkSeg seg;
    WalkTandem tan;
    WalkTrans tra;
    WalkInv inv;
};

struct WalkParams {
    int32_t min_mapq, min_sv, max_sv, qgt, qot, rgt, rot;
};

struct WalkRead {
    uint32_t aln_idx;        // the primary record
    uint32_t hap;
    int32_t read_len;        // primary.infer_read_length()
    uint32_t l_seq;          // stored query length of the primary
    const int32_t* contig_len;
    const int32_t* contig_lexrank;
    int32_t n_contig;
};

enum : uint32_t { WALK_ERR_BAD_TID = 1u, WALK_ERR_ASSERT = 2u, WALK_ERR_CAPACITY = 4u, WALK_ERR_NOSEQ = 8u };

struct WalkOut {
    svb_row* rows;           // nullptr: count only
    uint32_t cap;            // rows beyond `cap` are counted but not stored
    uint32_t n;
    uint32_t err;
    uint32_t ins_bytes;      // 4-bit packed bytes of the inserted sequences emitted so far (sizes the table's sequence pool)
};

SVB_HD int32_t wk_max(int32_t a, int32_t b) { return a > b ? a : b; }
SVB_HD int32_t wk_min(int32_t a, int32_t b) { return a < b ? a : b; }
SVB_HD long long wk_abs(long long a) { return a < 0 ? -a : a; }

SVB_HD void wk_blank(svb_row& r, const WalkRead& rd, uint32_t local) {
    r.flags = 0;
    r.genotype = SVB_GT_HOM;
    r.hap = static_cast<uint8_t>(rd.hap);
    r.src_tid = -1; r.src_start = 0; r.src_end = 0;
    r.dst_tid = -1; r.dst_start = 0; r.dst_end = 0;
    r.copies = 0;
    r.aln_idx = rd.aln_idx;
    r.seq_pos = 0; r.seq_len = 0;
    r.mate_aln = 0xFFFFFFFFu;
    r.ordinal = (static_cast<unsigned long long>(rd.aln_idx) << 32) | 0x80000000ull | local;
    r.reserved0 = 0;
}

SVB_HD void wk_push(WalkOut& o, const svb_row& r) {
    if (o.rows && o.n < o.cap) o.rows[o.n] = r;
    ++o.n;
}

// python slice s[start:stop] on a string of length n -> (pos, len)
SVB_HD void wk_pyslice(long long start, long long stop, long long n, uint32_t& pos, uint32_t& len) {
    if (start < 0) { start += n; if (start < 0) start = 0; } else if (start > n) start = n;
    if (stop < 0) { stop += n; if (stop < 0) stop = 0; } else if (stop > n) stop = n;
    pos = static_cast<uint32_t>(start);
    len = stop > start ? static_cast<uint32_t>(stop - start) : 0u;
}

// ---- constructors (SVCandidate.py) ----------------------------------------------------------
SVB_HD void wk_emit_del(WalkOut& o, const WalkRead& rd, int32_t tid, long long start, long long end) {
    if (end < start) o.err |= WALK_ERR_ASSERT;                                   // SVCandidate.py:40
    svb_row r; wk_blank(r, rd, o.n);
    r.type = SVB_DEL;
    r.src_tid = tid;
    r.src_start = static_cast<int32_t>(start < 0 ? 0 : start);                  // :44
    r.src_end = static_cast<int32_t>(end < rd.contig_len[tid] ? end : rd.contig_len[tid]);   // :46
    wk_push(o, r);
}
SVB_HD void wk_emit_ins(WalkOut& o, const WalkRead& rd, int32_t tid, long long start, long long end,
                        long long seq_start, long long seq_stop) {
    if (end < start) o.err |= WALK_ERR_ASSERT;                                   // SVCandidate.py:130
    if (rd.l_seq == 0) o.err |= WALK_ERR_NOSEQ;                                  // query_sequence is None -> TypeError
    svb_row r; wk_blank(r, rd, o.n);
    r.type = SVB_INS;
    r.dst_tid = tid;
    r.dst_start = static_cast<int32_t>(start < 0 ? 0 : start);                  // :134
    r.dst_end = static_cast<int32_t>(end < rd.contig_len[tid] ? end : rd.contig_len[tid]);   // :136
    wk_pyslice(seq_start, seq_stop, rd.l_seq, r.seq_pos, r.seq_len);
    o.ins_bytes += (r.seq_len + 1u) / 2u;
    wk_push(o, r);
}
SVB_HD void wk_emit_inv(WalkOut& o, const WalkRead& rd, int32_t tid, long long start, long long end, bool complete) {
    if (end < start) o.err |= WALK_ERR_ASSERT;                                   // SVCandidate.py:83
    svb_row r; wk_blank(r, rd, o.n);
    r.type = SVB_INV;
    r.flags = complete ? SVB_F_COMPLETE : 0;
    r.src_tid = tid;
    r.src_start = static_cast<int32_t>(start < 0 ? 0 : start);
    r.src_end = static_cast<int32_t>(end < rd.contig_len[tid] ? end : rd.contig_len[tid]);
    wk_push(o, r);
}
SVB_HD void wk_emit_tandem(WalkOut& o, const WalkRead& rd, int32_t tid, long long start, long long end, int32_t copies,
                           bool fully) {
    if (end < start) o.err |= WALK_ERR_ASSERT;                                   // SVCandidate.py:181
    svb_row r; wk_blank(r, rd, o.n);
    r.type = SVB_DUP_TAN;
    r.flags = fully ? SVB_F_FULLY_COVERED : 0;
    r.copies = copies;
    r.src_tid = tid;
    r.src_start = static_cast<int32_t>(start < 0 ? 0 : start);
    r.src_end = static_cast<int32_t>(end < rd.contig_len[tid] ? end : rd.contig_len[tid]);
    wk_push(o, r);
}
SVB_HD void wk_emit_dupint(WalkOut& o, const WalkRead& rd, int32_t stid, long long sstart, long long send, int32_t dtid,
                           long long dstart, long long dend) {
    if (send < sstart || dend < dstart) o.err |= WALK_ERR_ASSERT;                // SVCandidate.py:266-267
    svb_row r; wk_blank(r, rd, o.n);
    r.type = SVB_DUP_INT;
    r.src_tid = stid;
    r.src_start = static_cast<int32_t>(sstart < 0 ? 0 : sstart);
    r.src_end = static_cast<int32_t>(send < rd.contig_len[stid] ? send : rd.contig_len[stid]);
    r.dst_tid = dtid;
    r.dst_start = static_cast<int32_t>(dstart < 0 ? 0 : dstart);
    r.dst_end = static_cast<int32_t>(dend < rd.contig_len[dtid] ? dend : rd.contig_len[dtid]);
    wk_push(o, r);
}
// CandidateBreakend.__init__ (SVCandidate.py:351-376): (contig, pos) of the source must sort before
// the destination under python STRING order of the contig names; otherwise swap and flip directions.
SVB_HD void wk_fill_bnd(svb_row& r, const int32_t* contig_len, const int32_t* lexrank, int32_t tid1, long long pos1,
                        bool fwd1, int32_t tid2, long long pos2, bool fwd2) {
    const bool keep = lexrank[tid1] < lexrank[tid2] || (tid1 == tid2 && pos1 < pos2);
    int32_t st, dt;
    long long sp, dp;
    bool sf, df;
    if (keep) { st = tid1; sp = pos1; sf = fwd1; dt = tid2; dp = pos2; df = fwd2; }
    else { st = tid2; sp = pos2; sf = !fwd2; dt = tid1; dp = pos1; df = !fwd1; }
    r.type = SVB_BND;
    r.src_tid = st;
    r.src_start = static_cast<int32_t>(sp < 0 ? 0 : (sp < contig_len[st] ? sp : contig_len[st]));
    r.src_end = 0;
    r.dst_tid = dt;
    r.dst_start = static_cast<int32_t>(dp < 0 ? 0 : (dp < contig_len[dt] ? dp : contig_len[dt]));
    r.dst_end = 0;
    r.flags = static_cast<uint8_t>((r.flags & ~(SVB_F_SRC_FWD | SVB_F_DST_FWD)) | (sf ? SVB_F_SRC_FWD : 0) | (df ? SVB_F_DST_FWD : 0));
}
SVB_HD void wk_emit_bnd(WalkOut& o, const WalkRead& rd, WalkScratch* sc, uint32_t& n_tra, int32_t tid1, long long pos1,
                        bool fwd1, int32_t tid2, long long pos2, bool fwd2) {
    svb_row r; wk_blank(r, rd, o.n);
    wk_fill_bnd(r, rd.contig_len, rd.contig_lexrank, tid1, pos1, fwd1, tid2, pos2, fwd2);
    wk_push(o, r);
    // the raw, un-normalised tuple also feeds the interspersed-duplication pass (SVIM_inter.py:134,139,...)
    WalkTrans& t = sc[n_tra++].tra;
    t.dir1_fwd = fwd1; t.dir2_fwd = fwd2; t.tid1 = tid1; t.pos1 = static_cast<int32_t>(pos1);
    t.tid2 = tid2; t.pos2 = static_cast<int32_t>(pos2);
}

// reciprocal_overlap_distance (SVIM_inter.py:19-39), fp64 like the reference
SVB_HD double wk_inv_distance(const WalkInv& a, const WalkInv& b) {
    if (a.side == b.side) return 1.0;
    if (b.start >= a.end) return 1.0;
    if (a.start >= b.end) return 1.0;
    const int32_t m = wk_min(a.end, b.end);
    const double overlap = static_cast<double>(b.start >= a.start ? m - b.start : m - a.start);
    const double r1 = overlap / static_cast<double>(a.end - a.start);
    const double r2 = overlap / static_cast<double>(b.end - b.start);
    return 1.0 - (r1 < r2 ? r1 : r2);
}

// process_overlapping_inversions (SVIM_inter.py:42-60) on sc[first .. first+n).inv
SVB_HD void wk_flush_inversions(WalkOut& o, const WalkRead& rd, const WalkScratch* sc, uint32_t first, uint32_t n) {
    if (n == 0) return;
    if (n < 2) {
        const WalkInv& v = sc[first].inv;
        wk_emit_inv(o, rd, v.tid, v.start, v.end, false);
        return;
    }
    if (n > LINK_MAXN) { o.err |= WALK_ERR_CAPACITY; return; }
    double dist[LINK_MAXN * (LINK_MAXN - 1) / 2];
    int labels[LINK_MAXN];
    int m = 0;
    for (uint32_t i = 0; i + 1 < n; ++i)
        for (uint32_t j = i + 1; j < n; ++j) dist[m++] = wk_inv_distance(sc[first + i].inv, sc[first + j].inv);
    const int n_clusters = link_complete_fcluster(static_cast<int>(n), dist, 0.3, labels);
    for (int c = 1; c <= n_clusters; ++c) {
        int32_t tid = -1, lo = 0, hi = 0, members = 0;
        for (uint32_t i = 0; i < n; ++i) {
            if (labels[i] != c) continue;
            const WalkInv& v = sc[first + i].inv;
            if (members == 0) { tid = v.tid; lo = v.start; hi = v.end; }
            else { lo = wk_max(lo, v.start); hi = wk_min(hi, v.end); }
            ++members;
        }
        wk_emit_inv(o, rd, tid, lo, hi, members > 1);          // start = max(starts), end = min(ends)
    }
}

// sc[0..k) .seg holds the read's segments: primary first, then the SA segments that passed the
// mapq filter (SVIM_COLLECT.py:77).  Returns rows through `o`.
SVB_HD void walk_read(const WalkRead& rd, const WalkParams& p, WalkScratch* sc, uint32_t k, WalkOut& o) {
    // stable sort by (q_start, q_end)                                        SVIM_inter.py:83
    for (uint32_t i = 1; i < k; ++i) {
        const WalkSeg s = sc[i].seg;
        uint32_t j = i;
        while (j > 0 && (sc[j - 1].seg.q_start > s.q_start ||
                         (sc[j - 1].seg.q_start == s.q_start && sc[j - 1].seg.q_end > s.q_end))) {
            sc[j].seg = sc[j - 1].seg;
            --j;
        }
        sc[j].seg = s;
    }
    uint32_t n_tan = 0, n_tra = 0, n_inv = 0;
    const long long mn = p.min_sv, mx = p.max_sv;
    for (uint32_t i = 0; i + 1 < k; ++i) {
        const WalkSeg C = sc[i].seg, N = sc[i + 1].seg;
        if (C.tid < 0 || C.tid >= rd.n_contig || N.tid < 0 || N.tid >= rd.n_contig) {
            o.err |= WALK_ERR_BAD_TID;               // bam.get_reference_name / getrname raises (:99,:225-226)
            continue;
        }
        const long long dr = static_cast<long long>(N.q_start) - C.q_end;                        // :95
        if (C.tid == N.tid) {
            const int32_t chr = C.tid;
            if (C.rev == N.rev) {
                const long long dref = C.rev ? static_cast<long long>(C.ref_start) - N.ref_end
                                             : static_cast<long long>(N.ref_start) - C.ref_end;   // :103-106
                if (dr >= -static_cast<long long>(p.qot)) {                                      // :108
                    if (dref >= -static_cast<long long>(p.rot)) {                                // :110
                        const long long dev = dr - dref;
                        if (dev >= mn) {                                                         // :113
                            if (dref <= p.rgt) {
                                if (!C.rev)
                                    wk_emit_ins(o, rd, chr, C.ref_end, C.ref_end + dev, C.q_end, C.q_end + dev);      // :117-118
                                else {
                                    const long long s0 = static_cast<long long>(rd.read_len) - N.q_start;
                                    wk_emit_ins(o, rd, chr, C.ref_start, C.ref_start + dev, s0, s0 + dev);            // :120-121
                                }
                            }
                        } else if (-mx <= dev && dev <= -mn) {                                   // :123
                            if (dr <= p.qgt) {
                                if (!C.rev) wk_emit_del(o, rd, chr, C.ref_end, C.ref_end - dev);                       // :127
                                else wk_emit_del(o, rd, chr, N.ref_end, N.ref_end - dev);                              // :129
                            }
                        } else if (dev < -mx) {                                                  // :131
                            if (dr <= p.qgt) {
                                if (!C.rev) wk_emit_bnd(o, rd, sc, n_tra, chr, static_cast<long long>(C.ref_end) - 1, true, chr, N.ref_start, true);
                                else wk_emit_bnd(o, rd, sc, n_tra, chr, C.ref_start, false, chr, static_cast<long long>(N.ref_end) - 1, false);
                            }
                        }
                    } else if (dr <= p.qgt) {                                                    // :141-143
                        const long long dev = dr - dref;
                        if (dev >= mn) {
                            if (!C.rev) {
                                if (N.ref_end > C.ref_start) {                                   // :147 fully covered
                                    WalkTandem& t = sc[n_tan++].tan;
                                    t.tid = chr; t.start = N.ref_start; t.end = static_cast<int32_t>(N.ref_start + dev); t.fully = 1; t.dir_fwd = 1;
                                } else if (dref >= -mx) {                                        // :150
                                    WalkTandem& t = sc[n_tan++].tan;
                                    t.tid = chr; t.start = N.ref_start; t.end = static_cast<int32_t>(N.ref_start + dev); t.fully = 0; t.dir_fwd = 1;
                                } else {
                                    wk_emit_bnd(o, rd, sc, n_tra, chr, static_cast<long long>(C.ref_end) - 1, true, chr, N.ref_start, true);
                                }
                            } else {
                                if (N.ref_start < C.ref_end) {                                   // :160
                                    WalkTandem& t = sc[n_tan++].tan;
                                    t.tid = chr; t.start = C.ref_start; t.end = static_cast<int32_t>(C.ref_start + dev); t.fully = 1; t.dir_fwd = 0;
                                } else if (dref >= -mx) {                                        // :163
                                    WalkTandem& t = sc[n_tan++].tan;
                                    t.tid = chr; t.start = C.ref_start; t.end = static_cast<int32_t>(C.ref_start + dev); t.fully = 0; t.dir_fwd = 0;
                                } else {
                                    wk_emit_bnd(o, rd, sc, n_tra, chr, C.ref_start, false, chr, static_cast<long long>(N.ref_end) - 1, false);
                                }
                            }
                        }
                    }
                }
            } else {
                const bool in_window = -static_cast<long long>(p.qot) <= dr && dr <= p.qgt;     // :175,:201
                if (!C.rev && N.rev) {                                                           // :172
                    const long long dref = static_cast<long long>(N.ref_end) - C.ref_end;
                    const long long dev = dr - dref;
                    if (in_window) {
                        if (static_cast<long long>(N.ref_start) - C.ref_end >= -static_cast<long long>(p.rot)) {       // case 1
                            if (mn <= -dev && -dev <= mx) {
                                WalkInv& v = sc[n_inv++].inv;
                                v.tid = chr; v.start = C.ref_end; v.end = static_cast<int32_t>(C.ref_end - dev); v.side = 0;   // "left_fwd"
                            } else {
                                wk_emit_bnd(o, rd, sc, n_tra, chr, static_cast<long long>(C.ref_end) - 1, true, chr, static_cast<long long>(N.ref_end) - 1, false);
                            }
                        } else if (static_cast<long long>(C.ref_start) - N.ref_end >= -static_cast<long long>(p.rot)) { // case 3
                            if (mn <= dev && dev <= mx) {
                                WalkInv& v = sc[n_inv++].inv;
                                v.tid = chr; v.start = N.ref_end; v.end = static_cast<int32_t>(N.ref_end + dev); v.side = 0;   // "left_rev"
                            } else {
                                wk_emit_bnd(o, rd, sc, n_tra, chr, static_cast<long long>(C.ref_end) - 1, true, chr, static_cast<long long>(N.ref_end) - 1, false);
                            }
                        }
                    }
                }
                if (C.rev && !N.rev) {                                                           // :198
                    const long long dref = static_cast<long long>(N.ref_start) - C.ref_start;
                    const long long dev = dr - dref;
                    if (in_window) {
                        if (static_cast<long long>(N.ref_start) - C.ref_end >= -static_cast<long long>(p.rot)) {       // case 2
                            if (mn <= -dev && -dev <= mx) {
                                WalkInv& v = sc[n_inv++].inv;
                                v.tid = chr; v.start = C.ref_start; v.end = static_cast<int32_t>(C.ref_start - dev); v.side = 1; // "right_fwd"
                            } else {
                                wk_emit_bnd(o, rd, sc, n_tra, chr, C.ref_start, false, chr, N.ref_start, true);
                            }
                        } else if (static_cast<long long>(C.ref_start) - N.ref_end >= -static_cast<long long>(p.rot)) { // case 4
                            if (mn <= dev && dev <= mx) {
                                WalkInv& v = sc[n_inv++].inv;
                                v.tid = chr; v.start = N.ref_start; v.end = static_cast<int32_t>(N.ref_start + dev); v.side = 1; // "right_rev"
                            } else {
                                wk_emit_bnd(o, rd, sc, n_tra, chr, C.ref_start, false, chr, N.ref_start, true);
                            }
                        }
                    }
                }
            }
        } else if (dr >= -static_cast<long long>(p.qot) && dr <= p.qgt) {                        // :224-254
            if (C.rev == N.rev) {
                if (!C.rev) wk_emit_bnd(o, rd, sc, n_tra, C.tid, static_cast<long long>(C.ref_end) - 1, true, N.tid, N.ref_start, true);
                else wk_emit_bnd(o, rd, sc, n_tra, C.tid, C.ref_start, false, N.tid, static_cast<long long>(N.ref_end) - 1, false);
            } else {
                if (!C.rev) wk_emit_bnd(o, rd, sc, n_tra, C.tid, static_cast<long long>(C.ref_end) - 1, true, N.tid, static_cast<long long>(N.ref_end) - 1, false);
                else wk_emit_bnd(o, rd, sc, n_tra, C.tid, C.ref_start, false, N.tid, N.ref_start, true);
            }
        }
    }

    // ---- tandem duplications: runs of similar entries (SVIM_inter.py:260-290)
    if (n_tan) {
        int32_t cur_tid = sc[0].tan.tid;
        long long sum_s = sc[0].tan.start, sum_e = sc[0].tan.end;
        int32_t copies = 1, any_full = sc[0].tan.fully;
        const int32_t first_dir = sc[0].tan.dir_fwd;      // QUIRK: current_direction is never updated (:273 vs :283-287)
        for (uint32_t i = 1; i < n_tan; ++i) {
            const WalkTandem t = sc[i].tan;
            // is_similar(chr, mean(starts), mean(ends), ...) with the means kept exact: |sum - n*x| < 20*n
            const bool similar = cur_tid == t.tid && wk_abs(sum_s - static_cast<long long>(copies) * t.start) < 20ll * copies &&
                                 wk_abs(sum_e - static_cast<long long>(copies) * t.end) < 20ll * copies;
            if (similar && first_dir == t.dir_fwd) {
                sum_s += t.start; sum_e += t.end; ++copies; any_full |= t.fully;
            } else {
                wk_emit_tandem(o, rd, cur_tid, sum_s / copies, sum_e / copies, copies, any_full != 0);   // int(mean()) truncates
                cur_tid = t.tid; sum_s = t.start; sum_e = t.end; copies = 1; any_full = t.fully;
            }
        }
        wk_emit_tandem(o, rd, cur_tid, sum_s / copies, sum_e / copies, copies, any_full != 0);
    }

    // ---- interspersed duplications: pairs of translocations (SVIM_inter.py:292-320)
    for (uint32_t i = 0; i < n_tra; ++i) {
        const WalkTrans th = sc[i].tra;
        for (uint32_t j = 0; j < i; ++j) {
            const WalkTrans bf = sc[j].tra;
            if (bf.dir1_fwd != th.dir2_fwd || bf.dir2_fwd != th.dir1_fwd) continue;              // :303
            if (!(bf.tid1 == th.tid2 && wk_abs(static_cast<long long>(bf.pos1) - th.pos2) < 20)) continue;   // :305 (ends are 0 == 0)
            if (bf.tid2 != th.tid1) continue;                                                     // :307
            if (bf.dir2_fwd != bf.dir1_fwd) continue;                                             // :309
            if (bf.dir1_fwd) {
                const long long length = static_cast<long long>(th.pos1) + 1 - bf.pos2;
                if (mn <= length && length <= mx) {
                    const long long m = (static_cast<long long>(bf.pos1) + 1 + th.pos2) / 2;      // int(mean([a, b]))
                    wk_emit_dupint(o, rd, bf.tid2, bf.pos2, static_cast<long long>(th.pos1) + 1, bf.tid1, m, m + length);
                }
            } else {
                const long long length = static_cast<long long>(bf.pos2) + 1 - th.pos1;
                if (mn <= length && length <= mx) {
                    const long long m = (static_cast<long long>(bf.pos1) + th.pos2 + 1) / 2;
                    wk_emit_dupint(o, rd, bf.tid2, th.pos1, static_cast<long long>(bf.pos2) + 1, bf.tid1, m, m + length);
                }
            }
        }
    }

    // ---- inversions: sort, sweep overlapping runs, cluster each run (SVIM_inter.py:322-338)
    for (uint32_t i = 1; i < n_inv; ++i) {
        const WalkInv v = sc[i].inv;
        const int32_t rk = rd.contig_lexrank[v.tid];
        uint32_t j = i;
        while (j > 0) {
            const WalkInv& w = sc[j - 1].inv;
            const int32_t rw = rd.contig_lexrank[w.tid];
            const bool greater = rw > rk || (rw == rk && (w.start > v.start || (w.start == v.start && w.end > v.end)));
            if (!greater) break;
            sc[j].inv = w;
            --j;
        }
        sc[j].inv = v;
    }
    uint32_t run_first = 0, run_n = 0;
    int32_t run_max_end = 0;
    for (uint32_t i = 0; i < n_inv; ++i) {
        const WalkInv v = sc[i].inv;
        if (run_n == 0) {
            run_first = i; run_n = 1; run_max_end = v.end;
        } else if (v.tid == sc[run_first + run_n - 1].inv.tid && v.start < run_max_end) {
            ++run_n;
            run_max_end = wk_max(run_max_end, v.end);
        } else {
            wk_flush_inversions(o, rd, sc, run_first, run_n);
            run_n = 0;                                // QUIRK: the current inversion is dropped, not kept (:333-336)
        }
    }
    wk_flush_inversions(o, rd, sc, run_first, run_n);
}
