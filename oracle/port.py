"""TEST INFRASTRUCTURE -- CPU restatement of the SVIM-asm hot path on flat arrays (numpy / python).

This is the oracle that TRAVELS to the GPU box (the reference itself cannot: it lives under
/root/reference and needs pysam).  It restates, function by function,
    SVIM_intra.py:8-44     analyze_cigar_indel / analyze_alignment_indel  -> scan_*(), collect()
    SVIM_COLLECT.py:61-83  record filter and dispatch                      -> collect()
    SVIM_inter.py:62-340   analyze_read_segments                           -> walk()
    SVCandidate.py         constructors (clamps, asserts, BND normalisation)-> _mk_*()
    SVIM_COMBINE.py:15-366 form_partitions / compute_distance / pair_*     -> pair()
on the same record image the GPU consumes, producing rows with the fields of svb_row.
It is pinned against the unmodified reference (run through oracle/shims) by
tests/test_oracle_vs_reference.py in the build container and by the committed golden fixtures
under tests/golden/ everywhere else.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline may import it; the product never does.
"""
import ctypes
import os

import numpy as np

DEL, INV, INS, DUP_TAN, DUP_INT, BND = range(6)
TYPE_NAMES = ("DEL", "INV", "INS", "DUP_TAN", "DUP_INT", "BND")
F_COMPLETE, F_FULLY, F_CUTPASTE, F_SRC_FWD, F_DST_FWD = 1, 2, 4, 8, 16
GT_HOM, GT_HAP1, GT_HAP2 = 0, 1, 2
NO_MATE = 0xFFFFFFFF

ROW_DTYPE = np.dtype([("type", "u1"), ("flags", "u1"), ("genotype", "u1"), ("hap", "u1"),
                      ("src_tid", "<i4"), ("src_start", "<i4"), ("src_end", "<i4"),
                      ("dst_tid", "<i4"), ("dst_start", "<i4"), ("dst_end", "<i4"),
                      ("copies", "<i4"), ("aln_idx", "<u4"), ("seq_pos", "<u4"), ("seq_len", "<u4"),
                      ("mate_aln", "<u4"), ("ordinal", "<u8"), ("reserved0", "<u8")])

DEFAULTS = dict(min_mapq=20, min_sv_size=40, max_sv_size=100000, query_gap_tolerance=50,
                query_overlap_tolerance=50, reference_gap_tolerance=50, reference_overlap_tolerance=50,
                partition_max_distance=1000, max_edit_distance=200)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_clib = None


def clib():
    """The C helpers (oracle/csrc/oracle.c), built by oracle/Makefile; None if absent."""
    global _clib
    if _clib is None and os.path.exists(_SO):
        lib = ctypes.CDLL(_SO)
        lib.orc_edit_distance.restype = ctypes.c_int64
        lib.orc_edit_distance.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64]
        lib.orc_cigar_scan.restype = ctypes.c_int64
        lib.orc_cigar_scan.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64]
        _clib = lib
    return _clib


class Params(object):
    def __init__(self, **kw):
        for k, v in DEFAULTS.items():
            setattr(self, k, int(kw.get(k, v)))


# ---------------------------------------------------------------------------------------------
# a3: the CIGAR-op scan


def scan_python(ops, min_length):
    """SVIM_intra.py:8-30 as the reference runs it: one interpreter iteration per op."""
    ref = read = 0
    found = []
    for packed in ops:
        code = packed & 15
        n = packed >> 4
        if code == 0 or code == 7 or code == 8:      # M, =, X advance both (:14-16, :27-29)
            ref += n
            read += n
        elif code == 1:                                # I (:17-20)
            if n >= min_length:
                found.append((ref, read, n, "INS"))
            read += n
        elif code == 2:                                # D (:21-24)
            if n >= min_length:
                found.append((ref, read, n, "DEL"))
            ref += n
        elif code == 4:                                # S (:25-26)
            read += n
        # N, H, P, B and the pad code are ignored: N does not advance the reference (QUIRK App. B#1)
    return found


def scan_numpy(ops, min_length):
    """Same result as scan_python through prefix sums (used for large inputs)."""
    ops = np.asarray(ops, dtype=np.uint32)
    code = ops & 15
    n = (ops >> 4).astype(np.int64)
    adv_ref = np.where((code == 0) | (code == 2) | (code == 7) | (code == 8), n, 0)
    adv_read = np.where((code == 0) | (code == 1) | (code == 4) | (code == 7) | (code == 8), n, 0)
    pos_ref = np.cumsum(adv_ref) - adv_ref
    pos_read = np.cumsum(adv_read) - adv_read
    hit = np.nonzero(((code == 1) | (code == 2)) & (n >= min_length))[0]
    return [(int(pos_ref[i]), int(pos_read[i]), int(n[i]), "INS" if code[i] == 1 else "DEL") for i in hit]


# ---------------------------------------------------------------------------------------------
# a10: constructors


class OracleAbort(Exception):
    """The reference would have raised (assert / ValueError) and aborted the run."""


def _blank(hap, aln_idx, ordinal):
    r = np.zeros((), dtype=ROW_DTYPE)
    r["hap"] = hap
    r["src_tid"] = r["dst_tid"] = -1
    r["aln_idx"] = aln_idx
    r["mate_aln"] = NO_MATE
    r["ordinal"] = ordinal
    return r


def _span_row(kind, tid, start, end, clen, hap, aln_idx, ordinal, dest=False):
    if end < start:
        raise OracleAbort("end smaller than start")                   # SVCandidate.py:40,83,130,181
    r = _blank(hap, aln_idx, ordinal)
    r["type"] = kind
    pre = "dst" if dest else "src"
    r[pre + "_tid"] = tid
    r[pre + "_start"] = max(0, start)
    r[pre + "_end"] = min(clen[tid], end)
    return r


def _trunc_div(a, b):
    """int(mean(...)): exact quotient truncated toward zero (SVIM_inter.py:282,290,313,317)."""
    return a // b if a >= 0 else -((-a) // b)


def _pyslice(start, stop, n):
    lo, hi, _ = slice(start, stop).indices(n)
    return lo, max(0, hi - lo)


def bnd_fields(contig_len, rank, tid1, pos1, fwd1, tid2, pos2, fwd2):
    """CandidateBreakend.__init__ (SVCandidate.py:351-376): normalise by python string order of the names."""
    if rank[tid1] < rank[tid2] or (tid1 == tid2 and pos1 < pos2):
        st, sp, sf, dt, dp, df = tid1, pos1, fwd1, tid2, pos2, fwd2
    else:
        st, sp, sf, dt, dp, df = tid2, pos2, not fwd2, tid1, pos1, not fwd1
    return st, min(contig_len[st], max(0, sp)), sf, dt, min(contig_len[dt], max(0, dp)), df


# ---------------------------------------------------------------------------------------------
# a5-a9: the split-alignment walk of one read


def _inv_distance(a, b):
    """reciprocal_overlap_distance (SVIM_inter.py:19-39)."""
    if a[3] == b[3]:
        return 1.0
    if b[1] >= a[2] or a[1] >= b[2]:
        return 1.0
    top = min(a[2], b[2])
    overlap = top - b[1] if b[1] >= a[1] else top - a[1]
    return 1.0 - min(overlap / float(a[2] - a[1]), overlap / float(b[2] - b[1]))


def cluster_labels(condensed, threshold):
    """linkage(method="complete") + fcluster(criterion="distance") with the installed scipy."""
    from scipy.cluster.hierarchy import fcluster, linkage
    return list(fcluster(linkage(np.asarray(condensed, dtype=np.float64), method="complete"), threshold,
                         criterion="distance"))


def walk(segments, read_len, l_seq, p, contig_len, rank, hap, aln_idx):
    """analyze_read_segments on (q_start, q_end, tid, ref_start, ref_end, rev) tuples, primary first."""
    rows = []
    segs = sorted(segments, key=lambda s: (s[0], s[1]))                # SVIM_inter.py:83
    n_contig = len(contig_len)
    tandem, trans, inversions = [], [], []

    def ordinal():
        return (aln_idx << 32) | 0x80000000 | len(rows)

    def add_bnd(t1, p1, f1, t2, p2, f2):
        r = _blank(hap, aln_idx, ordinal())
        r["type"] = BND
        st, sp, sf, dt, dp, df = bnd_fields(contig_len, rank, t1, p1, f1, t2, p2, f2)
        r["src_tid"], r["src_start"], r["dst_tid"], r["dst_start"] = st, sp, dt, dp
        r["flags"] = (F_SRC_FWD if sf else 0) | (F_DST_FWD if df else 0)
        rows.append(r)
        trans.append((f1, f2, t1, p1, t2, p2))

    def add_ins(tid, start, end, s0, s1):
        if l_seq == 0:
            raise OracleAbort("query_sequence is None")
        r = _span_row(INS, tid, start, end, contig_len, hap, aln_idx, ordinal(), dest=True)
        r["seq_pos"], r["seq_len"] = _pyslice(s0, s1, l_seq)
        rows.append(r)

    mn, mx = p.min_sv_size, p.max_sv_size
    for cur, nxt in zip(segs, segs[1:]):
        cq0, cq1, ct, cr0, cr1, crev = cur
        nq0, nq1, nt, nr0, nr1, nrev = nxt
        if not (0 <= ct < n_contig and 0 <= nt < n_contig):
            raise OracleAbort("reference_id out of range")              # get_reference_name raises
        gap = nq0 - cq1                                                  # :95
        if ct == nt:
            if crev == nrev:
                dref = (cr0 - nr1) if crev else (nr0 - cr1)              # :103-106
                if gap < -p.query_overlap_tolerance:
                    continue
                if dref >= -p.reference_overlap_tolerance:
                    dev = gap - dref
                    if dev >= mn:
                        if dref <= p.reference_gap_tolerance:
                            if not crev:
                                add_ins(ct, cr1, cr1 + dev, cq1, cq1 + dev)                     # :117-118
                            else:
                                add_ins(ct, cr0, cr0 + dev, read_len - nq0, read_len - nq0 + dev)  # :120-121
                    elif -mx <= dev <= -mn:
                        if gap <= p.query_gap_tolerance:
                            anchor = nr1 if crev else cr1
                            rows.append(_span_row(DEL, ct, anchor, anchor - dev, contig_len, hap, aln_idx, ordinal()))
                    elif dev < -mx:
                        if gap <= p.query_gap_tolerance:
                            if not crev:
                                add_bnd(ct, cr1 - 1, True, ct, nr0, True)
                            else:
                                add_bnd(ct, cr0, False, ct, nr1 - 1, False)
                elif gap <= p.query_gap_tolerance:                       # reference overlap (:141-168)
                    dev = gap - dref
                    if dev >= mn:
                        if not crev:
                            if nr1 > cr0:
                                tandem.append((ct, nr0, nr0 + dev, True, True))
                            elif dref >= -mx:
                                tandem.append((ct, nr0, nr0 + dev, False, True))
                            else:
                                add_bnd(ct, cr1 - 1, True, ct, nr0, True)
                        else:
                            if nr0 < cr1:
                                tandem.append((ct, cr0, cr0 + dev, True, False))
                            elif dref >= -mx:
                                tandem.append((ct, cr0, cr0 + dev, False, False))
                            else:
                                add_bnd(ct, cr0, False, ct, nr1 - 1, False)
            else:
                window = -p.query_overlap_tolerance <= gap <= p.query_gap_tolerance
                if not window:
                    continue
                if not crev:                                             # forward -> reverse (:172-193)
                    dev = gap - (nr1 - cr1)
                    if nr0 - cr1 >= -p.reference_overlap_tolerance:
                        if mn <= -dev <= mx:
                            inversions.append((ct, cr1, cr1 - dev, 0))
                        else:
                            add_bnd(ct, cr1 - 1, True, ct, nr1 - 1, False)
                    elif cr0 - nr1 >= -p.reference_overlap_tolerance:
                        if mn <= dev <= mx:
                            inversions.append((ct, nr1, nr1 + dev, 0))
                        else:
                            add_bnd(ct, cr1 - 1, True, ct, nr1 - 1, False)
                else:                                                    # reverse -> forward (:198-219)
                    dev = gap - (nr0 - cr0)
                    if nr0 - cr1 >= -p.reference_overlap_tolerance:
                        if mn <= -dev <= mx:
                            inversions.append((ct, cr0, cr0 - dev, 1))
                        else:
                            add_bnd(ct, cr0, False, ct, nr0, True)
                    elif cr0 - nr1 >= -p.reference_overlap_tolerance:
                        if mn <= dev <= mx:
                            inversions.append((ct, nr0, nr0 + dev, 1))
                        else:
                            add_bnd(ct, cr0, False, ct, nr0, True)
        elif -p.query_overlap_tolerance <= gap <= p.query_gap_tolerance:    # different contigs (:224-254)
            if crev == nrev:
                if not crev:
                    add_bnd(ct, cr1 - 1, True, nt, nr0, True)
                else:
                    add_bnd(ct, cr0, False, nt, nr1 - 1, False)
            elif not crev:
                add_bnd(ct, cr1 - 1, True, nt, nr1 - 1, False)
            else:
                add_bnd(ct, cr0, False, nt, nr0, True)

    # tandem duplications (:260-290); exact means via integer sums
    if tandem:
        def flush(run):
            n = len(run)
            s = sum(t[1] for t in run)
            e = sum(t[2] for t in run)
            r = _span_row(DUP_TAN, run[0][0], _trunc_div(s, n), _trunc_div(e, n), contig_len, hap, aln_idx, ordinal())
            r["copies"] = n
            r["flags"] = F_FULLY if any(t[3] for t in run) else 0
            rows.append(r)
        first_dir = tandem[0][4]                       # QUIRK: never updated afterwards (:273 vs :283-287)
        run = [tandem[0]]
        for t in tandem[1:]:
            n = len(run)
            s = sum(x[1] for x in run)
            e = sum(x[2] for x in run)
            close = run[0][0] == t[0] and abs(s - n * t[1]) < 20 * n and abs(e - n * t[2]) < 20 * n
            if close and first_dir == t[4]:
                run.append(t)
            else:
                flush(run)
                run = [t]
        flush(run)

    # interspersed duplications (:292-320)
    for i, (td1, td2, tt1, tp1, tt2, tp2) in enumerate(trans):
        for bd1, bd2, bt1, bp1, bt2, bp2 in trans[:i]:
            if bd1 != td2 or bd2 != td1:
                continue
            if not (bt1 == tt2 and abs(bp1 - tp2) < 20):
                continue
            if bt2 != tt1 or bd2 != bd1:
                continue
            if bd1:
                length = tp1 + 1 - bp2
                src0, src1, mid = bp2, tp1 + 1, (bp1 + 1 + tp2)
            else:
                length = bp2 + 1 - tp1
                src0, src1, mid = tp1, bp2 + 1, (bp1 + tp2 + 1)
            if not (mn <= length <= mx):
                continue
            m = _trunc_div(mid, 2)
            if src1 < src0 or length < 0:
                raise OracleAbort("interspersed duplication end smaller than start")
            r = _blank(hap, aln_idx, ordinal())
            r["type"] = DUP_INT
            r["src_tid"], r["src_start"], r["src_end"] = bt2, max(0, src0), min(contig_len[bt2], src1)
            r["dst_tid"], r["dst_start"], r["dst_end"] = bt1, max(0, m), min(contig_len[bt1], m + length)
            rows.append(r)

    # inversions (:322-338, :42-60)
    def flush_inv(active):
        if not active:
            return
        if len(active) < 2:
            groups = [active]
        else:
            cond = [_inv_distance(active[i], active[j]) for i in range(len(active) - 1) for j in range(i + 1, len(active))]
            labels = cluster_labels(cond, 0.3)
            groups = [[] for _ in range(max(labels))]
            for item, lab in zip(active, labels):
                groups[lab - 1].append(item)
        for g in groups:
            r = _span_row(INV, g[0][0], max(i[1] for i in g), min(i[2] for i in g), contig_len, hap, aln_idx, ordinal())
            r["flags"] = F_COMPLETE if len(g) > 1 else 0
            rows.append(r)

    active = []
    for item in sorted(inversions, key=lambda v: (rank[v[0]], v[1], v[2])):
        if not active:
            active.append(item)
        elif item[0] == active[-1][0] and item[1] < max(i[2] for i in active):
            active.append(item)
        else:
            flush_inv(active)
            active = []                                # QUIRK: `item` itself is dropped (:333-336)
    flush_inv(active)
    return rows


# ---------------------------------------------------------------------------------------------
# a1/a2/a4: collect over a record image


def _rank(names):
    order = sorted(range(len(names)), key=lambda i: names[i])
    rank = [0] * len(names)
    for r, i in enumerate(order):
        rank[i] = r
    return rank


def _clip_lengths(ops, l_seq):
    """pysam query_alignment_start / query_alignment_end."""
    qas = 0
    for v in ops:
        code = v & 15
        if code == 5:
            continue
        if code == 4:
            qas += v >> 4
        else:
            break
    qae = l_seq
    if qae == 0:
        for v in ops:
            code = v & 15
            if code in (0, 1, 7, 8) or (code == 4 and qae == 0):
                qae += v >> 4
    else:
        for v in ops[:0:-1]:
            code = v & 15
            if code == 5:
                continue
            if code == 4:
                qae -= v >> 4
            else:
                break
    return qas, qae


def collect(host, p, hap=0, scan=None):
    """analyze_alignment_file_coordsorted (SVIM_COLLECT.py:61-83) over a record image.

    `host` needs: hdr (tid,pos,flag,mapq,n_cigar,cigar_off,l_seq,sa_first), cigar, seg, sa_count,
    contig_names, contig_lengths.  Returns a structured array of rows in the reference's order."""
    scan = scan or scan_numpy
    clen = [int(x) for x in host.contig_lengths]
    rank = _rank(list(host.contig_names))
    hdr = host.hdr
    rows = []
    n = hdr.shape[0]
    # bam.fetch(contig) in header order: records grouped by tid, file order inside (stable)
    order = np.argsort(hdr["tid"], kind="stable") if n and np.any(np.diff(hdr["tid"].astype(np.int64)) < 0) else range(n)
    for i in order:
        h = hdr[i]
        tid, flag = int(h["tid"]), int(h["flag"])
        if tid < 0:
            continue                                   # never returned by fetch(contig)
        if flag & 0x4 or flag & 0x100 or int(h["mapq"]) < p.min_mapq:    # :71
            continue
        lo = int(h["cigar_off"])
        ops = host.cigar[lo:lo + int(h["n_cigar"])]
        pos, l_seq = int(h["pos"]), int(h["l_seq"])
        for k, (pr, pq, ln, kind) in enumerate(scan(ops, p.min_sv_size)):
            # ordinal: the GPU derives it from the op index; only the ORDER matters for comparisons
            if kind == "DEL":
                r = _span_row(DEL, tid, pos + pr, pos + pr + ln, clen, hap, i, (int(i) << 32) | k)
                r["seq_pos"] = pq                    # pos_read travels with the row (no sequence: seq_len 0)
                rows.append(r)
            else:
                if l_seq == 0:
                    raise OracleAbort("query_sequence is None")
                r = _span_row(INS, tid, pos + pr, pos + pr + ln, clen, hap, i, (int(i) << 32) | k, dest=True)
                r["seq_pos"], r["seq_len"] = _pyslice(pq, pq + ln, l_seq)
                rows.append(r)
        if flag & 0x800:                               # supplementary records: indels only (:73-74)
            continue
        n_sa = int(host.sa_count[i]) if host.sa_count.shape[0] else 0
        code = ops & 15
        if n_sa == 0 or int((ops[code == 5] >> 4).sum()) > 0:           # no SA / hard clips (:11-16)
            continue
        lens = (ops >> 4).astype(np.int64)
        ref_span = int(lens[(code == 0) | (code == 2) | (code == 3) | (code == 7) | (code == 8)].sum())
        read_len = int(lens[(code == 0) | (code == 1) | (code == 4) | (code == 7) | (code == 8) | (code == 5)].sum())
        qas, qae = _clip_lengths(ops.tolist(), l_seq)
        rev = bool(flag & 0x10)
        prim = (read_len - qae, read_len - qas, tid, pos, pos + (ref_span or 1), rev) if rev else \
               (qas, qae, tid, pos, pos + (ref_span or 1), rev)
        segments = [prim]
        first = int(h["sa_first"])
        for g in host.seg[first:first + n_sa]:
            if int(g["mapq"]) < p.min_mapq:             # :77
                continue
            if g["is_reverse"]:
                segments.append((int(g["read_len"]) - int(g["q_aend"]), int(g["read_len"]) - int(g["q_astart"]),
                                 int(g["tid"]), int(g["pos"]), int(g["ref_end"]), True))
            else:
                segments.append((int(g["q_astart"]), int(g["q_aend"]), int(g["tid"]), int(g["pos"]),
                                 int(g["ref_end"]), False))
        if len(segments) > 1:
            rows.extend(walk(segments, read_len, l_seq, p, clen, rank, hap, int(i)))
    out = np.zeros(len(rows), dtype=ROW_DTYPE)
    for k, r in enumerate(rows):
        out[k] = r
    return out


# ---------------------------------------------------------------------------------------------
# a11-a16: diploid pairing


def edit_distance(a, b):
    lib = clib()
    if lib is not None:
        return int(lib.orc_edit_distance(a, len(a), b, len(b)))
    if len(a) < len(b):
        a, b = b, a
    if not b:
        return len(a)
    bv = np.frombuffer(b, dtype=np.uint8)
    idx = np.arange(len(b) + 1, dtype=np.int64)
    prev = idx.copy()
    for i, ch in enumerate(a, 1):
        row = np.empty_like(prev)
        row[0] = i
        row[1:] = np.minimum(prev[:-1] + (bv != ch), prev[1:] + 1)
        prev = np.minimum.accumulate(row - idx) + idx
    return int(prev[-1])


_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def key_of(r):
    """Candidate.get_key (SVCandidate.py:17-19,147-148,292-293,386-387) -> (contig tid, position)."""
    t = int(r["type"])
    if t in (DEL, INV, DUP_TAN):
        return int(r["src_tid"]), (int(r["src_start"]) + int(r["src_end"])) // 2
    if t in (INS, DUP_INT):
        return int(r["dst_tid"]), int(r["dst_start"])
    return int(r["src_tid"]), int(r["src_start"])


def haplotype_string(r, fetch, clen, region, seq):
    """One side of compute_distance (SVIM_COMBINE.py:43-100): the window with the variant applied."""
    t = int(r["type"])
    lo, hi = region
    if t in (DEL, INV, DUP_TAN):
        tid, s, e = int(r["src_tid"]), int(r["src_start"]), int(r["src_end"])
        left, right = fetch(tid, lo, s), fetch(tid, e, hi)
        if t == DEL:
            mid = b""
        elif t == INV:
            mid = fetch(tid, s, e)[::-1].translate(_COMP)
        else:
            mid = fetch(tid, s, e) * (int(r["copies"]) + 1)
        return left + mid + right
    tid, s = int(r["dst_tid"]), int(r["dst_start"])
    mid = seq if t == INS else fetch(int(r["src_tid"]), int(r["src_start"]), int(r["src_end"]))
    return fetch(tid, lo, s) + mid + fetch(tid, s, hi)


def compute_distance(r1, r2, fetch, clen, seq1, seq2):
    t = int(r1["type"])
    if t in (DEL, INV, DUP_TAN):
        tid = int(r1["src_tid"])
        lo = max(0, min(int(r1["src_start"]), int(r2["src_start"])) - 100)
        hi = min(clen[tid], max(int(r1["src_end"]), int(r2["src_end"])) + 100)
    else:
        tid = int(r1["dst_tid"])
        lo = max(0, min(int(r1["dst_start"]), int(r2["dst_start"])) - 100)
        hi = min(clen[tid], max(int(r1["dst_start"]), int(r2["dst_start"])) + 100)
    return edit_distance(haplotype_string(r1, fetch, clen, (lo, hi), seq1),
                         haplotype_string(r2, fetch, clen, (lo, hi), seq2))


def pair(rows1, rows2, host1, host2, ref_fetch, p):
    """pair_candidates (SVIM_COMBINE.py:164-366).  ref_fetch(tid, start, end) -> upper-cased bytes,
    python-slice semantics of FastaFile.fetch ('' when start >= end, clipped at the contig end)."""
    clen = [int(x) for x in host1.contig_lengths]
    rank = _rank(list(host1.contig_names))
    hosts = {1: host1, 2: host2}
    out = []

    def ins_seq(hap, r):
        if int(r["type"]) != INS:
            return b""
        return hosts[hap].sequence_slice(int(r["aln_idx"]), int(r["seq_pos"]), int(r["seq_len"])).encode("ascii")

    def fetch(tid, s, e):
        s = max(0, s)
        e = min(clen[tid], e)
        return ref_fetch(tid, s, e) if e > s else b""

    for t in (DEL, INV, INS, DUP_TAN, DUP_INT, BND):
        items = [(1, r) for r in rows1 if r["type"] == t] + [(2, r) for r in rows2 if r["type"] == t]
        items.sort(key=lambda it: (rank[key_of(it[1])[0]], key_of(it[1])[1]))        # form_partitions :17 (stable)
        partitions, cur = [], []
        for it in items:
            if cur:
                k0, k1 = key_of(cur[-1][1]), key_of(it[1])
                if k0[0] != k1[0] or abs(k0[1] - k1[1]) > p.partition_max_distance:
                    partitions.append(cur)
                    cur = []
            cur.append(it)
        if cur:
            partitions.append(cur)
        for part in partitions:
            n = len(part)
            if n < 2:
                clusters = [part]
            elif n > 10:
                continue                                                              # dropped (:126-128,:151-152)
            else:
                cond = []
                for i in range(n - 1):
                    for j in range(i + 1, n):
                        (h1, a), (h2, b) = part[i], part[j]
                        if t == BND:                                                  # :105-117
                            same_dirs = (a["flags"] & (F_SRC_FWD | F_DST_FWD)) == (b["flags"] & (F_SRC_FWD | F_DST_FWD))
                            if h1 != h2 and same_dirs:
                                d = (abs(int(a["src_start"]) - int(b["src_start"])) +
                                     abs(int(a["dst_start"]) - int(b["dst_start"]))) / 3000
                            else:
                                d = 99999
                        elif h1 == h2:
                            d = 1000000000                                            # :40-41
                        else:
                            d = compute_distance(a, b, fetch, clen, ins_seq(h1, a), ins_seq(h2, b))
                        cond.append(d)
                labels = cluster_labels(cond, 0.3 if t == BND else p.max_edit_distance)
                clusters = [[] for _ in range(max(labels))]
                for it, lab in zip(part, labels):
                    clusters[lab - 1].append(it)
            for cl in clusters:
                if len(cl) not in (1, 2):
                    continue                                                          # logged as error, skipped
                hap0, first = cl[0]
                r = first.copy()
                r["hap"] = hap0
                r["mate_aln"] = NO_MATE
                if len(cl) == 1:
                    r["genotype"] = GT_HAP1 if hap0 == 1 else GT_HAP2
                else:
                    second = cl[1][1]
                    r["genotype"] = GT_HOM
                    r["mate_aln"] = int(second["aln_idx"])
                    r["flags"] = int(first["flags"]) | (int(second["flags"]) & (F_COMPLETE | F_FULLY | F_CUTPASTE)) \
                        if t != BND else int(first["flags"])
                    if t == DUP_TAN:
                        r["copies"] = round((int(first["copies"]) + int(second["copies"])) / 2)   # banker's (:290)
                if t == BND:                                                          # re-run of the constructor (:342-363)
                    st, sp, sf, dt, dp, df = bnd_fields(clen, rank, int(r["src_tid"]), int(r["src_start"]),
                                                        bool(r["flags"] & F_SRC_FWD), int(r["dst_tid"]),
                                                        int(r["dst_start"]), bool(r["flags"] & F_DST_FWD))
                    r["src_tid"], r["src_start"], r["dst_tid"], r["dst_start"] = st, sp, dt, dp
                    r["flags"] = (F_SRC_FWD if sf else 0) | (F_DST_FWD if df else 0)
                out.append(r)
    res = np.zeros(len(out), dtype=ROW_DTYPE)
    for k, r in enumerate(out):
        res[k] = r
        res[k]["ordinal"] = k
    return res


def collect_indels_vectorised(host, p, hap=0):
    """The indel rows of collect() for a whole record image at once (numpy prefix sums over the flat op array);
    same rows, same order, usable at whole-genome size where the per-record loop would take minutes."""
    hdr = host.hdr
    n = hdr.shape[0]
    ops = np.asarray(host.cigar, dtype=np.uint32)
    code = ops & 15
    ln = (ops >> 4).astype(np.int64)
    adv_ref = np.where((code == 0) | (code == 2) | (code == 7) | (code == 8), ln, 0)
    adv_read = np.where((code == 0) | (code == 1) | (code == 4) | (code == 7) | (code == 8), ln, 0)
    cum_ref = np.cumsum(adv_ref) - adv_ref
    cum_read = np.cumsum(adv_read) - adv_read
    hit = np.nonzero(((code == 1) | (code == 2)) & (ln >= p.min_sv_size))[0]
    starts = hdr["cigar_off"].astype(np.int64)
    aln = np.searchsorted(starts, hit, side="right") - 1
    # pad ops sit at the end of a run: searchsorted on the run starts is enough
    flag, mapq, tid = hdr["flag"][aln], hdr["mapq"][aln].astype(np.int64), hdr["tid"][aln].astype(np.int64)
    ok = (tid >= 0) & ((flag & 0x4) == 0) & ((flag & 0x100) == 0) & (mapq >= p.min_mapq)
    hit, aln, tid = hit[ok], aln[ok], tid[ok]
    base = starts[aln]
    pos_ref = cum_ref[hit] - cum_ref[base]
    pos_read = cum_read[hit] - cum_read[base]
    length = ln[hit]
    is_del = code[hit] == 2
    start = hdr["pos"][aln].astype(np.int64) + pos_ref
    clen = np.asarray(host.contig_lengths, dtype=np.int64)[tid]
    cs, ce = np.maximum(0, start), np.minimum(clen, start + length)
    l_seq = hdr["l_seq"][aln].astype(np.int64)
    out = np.zeros(hit.shape[0], dtype=ROW_DTYPE)
    out["type"] = np.where(is_del, DEL, INS)
    out["hap"] = hap
    out["src_tid"] = np.where(is_del, tid, -1)
    out["src_start"] = np.where(is_del, cs, 0)
    out["src_end"] = np.where(is_del, ce, 0)
    out["dst_tid"] = np.where(is_del, -1, tid)
    out["dst_start"] = np.where(is_del, 0, cs)
    out["dst_end"] = np.where(is_del, 0, ce)
    out["aln_idx"] = aln
    out["seq_pos"] = pos_read
    out["seq_len"] = np.where(is_del | (pos_read >= l_seq), 0, np.minimum(length, l_seq - pos_read))
    out["mate_aln"] = NO_MATE
    out["ordinal"] = (aln.astype(np.uint64) << np.uint64(32)) | (hit - base).astype(np.uint64)
    return out
