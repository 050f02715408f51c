"""TEST INFRASTRUCTURE -- the oracle's own record image: plain numpy, no product code.

`OracleBatch.from_record_batch` gives oracle/port.py the attribute surface it reads (hdr, cigar, seg, sa_count,
seq4, seq_off, contig names / lengths, sequence_slice) without touching libsvimasm_b200: the SA:Z text is parsed
here in python, following retrieve_other_alignments (reference SVIM_COLLECT.py:8-58) and the pysam properties the
walk reads afterwards (reference_end, query_alignment_start / _end with l_qseq == 0,
infer_read_length).  bench.py's reference arm and cpu_baseline leg run on it, so that nothing of the product is
mapped into the process that times the CPU path; tests/test_oracle_image_cpu.py holds it against the product's
C++ parser (svb_parse_sa) field by field.

`closed_sample` picks the bounded CPU sample of a whole-genome batch: the records of a few contigs plus every
primary elsewhere whose SA tag names one of them, so that the candidates KEYED on those contigs are exactly the ones
the full run produces there (walk-derived candidates can land on a contig other than their primary's).
"""
import re

import numpy as np

HDR_DTYPE = np.dtype([("tid", "<i4"), ("pos", "<i4"), ("flag", "<u2"), ("mapq", "u1"), ("reserved0", "u1"),
                      ("n_cigar", "<u4"), ("cigar_off", "<u8"), ("l_seq", "<u4"), ("sa_first", "<u4")])
SEG_DTYPE = np.dtype([("tid", "<i4"), ("pos", "<i4"), ("is_reverse", "u1"), ("mapq", "u1"), ("reserved0", "<u2"),
                      ("ref_end", "<i4"), ("q_astart", "<i4"), ("q_aend", "<i4"), ("read_len", "<i4"),
                      ("reserved1", "<i4")])
NT16 = "=ACMGRSVTWYHKDBN"
_NT16_LUT = np.frombuffer(NT16.encode("ascii"), dtype=np.uint8)
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=XB])")      # pysam's cigarstring setter keeps every <digits><op> it finds
_OPS = "MIDNSHP=XB"


def parse_sa(text, contig_names):
    """SA:Z value -> list of (tid, pos, is_reverse, mapq, ref_end, q_astart, q_aend, read_len), one per element that
    retrieve_other_alignments turns into a pseudo alignment (SVIM_COLLECT.py:18-55).  Raises ValueError where int() would."""
    tid_of = {}
    for t, n in enumerate(contig_names):
        tid_of.setdefault(n, t)
    out = []
    for element in text.split(";"):
        fields = element.split(",")
        if len(fields) != 6:                                     # :22-23
            continue
        pos, mapq = int(fields[1]), int(fields[4])               # int(): surrounding blanks and a sign are accepted
        int(fields[5])
        if not (0 <= mapq <= 255):                               # OverflowError -> 0 (:42-45)
            mapq = 0
        ref_span = read_len = qas = qae = 0
        lead, overflow = True, False
        for num, op in _CIGAR_RE.findall(fields[3]):
            n, code = int(num), _OPS.index(op)
            if n >= (1 << 28):                                   # OverflowError: logged, element skipped (:48-50)
                overflow = True
                break
            if code in (0, 2, 3, 7, 8):
                ref_span += n
            if code in (0, 1, 4, 7, 8, 5):
                read_len += n
            if lead:
                if code == 4:
                    qas += n
                elif code != 5:
                    lead = False
            if code in (0, 1, 7, 8) or (code == 4 and qae == 0):
                qae += n
        if overflow:
            continue
        p0 = pos - 1                                             # :41
        out.append((tid_of.get(fields[0], -1), p0, 0 if fields[2] == "+" else 1, mapq, p0 + (ref_span if ref_span > 0 else 1),
                    qas, qae, read_len))
    return out


class OracleBatch(object):
    """Flat host image of one BAM file for the oracle (same arrays as the product's HostBatch)."""

    @classmethod
    def from_record_batch(cls, rb):
        self = cls()
        n = rb.n_aln
        self.contig_names = list(rb.contig_names)
        self.contig_lengths = np.ascontiguousarray(rb.contig_lengths, dtype=np.int32)
        hdr = np.zeros(n, dtype=HDR_DTYPE)
        hdr["tid"], hdr["pos"], hdr["flag"], hdr["mapq"] = rb.tid, rb.pos, rb.flag, rb.mapq
        hdr["n_cigar"], hdr["cigar_off"], hdr["l_seq"] = rb.n_cigar, rb.cigar_off[:-1], rb.l_seq
        self.cigar = np.ascontiguousarray(rb.cigar, dtype=np.uint32)
        self.seq4 = np.ascontiguousarray(rb.seq4, dtype=np.uint8)
        self.seq_off = np.ascontiguousarray(rb.seq_off, dtype=np.uint64)
        self.sa_count = np.zeros(n, dtype=np.uint32)
        segs = []
        for i in sorted(rb.sa):
            got = parse_sa(rb.sa[i], self.contig_names)
            self.sa_count[i] = len(got)
            segs.extend(got)
        hdr["sa_first"] = (np.concatenate(([0], np.cumsum(self.sa_count)[:-1])) if n else np.zeros(0)).astype(np.uint32)
        self.hdr = hdr
        self.seg = np.zeros(len(segs), dtype=SEG_DTYPE)
        for k, name in enumerate(("tid", "pos", "is_reverse", "mapq", "ref_end", "q_astart", "q_aend", "read_len")):
            self.seg[name] = [s[k] for s in segs]
        self._names = list(rb.names)
        return self

    @property
    def n_aln(self):
        return int(self.hdr.shape[0])

    @property
    def n_ops(self):
        return int(self.hdr["n_cigar"].sum(dtype=np.uint64))

    def query_name(self, i):
        return self._names[int(i)]

    def sequence_slice(self, i, start, length):
        """query_sequence[start:start+length] of record i (start / length already python-slice normalised)."""
        if length <= 0:
            return ""
        nib0 = 2 * int(self.seq_off[i]) + int(start)
        b0, b1 = nib0 // 2, (nib0 + int(length) + 1) // 2
        raw = self.seq4[b0:b1]
        nib = np.empty(raw.shape[0] * 2, dtype=np.uint8)
        nib[0::2] = raw >> 4
        nib[1::2] = raw & 15
        s = nib0 - 2 * b0
        return _NT16_LUT[nib[s:s + int(length)]].tobytes().decode("ascii")


def closed_sample(rb, tids):
    """Indices (ascending) of the records on contigs `tids` plus every record elsewhere whose SA tag names one of them."""
    wanted = set(int(t) for t in tids)
    names = set(rb.contig_names[t] for t in wanted)
    take = np.isin(rb.tid, list(wanted))
    for i, text in rb.sa.items():
        if not take[i] and any(el.split(",")[0] in names for el in text.split(";")):
            take[i] = True
    return np.nonzero(take)[0]


def key_contig(rows):
    """tid of Candidate.get_key() (SVCandidate.py:17-19,147-148,292-293,386-387): INS / DUP_INT key on the destination."""
    by_dest = (rows["type"] == 2) | (rows["type"] == 4)
    return np.where(by_dest, rows["dst_tid"], rows["src_tid"])


COMPARE_FIELDS = ("type", "flags", "genotype", "hap", "src_tid", "src_start", "src_end", "dst_tid", "dst_start", "dst_end",
                  "copies", "aln_idx", "seq_pos", "seq_len", "mate_aln")


def compare_on_contigs(full_rows, sample_rows, tids, idx1, idx2):
    """Paired rows of the FULL run against the paired rows of the closed sample (records idx1 / idx2 of haplotype 1 / 2),
    both restricted to the candidates keyed on `tids`: partitions never span contigs (SVIM_COMBINE.py:24-26), so the two
    must agree row for row, in order, once the sample's record indices are mapped back.  Returns (rows compared, None) or
    (rows compared, description of the first difference)."""
    tids = np.asarray(sorted(int(t) for t in tids))
    a = full_rows[np.isin(key_contig(full_rows), tids)]
    b = sample_rows[np.isin(key_contig(sample_rows), tids)].copy()
    maps = {1: np.asarray(idx1, dtype=np.int64), 2: np.asarray(idx2, dtype=np.int64)}
    hap = b["hap"].astype(np.int64)
    own = np.where(hap == 2, maps[2][np.minimum(b["aln_idx"], maps[2].shape[0] - 1)] if maps[2].shape[0] else 0,
                   maps[1][np.minimum(b["aln_idx"], maps[1].shape[0] - 1)] if maps[1].shape[0] else 0)
    has_mate = b["mate_aln"] != 0xFFFFFFFF
    m = np.where(has_mate, b["mate_aln"], 0).astype(np.int64)
    mate = np.where(hap == 2, maps[1][np.minimum(m, maps[1].shape[0] - 1)] if maps[1].shape[0] else 0,
                    maps[2][np.minimum(m, maps[2].shape[0] - 1)] if maps[2].shape[0] else 0)
    b["aln_idx"] = own
    b["mate_aln"] = np.where(has_mate, mate, 0xFFFFFFFF)
    if a.shape[0] != b.shape[0]:
        return int(a.shape[0]), "row counts differ: %d (full run) vs %d (sample)" % (a.shape[0], b.shape[0])
    for f in COMPARE_FIELDS:
        bad = np.nonzero(a[f] != b[f])[0]
        if bad.size:
            i = int(bad[0])
            return int(a.shape[0]), "row %d field %s: %r (full run) vs %r (sample)" % (i, f, a[f][i], b[f][i])
    return int(a.shape[0]), None
