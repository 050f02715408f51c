"""The slice of pysam.AlignmentFile / AlignedSegment the reference's call sites use,
served from the C++ ingest's flat record image instead of htslib.

The hot path never touches these objects: `analyze_alignment_file_coordsorted` hands the whole
`HostBatch` to the GPU.  They exist so that the per-alignment seams of the reference
(`retrieve_other_alignments`, `analyze_alignment_indel`, `analyze_read_segments`) keep their signatures.
"""
import os
import re

import numpy as np

from .engine import HostBatch

_OPS = "MIDNSHP=XB"
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=XB])")


class AlignedSegment(object):
    """One alignment: either a view of record `index` of a HostBatch, or a free-standing pseudo record
    (what retrieve_other_alignments builds from an SA entry, SVIM_COLLECT.py:33-55)."""

    def __init__(self, host=None, index=None):
        self._host, self._index = host, index
        self._cigar = None
        self._sa = None
        self.query_name = None
        self.flag = 0
        self.reference_id = -1
        self.reference_start = -1
        self.mapping_quality = 0
        self._l_seq = 0
        if host is not None:
            h = host.hdr[index]
            self.query_name = host.query_name(index)
            self.flag, self.reference_id, self.reference_start = int(h["flag"]), int(h["tid"]), int(h["pos"])
            self.mapping_quality, self._l_seq = int(h["mapq"]), int(h["l_seq"])

    is_unmapped = property(lambda self: bool(self.flag & 0x4))
    is_reverse = property(lambda self: bool(self.flag & 0x10))
    is_secondary = property(lambda self: bool(self.flag & 0x100))
    is_supplementary = property(lambda self: bool(self.flag & 0x800))

    @property
    def cigartuples(self):
        if self._cigar is None and self._host is not None:
            h = self._host.hdr[self._index]
            lo = int(h["cigar_off"])
            run = self._host.cigar[lo:lo + int(h["n_cigar"])]
            self._cigar = [(int(v & 15), int(v >> 4)) for v in run]
        return self._cigar or None

    @property
    def cigarstring(self):
        ct = self.cigartuples
        return "".join("%d%s" % (n, _OPS[op]) for op, n in ct) if ct else None

    @cigarstring.setter
    def cigarstring(self, text):
        self._cigar = [(_OPS.index(letter), int(num)) for num, letter in _CIGAR_RE.findall(text or "")]
        if any(n >= (1 << 28) for _, n in self._cigar):
            raise OverflowError("value too large to convert to uint32_t")

    def get_cigar_stats(self):
        bases, blocks = [0] * 11, [0] * 11
        for op, n in self.cigartuples or []:
            bases[op] += n
            blocks[op] += 1
        return bases, blocks

    @property
    def query_sequence(self):
        if self._host is None or self._l_seq == 0:
            return None
        return self._host.sequence_slice(self._index, 0, self._l_seq)

    @property
    def reference_end(self):
        ct = self.cigartuples
        if self.is_unmapped or not ct:
            return None
        span = sum(n for op, n in ct if op in (0, 2, 3, 7, 8))
        return self.reference_start + (span or 1)

    def infer_read_length(self):
        ct = self.cigartuples
        return sum(n for op, n in ct if op in (0, 1, 4, 5, 7, 8)) if ct else None

    @property
    def query_alignment_start(self):
        start = 0
        for op, n in self.cigartuples or []:
            if op == 5:
                continue
            if op != 4:
                break
            start += n
        return start

    @property
    def query_alignment_end(self):
        ct = self.cigartuples or []
        end = self._l_seq
        if end == 0:
            for op, n in ct:
                if op in (0, 1, 7, 8) or (op == 4 and end == 0):
                    end += n
            return end
        for op, n in reversed(ct[1:]):
            if op == 5:
                continue
            if op != 4:
                break
            end -= n
        return end

    def get_tag(self, name):
        if name == "SA" and self._host is not None:
            text = self._host.sa_text(self._index)
            if text is not None:
                return text
        raise KeyError("tag '%s' not present" % name)


class AlignmentFile(object):
    """BAM file opened through the device ingest (svb_bam_open_device: BGZF inflate and record split on the GPU; the
    record image is then already resident for analyze_alignment_file_coordsorted) or, with SVIM_ASM_B200_INGEST=host
    or when no engine is passed, through the multi-threaded C++ host ingest (svb_bam_open)."""

    def __init__(self, path, mode="rb", threads=0, engine=None):
        self.filename = path
        self.records = None                      # device record image of a device ingest
        if engine is not None and os.environ.get("SVIM_ASM_B200_INGEST", "device") != "host":
            self.host, self.records = HostBatch.from_bam_device(engine, path)
        else:
            self.host = HostBatch.from_bam(path, threads)
        self.references = tuple(self.host.contig_names)
        self.lengths = tuple(int(x) for x in self.host.contig_lengths)
        self.nreferences = len(self.references)
        self._tid = {name: i for i, name in enumerate(self.references)}
        self.header = {"HD": {"SO": self.host.sort_order}} if self.host.sort_order else {"HD": {}}

    def check_index(self):
        stem = os.path.splitext(self.filename)[0]
        if any(os.path.exists(p) for p in (self.filename + ".bai", self.filename + ".csi", stem + ".bai")):
            return True
        raise ValueError("mapping information not recorded in index or index not available")

    def get_tid(self, name):
        return self._tid.get(name, -1)

    def get_reference_name(self, tid):
        if not 0 <= tid < self.nreferences:
            raise ValueError("reference_id %i out of range 0<=tid<%i" % (tid, self.nreferences))
        return self.references[tid]

    getrname = get_reference_name

    def get_reference_length(self, name):
        return self.lengths[self._tid[name]]

    def fetch(self, contig=None, until_eof=False, **kwargs):
        idx = range(self.host.n_aln)
        if contig is not None:
            tid = self._tid[contig]
            idx = np.nonzero(self.host.hdr["tid"] == tid)[0]
        return (AlignedSegment(self.host, int(i)) for i in idx)

    def close(self):
        self.host.close()
