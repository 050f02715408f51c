"""Seams of reference src/svim_asm/SVIM_COLLECT.py.

`analyze_alignment_file_coordsorted` is the hot path: the whole BAM image goes to the GPU in one call
(svb_load_records + svb_collect: cigar_scan, segment_walk, ordered merge); the per-record python loop of
the reference (SVIM_COLLECT.py:64-82) does not exist here."""
import logging

from .bamfile import AlignedSegment
from .engine import make_params
from .runtime import get_engine
import numpy as np

from .SVCandidate import TYPE_NAMES, candidates_from_rows, decode_pool


class CandidateList(list):
    """list of Candidate objects that also remembers the device table it came from, so that pair_candidates and the
    VCF writer can stay on the GPU.  The python objects are only built when somebody looks at them (iteration,
    indexing, ...): a file -> variants.vcf run never does.  Mutating the list cuts the link to the device table."""
    table = None           # engine.Table the rows live in
    rows = None            # numpy copy of this list's rows (svb_row)
    row_index = None       # their indices in `table`
    records = None         # record image of a collect (one haplotype)
    records_by_hap = None  # {haplotype slot: record image} -- where the INS rows' query bases are
    host = None
    _pending = None        # arguments of candidates_from_rows until the objects are built
    _built_from = None
    _sequences = None
    _sequence_source = None

    @property
    def sequences(self):
        """decode_pool() of the table's INS rows (device ingest: the query bases stay in HBM and only the inserted ones
        are downloaded, when the objects are built)."""
        if self._sequence_source is not None:
            source, self._sequence_source = self._sequence_source, None
            self._sequences = source()
        return self._sequences

    @classmethod
    def from_rows(cls, rows, hosts, contig_names, contig_lengths, sequences, table, records_by_hap, row_index=None):
        self = cls()
        self.rows, self.table, self.records_by_hap = rows, table, records_by_hap
        self.row_index = np.arange(rows.shape[0], dtype=np.uint32) if row_index is None else row_index
        self._pending = (hosts, contig_names, contig_lengths, sequences)
        return self

    def _materialize(self):
        if self._pending is not None:
            hosts, names, lengths, sequences = self._pending
            self._pending = None
            if callable(sequences):
                sequences = sequences()
            list.extend(self, candidates_from_rows(self.rows, hosts, names, lengths, sequences=sequences))
            self._built_from = (hosts, names, lengths, sequences)

    def _unlink(self):
        self._materialize()
        self.table = self.rows = self.row_index = None

    def __len__(self):
        return int(self.rows.shape[0]) if self._pending is not None else list.__len__(self)

    def __bool__(self):
        return len(self) > 0

    def of_type(self, kind):
        """The candidates of one class, in order (what svim-asm:133-148 does with six list comprehensions); the result
        keeps the link to the device table."""
        if self.rows is None:
            return CandidateList(c for c in self if c.type == kind)
        keep = np.nonzero(self.rows["type"] == TYPE_NAMES.index(kind))[0]
        args = self._pending if self._pending is not None else self._built_from
        out = CandidateList.from_rows(self.rows[keep], args[0], args[1], args[2], args[3], self.table, self.records_by_hap,
                                      self.row_index[keep])
        if self._pending is None:                      # objects exist already: share them
            out._pending = None
            out._built_from = args
            list.extend(out, [list.__getitem__(self, int(i)) for i in keep])
        return out


def _reader(name):
    def method(self, *args, **kwargs):
        self._materialize()
        return getattr(list, name)(self, *args, **kwargs)
    method.__name__ = name
    return method


def _writer(name):
    def method(self, *args, **kwargs):
        self._unlink()
        return getattr(list, name)(self, *args, **kwargs)
    method.__name__ = name
    return method


for _name in ("__iter__", "__getitem__", "__contains__", "__reversed__", "__eq__", "__ne__", "__add__", "__mul__", "__rmul__",
              "__repr__", "__lt__", "__le__", "__gt__", "__ge__", "index", "count", "copy"):
    setattr(CandidateList, _name, _reader(_name))
for _name in ("__setitem__", "__delitem__", "__iadd__", "__imul__", "append", "extend", "insert", "pop", "remove", "sort", "reverse",
              "clear"):
    setattr(CandidateList, _name, _writer(_name))
CandidateList.__hash__ = None


def retrieve_other_alignments(main_alignment, bam):
    """SVIM_COLLECT.py:8-58: pseudo alignments from the SA tag ([] with hard clips or without SA)."""
    if main_alignment.get_cigar_stats()[0][5] > 0:
        return []
    try:
        entries = main_alignment.get_tag("SA").split(";")
    except KeyError:
        return []
    others = []
    for entry in entries:
        fields = entry.split(",")
        if len(fields) != 6:
            continue
        rname, pos, strand, cigar, mapq, nm = fields[0], int(fields[1]), fields[2], fields[3], int(fields[4]), int(fields[5])
        seg = AlignedSegment()
        seg.query_name = main_alignment.query_name
        seg.flag = 2048 if strand == "+" else 2064
        seg.reference_id = bam.get_tid(rname)
        seg.reference_start = pos - 1
        seg.mapping_quality = mapq if 0 <= mapq <= 255 else 0
        try:
            seg.cigarstring = cigar
        except OverflowError:
            logging.error("OverflowError while retrieving supplementary CIGAR string. Read name: {0}, Position: {1}, "
                          "CIGAR: {2}".format(rname, pos, cigar))
            continue
        others.append(seg)
    return others


def analyze_alignment_file_coordsorted(bam, options):
    """SVIM_COLLECT.py:61-83 on the GPU.  `bam` is a svim_asm_b200.bamfile.AlignmentFile."""
    for contig in bam.references:
        logging.info("Processing chromosome {0}...".format(contig))
    eng = get_engine()
    host = bam.host
    records = getattr(bam, "records", None) or eng.load_records(host)     # a device ingest left the records in HBM
    hap = getattr(options, "_haplotype", 0)
    table = eng.collect(records, make_params(options), hap=hap)
    rows = table.to_numpy()
    out = CandidateList.from_rows(rows, {hap: host}, list(bam.references), list(bam.lengths), None, table, {hap: records})
    out.records, out.host = records, host
    if getattr(records, "has_sequences", False):       # device ingest: only the inserted bases come to the host
        def inserted():
            table.gather_sequences(records)
            pool, starts = table.pool_to_numpy()
            return decode_pool(rows, pool, starts)
        out._sequence_source = inserted
        out._pending = out._pending[:3] + ((lambda: {hap: out.sequences}),)
    return out
