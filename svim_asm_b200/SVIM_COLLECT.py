"""Seams of reference src/svim_asm/SVIM_COLLECT.py.

`analyze_alignment_file_coordsorted` is the hot path: the whole BAM image goes to the GPU in one call
(svb_load_records + svb_collect: cigar_scan, segment_walk, ordered merge); the per-record python loop of
the reference (SVIM_COLLECT.py:64-82) does not exist here."""
import logging

from .bamfile import AlignedSegment
from .engine import make_params
from .runtime import get_engine
from .SVCandidate import candidates_from_rows, decode_pool


class CandidateList(list):
    """list of Candidate objects that also remembers the device table it was materialised from, so that
    pair_candidates can stay on the GPU."""
    table = None
    records = None
    host = None
    sequences = None       # decode_pool() of the table's INS rows (device ingest: the query bases stay in HBM)


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
    sequences = None
    if getattr(records, "has_sequences", False):       # device ingest: only the inserted bases come to the host
        table.gather_sequences(records)
        pool, starts = table.pool_to_numpy()
        sequences = decode_pool(rows, pool, starts)
    out = CandidateList(candidates_from_rows(rows, {hap: host}, list(bam.references), list(bam.lengths),
                                             sequences={hap: sequences} if sequences is not None else None))
    out.table, out.records, out.host, out.sequences = table, records, host, sequences
    return out
