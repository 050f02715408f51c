"""Seams of reference src/svim_asm/SVIM_inter.py, computed by the segment_walk kernel (csrc/walk.cuh)."""
import numpy as np

from .SVCandidate import candidates_from_rows
from .SVIM_intra import _collect_single


def is_similar(chr1, start1, end1, chr2, start2, end2):
    """SVIM_inter.py:12-16 (strict inequalities)."""
    return chr1 == chr2 and abs(start1 - start2) < 20 and abs(end1 - end2) < 20


def analyze_read_segments(primary, supplementaries, bam, options):
    """SVIM_inter.py:62-340: candidates from the primary alignment and its (already filtered) SA segments."""
    class _Unfiltered(object):
        pass
    # the caller has filtered the supplementaries by mapq (SVIM_COLLECT.py:77): keep them all here
    opts = _Unfiltered()
    opts.__dict__.update(vars(options))
    opts.min_mapq = 0
    segs = list(supplementaries)
    if not segs:
        return []
    saved = primary.flag
    try:
        primary.flag = saved & ~0x800          # the walk is defined for the primary record
        rows, host = _collect_single(primary, bam, opts, segs)
    finally:
        primary.flag = saved
    rows = rows[(rows["ordinal"] & np.uint64(0x80000000)) != 0]
    return candidates_from_rows(rows, {0: host}, list(bam.references), list(bam.lengths))
