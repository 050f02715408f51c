"""TEST INFRASTRUCTURE -- stand-in for the `edlib` module (absent from this image).

The reference calls `edlib.align(a, b)["editDistance"]` only
(SVIM_COMBINE.py:50,64,76,88,100): edlib defaults mode="NW", task="distance",
k=-1, i.e. the unit-cost global Levenshtein distance, which is a unique
number, so any exact implementation agrees [ext].  Uses the oracle's C helper
when it has been built (oracle/Makefile), else a numpy row DP.
"""
import ctypes
import os

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_so = os.path.join(_here, "..", "_build", "liboracle.so")
_lib = None
if os.path.exists(_so):
    _lib = ctypes.CDLL(_so)
    _lib.orc_edit_distance.restype = ctypes.c_int64
    _lib.orc_edit_distance.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64]


def _levenshtein_numpy(a, b):
    if len(a) < len(b):
        a, b = b, a
    if not b:
        return len(a)
    bv = np.frombuffer(b, dtype=np.uint8)
    idx = np.arange(len(b) + 1, dtype=np.int64)
    prev = idx.copy()
    for i, ch in enumerate(a, 1):
        sub = prev[:-1] + (bv != ch)
        cand = np.minimum(sub, prev[1:] + 1)
        # left-to-right insertion chain: cur[j] = min(cand[j], cur[j-1] + 1)
        row = np.empty_like(prev)
        row[0] = i
        row[1:] = cand
        row = np.minimum.accumulate(row - idx) + idx
        prev = row
    return int(prev[-1])


def align(query, target, mode="NW", task="distance", k=-1, **kwargs):
    if mode != "NW":
        raise NotImplementedError("shim only provides global alignment")
    qa = query.encode("latin-1") if isinstance(query, str) else bytes(query)
    ta = target.encode("latin-1") if isinstance(target, str) else bytes(target)
    if _lib is not None:
        dist = int(_lib.orc_edit_distance(qa, len(qa), ta, len(ta)))
    else:
        dist = _levenshtein_numpy(qa, ta)
    if k >= 0 and dist > k:
        dist = -1
    return {"editDistance": dist, "alphabetLength": len(set(qa) | set(ta)), "locations": [(None, len(ta) - 1)], "cigar": None}
