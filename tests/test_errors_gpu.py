"""Error behaviour through the C ABI mirrors where the reference itself would raise."""
import numpy as np
import pytest

from svim_asm_b200.engine import HostBatch, make_params
from tests import util



@pytest.mark.gpu
def test_unknown_sa_contig_raises_like_get_reference_name(engine):
    # retrieve_other_alignments gives reference_id -1 for an unknown rname (SVIM_COLLECT.py:40); the walk then calls
    # bam.get_reference_name(-1), which raises ValueError in pysam (SVIM_inter.py:99,225)
    recs = [dict(tid=0, pos=1000, cigar=[(0, 5000), (4, 5000)], sa="chrUnknown,20000,+,5000S5000M,60,0;")]
    rb = util.batch_from_records(["c1"], [100000], recs)
    rec = engine.load_records(HostBatch.from_record_batch(rb))
    with pytest.raises(RuntimeError, match="reference_id out of range"):
        engine.collect(rec, make_params())
    # the context stays usable afterwards
    ok = util.batch_from_records(["c1"], [100000], [dict(tid=0, pos=10, cigar=[(0, 100), (2, 50), (0, 100)])])
    assert len(engine.collect(engine.load_records(HostBatch.from_record_batch(ok)), make_params())) == 1


@pytest.mark.gpu
def test_insertion_without_sequence_raises(engine):
    # alignment.query_sequence is None when no sequence is stored: the slice at SVIM_intra.py:42 raises TypeError
    recs = [dict(tid=0, pos=10, cigar=[(0, 100), (1, 60), (0, 100)], l_seq=0)]
    rb = util.batch_from_records(["c1"], [100000], recs)
    rec = engine.load_records(HostBatch.from_record_batch(rb))
    with pytest.raises(RuntimeError):
        engine.collect(rec, make_params())


@pytest.mark.gpu
def test_bad_arguments_are_rejected(engine):
    from svim_asm_b200 import _lib
    import ctypes
    out = ctypes.c_void_p()
    hdr = np.zeros(1, dtype=_lib.HDR_DTYPE)
    hdr["n_cigar"], hdr["cigar_off"] = 3, 2                      # run not 16-byte aligned
    ops = np.full(8, 15, dtype=np.uint32)
    rc = _lib.lib.svb_load_records(engine.handle, hdr.ctypes.data, 1, ops.ctypes.data, 8, None, None, 0,
                                   np.array([100], np.int32).ctypes.data, np.array([0], np.int32).ctypes.data, 1,
                                   ctypes.byref(out))
    assert rc == -1 and b"aligned" in _lib.lib.svb_last_error(engine.handle)


def test_malformed_sa_integer_is_an_error_at_ingest():
    # int(fields[1]) raises ValueError in the reference (SVIM_COLLECT.py:25): the parser reports it instead of guessing
    from svim_asm_b200 import _lib
    import ctypes
    names = (ctypes.c_char_p * 1)(b"c1")
    seg = np.zeros(4, dtype=_lib.SEG_DTYPE)
    assert _lib.lib.svb_parse_sa(b"c1,12x,+,10M,60,0;", names, 1, seg.ctypes.data, 4) == -5
    assert _lib.lib.svb_parse_sa(b"c1,12,+,10M,60,0;c1,5,-,3S7M,300,1;bad;c1,1,+,1M,2,3,4;", names, 1, seg.ctypes.data, 4) == 2
    assert seg[1]["mapq"] == 0 and seg[1]["is_reverse"] == 1 and seg[0]["pos"] == 11
