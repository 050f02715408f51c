"""The oracle's own record image (oracle/hostimage.py: python SA parser restating retrieve_other_alignments,
SVIM_COLLECT.py:8-58) against the product's host image (svb_parse_sa, C++), and the closed CPU sample bench.py's
parity check relies on."""
import numpy as np
import pytest

from oracle import hostimage, port
from svim_asm_b200 import synth
from svim_asm_b200.engine import HostBatch


def _same_image(rb):
    a, b = hostimage.OracleBatch.from_record_batch(rb), HostBatch.from_record_batch(rb)
    for name in ("hdr", "seg"):
        x, y = getattr(a, name), getattr(b, name)
        assert x.shape == y.shape
        for f in x.dtype.names:
            if not f.startswith("reserved"):
                assert np.array_equal(x[f], y[f]), (name, f)
    for name in ("cigar", "sa_count", "seq4", "seq_off", "contig_lengths"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    return a, b


def test_oracle_image_matches_product_image():
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2"], [300000, 200000, 250000], 120, 6e4, 4242, sv_per_event=6e-3,
                            split_fraction=0.5, sv_max=1500)
    rb1, rb2 = synth.make_diploid(cfg)
    for rb in (rb1, rb2):
        a, b = _same_image(rb)
        assert a.seg.shape[0] > 20
        p = port.Params()
        ra, rb_ = port.collect(a, p, hap=1), port.collect(b, p, hap=1)
        assert ra.tobytes() == rb_.tobytes()


@pytest.mark.parametrize("text,want", [
    ("chr2,101,+,10S90M,60,3;", [(1, 100, 0, 60, 190, 10, 100, 100)]),
    ("chr2,101,-,10S90M5H,60,3", [(1, 100, 1, 60, 190, 10, 100, 105)]),
    ("chrZ,5,+,4M,-400,0;chr1,7,+,4M,1,0,extra;", [(-1, 4, 0, 0, 8, 0, 4, 4)]),
    ("chr1, 12 ,+,5H3S20M2D4I1N2=1X7S, 255 ,0", [(0, 11, 0, 255, 37, 3, 30, 42)]),
    ("chr1,3,+,,9,0", [(0, 2, 0, 9, 3, 0, 0, 0)]),
    ("chr1,3,+,300000000M,9,0;chr1,3,*,4M,9,0", [(0, 2, 1, 9, 6, 0, 4, 4)]),
])
def test_sa_parser_cases(text, want):
    names = ["chr1", "chr2"]
    assert hostimage.parse_sa(text, names) == want
    # and the product parser agrees
    import ctypes
    from svim_asm_b200 import _lib
    tmp = np.zeros(8, dtype=_lib.SEG_DTYPE)
    arr = (ctypes.c_char_p * 2)(*[n.encode() for n in names])
    cnt = _lib.lib.svb_parse_sa(text.encode(), arr, 2, tmp.ctypes.data, 8)
    got = [tuple(int(tmp[k][f]) for f in ("tid", "pos", "is_reverse", "mapq", "ref_end", "q_astart", "q_aend", "read_len"))
           for k in range(cnt)]
    assert got == want


def test_sa_parser_raises_like_int():
    with pytest.raises(ValueError):
        hostimage.parse_sa("chr1,x,+,4M,9,0", ["chr1"])


def test_closed_sample_reproduces_full_run_on_its_contigs():
    """Pairing the closed sample gives, on the sample's contigs, exactly the rows of the full run (the property bench.py's
    parity_check and tests/test_scale_gpu.py use at whole-genome size)."""
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2", "chr3"], [300000, 200000, 250000, 150000], 160, 8e4, 777,
                            sv_per_event=6e-3, split_fraction=0.6, sv_max=1200)
    rb1, rb2 = synth.make_diploid(cfg)
    ref = synth.random_reference(cfg)
    names = cfg.contig_names

    def fetch(tid, s, e):
        return ref[names[tid]][s:e].tobytes()
    p = port.Params(max_sv_size=2500)      # keeps the plain-DP edit distances of the oracle short
    f1, f2 = hostimage.OracleBatch.from_record_batch(rb1), hostimage.OracleBatch.from_record_batch(rb2)
    full = port.pair(port.collect(f1, p, hap=1), port.collect(f2, p, hap=2), f1, f2, fetch, p)
    tids = [1, 3]
    i1, i2 = hostimage.closed_sample(rb1, tids), hostimage.closed_sample(rb2, tids)
    assert np.any(~np.isin(rb1.tid[i1], tids)) or np.any(~np.isin(rb2.tid[i2], tids))      # the closure adds records
    s1, s2 = hostimage.OracleBatch.from_record_batch(rb1.subset(i1)), hostimage.OracleBatch.from_record_batch(rb2.subset(i2))
    sample = port.pair(port.collect(s1, p, hap=1), port.collect(s2, p, hap=2), s1, s2, fetch, p)
    n, diff = hostimage.compare_on_contigs(full, sample, tids, i1, i2)
    assert diff is None, diff
    assert n > 20
    # without the closure the comparison must be able to fail: walk rows keyed on the sample's contigs go missing
    broken = full.copy()
    broken["src_start"][np.nonzero(np.isin(hostimage.key_contig(broken), tids))[0][0]] += 1
    assert hostimage.compare_on_contigs(broken, sample, tids, i1, i2)[1] is not None
