"""Pins the travelling CPU oracle (oracle/port.py) to the UNMODIFIED reference, run in place through oracle/shims.
Only runs where /root/reference exists (the build container); everywhere else the committed golden fixtures
(tests/test_golden_cpu.py) carry the same guarantee."""
import os

import numpy as np
import pytest

from oracle import port, refrun
from svim_asm_b200 import bamio, synth
from tests import util


def _cfg(seed, multi):
    return synth.SynthConfig(["chr1", "chr10", "chr2"] if multi else ["chrA"], [400000, 300000, 350000] if multi else [900000],
                             60, 4e4, seed, sv_per_event=5e-3, split_fraction=0.5)


@pytest.mark.reference
@pytest.mark.parametrize("seed", [1, 2, 5, 8])
def test_collect_equals_reference(tmp_path, built_library, seed):
    from svim_asm_b200.engine import HostBatch
    rb = synth.make_haploid(_cfg(seed, seed % 2 == 0))
    bam = str(tmp_path / "h.bam")
    bamio.write_bam(bam, rb)
    opts = refrun.parse_options(["haploid", str(tmp_path / "out"), bam, "ref.fa"])
    cands, _ = refrun.collect(bam, opts)
    want = [refrun.canon(c) for c in cands]
    for host in (HostBatch.from_record_batch(rb), HostBatch.from_bam(bam)):      # array path and C++ ingest path
        rows = port.collect(host, port.Params())
        assert [util.canon_row(r, {0: host}, host.contig_names) for r in rows] == want
    assert len(want) > 30


@pytest.mark.reference
@pytest.mark.parametrize("seed,extra", [(31, []), (77, ["--partition_max_distance", "3000", "--max_edit_distance", "150"])])
def test_pair_equals_reference(tmp_path, built_library, oracle_clib, seed, extra):
    from svim_asm_b200.engine import HostBatch
    cfg = (synth.SynthConfig(["chrA"], [120000], 12, 1.5e4, 77, sv_per_event=6e-2, split_fraction=0.3, sv_max=300) if seed == 77
           else synth.SynthConfig(["chr1", "chr10", "chr2"], [400000, 300000, 350000], 90, 6e4, seed, sv_per_event=6e-3,
                                  split_fraction=0.5, sv_max=3000))
    rb1, rb2 = synth.make_diploid(cfg)
    ref = synth.random_reference(cfg)
    b1, b2, fa = str(tmp_path / "h1.bam"), str(tmp_path / "h2.bam"), str(tmp_path / "ref.fa")
    bamio.write_bam(b1, rb1)
    bamio.write_bam(b2, rb2)
    bamio.write_fasta(fa, ref, cfg.contig_names)
    mods = refrun.modules()
    opts = refrun.parse_options(["diploid", str(tmp_path / "out"), b1, b2, fa] + extra)
    c1, bam1 = refrun.collect(b1, opts)
    c2, _ = refrun.collect(b2, opts)
    paired = mods["SVIM_COMBINE"].pair_candidates(c1, c2, mods["pysam"].FastaFile(fa), bam1, opts)
    want = [refrun.canon(c) for c in paired]
    h1, h2 = HostBatch.from_record_batch(rb1), HostBatch.from_record_batch(rb2)
    p = port.Params(partition_max_distance=opts.partition_max_distance, max_edit_distance=opts.max_edit_distance)
    rows = port.pair(port.collect(h1, p, 1), port.collect(h2, p, 2), h1, h2,
                     lambda tid, s, e: ref[cfg.contig_names[tid]][s:e].tobytes(), p)
    assert [util.canon_row(r, {1: h1, 2: h2}, cfg.contig_names) for r in rows] == want
    assert len(want) >= 20


@pytest.mark.reference
def test_reference_unit_tests_pass_on_the_shims():
    """The reference's own live tests (src/tests/test_intra.py, test_inter.py, test_satag.py) under the shims;
    test_satag_extraction_complete is stale in the reference itself and is expected to fail."""
    import subprocess
    import sys
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(os.path.dirname(refrun.HERE), "oracle", "shims"),
                                                       refrun.REFERENCE_SRC]))
    out = subprocess.run([sys.executable, "-m", "unittest", "tests.test_intra", "tests.test_inter", "tests.test_satag"],
                         cwd="/tmp", env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "Ran 7 tests" in out and "failures=1" in out and "test_satag_extraction_complete" in out, out


def test_vectorised_oracle_equals_per_record_oracle(built_library):
    from svim_asm_b200.engine import HostBatch
    for seed in (3, 4):
        host = HostBatch.from_record_batch(synth.make_haploid(_cfg(seed, seed % 2 == 0)))
        a = port.collect(host, port.Params())
        a = a[(a["ordinal"] & np.uint64(0x80000000)) == 0]
        b = port.collect_indels_vectorised(host, port.Params())
        assert util.rows_equal(a, b) is None            # (the per-record port numbers ordinals by indel, not by op: order only)
        assert np.all(np.diff(b["ordinal"].astype(np.int64)) > 0)
