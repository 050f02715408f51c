"""File -> VCF through the drop-in CLI (`svim-asm haploid|diploid`), byte-identical to the VCF the unmodified
reference wrote for the same files (tests/golden/make_golden.py); only the wall-clock ##fileDate line is masked."""
import json
import os

import numpy as np
import pytest

from oracle import port
from tests import util
from tests.test_golden_cpu import GOLDEN, _NpzHost, _as_tuples

pytestmark = pytest.mark.gpu


def _masked(path):
    lines = open(path).read().split("\n")
    return "\n".join("##fileDate=MASKED" if ln.startswith("##fileDate=") else ln for ln in lines)


def _run(tmp_path, argv):
    from svim_asm_b200 import cli
    out = str(tmp_path / "out")
    cli.main([argv[0], out] + argv[1:])
    return _masked(os.path.join(out, "variants.vcf"))


def test_haploid_vcf_is_byte_identical(tmp_path, built_library):
    d = os.path.join(GOLDEN, "haploid")
    got = _run(tmp_path, ["haploid", os.path.join(d, "h.bam"), os.path.join(d, "ref.fa")])
    assert got == open(os.path.join(d, "variants.vcf")).read()


def test_haploid_vcf_with_options(tmp_path, built_library):
    d = os.path.join(GOLDEN, "haploid")
    got = _run(tmp_path, ["haploid", os.path.join(d, "h.bam"), os.path.join(d, "ref.fa"), "--min_sv_size", "50",
                          "--symbolic_alleles", "--query_names", "--tandem_duplications_as_insertions", "--sample", "NA12878"])
    assert got == open(os.path.join(GOLDEN, "haploid_opts", "variants.vcf")).read()


@pytest.mark.parametrize("ingest", ["device", "host"])
def test_diploid_vcf_is_byte_identical(tmp_path, built_library, monkeypatch, ingest):
    monkeypatch.setenv("SVIM_ASM_B200_INGEST", ingest)     # BGZF inflate + record split on the GPU / zlib on the host
    d = os.path.join(GOLDEN, "diploid")
    got = _run(tmp_path, ["diploid", os.path.join(d, "h1.bam"), os.path.join(d, "h2.bam"), os.path.join(d, "ref.fa"),
                          "--query_names"])
    assert got == open(os.path.join(d, "variants.vcf")).read()


@pytest.mark.parametrize("stem", ["chimeric_read", "chimeric_read_errors"])
def test_reference_bam_fixtures_on_gpu(engine, stem):
    from svim_asm_b200.engine import HostBatch, make_params
    z = _NpzHost(os.path.join(GOLDEN, stem + ".npz"))
    host = HostBatch()
    host.contig_names, host.contig_lengths = z.contig_names, np.ascontiguousarray(z.contig_lengths)
    host.hdr, host.cigar, host.seg, host.sa_count = (np.ascontiguousarray(z.hdr), np.ascontiguousarray(z.cigar),
                                                     np.ascontiguousarray(z.seg), np.ascontiguousarray(z.sa_count))
    host.seq4, host.seq_off, host._names = np.ascontiguousarray(z.seq4), np.ascontiguousarray(z.seq_off), z._names
    rec = engine.load_records(host)
    for min_sv in ("40", "2"):
        rows = engine.collect(rec, make_params(min_sv_size=int(min_sv))).to_numpy()
        got = [util.canon_row(r, {0: host}, host.contig_names) for r in rows]
        assert got == _as_tuples(z.expected[min_sv])
