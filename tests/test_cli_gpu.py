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


def _permuted_contigs(rb, perm):
    """The same alignments in a BAM whose header lists the contigs in another order (records re-sorted by the new tid)."""
    from svim_asm_b200 import synth
    new_tid_of = np.zeros(len(perm), dtype=np.int32)
    new_tid_of[np.asarray(perm)] = np.arange(len(perm), dtype=np.int32)           # old tid -> new tid
    order = np.argsort(new_tid_of[rb.tid], kind="stable")
    sub = rb.subset(order)
    return synth.RecordBatch([rb.contig_names[t] for t in perm], np.asarray(rb.contig_lengths)[list(perm)].astype(np.int32),
                             new_tid_of[sub.tid], sub.pos, sub.flag, sub.mapq, sub.n_cigar, sub.cigar_off, sub.l_seq, sub.seq_off,
                             sub.cigar, sub.seq4, sub.names, sub.sa)


def test_second_bam_with_another_contig_order(tmp_path, built_library):
    """The reference pairs by contig NAME and takes the lengths from the first BAM (SVIM_COMBINE.py:164-366): a second BAM
    whose @SQ lines come in another order must give the same VCF (rows carry tids of their own file)."""
    from svim_asm_b200 import bamio, synth
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2"], [300000, 200000, 250000], 60, 4e4, 515, sv_per_event=8e-3,
                            split_fraction=0.5, sv_max=900)
    rb1, rb2 = synth.make_diploid(cfg)
    ref = synth.random_reference(cfg)
    d = tmp_path / "in"
    d.mkdir()
    bamio.write_fasta(str(d / "ref.fa"), ref, cfg.contig_names)
    bamio.write_bam(str(d / "h1.bam"), rb1)
    bamio.write_bam(str(d / "h2.bam"), rb2)
    bamio.write_bam(str(d / "h2p.bam"), _permuted_contigs(rb2, [2, 0, 1]))
    same = _run(tmp_path / "a", ["diploid", str(d / "h1.bam"), str(d / "h2.bam"), str(d / "ref.fa"), "--query_names"])
    perm = _run(tmp_path / "b", ["diploid", str(d / "h1.bam"), str(d / "h2p.bam"), str(d / "ref.fa"), "--query_names"])
    assert same.count("\n") > 60
    assert perm == same


def test_fasta_without_a_contig_that_carries_candidates(tmp_path, built_library):
    """reference.fetch raises KeyError for a contig the FASTA lacks (SVIM_COMBINE.py:48); the device copy must not read it
    as an empty contig."""
    from svim_asm_b200 import bamio, cli, synth
    cfg = synth.SynthConfig(["chr1", "chr2"], [200000, 150000], 40, 3e4, 616, sv_per_event=8e-3, split_fraction=0.3, sv_max=600)
    rb1, rb2 = synth.make_diploid(cfg)
    ref = synth.random_reference(cfg)
    d = tmp_path / "in"
    d.mkdir()
    bamio.write_fasta(str(d / "ref.fa"), ref, ["chr1"])
    bamio.write_bam(str(d / "h1.bam"), rb1)
    bamio.write_bam(str(d / "h2.bam"), rb2)
    with pytest.raises(KeyError):
        cli.main(["diploid", str(tmp_path / "out"), str(d / "h1.bam"), str(d / "h2.bam"), str(d / "ref.fa")])
