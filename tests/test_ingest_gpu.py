"""Device ingest (csrc/bam_device.cu: BGZF inflate, record chase, field extraction, copies on the GPU) against the host
ingest (csrc/bam_ingest.cpp, zlib): the same BAM file must give the same record image, bit for bit, and the same
candidates.  Stands in for pysam.AlignmentFile + bam.fetch + cigartuples of the reference (svim-asm:63,85-86;
SVIM_COLLECT.py:65; SVIM_intra.py:37), so the host ingest (itself pinned by the reference's BAM fixtures in
tests/test_golden_cpu.py) is the oracle here."""
import numpy as np
import pytest

from svim_asm_b200 import bamio, synth
from svim_asm_b200.engine import HostBatch, make_params

pytestmark = pytest.mark.gpu


def _compare(engine, path):
    want = HostBatch.from_bam(path)
    got, records = HostBatch.from_bam_device(engine, path)
    assert got.contig_names == want.contig_names and got.sort_order == want.sort_order
    assert np.array_equal(got.contig_lengths, want.contig_lengths)
    assert got.hdr.tobytes() == want.hdr.tobytes()
    assert got.seg.tobytes() == want.seg.tobytes() and np.array_equal(got.sa_count, want.sa_count)
    assert np.array_equal(got.seq_off, want.seq_off)
    for i in range(0, want.n_aln, max(1, want.n_aln // 50)):
        assert got.query_name(i) == want.query_name(i) and got.sa_text(i) == want.sa_text(i)
    # the device-resident arrays, downloaded
    assert np.array_equal(got.cigar, want.cigar)
    assert np.array_equal(got.seq4, want.seq4)
    # and what the hot path makes of them
    params = make_params()
    rows_dev = engine.collect(records, params, hap=1).to_numpy()
    rec_host = engine.load_records(want, with_sequences=True)
    rows_host = engine.collect(rec_host, params, hap=1).to_numpy()
    assert rows_dev.tobytes() == rows_host.tobytes()
    return want


@pytest.mark.parametrize("level", [0, 1, 6])
def test_device_ingest_matches_host_ingest(engine, tmp_path, level):
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2"], [400000, 300000, 350000], 80, 6e4, 51 + level, sv_per_event=5e-3,
                            split_fraction=0.5, sv_max=1500)
    rb = synth.make_haploid(cfg)
    path = str(tmp_path / "a.bam")
    bamio.write_bam(path, rb, level=level)
    want = _compare(engine, path)
    assert want.n_aln > 60 and want.seg.shape[0] > 10


def test_device_ingest_long_cigar_tag_and_many_members(engine, tmp_path):
    # an alignment with more than 65535 ops is stored as placeholder + CG:B,I (htslib restores it transparently)
    cfg = synth.config_c2()
    cfg.target_ops = 1.2e6
    cfg.n_aln = 200
    cfg.giant_ops = 120_000
    rb = synth.make_haploid(cfg)
    path = str(tmp_path / "giant.bam")
    bamio.write_bam(path, rb, level=1)
    want = _compare(engine, path)
    assert int(want.hdr["n_cigar"].max()) > 65535


def test_device_ingest_rejects_corrupt_files(engine, tmp_path):
    cfg = synth.SynthConfig(["chrA"], [200000], 20, 1e4, 5)
    rb = synth.make_haploid(cfg)
    path = str(tmp_path / "ok.bam")
    bamio.write_bam(path, rb, level=6)
    raw = bytearray(open(path, "rb").read())
    bad = str(tmp_path / "bad.bam")
    raw[len(raw) // 2] ^= 0x55                      # inside a deflate stream
    raw[len(raw) // 2 + 1] ^= 0xAA
    open(bad, "wb").write(bytes(raw))
    with pytest.raises((RuntimeError, IOError)):
        HostBatch.from_bam_device(engine, bad)
    open(bad, "wb").write(b"not a bam file at all")
    with pytest.raises((RuntimeError, IOError)):
        HostBatch.from_bam_device(engine, bad)
    # the context stays usable
    got, records = HostBatch.from_bam_device(engine, path)
    assert got.n_aln == rb.n_aln


def test_reference_from_fasta_file_matches_host_loader(engine, tmp_path):
    """svb_ref_load_fasta (the file goes to HBM as it is; line terminators dropped and bases upper-cased by a kernel) against
    FastaFile.load_upper + svb_ref_load: same bases, same symbol classes; contig order of the BAM header, a contig the
    FASTA lacks, lower-case and IUPAC letters, a last line shorter than the others."""
    from svim_asm_b200.fasta import FastaFile
    rng = np.random.default_rng(8)
    names = ["chrB", "chrA", "chrC"]
    seqs = {n: bytes(rng.choice(list(b"ACGTacgtNnRY"), size).astype(np.uint8)) for n, size in zip(names, (100003, 61, 7000))}
    path = str(tmp_path / "ref.fa")
    with open(path, "w") as fh, open(path + ".fai", "w") as fai:
        for n in names:
            fh.write(">%s some description\n" % n)
            off = fh.tell()
            s = seqs[n].decode()
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")
            fai.write("%s\t%d\t%d\t60\t61\n" % (n, len(s), off))
    fa = FastaFile(path)
    order = ["chrA", "missing", "chrC", "chrB"]
    bases, offsets = fa.load_upper(order)
    want = engine.load_reference(bases, offsets).to_numpy()
    got = engine.load_reference_fasta(path, fa.fai_rows(order)).to_numpy()
    assert got[0].tobytes() == want[0].tobytes() == b"".join(seqs[n].upper() for n in order if n in seqs)
    assert np.array_equal(got[1], want[1])
