"""bamio.read_reference_names: the CLI peeks at the first BAM's contig names to start the FASTA load early."""
import numpy as np

from svim_asm_b200 import bamio, synth


def _batch(names, lengths):
    cfg = synth.SynthConfig(list(names), list(lengths), 40, 3e4, 77, sv_per_event=5e-3, split_fraction=0.3, sv_max=500)
    return synth.make_haploid(cfg)


def test_names_of_a_small_header(tmp_path):
    rb = _batch(["chr1", "chr10", "chr2"], [90000, 60000, 70000])
    path = str(tmp_path / "a.bam")
    bamio.write_bam(path, rb)
    assert bamio.read_reference_names(path) == ["chr1", "chr10", "chr2"]


def test_header_spanning_several_members(tmp_path):
    """6,000 contigs: the reference table alone is larger than one 64 KB BGZF member."""
    names = ["scaffold_%05d_of_some_assembly" % i for i in range(6000)]
    rb = _batch(names[:3], [90000, 60000, 70000])
    rb = synth.RecordBatch(names, np.asarray([90000, 60000, 70000] + [1000] * (len(names) - 3), dtype=np.int32), rb.tid, rb.pos,
                           rb.flag, rb.mapq, rb.n_cigar, rb.cigar_off, rb.l_seq, rb.seq_off, rb.cigar, rb.seq4, rb.names, rb.sa)
    path = str(tmp_path / "b.bam")
    bamio.write_bam(path, rb)
    assert bamio.read_reference_names(path) == names


def test_not_a_bam(tmp_path):
    p = tmp_path / "x.bam"
    p.write_bytes(b"not a bam file at all, just text\n" * 10)
    assert bamio.read_reference_names(str(p)) is None
    assert bamio.read_reference_names(str(tmp_path / "missing.bam")) is None
    q = tmp_path / "trunc.bam"
    rb = _batch(["chr1", "chr2"], [90000, 60000])
    bamio.write_bam(str(tmp_path / "ok.bam"), rb)
    q.write_bytes((tmp_path / "ok.bam").read_bytes()[:40])
    assert bamio.read_reference_names(str(q)) is None


def test_member_table_in_pieces(tmp_path, built_library):
    """The device ingest walks the list of BGZF member headers in pieces (one thread per range of the file, each finding
    its own first member): same table as the serial walk, on a file large enough for several ranges."""
    import ctypes
    from svim_asm_b200 import _lib
    cfg = synth.SynthConfig(["chr1", "chr2", "chr3"], [4000000, 3000000, 2500000], 700, 1.2e7, 99, sv_per_event=2e-3, split_fraction=0.2, sv_max=800)
    rb = synth.make_haploid(cfg)
    path = str(tmp_path / "big.bam")
    bamio.write_bam(path, rb, level=1)
    import os
    assert os.path.getsize(path) > 4 * (8 << 20)        # four ranges of at least 8 MB
    got = []
    for n_threads in (1, 2, 5, 8):
        out = (ctypes.c_uint64 * 3)()
        assert _lib.lib.svb_bgzf_member_table_check(path.encode(), n_threads, out) == 0
        got.append(tuple(out))
    assert got[0][0] > 600 and all(g == got[0] for g in got)
