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


def _raw_bam(path, names, lengths, records, level=6, block=0xFF00):
    """BAM file from explicit record dicts (refID, pos, flag, mapq, name, cigar [(op, len)], seq4 bytes, l_seq, tags bytes):
    lets the tests put every auxiliary field type in front of the tag walker."""
    import struct
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % nl for nl in zip(names, lengths))
    stream = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(names)))
    for n, ln in zip(names, lengths):
        stream += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", ln)
    for r in records:
        name = r["name"].encode() + b"\0"
        cig = b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in r["cigar"])
        l_seq = r["l_seq"]
        core = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(name), r.get("mapq", 60), 4680, len(r["cigar"]), r.get("flag", 0),
                           l_seq, -1, -1, 0)
        body = core + name + cig + r["seq4"] + b"\xff" * l_seq + r.get("tags", b"")
        stream += struct.pack("<i", len(body)) + body
    with open(path, "wb") as out:
        for lo in range(0, len(stream), block):
            out.write(bamio._bgzf_block(bytes(stream[lo:lo + block]), level))
        out.write(bamio._EOF_BLOCK)


def test_device_ingest_tag_types_unmapped_and_empty_files(engine, tmp_path):
    import struct
    rng = np.random.default_rng(12)
    tags_all = (b"NMC\x05" + b"ASc\xfb" + b"XSs" + struct.pack("<h", -300) + b"YSS" + struct.pack("<H", 60000) + b"XIi" + struct.pack("<i", -5)
                + b"YII" + struct.pack("<I", 4000000000) + b"XFf" + struct.pack("<f", 1.5) + b"XAAq" + b"MDZ10A5^AC6\0" + b"XHH1AE301\0"
                + b"XBBc" + struct.pack("<i", 3) + b"\x01\x02\x03" + b"YBBS" + struct.pack("<i", 2) + struct.pack("<HH", 7, 8)
                + b"ZBBf" + struct.pack("<i", 1) + struct.pack("<f", 2.0))
    sa = b"SAZchrB,500,-,100S300M,60,3;chrA,9000,+,300M100S,13,0;\0"

    def rec(tid, pos, cigar, tags=b"", flag=0, l_seq=None, name="r"):
        n = sum(ln for op, ln in cigar if op in (0, 1, 4, 7, 8)) if l_seq is None else l_seq
        return dict(tid=tid, pos=pos, flag=flag, name=name, cigar=cigar, l_seq=n,
                    seq4=bytes(rng.integers(0, 256, (n + 1) // 2, dtype=np.uint8)), tags=tags)
    records = [rec(0, 100, [(0, 300), (2, 50), (0, 100)], tags_all + sa, name="with_every_tag_type"),
               rec(0, 700, [(4, 100), (0, 300)], sa + tags_all, name="sa_first"),
               rec(0, 900, [(0, 50), (1, 45), (0, 50)], b"", name="no_tags"),
               rec(1, 10, [], b"", l_seq=0, name="no_cigar_no_seq"),                     # '*' CIGAR and sequence
               rec(1, 20, [(0, 10)], tags_all, flag=256, name="secondary"),
               rec(-1, -1, [], b"XAAz", flag=4, l_seq=7, name="unmapped_at_the_end")]
    path = str(tmp_path / "tags.bam")
    _raw_bam(path, ["chrA", "chrB"], [100000, 50000], records, block=137)                  # many tiny members: records span them
    _compare(engine, path)
    # a header-only file, and one whose only member is the header
    _raw_bam(str(tmp_path / "empty.bam"), ["chrA"], [1000], [])
    got, recs = HostBatch.from_bam_device(engine, str(tmp_path / "empty.bam"))
    assert got.n_aln == 0 and got.contig_names == ["chrA"] and len(engine.collect(recs, make_params())) == 0
    # a truncated record and an unknown tag type are refused like on the host
    bad = str(tmp_path / "badtag.bam")
    _raw_bam(bad, ["chrA"], [1000], [rec(0, 1, [(0, 10)], b"XQq\x01", name="bad_type")])
    with pytest.raises((RuntimeError, IOError)):
        HostBatch.from_bam_device(engine, bad)
    with pytest.raises((RuntimeError, IOError)):
        HostBatch.from_bam(bad)


def test_record_chase_from_the_index_and_without(engine, tmp_path, monkeypatch):
    """Record boundaries: in parallel from the record starts the .bai names (every segment must land on the next start),
    serially without an index, and serially again when the index belongs to another file (stale index -> fall-back)."""
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2"], [900000, 700000, 800000], 700, 2.5e5, 97, sv_per_event=3e-3,
                            split_fraction=0.3, sv_max=800)
    rb = synth.make_haploid(cfg)
    path = str(tmp_path / "a.bam")
    bamio.write_bam(path, rb, level=1)
    want = _compare(engine, path)                              # with a.bam.bai: the indexed walk
    assert want.n_aln > 500
    monkeypatch.setenv("SVB_INGEST_SERIAL_CHASE", "1")
    _compare(engine, path)                                     # forced serial walk
    monkeypatch.delenv("SVB_INGEST_SERIAL_CHASE")
    import os
    os.remove(path + ".bai")
    _compare(engine, path)                                     # no index at all
    other = synth.make_haploid(synth.SynthConfig(["chr1", "chr10", "chr2"], [900000, 700000, 800000], 650, 2.4e5, 98, sv_per_event=3e-3,
                                                 split_fraction=0.3, sv_max=800))
    bamio.write_bam(str(tmp_path / "b.bam"), other, level=1)
    os.replace(str(tmp_path / "b.bam.bai"), path + ".bai")     # an index of another file
    _compare(engine, path)
