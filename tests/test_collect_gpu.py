"""GPU parity of svb_collect (K2 cigar_scan + K4 segment_walk + K5 merge) against the CPU oracle.

Reference path under test: SVIM_COLLECT.py:61-83 -> SVIM_intra.py:8-44 + SVIM_inter.py:62-340.
Bit-exact: every integer field of every row, and the row order."""
import numpy as np
import pytest

from oracle import port
from svim_asm_b200 import synth
from svim_asm_b200.engine import HostBatch, make_params
from tests import util

pytestmark = pytest.mark.gpu

KATS = [  # reference src/tests/test_intra.py:8-22
    ([(5, 10), (4, 20), (0, 10), (7, 10), (8, 5), (0, 5), (1, 50), (0, 30), (4, 25), (5, 15)], [(30, 50, 50, "INS")]),
    ([(5, 10), (4, 20), (0, 30), (2, 50), (0, 30), (4, 25), (5, 15)], [(30, 50, 50, "DEL")]),
    ([(5, 10), (4, 20), (0, 30), (2, 40), (1, 50), (0, 30), (4, 25), (5, 15)], [(30, 50, 40, "DEL"), (70, 50, 50, "INS")]),
    ([(5, 10), (4, 20), (0, 30), (1, 40), (2, 50), (0, 30), (4, 25), (5, 15)], [(30, 50, 40, "INS"), (30, 90, 50, "DEL")]),
]


@pytest.mark.parametrize("variant", [0, 1])
def test_reference_known_answers(engine, variant):
    engine.set_scan_variant(variant)
    for tuples, want in KATS:
        assert engine.cigar_indel(tuples, 30) == want
    assert engine.cigar_indel([], 30) == []
    assert engine.cigar_indel([(3, 100), (2, 40)], 40) == [(0, 0, 40, "DEL")]      # N does not advance (App. B#1)
    engine.set_scan_variant(0)


def _compare(engine, rb, variants=(0, 1), **params):
    host = HostBatch.from_record_batch(rb)
    want = port.collect(host, port.Params(**params))
    rec = engine.load_records(host)
    for v in variants:
        engine.set_scan_variant(v)
        got = engine.collect(rec, make_params(**params)).to_numpy()
        diff = util.rows_equal(got, want)
        assert diff is None, (v, diff)
        assert np.all(np.diff(got["ordinal"].astype(np.uint64)) > 0) if got.shape[0] > 1 else True
    engine.set_scan_variant(0)
    rec.free()
    return want


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_small_random_batches(engine, seed):
    multi = seed % 2 == 0
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2"] if multi else ["chrA"],
                            [400000, 300000, 350000] if multi else [900000], 60, 4e4, seed,
                            sv_per_event=5e-3, split_fraction=0.5)
    want = _compare(engine, synth.make_haploid(cfg))
    assert want.shape[0] > 20


def test_many_tiny_alignments(engine):
    # several alignment heads per 1024-op chunk: the generic per-piece path of the scan
    rng = np.random.default_rng(9)
    recs = []
    pos = 10
    for i in range(3000):
        k = int(rng.integers(0, 9))
        cig = []
        for _ in range(k):
            op = int(rng.choice([0, 1, 2, 3, 4, 5, 7, 8]))
            cig.append((op, int(rng.choice([1, 5, 39, 40, 41, 300]))))
        flag = int(rng.choice([0, 16, 4, 256, 2048]))
        recs.append(dict(tid=int(i >= 1500), pos=pos % 400000, flag=flag, mapq=int(rng.choice([60, 19, 20])), cigar=cig,
                         l_seq=100000))
        pos += 97
    recs.sort(key=lambda r: (r["tid"], r["pos"]))
    rb = util.batch_from_records(["c1", "c2"], [500000, 500000], recs)
    want = _compare(engine, rb)
    assert want.shape[0] > 500


def test_dense_events_and_capacity_retry(engine):
    # every op emits: the output is far larger than the first capacity guess
    cig = [(1 if i % 2 else 2, 50) for i in range(20000)]
    recs = [dict(tid=0, pos=100, cigar=cig, l_seq=2_000_000), dict(tid=0, pos=200, cigar=cig[:777], l_seq=100000)]
    rb = util.batch_from_records(["c1"], [3_000_000], recs)
    want = _compare(engine, rb)
    assert want.shape[0] == 20777


def test_contig_end_clamp_and_min_size(engine):
    recs = [dict(tid=0, pos=990, cigar=[(0, 5), (1, 100), (2, 60), (0, 5)], l_seq=110),     # INS/DEL ends clamp to 1000
            dict(tid=0, pos=0, cigar=[(2, 40), (0, 10), (1, 39), (0, 1)], l_seq=50)]
    rb = util.batch_from_records(["c1"], [1000], recs)
    _compare(engine, rb)
    _compare(engine, rb, min_sv_size=1)
    _compare(engine, rb, min_sv_size=0)
    _compare(engine, rb, min_mapq=61)


def test_chr20_scale_with_giant_alignment(engine):
    cfg = synth.config_c2()
    cfg.target_ops = 3.0e6
    cfg.n_aln = 600
    cfg.giant_ops = 200_000                     # > 65535 ops and > 24 tiles: long look-back chains
    rb = synth.make_haploid(cfg)
    want = _compare(engine, rb)
    assert int(rb.n_cigar.max()) > 150_000 and want.shape[0] > 100


@pytest.mark.parametrize("threads", [32, 64])
def test_unit_scan_lookback_with_small_ctas(engine, threads, monkeypatch):
    # the carry / row-base scan over the unit aggregates is a chained scan over CTAs (1024 units each at full size); with
    # 32-thread CTAs a 3M-op batch already runs 45 CTAs, i.e. two look-back windows, with alignments that span many units
    monkeypatch.setenv("SVB_UNIT_SCAN_THREADS", str(threads))
    cfg = synth.config_c2()
    cfg.target_ops = 3.0e6
    cfg.n_aln = 300
    cfg.giant_ops = 400_000
    rb = synth.make_haploid(cfg)
    want = _compare(engine, rb)
    assert want.shape[0] > 100


def test_empty_and_degenerate_batches(engine):
    """No records at all; records without CIGAR ops (unmapped-style, '*'); everything filtered out."""
    empty = util.batch_from_records(["c1"], [1000], [])
    assert _compare(engine, empty).shape[0] == 0
    recs = [dict(tid=0, pos=5, cigar=[]), dict(tid=0, pos=10, cigar=[(0, 50), (2, 45), (0, 5)]), dict(tid=0, pos=20, cigar=[]),
            dict(tid=0, pos=30, cigar=[], flag=4), dict(tid=0, pos=40, cigar=[(1, 60)], l_seq=60)]
    want = _compare(engine, util.batch_from_records(["c1"], [1000], recs))
    assert want.shape[0] == 2
    filtered = [dict(tid=0, pos=10, cigar=[(0, 50), (2, 45), (0, 5)], mapq=3), dict(tid=0, pos=90, cigar=[(2, 100)], flag=256)]
    assert _compare(engine, util.batch_from_records(["c1"], [1000], filtered)).shape[0] == 0


@pytest.mark.parametrize("n_ops", [1023, 1024, 1025, 2047, 2048, 2049, 4 * 1024 * 2, 4 * 1024 * 2 + 1, 16 * 1024 + 3])
def test_alignment_lengths_around_chunk_and_unit_boundaries(engine, n_ops):
    """Runs that end exactly at, one before and one after a 1024-op chunk / a unit of chunks, each followed by a second
    alignment (so a head falls on the boundary), with emitting ops on both sides of every boundary."""
    rng = np.random.default_rng(n_ops)
    def cig(n):
        ops = [(7, int(x)) for x in rng.integers(1, 30, n)]
        for i in (0, 1, n // 2, max(n - 2, 0), n - 1):
            ops[i] = (int(rng.choice([1, 2])), int(rng.integers(40, 90)))
        for i in range(1000, n, 1024):                   # right before / at / after every chunk boundary
            for k in (23, 24, 25):
                if i + k < n:
                    ops[i + k] = (int(rng.choice([1, 2])), 50)
        return ops
    recs = [dict(tid=0, pos=100, cigar=cig(n_ops)), dict(tid=0, pos=5000, cigar=cig(1500)), dict(tid=0, pos=9000, cigar=cig(7))]
    for r in recs:
        r["l_seq"] = sum(ln for op, ln in r["cigar"] if op in (0, 1, 4, 7, 8))
    want = _compare(engine, util.batch_from_records(["c1"], [50_000_000], recs))
    assert want.shape[0] > 10


def test_very_long_ops_do_not_overflow_packed_sums(engine):
    # 28-bit lengths: eight such ops in two rows of one lane still fit the 31-bit field of the packed accumulator
    big = (1 << 28) - 1
    cigar = [(7, big)] * 7 + [(2, 60), (7, 5), (1, 50), (7, big)] + [(7, 1)] * 40
    recs = [dict(tid=0, pos=0, cigar=cigar, l_seq=10)]
    rb = util.batch_from_records(["c1"], [2**31 - 1], recs)
    host = HostBatch.from_record_batch(rb)
    rec = engine.load_records(host)
    got = engine.collect(rec, make_params()).to_numpy()
    want = port.collect(host, port.Params())
    assert util.rows_equal(got, want) is None, util.rows_equal(got, want)
    assert got.shape[0] == 2


def test_collect2_equals_two_collects(engine):
    """svb_collect2 (both haplotypes, one host synchronisation, sequence pools sized on the device) against two svb_collect
    calls + svb_table_gather_sequences: same rows, same pools; also with the query sequences read in place from pinned host
    memory (svb_records_map_sequences_host) instead of uploaded."""
    from svim_asm_b200.bench_util import pinned_host
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2"], [400000, 300000, 350000], 110, 7e4, 808, sv_per_event=7e-3,
                            split_fraction=0.5, sv_max=2500)
    rb1, rb2 = synth.make_diploid(cfg)
    h1, h2 = pinned_host(HostBatch.from_record_batch(rb1)), pinned_host(HostBatch.from_record_batch(rb2))
    params = make_params()
    r1, r2 = engine.load_records(h1, with_sequences=True), engine.load_records(h2, with_sequences=True)
    want = []
    for k, r in ((1, r1), (2, r2)):
        t = engine.collect(r, params, hap=k)
        t.gather_sequences(r)
        want.append((t.to_numpy(), t.pool_to_numpy()))
    assert want[0][0].shape[0] > 60 and np.any(want[0][0]["ordinal"] & np.uint64(0x80000000)) and want[0][1][0].shape[0] > 500
    for mapped in (False, True):
        if mapped:
            r1, r2 = engine.load_records(h1), engine.load_records(h2)
            engine.map_sequences_host(r1)
            engine.map_sequences_host(r2)
        for with_pools in (False, True):
            t1, t2 = engine.collect2(r1, r2, params, with_pools=with_pools)
            for t, (rows, (pool, off)) in zip((t1, t2), want):
                assert t.to_numpy().tobytes() == rows.tobytes()
                if with_pools:
                    got_pool, got_off = t.pool_to_numpy()
                    assert got_pool.tobytes() == pool.tobytes() and np.array_equal(got_off, off)
    # an empty haplotype next to a full one
    empty = HostBatch.from_record_batch(rb1.subset(np.zeros(0, dtype=np.int64)))
    re_ = engine.load_records(empty, with_sequences=True)
    t1, t2 = engine.collect2(r1, re_, params, with_pools=True)
    assert len(t2) == 0 and t1.to_numpy().tobytes() == want[0][0].tobytes()


def test_walk_with_many_segments_per_read(engine):
    """Reads split into many segments: more walk rows per primary than the count pass stages (the write pass recomputes
    those), more segments than the shared-memory scratch holds (global scratch), next to ordinary two-segment reads."""
    names, lengths = ["chr1", "chr2"], [3_000_000, 2_000_000]
    records = []

    def chain(name, n_seg, start, seg=10000, gap=500):
        # n_seg forward segments, each `gap` further along the reference than the read: n_seg - 1 deletions (App. D "DEL fwd")
        L = n_seg * seg
        segs = [(i * seg, (i + 1) * seg, start + i * (seg + gap)) for i in range(n_seg)]

        def cigar(q0, q1):
            ops = ([(4, q0)] if q0 else []) + [(7, q1 - q0)] + ([(4, L - q1)] if L - q1 else [])
            return ops
        sa = ";".join("chr1,%d,+,%s,60,0" % (p + 1, "".join("%d%s" % (ln, "MIDNSHP=X"[op]) for op, ln in cigar(q0, q1)))
                      for q0, q1, p in segs[1:]) + ";"
        records.append(dict(tid=0, pos=segs[0][2], cigar=cigar(*segs[0][:2]), sa=sa, name=name))
    chain("two", 2, 100000)
    chain("twelve", 12, 400000)          # 11 rows > the 8 staged ones, 12 segments > the 8 shared-memory slots
    chain("nine", 9, 900000)             # 8 rows: exactly the staged capacity, 9 segments > 8 slots
    chain("five", 5, 1500000)
    records.sort(key=lambda r: r["pos"])
    rb = util.batch_from_records(names, lengths, records)
    host = HostBatch.from_record_batch(rb)
    rec = engine.load_records(host, with_sequences=True)
    got = engine.collect(rec, make_params()).to_numpy()
    want = port.collect(host, port.Params())
    assert util.rows_equal(got, want) is None, util.rows_equal(got, want)
    assert want.shape[0] == 1 + 11 + 8 + 4 and set(want["type"].tolist()) == {0}
