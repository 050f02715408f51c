"""GPU parity of the pairing stage (K6 sort, K7 partitions, K8 edit distance, K9 clustering) vs the CPU oracle.

Reference path: SVIM_COMBINE.py:15-366 (form_partitions, compute_distance, pair_haplotypes[_breakends],
pair_candidates).  Bit-exact rows in the reference's order."""
import numpy as np
import pytest

from oracle import port
from svim_asm_b200 import synth
from svim_asm_b200.engine import HostBatch, make_params
from tests import util

pytestmark = pytest.mark.gpu


def _mutate(rng, s, n_edits, alphabet):
    b = list(s)
    for _ in range(n_edits):
        kind = int(rng.integers(0, 3))
        pos = int(rng.integers(0, max(1, len(b))))
        if kind == 0 and b:
            b[pos % len(b)] = int(rng.choice(alphabet))
        elif kind == 1:
            b.insert(pos, int(rng.choice(alphabet)))
        elif b:
            del b[pos % len(b)]
    return bytes(b)


def test_edit_distance_matches_plain_dp(engine, oracle_clib):
    rng = np.random.default_rng(21)
    alphabet = list(b"ACGT")
    pairs = [(b"", b""), (b"", b"ACGT"), (b"ACGT", b""), (b"kitten", b"sitting"), (b"A" * 64, b"A" * 64), (b"A" * 65, b"C" * 63)]
    for _ in range(300):
        m = int(rng.integers(1, 700))
        a = bytes(rng.choice(alphabet, m).tolist())
        if rng.random() < 0.6:
            b = _mutate(rng, a, int(rng.integers(0, 40)), alphabet)
        else:
            b = bytes(rng.choice(alphabet, int(rng.integers(1, 700))).tolist())
        pairs.append((a, b))
    # several 2048-row stripes (horizontal deltas parked in HBM between stripes)
    for m, n in ((2048, 2048), (2049, 2500), (5000, 4100), (6200, 9000)):
        a = bytes(rng.choice(alphabet + list(b"N"), m).tolist())
        pairs.append((a, _mutate(rng, a, 150, alphabet) if n == 4100 else bytes(rng.choice(alphabet, n).tolist())))
    got = engine.edit_distance(pairs)
    want = [port.edit_distance(a, b) for a, b in pairs]
    assert got.tolist() == want


def test_edit_distance_long_patterns_window_and_stripes(engine, oracle_clib):
    """Patterns longer than one 2048-row stripe: the sliding window settles distances inside its band (about 1000
    either side), everything else falls through to the striped band attempts.  Distances below, at and beyond the
    band edge; length differences that narrow the band or rule the window out; alphabets with N."""
    rng = np.random.default_rng(23)
    alphabet = list(b"ACGT")
    pairs = []
    for m, edits in ((1025, 0), (1100, 30), (1500, 200), (1900, 600), (2048, 480), (2049, 0), (2100, 3), (3000, 120), (4500, 450),
                     (4500, 520), (5200, 900), (5200, 1400), (7000, 2500), (9900, 520), (9900, 400)):
        a = bytes(rng.choice(alphabet + list(b"N"), m).tolist())
        pairs.append((a, _mutate(rng, a, edits, alphabet)))
    a = bytes(rng.choice(alphabet, 6000).tolist())
    for extra in (700, 1500, 1990, 2100, 5000):              # n - m narrows the band, then excludes the window
        cut = int(rng.integers(0, len(a)))
        pairs.append((a, a[:cut] + bytes(rng.choice(alphabet, extra).tolist()) + a[cut:]))
    pairs.append((bytes(rng.choice(alphabet, 4000).tolist()), bytes(rng.choice(alphabet, 4100).tolist())))   # unrelated
    # short pattern against a long text: transposed full tables, or the text split between the two warps of a CTA
    for m, n in ((300, 9000), (1300, 10900), (820, 9900), (2000, 2300), (2218, 2425), (64, 5000), (1, 4000), (2000, 12000),
                 (1500, 4100), (2048, 4096), (1999, 4097)):
        pairs.append((bytes(rng.choice(alphabet, m).tolist()), bytes(rng.choice(alphabet, n).tolist())))
    core = bytes(rng.choice(alphabet, 1800).tolist())                   # related strings with a long private insertion
    pairs.append((core, core[:900] + bytes(rng.choice(alphabet, 5000).tolist()) + _mutate(rng, core[900:], 30, alphabet)))
    pairs.append((b"A" * 5000, b"A" * 4990 + b"C" * 10))
    pairs.append((b"AC" * 3000, b"CA" * 3000))
    got = engine.edit_distance(pairs)
    want = [port.edit_distance(x, y) for x, y in pairs]
    assert got.tolist() == want


def test_edit_distance_bounded_matches_plain_dp(engine, oracle_clib):
    """svb_edit_distance_bounded = edlib.align(a, b, k=t): the thresholded wavefront kernel svb_pair runs on every pair
    (wfa.cu).  Exact whenever the distance is <= t, -1 otherwise; every threshold the kernel accepts, distances at the
    threshold, empty strings, low-complexity strings (many tied diagonals), length differences around t."""
    rng = np.random.default_rng(29)
    alphabet = list(b"ACGT")
    pairs = [(b"", b""), (b"", b"ACGT"), (b"ACGT", b""), (b"kitten", b"sitting"), (b"A" * 64, b"A" * 64), (b"A" * 65, b"C" * 63),
             (b"A" * 500, b"A" * 490 + b"C" * 10), (b"AC" * 300, b"CA" * 300), (b"ACGT" * 100, b"ACGT" * 99)]
    for _ in range(400):
        m = int(rng.integers(1, 900))
        a = bytes(rng.choice(alphabet, m).tolist())
        if rng.random() < 0.7:
            b = _mutate(rng, a, int(rng.integers(0, 260)), alphabet)
        else:
            b = bytes(rng.choice(alphabet, int(rng.integers(1, 900))).tolist())
        pairs.append((a, b))
    want = [port.edit_distance(a, b) for a, b in pairs]
    for t in (0, 1, 10, 200, 201, 1024):
        got = engine.edit_distance(pairs, max_distance=t)
        assert got.tolist() == [d if d <= t else -1 for d in want], t
    d199 = [(a, b) for (a, b), d in zip(pairs, want) if 150 <= d <= 260]
    assert len(d199) > 10                        # the cut itself is exercised (200 merges, 201 does not)


def test_edit_distance_bounded_at_the_distance_itself(engine, oracle_clib):
    """Thresholds right at, below and above every pair's distance, even and odd: the last round of the bidirectional run
    tests the total t itself, on the outermost diagonal the pruning leaves (a one-sided pruning bound lost exactly that
    diagonal for even t: found by the host statement of the rounds, tests/test_edit_core.py)."""
    rng = np.random.default_rng(31)
    alphabet = list(b"ACGT")
    pairs = [(b"ACGT" * 100, b"ACGT" * 99), (b"ACGT" * 100 + b"T", b"G" + b"ACGT" * 99)]
    for _ in range(300):
        m = int(rng.integers(2, 500))
        a = bytes(rng.choice(alphabet, m).tolist())
        b = _mutate(rng, a, int(rng.integers(1, 70)), alphabet)
        if rng.random() < 0.5 and len(b) > 1:                 # ends that differ: nothing to trim, the rounds see the whole pair
            a = b"A" + a + b"C"
            b = b"G" + b + b"T"
        pairs.append((a, b))
    want = [port.edit_distance(a, b) for a, b in pairs]
    for t in sorted({x for d in want for x in (d - 1, d, d + 1) if 0 <= x <= 1024}):
        got = engine.edit_distance(pairs, max_distance=t)
        assert got.tolist() == [d if d <= t else -1 for d in want], t


def test_edit_distance_bounded_long_pairs(engine, oracle_clib):
    """Long haplotype pairs: the whole-warp match extension, the one-CTA-per-SM stage for strings beyond the small
    shared-memory window, and the hand-over to the exact kernel for strings beyond shared memory altogether."""
    rng = np.random.default_rng(30)
    alphabet = list(b"ACGT")
    pairs = []
    for m, edits in ((9900, 0), (9900, 3), (9900, 150), (9900, 199), (9900, 230), (9900, 520), (12500, 40), (30000, 120), (30000, 400),
                     (60000, 90), (100000, 20), (130000, 15)):
        a = bytes(rng.choice(alphabet + list(b"N"), m).tolist())
        pairs.append((a, _mutate(rng, a, edits, alphabet)))
    a = bytes(rng.choice(alphabet, 8000).tolist())
    pairs.append((a, a[:4000] + bytes(rng.choice(alphabet, 150).tolist()) + a[4000:]))          # one 150-base insertion
    pairs.append((a, a[:4000] + bytes(rng.choice(alphabet, 250).tolist()) + a[4000:]))          # length difference beyond t
    pairs.append((bytes(rng.choice(alphabet, 5000).tolist()), bytes(rng.choice(alphabet, 5100).tolist())))      # unrelated
    pairs.append((b"ACGT" * 3000, b"ACGT" * 2990 + b"TTTT" * 10))
    want = []
    for x, y in pairs:
        want.append(port.edit_distance(x, y) if len(x) * len(y) <= 4e8 else None)
    got = engine.edit_distance(pairs, max_distance=200)
    exact = engine.edit_distance(pairs)
    for g, e, w in zip(got.tolist(), exact.tolist(), want):
        if w is not None:
            assert e == w
        assert g == (e if e <= 200 else -1)


def test_edit_distance_parked_delta_stride_grows(built_library, oracle_clib, monkeypatch):
    """A table of several stripes parks one byte per text column in a per-worker slice; a pair whose text is longer than
    the slice makes the call grow the slices and run again (the pool no longer scales with the longest pair
    times every CTA)."""
    from svim_asm_b200.engine import Engine
    monkeypatch.setenv("SVB_ED_STRIDE", "2048")
    eng = Engine(0)
    try:
        rng = np.random.default_rng(31)
        alphabet = list(b"ACGT")
        pairs = [(bytes(rng.choice(alphabet, 5000).tolist()), bytes(rng.choice(alphabet, 4100).tolist())),
                 (bytes(rng.choice(alphabet, 300).tolist()), bytes(rng.choice(alphabet, 280).tolist()))]
        got = eng.edit_distance(pairs)
        assert got.tolist() == [port.edit_distance(a, b) for a, b in pairs]
    finally:
        eng.close()


def test_cluster_labels_match_scipy(engine):
    rng = np.random.default_rng(22)
    problems = []
    for _ in range(3000):
        n = int(rng.integers(2, 11))
        m = n * (n - 1) // 2
        if rng.random() < 0.7:
            problems.append(rng.choice([1e9, 200.0, 300.0, 150.0, 0.0, 201.0, 99999.0], size=m))
        else:
            problems.append(np.round(rng.uniform(0, 400, m)))
    got = engine.cluster_labels(problems, 200.0)
    for d, g in zip(problems, got):
        assert g == [int(x) for x in port.cluster_labels(d, 200.0)]


def _reference_arrays(cfg):
    ref = synth.random_reference(cfg)
    off = np.zeros(len(cfg.contig_names) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([ref[n].shape[0] for n in cfg.contig_names])
    return np.concatenate([ref[n] for n in cfg.contig_names]), off


def _pair_both(engine, cfg, **params):
    rb1, rb2 = synth.make_diploid(cfg)
    h1, h2 = HostBatch.from_record_batch(rb1), HostBatch.from_record_batch(rb2)
    bases, off = _reference_arrays(cfg)
    p = port.Params(**params)
    w1, w2 = port.collect(h1, p, hap=1), port.collect(h2, p, hap=2)

    def fetch(tid, s, e):
        return bases[int(off[tid]) + s:int(off[tid]) + e].tobytes()
    want = port.pair(w1, w2, h1, h2, fetch, p)
    r1, r2 = engine.load_records(h1, with_sequences=True), engine.load_records(h2, with_sequences=True)
    ref = engine.load_reference(bases, off)
    dp = make_params(**params)
    t1, t2 = engine.collect(r1, dp, hap=1), engine.collect(r2, dp, hap=2)
    assert util.rows_equal(t1.to_numpy(), w1) is None and util.rows_equal(t2.to_numpy(), w2) is None
    got = engine.pair(t1, t2, r1, r2, ref, dp).to_numpy()
    return got, want


@pytest.mark.parametrize("seed", [31, 32, 33])
def test_diploid_pairing_small(engine, oracle_clib, seed):
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2"], [400000, 300000, 350000], 90, 6e4, seed, sv_per_event=6e-3,
                            split_fraction=0.5, sv_max=3000)
    got, want = _pair_both(engine, cfg)
    diff = util.rows_equal(got, want)
    assert diff is None, diff
    assert {0, 1, 2} <= set(int(g) for g in got["genotype"])        # 1/1, 1/0 and 0/1 all occur
    assert want.shape[0] > 60


def test_diploid_pairing_tight_partitions(engine, oracle_clib):
    # many candidates close together: large partitions (some > 10 are dropped), chains, label order
    cfg = synth.SynthConfig(["chrA"], [120000], 12, 1.5e4, 77, sv_per_event=6e-2, split_fraction=0.3, sv_max=300)
    got, want = _pair_both(engine, cfg, partition_max_distance=3000, max_edit_distance=150)
    assert util.rows_equal(got, want) is None, util.rows_equal(got, want)


def test_sequence_pool_sources_agree(engine):
    """The inserted bases of the INS rows reach the table three ways: gathered on the device from resident query sequences,
    gathered on the host from pageable buffers, and read in place from pinned host buffers by the gather kernel.  Same
    bytes each time (they are candidate.sequence of SVIM_intra.py:42 / SVIM_COMBINE.py:70-75)."""
    from svim_asm_b200.bench_util import pinned_host
    cfg = synth.SynthConfig(["chr1", "chr2"], [400000, 300000], 80, 5e4, 77, sv_per_event=8e-3, split_fraction=0.3, sv_max=3000)
    rb = synth.make_haploid(cfg)
    host = HostBatch.from_record_batch(rb)
    rec = engine.load_records(host, with_sequences=True)
    params = make_params()
    t = engine.collect(rec, params, hap=1)
    t.gather_sequences(rec)
    pool_dev, off_dev = t.pool_to_numpy()
    t.attach_sequences_host(host)                        # numpy arrays: pageable
    pool_host, off_host = t.pool_to_numpy()
    pinned = pinned_host(HostBatch.from_record_batch(rb))
    t.attach_sequences_host(pinned)                      # pinned: zero-copy gather
    pool_pin, off_pin = t.pool_to_numpy()
    assert pool_dev.shape[0] > 1000
    assert np.array_equal(off_dev, off_host) and np.array_equal(off_dev, off_pin)
    assert pool_dev.tobytes() == pool_host.tobytes() == pool_pin.tobytes()
