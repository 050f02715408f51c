"""Host checks of the edit-distance cores shared with the CUDA kernel (csrc/edit_core.cuh, compiled by g++ in
tests/hostcheck): the 64-row block step in stripes, and the sliding window over the Ukkonen band, both against the
plain DP of the oracle (what edlib.align(a, b)["editDistance"] returns, reference SVIM_COMBINE.py:50-100)."""
import numpy as np
import pytest

from tests import hostcheck


def _mutate(rng, s, n_edits):
    b = bytearray(s)
    for _ in range(n_edits):
        kind = int(rng.integers(0, 3))
        pos = int(rng.integers(0, max(1, len(b))))
        if kind == 0 and b:
            b[pos % len(b)] = int(rng.choice(list(b"ACGT")))
        elif kind == 1:
            b.insert(pos, int(rng.choice(list(b"ACGT"))))
        elif b:
            del b[pos % len(b)]
    return bytes(b)


def test_block_step_in_stripes(oracle_clib):
    lib = hostcheck.load()
    rng = np.random.default_rng(5)
    for _ in range(60):
        m = int(rng.integers(1, 400))
        a = bytes(rng.choice(list(b"ACGT"), m).tolist())
        b = _mutate(rng, a, int(rng.integers(0, 60))) if rng.random() < 0.7 else bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 400))).tolist())
        want = oracle_clib.orc_edit_distance(a, len(a), b, len(b))
        for blocks in (1, 2, 32):
            assert lib.hc_myers(a, len(a), b, len(b), blocks) == want


@pytest.mark.parametrize("bw", [64, 32])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sliding_window_exact_within_band(oracle_clib, seed, bw):
    """window value == distance whenever it is <= K; never below the distance; the widest K one warp covers is accepted."""
    lib = hostcheck.load()
    rng = np.random.default_rng(seed)
    cases = 0
    for trial in range(14):
        m = int(rng.integers(2049, 5200)) if trial % 2 else int(rng.integers(65, 2500))
        a = bytes(rng.choice(list(b"ACGT"), m).tolist())
        b = _mutate(rng, a, int(rng.integers(0, 700)))
        if rng.random() < 0.3:                                   # a long private insertion in the middle
            cut = int(rng.integers(0, len(b)))
            b = b[:cut] + bytes(rng.choice(list(b"ACGT"), int(rng.integers(100, 900))).tolist()) + b[cut:]
        p, t = (a, b) if len(a) <= len(b) else (b, a)
        if not p:
            continue
        want = oracle_clib.orc_edit_distance(p, len(p), t, len(t))
        kmax = lib.hc_win_kmax(len(p), len(t), bw)
        for K in sorted({1, 64, 300, kmax} - {0}):
            if K > kmax:
                continue
            got = lib.hc_myers_window(p, len(p), t, len(t), K, bw)
            assert got >= want, (len(p), len(t), K, got, want)
            if want <= K:
                assert got == want, (len(p), len(t), K, got, want)
                cases += 1
            if got <= K:
                assert got == want
    assert cases > 8


@pytest.mark.parametrize("bw", [64, 32])
def test_sliding_window_edge_shapes(oracle_clib, bw):
    lib = hostcheck.load()
    rng = np.random.default_rng(11)
    for m, n in [(1, 1), (1, 300), (64, 64), (65, 64 + 65), (128, 128), (2048, 2048), (2049, 2049), (4096, 4100), (63, 1500)]:
        p = bytes(rng.choice(list(b"AC"), m).tolist())
        t = (p + bytes(rng.choice(list(b"AC"), n - m).tolist())) if n > m else p
        t = _mutate(rng, t, 5)
        if len(t) < len(p):
            p, t = t, p
        kmax = lib.hc_win_kmax(len(p), len(t), bw)
        if kmax == 0:
            continue
        want = oracle_clib.orc_edit_distance(p, len(p), t, len(t))
        got = lib.hc_myers_window(p, len(p), t, len(t), kmax, bw)
        assert got >= want
        if want <= kmax:
            assert got == want, (m, n, got, want)


def test_sentinel_padding_keeps_the_distance(oracle_clib):
    """window_pass32 pads both strings with a symbol that matches only itself up to a multiple of 32 rows:
    D(P + S^k, T + S^k) = D(P, T), and the padded window is exact within the same band."""
    lib = hostcheck.load()
    rng = np.random.default_rng(17)
    for _ in range(12):
        m = int(rng.integers(1025, 4000))
        a = bytes(rng.choice(list(b"ACGT"), m).tolist())
        b = _mutate(rng, a, int(rng.integers(0, 400)))
        p, t = (a, b) if len(a) <= len(b) else (b, a)
        want = oracle_clib.orc_edit_distance(p, len(p), t, len(t))
        k = (32 - len(p) % 32) % 32
        pp, tt = p + b"#" * k, t + b"#" * k
        assert oracle_clib.orc_edit_distance(pp, len(pp), tt, len(tt)) == want
        kmax = lib.hc_win_kmax(len(p), len(t), 32)
        if kmax and want <= kmax:
            assert lib.hc_myers_window(pp, len(pp), tt, len(tt), kmax, 32) == want


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_split_window_meets_in_the_middle(oracle_clib, seed):
    """Two halves (forward from the top rows, backward from the bottom rows of the reversed strings) meeting at the middle
    row: min_j F(j) + B(n - j) is the distance whenever it is inside the band, and never below it."""
    lib = hostcheck.load()
    rng = np.random.default_rng(seed)
    exact = 0
    for trial in range(10):
        m = int(rng.integers(1100, 5000))
        a = bytes(rng.choice(list(b"ACGT"), m).tolist())
        b = _mutate(rng, a, int(rng.integers(0, 450)))
        if trial % 3 == 0:
            cut = int(rng.integers(0, len(b)))
            b = b[:cut] + bytes(rng.choice(list(b"ACGT"), int(rng.integers(50, 400))).tolist()) + b[cut:]
        p, t = (a, b) if len(a) <= len(b) else (b, a)
        want = oracle_clib.orc_edit_distance(p, len(p), t, len(t))
        kmax = lib.hc_win_kmax(len(p) + 31, len(t) + 31, 32)
        for K in sorted({40, 200, kmax} - {0}):
            if K > kmax:
                continue
            got = lib.hc_myers_window_split(p, len(p), t, len(t), K)
            if got == -1:
                continue
            assert got == -2 or got >= want, (len(p), len(t), K, got, want)
            if want <= K:
                assert got == want, (len(p), len(t), K, got, want)
                exact += 1
    assert exact > 8


def test_wavefront_threshold_core_matches_plain_dp():
    """wfa_core.cuh (the thresholded wavefront distance of wfa.cu) through its serial driver: the distance when it is <= t,
    -1 otherwise; thresholds 0 .. 200, related / unrelated / empty / low-complexity strings."""
    from oracle import port
    lib = hostcheck.load()
    rng = np.random.default_rng(5)
    alphabet = list(b"ACGT")

    def mutate(s, k):
        b = list(s)
        for _ in range(k):
            kind, pos = int(rng.integers(0, 3)), int(rng.integers(0, max(1, len(b))))
            if kind == 0 and b:
                b[pos % len(b)] = int(rng.choice(alphabet))
            elif kind == 1:
                b.insert(pos, int(rng.choice(alphabet)))
            elif b:
                del b[pos % len(b)]
        return bytes(b)
    cases = [(b"A" * 500, b"A" * 490 + b"C" * 10, 20), (b"AC" * 300, b"CA" * 300, 5), (b"ACGT" * 100, b"ACGT" * 99, 4),
             (b"A" * 100, b"A" * 100, 0), (b"", b"", 0), (b"", b"AAA", 3), (b"", b"AAA", 2), (b"AAA", b"", 3), (b"A", b"C", 1), (b"A", b"C", 0),
             (b"ACGTACGTAC", b"ACGTACGTAC", 0), (b"ACGTACGTAC", b"TACGTACGTA", 2), (b"ACGTACGTAC", b"TACGTACGTA", 1)]
    for _ in range(1500):
        a = bytes(rng.choice(alphabet, int(rng.integers(0, 400))).tolist())
        b = mutate(a, int(rng.integers(0, 60))) if rng.random() < 0.7 else bytes(rng.choice(alphabet, int(rng.integers(0, 400))).tolist())
        cases.append((a, b, int(rng.choice([0, 1, 2, 3, 5, 10, 30, 31, 200]))))
    # thresholds right at, below and above the distance (the last round of the bidirectional run tests the total t itself)
    for a, b, _t in list(cases[13:613]):
        d = port.edit_distance(a, b)
        for t in (d - 1, d, d + 1):
            if 0 <= t <= 300:
                cases.append((a, b, t))
    for a, b, t in cases:
        d = port.edit_distance(a, b)
        assert lib.hc_wfa(a, len(a), b, len(b), t) == (d if d <= t else -1), (len(a), len(b), t, d)
        # the bidirectional run (forward and backward waves meeting in the middle: what the kernel does)
        assert lib.hc_wfa_bidir(a, len(a), b, len(b), t) == (d if d <= t else -1), ("bidirectional", len(a), len(b), t, d)
        # the kernel's round structure: branch-free clamped recurrence from "wave -1", the overlap tests of the totals 2r - 2
        # and 2r - 1 folded into the forward wave of round r
        assert lib.hc_wfa_rounds(a, len(a), b, len(b), t) == (d if d <= t else -1), ("rounds", len(a), len(b), t, d)
