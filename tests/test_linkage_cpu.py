"""linkage.cuh on the host: the scipy-compatible labels of the plain run, and the interval run pair.cu uses to decide
which far pairs need their exact edit distance (resolve_kernel): whenever it says "determined", EVERY choice of values
inside the intervals must give those labels with scipy itself."""
import numpy as np
from scipy.cluster.hierarchy import fcluster, linkage

from tests import hostcheck


def _scipy_labels(d, t):
    return [int(x) for x in fcluster(linkage(np.asarray(d, dtype=np.float64), method="complete"), t, criterion="distance")]


def test_plain_labels_match_scipy():
    lib = hostcheck.load()
    rng = np.random.default_rng(7)
    for _ in range(2000):
        n = int(rng.integers(2, 11))
        m = n * (n - 1) // 2
        d = rng.choice([1e9, 200.0, 300.0, 150.0, 0.0, 201.0, 99999.0], size=m) if rng.random() < 0.7 else np.round(rng.uniform(0, 400, m))
        d = np.ascontiguousarray(d, dtype=np.float64)
        labels = np.zeros(n, dtype=np.int32)
        lib.hc_cluster_labels(d.ctypes.data, n, 200.0, labels.ctypes.data)
        assert labels.tolist() == _scipy_labels(d, 200.0)


def test_interval_run_is_sound():
    """Distances as the wavefront kernel leaves them: exact values <= t, the same-haplotype constant 1e9, and far intervals
    [lo, hi] with t < lo.  'Determined' must hold for the corners and for random interior points of the intervals."""
    lib = hostcheck.load()
    rng = np.random.default_rng(8)
    t = 200.0
    determined = undetermined = 0
    for _ in range(3000):
        n = int(rng.integers(2, 8))
        m = n * (n - 1) // 2
        hap = rng.integers(1, 3, n)
        lo, hi = np.zeros(m), np.zeros(m)
        k = 0
        for i in range(n - 1):
            for j in range(i + 1, n):
                if hap[i] == hap[j]:
                    lo[k] = hi[k] = 1e9
                elif rng.random() < 0.5:
                    lo[k] = hi[k] = float(rng.integers(0, 201))
                else:
                    a = float(rng.choice([201, 201, 250, 900, 5000]))
                    lo[k], hi[k] = a, a + float(rng.choice([0, 50, 400, 9000]))
                k += 1
        labels = np.zeros(n, dtype=np.int32)
        ok = lib.hc_labels_determined(lo.ctypes.data, hi.ctypes.data, n, t, labels.ctypes.data)
        if not ok:
            undetermined += 1
            continue
        determined += 1
        want = labels.tolist()
        trials = [lo, hi] + [np.floor(lo + rng.random(m) * (hi - lo + 1)).clip(lo, hi) for _ in range(6)]
        # adversarial: far values pushed to alternate ends
        alt = np.where(np.arange(m) % 2 == 0, lo, hi)
        trials += [alt, np.where(np.arange(m) % 2 == 0, hi, lo)]
        for d in trials:
            assert _scipy_labels(d, t) == want, (lo, hi, d)
    assert determined > 1500 and undetermined > 100


def test_interval_run_flags_the_app_d_case():
    """one hap-2 candidate far from three hap-1 candidates: the order of the far distances decides the
    labels (A2, B, A, A3), so overlapping intervals must NOT be reported as determined; disjoint ones may."""
    lib = hostcheck.load()
    # points: A, A3, A2 (hap 1), B (hap 2): pairs (A,A3) (A,A2) (A,B) (A3,A2) (A3,B) (A2,B)
    lo = np.array([1e9, 1e9, 300.0, 1e9, 300.0, 300.0])
    hi = np.array([1e9, 1e9, 900.0, 1e9, 900.0, 900.0])
    labels = np.zeros(4, dtype=np.int32)
    assert lib.hc_labels_determined(lo.ctypes.data, hi.ctypes.data, 4, 200.0, labels.ctypes.data) == 0
    lo = np.array([1e9, 1e9, 300.0, 1e9, 500.0, 700.0])
    hi = np.array([1e9, 1e9, 400.0, 1e9, 600.0, 800.0])
    assert lib.hc_labels_determined(lo.ctypes.data, hi.ctypes.data, 4, 200.0, labels.ctypes.data) == 1
    assert labels.tolist() == _scipy_labels(lo, 200.0) == _scipy_labels(hi, 200.0)
