"""Latency of the thresholded wavefront kernel (K8, csrc/wfa.cu) on hand-made pairs: one long far pair alone against many at
once -- is a launch bound by the latency of one pair's rounds or by the issue slots its SM shares with other pairs?
Needs a GPU:  python tools/perf_wfa.py"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from svim_asm_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--length", type=int, default=9748)
ap.add_argument("--t", type=int, default=200)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
rng = np.random.default_rng(5)
ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def mutated(n, divergence):
    a = ALPHA[rng.integers(0, 4, n)]
    b = a.copy()
    hit = rng.random(n) < divergence
    b[hit] = ALPHA[(np.searchsorted(ALPHA, b[hit]) + rng.integers(1, 4, int(hit.sum()))) % 4]
    return a.tobytes(), b.tobytes()


def unrelated(n):
    return ALPHA[rng.integers(0, 4, n)].tobytes(), ALPHA[rng.integers(0, 4, n)].tobytes()


eng = Engine(0)
for label, make in (("same length, 4 %% substitutions (d ~ %d)" % int(args.length * 0.04), lambda: mutated(args.length, 0.04)),
                    ("same length, 1.5 %% substitutions (d ~ %d)" % int(args.length * 0.015), lambda: mutated(args.length, 0.015)),
                    ("unrelated", lambda: unrelated(args.length))):
    print(label)
    for n in (1, 8, 148, 592, 1184):
        pairs = [make() for _ in range(n)]
        eng.edit_distance(pairs, max_distance=args.t)
        eng.timing_reset()
        for _ in range(args.reps):
            out = eng.edit_distance(pairs, max_distance=args.t)
        ms = eng.timing()["edit_distance"][0] / args.reps
        print("  %5d pairs: %.3f ms per launch pair (%d within t)" % (n, ms, int((out >= 0).sum())), flush=True)
