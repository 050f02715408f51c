"""Stage-by-stage timeline of the e2e step of bench.py (host buffers in, paired rows out), each stage synchronised, next to
the raw pinned host->device copy rate of the same bytes.  Needs a GPU:  python tools/perf_e2e.py --scale 1.0"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from svim_asm_b200.bench_util import pin, pinned_host
from svim_asm_b200.engine import Engine, HostBatch, make_params

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--iters", type=int, default=4)
args = ap.parse_args()

cfg, rb1, rb2, bases, off = bench.build_workload(args.scale)
h1, h2 = pinned_host(HostBatch.from_record_batch(rb1)), pinned_host(HostBatch.from_record_batch(rb2))
eng = Engine(0)
params = make_params()
bases_p, keep = pin(bases)
ref = eng.load_reference(bases_p, off)

# raw copy rate of one CIGAR array
src = torch.from_numpy(h1.cigar)
dst = torch.empty(src.shape, dtype=src.dtype, device="cuda")
for _ in range(2):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 4
print("raw pinned H2D: %.1f MB in %.2f ms = %.1f GB/s (pinned: %s)" % (src.nbytes / 1e6, dt * 1e3, src.nbytes / dt / 1e9, src.is_pinned()))
del dst

stages = ["load h1", "load h2", "collect h1", "collect h2", "attach seq", "pair", "rows to host", "free"]
acc = np.zeros(len(stages))
for it in range(args.iters + 2):
    ts = [time.perf_counter()]

    def tick():
        eng.synchronize()
        ts.append(time.perf_counter())
    r1 = eng.load_records(h1); tick()
    r2 = eng.load_records(h2); tick()
    t1 = eng.collect(r1, params, hap=1); tick()
    t2 = eng.collect(r2, params, hap=2); tick()
    t1.attach_sequences_host(h1); t2.attach_sequences_host(h2); tick()
    paired = eng.pair(t1, t2, r1, r2, ref, params); tick()
    rows = paired.to_numpy(); tick()
    for obj in (t1, t2, paired, r1, r2):
        obj.free()
    tick()
    if it >= 2:
        acc += np.diff(ts)
for s, v in zip(stages, acc / args.iters * 1e3):
    print("%-14s %8.3f ms" % (s, v))
print("%-14s %8.3f ms" % ("total", acc.sum() / args.iters * 1e3))
