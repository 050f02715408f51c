"""Quick device timing of svb_collect (cigar_scan variants) on a synthetic haploid whole-genome batch."""
import argparse
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from svim_asm_b200 import synth
from svim_asm_b200.engine import Engine, HostBatch, make_params

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.5)
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()

cfg = synth.config_c3(scale=args.scale)
cfg.with_sequence = False
t0 = time.time()
rb = synth.make_haploid(cfg)
host = HostBatch.from_record_batch(rb)
print("generated %d alignments, %d ops in %.1fs" % (host.n_aln, host.n_ops, time.time() - t0), flush=True)
eng = Engine(0)
rec = eng.load_records(host)
params = make_params()
for variant in (0, 1):
    eng.set_scan_variant(variant)
    for _ in range(3):
        t = eng.collect(rec, params)
    n_rows = len(t)
    eng.timing_reset()
    w0 = time.time()
    for _ in range(args.iters):
        t = eng.collect(rec, params)
    eng.synchronize()
    wall = (time.time() - w0) / args.iters * 1e3
    tm = eng.timing()
    ms, launches = tm["cigar_scan"]
    per = ms / max(launches, 1)
    alg = 4.0 * host.cigar.shape[0] + 32.0 * host.n_aln + 64.0 * n_rows
    print(json.dumps({"variant": "tma" if variant == 0 else "ldg", "rows": n_rows, "scan_ms": per,
                      "scan_GBps": alg / per / 1e6, "ops_per_s": host.n_ops / per * 1e3, "collect_wall_ms": wall,
                      "finalize_ms": tm["cigar_scan_finalize"][0] / max(tm["cigar_scan_finalize"][1], 1),
                      "walk_ms": tm["segment_walk"][0] / args.iters, "merge_ms": tm["merge"][0] / args.iters}), flush=True)
