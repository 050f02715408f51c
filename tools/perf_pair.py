"""Per-job profile of the edit-distance kernel on the bench workload: which pairs are the long poles.

  python tools/perf_pair.py --scale 0.25     (needs a GPU; uses the SVB_ED_PROFILE dump of csrc/edit_distance.cu)"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from svim_asm_b200.engine import Engine, HostBatch, make_params

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.25)
ap.add_argument("--out", default="gpurun_out/ed_profile.npy")
args = ap.parse_args()

cfg, rb1, rb2, bases, off = bench.build_workload(args.scale)
eng = Engine(0)
hosts = [HostBatch.from_record_batch(rb) for rb in (rb1, rb2)]
recs = [eng.load_records(h, with_sequences=True) for h in hosts]
ref = eng.load_reference(bases, off)
params = make_params()
tabs = [eng.collect(r, params, hap=k + 1) for k, r in enumerate(recs)]
for _ in range(2):
    eng.pair(tabs[0], tabs[1], recs[0], recs[1], ref, params)
eng.timing_reset()
for _ in range(5):
    eng.pair(tabs[0], tabs[1], recs[0], recs[1], ref, params)
t = eng.timing()
stats = eng.pair_stats()
print("pair x5: K8 (wavefront x2 + resolve + exact kernel) %.3f ms per call, sort %.3f, cluster %.3f" % (
    t["edit_distance"][0] / 5, t["sort"][0] / 5, t["cluster"][0] / 5))
print("pair stats: %(partitions)d partitions, %(pairs)d cross-haplotype pairs, %(exact_pairs)d needed the exact kernel, "
      "%(table_cells).3g full-table cells" % stats)
print("K8: %.3g equivalent cell updates per second" % (stats["table_cells"] / (t["edit_distance"][0] / 5 * 1e-3)))
wpath = "/tmp/wfa_profile.bin"
os.environ["SVB_WFA_PROFILE"] = wpath
eng.pair(tabs[0], tabs[1], recs[0], recs[1], ref, params)
del os.environ["SVB_WFA_PROFILE"]
w = np.fromfile(wpath, dtype=np.uint32).reshape(-1, 4)
w = w[w[:, 3] > 0]
cyc = w[:, 3].astype(np.float64)
waves = w[:, 2] & 0xFFFF
trimmed = waves == 0xFFFF
print("wavefront kernel: %d jobs profiled, sum of job cycles %.3g, max %.3g (%.3f ms at 1.9 GHz); settled by trimming / length: %d (cycles sum %.3g, max %.3g)"
      % (w.shape[0], cyc.sum(), cyc.max(), cyc.max() / 1.9e6, trimmed.sum(), cyc[trimmed].sum(), cyc[trimmed].max() if trimmed.any() else 0))
print("     la     lb  waves stage   cycles")
for i in np.argsort(-cyc)[:20]:
    print("%7d %6d %6s %5d %8d" % (w[i, 0], w[i, 1], "-" if trimmed[i] else str(int(waves[i])), int(w[i, 2] >> 16) if not trimmed[i] else 0, w[i, 3]))
ran = ~trimmed
if ran.any():
    print("jobs that ran waves: %d, mean cycles %.3g, mean waves %.1f" % (ran.sum(), cyc[ran].mean(), waves[ran].mean()))
path = "/tmp/ed_profile.bin"
os.environ["SVB_ED_PROFILE"] = path
eng.timing_reset()
eng.pair(tabs[0], tabs[1], recs[0], recs[1], ref, params)
del os.environ["SVB_ED_PROFILE"]
if not os.path.exists(path) or stats["exact_pairs"] == 0:
    print("no pair went to the exact kernel")
    sys.exit(0)
prof = np.fromfile(path, dtype=np.uint32).reshape(-1, 4)
prof = prof[prof[:, 3] > 0]
np.save(args.out, prof)
print("exact kernel: %d jobs" % prof.shape[0])
cyc = prof[:, 3].astype(np.float64)
print("sum of job cycles %.3g, max %.3g (%.3f ms at 1.9 GHz)" % (cyc.sum(), cyc.max(), cyc.max() / 1.9e6))
print("   rows   cols  steps   cycles  cyc/step")
for i in np.argsort(-cyc)[:25]:
    m, n, st, c = (int(x) for x in prof[i])
    print("%7d %6d %6d %8d %8.1f" % (m, n, st, c, c / max(st, 1)))
