"""File -> record image: the device ingest (BGZF inflate + record split on the GPU) next to the host ingest
(multi-threaded zlib + record pass + upload), on a synthetic haploid BAM of whole-genome shape.
Needs a GPU:  python tools/perf_ingest.py --scale 0.5"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from svim_asm_b200 import bamio, synth
from svim_asm_b200.engine import Engine, HostBatch, make_params

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.5)
ap.add_argument("--level", type=int, default=1)
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()

cfg = synth.config_c3(scale=args.scale)
t0 = time.time()
rb = synth.make_haploid(cfg)
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "hap.bam")
bamio.write_bam(path, rb, level=args.level, index=False)
fsize = os.path.getsize(path)
print("wrote %s: %d alignments, %d ops, %.1f MB (level %d) in %.0fs" % (path, rb.n_aln, rb.n_ops, fsize / 1e6, args.level, time.time() - t0), flush=True)
del rb
eng = Engine(0)
params = make_params()

res = {"file_MB": fsize / 1e6}
host_times, dev_times = [], []
for it in range(args.iters):
    t0 = time.perf_counter()
    host = HostBatch.from_bam(path)
    t1 = time.perf_counter()
    rec = eng.load_records(host, with_sequences=True)
    eng.synchronize()
    t2 = time.perf_counter()
    n_host = len(eng.collect(rec, params))
    rec.free()
    host_times.append((t1 - t0, t2 - t1))
    host.close()
    t0 = time.perf_counter()
    dhost, drec = HostBatch.from_bam_device(eng, path)
    eng.synchronize()
    t1 = time.perf_counter()
    n_dev = len(eng.collect(drec, params))
    assert n_dev == n_host
    tm = eng.ingest_timings()
    dev_times.append((t1 - t0, tm))
    drec.free()
    dhost.close()
host_ingest = min(h[0] for h in host_times)
host_upload = min(h[1] for h in host_times)
dev_total = min(d[0] for d in dev_times)
tm = min(dev_times, key=lambda d: d[0])[1]
res.update({"host_ingest_ms": host_ingest * 1e3, "host_upload_ms": host_upload * 1e3, "host_threads": os.cpu_count(),
            "device_ingest_ms": dev_total * 1e3, "device_stages_ms": tm, "inflated_MB": tm["inflated_bytes"] / 1e6,
            "inflate_GBps_out": tm["inflated_bytes"] / tm["inflate"] / 1e6 if tm["inflate"] else None,
            "speedup_file_to_records": (host_ingest + host_upload) / dev_total})
print(json.dumps(res, indent=1))
