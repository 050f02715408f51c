"""Device ingest of one synthetic BAM, for profiling bgzf_inflate_kernel (ncu -k regex:bgzf_inflate).
    python tools/perf_inflate.py --scale 0.25"""
import argparse
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svim_asm_b200 import bamio, synth
from svim_asm_b200.engine import Engine, HostBatch

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.25)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--file", default=None, help="reuse this BAM if it exists, else write the synthetic one there (several library variants, one file)")
args = ap.parse_args()
lengths = [max(100000, int(x * args.scale)) for x in synth.HG38_LENGTHS]
cfg = synth.SynthConfig(list(synth.HG38_NAMES), lengths, max(24, int(40000 * args.scale)), 2.0e8 * args.scale, 1004,
                        giant_ops=int(1_000_000 * min(1.0, args.scale * 4)))
path = args.file or os.path.join(tempfile.mkdtemp(), "h.bam")
if not os.path.exists(path):
    rb = synth.make_haploid(cfg)
    bamio.write_bam(path, rb, level=1)
    del rb
eng = Engine(0)
for _ in range(args.runs):
    host, rec = HostBatch.from_bam_device(eng, path)
    tm = eng.ingest_timings()
    print({k: round(v, 3) for k, v in tm.items()}, "-> %.1f GB/s inflated (upload + inflate span)" % (tm["inflated_bytes"] / max(tm["inflate"], 1e-9) / 1e6), flush=True)
    rec.free()
    host.close()
eng.close()
