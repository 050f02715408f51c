"""File -> variants.vcf through the drop-in CLI on a synthetic diploid sample (a tenth of a human genome by default), with
a profile of where the wall time goes (ingest, COLLECT, PAIR, candidate objects, VCF text).
Needs a GPU:  python tools/perf_cli.py --scale 0.1"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from svim_asm_b200 import bamio, cli, synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--dir", default=None, help="where the input files are written (default: the system temp directory)")
args = ap.parse_args()

lengths = [max(100000, int(x * args.scale)) for x in synth.HG38_LENGTHS]
cfg = synth.SynthConfig(list(synth.HG38_NAMES), lengths, max(24, int(40000 * args.scale)), 2.0e8 * args.scale, 1004,
                        giant_ops=int(1_000_000 * min(1.0, args.scale * 4)))
t0 = time.time()
rb1, rb2 = synth.make_diploid(cfg)
ref = synth.random_reference(cfg)
tmp = tempfile.mkdtemp(dir=args.dir)
p1, p2, pf = os.path.join(tmp, "h1.bam"), os.path.join(tmp, "h2.bam"), os.path.join(tmp, "ref.fa")
bamio.write_bam(p1, rb1, level=1)
bamio.write_bam(p2, rb2, level=1)
bamio.write_fasta(pf, ref, cfg.contig_names)
print("inputs: 2 x %d alignments, %d + %d ops, BAMs %.0f + %.0f MB, FASTA %.0f MB, written in %.0fs" % (
    rb1.n_aln, rb1.n_ops, rb2.n_ops, os.path.getsize(p1) / 1e6, os.path.getsize(p2) / 1e6, os.path.getsize(pf) / 1e6, time.time() - t0), flush=True)
del rb1, rb2, ref
for run in range(args.runs):
    out = os.path.join(tmp, "out%d" % run)
    prof = cProfile.Profile()
    t0 = time.perf_counter()
    prof.enable()
    cli.main(["diploid", out, p1, p2, pf])
    prof.disable()
    wall = time.perf_counter() - t0
    n_lines = sum(1 for ln in open(os.path.join(out, "variants.vcf")) if not ln.startswith("#"))
    print("run %d: %.3f s wall, %d VCF records" % (run, wall, n_lines), flush=True)
    if run == args.runs - 1:
        from svim_asm_b200.runtime import get_engine
        print("stages of the last device ingest (ms):", {k: round(v, 2) for k, v in get_engine().ingest_timings().items()}, flush=True)
        s = io.StringIO()
        pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(28)
        print("\n".join(s.getvalue().split("\n")[:60]))

# the same run with the python writer (Candidate objects, one reference.fetch per allele): identical bytes, and its time
def body(path):
    return [ln for ln in open(path).read().split("\n") if not ln.startswith("##fileDate")]
os.environ["SVIM_ASM_B200_VCF"] = "host"
out = os.path.join(tmp, "out_host_writer")
t0 = time.perf_counter()
cli.main(["diploid", out, p1, p2, pf])
wall = time.perf_counter() - t0
same = body(os.path.join(out, "variants.vcf")) == body(os.path.join(tmp, "out%d" % (args.runs - 1), "variants.vcf"))
print("python writer: %.3f s wall; VCF identical to the device writer's: %s (%.1f MB)" % (
    wall, same, os.path.getsize(os.path.join(out, "variants.vcf")) / 1e6), flush=True)
assert same
