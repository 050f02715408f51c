"""TEST INFRASTRUCTURE (lives under oracle/: nothing in the product imports it).

The three CPU figures of BASELINE.md section 3 for the UNMODIFIED reference, run in the build container
(/root/reference through oracle/refrun.py; pysam / edlib are this repo's pure-python stand-ins when the real packages
are absent, so "file ->" figures include OUR BAM decoder, not htslib; the loop-only figure is the reference's own code on
pre-materialised tuples and does not depend on the stand-ins).  One core.

    python oracle/cpu_reference_figures.py [--config C1|C2]
"""
import argparse
import os
import platform
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import refrun
from svim_asm_b200 import bamio, synth

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
args = ap.parse_args()
assert refrun.available(), "needs the reference tree"
if args.config == "C1":
    cfg = synth.SynthConfig(["chr1"], [1_000_000], 50, 5e4, 1001, split_fraction=0.2)
else:
    cfg = synth.SynthConfig(["chr20"], [64_444_167], 2000, 1.0e7, 1002, giant_ops=250_000)
rb = synth.make_haploid(cfg)
tmp = tempfile.mkdtemp()
bam_path, fa_path = os.path.join(tmp, "h.bam"), os.path.join(tmp, "ref.fa")
bamio.write_bam(bam_path, rb, level=6)
bamio.write_fasta(fa_path, synth.random_reference(cfg), cfg.contig_names)
mods = refrun.modules()
real = {m: not getattr(sys.modules.get(m), "__file__", "").startswith(os.path.join(os.path.dirname(refrun.__file__), "shims"))
        for m in ("pysam", "edlib")}
print("config %s: %d alignments, %d CIGAR ops; host %s, %d cores visible, 1 used; real pysam %s, real edlib %s" % (
    args.config, rb.n_aln, rb.n_ops, platform.processor() or platform.machine(), os.cpu_count(), real["pysam"], real["edlib"]))
opts = refrun.parse_options(["haploid", os.path.join(tmp, "out"), bam_path, fa_path])

t0 = time.perf_counter()
cands, bam = refrun.collect(bam_path, opts)
t_collect = time.perf_counter() - t0
print("file -> candidates (analyze_alignment_file_coordsorted, SVIM_COLLECT.py:61-83): %.2f s, %d candidates, %.3g alignments/s, %.3g ops/s"
      % (t_collect, len(cands), rb.n_aln / t_collect, rb.n_ops / t_collect))

tuples = []
for i in range(rb.n_aln):
    lo, n = int(rb.cigar_off[i]), int(rb.n_cigar[i])
    ops = rb.cigar[lo:lo + n]
    tuples.append(list(zip((ops & 15).tolist(), (ops >> 4).tolist())))
fn = mods["SVIM_intra"].analyze_cigar_indel
t0 = time.perf_counter()
found = sum(len(fn(t, opts.min_sv_size)) for t in tuples)
t_loop = time.perf_counter() - t0
print("loop only (analyze_cigar_indel on pre-materialised tuples, SVIM_intra.py:8-30): %.2f s, %d indels, %.3g ops/s"
      % (t_loop, found, rb.n_ops / t_loop))

t0 = time.perf_counter()
refrun.run_cli(["haploid", os.path.join(tmp, "out"), bam_path, fa_path])
t_cli = time.perf_counter() - t0
n_rec = sum(1 for ln in open(os.path.join(tmp, "out", "variants.vcf")) if not ln.startswith("#"))
print("file -> variants.vcf (svim-asm haploid): %.2f s, %d records, %.3g alignments/s, %.3g ops/s"
      % (t_cli, n_rec, rb.n_aln / t_cli, rb.n_ops / t_cli))
