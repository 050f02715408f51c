"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/src/svim_asm, through oracle/shims) in the build container.

    python tests/golden/make_golden.py

Outputs (all small, committed):
  haploid/  h.bam(+.bai) ref.fa(+.fai) variants.vcf candidates.json      `svim-asm haploid`, defaults
  haploid_opts/ variants.vcf                                              same input, --min_sv_size 50 --symbolic_alleles
                                                                          --query_names --tandem_duplications_as_insertions
  diploid/  h1.bam h2.bam ref.fa variants.vcf candidates.json             `svim-asm diploid`, defaults + --query_names
  chimeric_read.npz / chimeric_read_errors.npz                            record images of the reference's own BAM
                                                                          fixtures (src/tests/*.bam) + the candidates the
                                                                          reference finds in them
The ##fileDate header line is wall-clock (SVIM_COMBINE.py:395) and is replaced by a fixed string.
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refrun                                   # noqa: E402
from svim_asm_b200 import bamio, synth                      # noqa: E402
from svim_asm_b200.engine import HostBatch                  # noqa: E402


def mask_date(path):
    lines = open(path).read().split("\n")
    lines = ["##fileDate=MASKED" if ln.startswith("##fileDate=") else ln for ln in lines]
    open(path, "w").write("\n".join(lines))


def canon_json(cands):
    return [list(refrun.canon(c)[:-1]) + [list(refrun.canon(c)[-1])] for c in cands]


def run_cli(workdir, argv, keep_as):
    out = os.path.join(workdir, "_out")
    shutil.rmtree(out, ignore_errors=True)
    refrun.run_cli([argv[0], out] + argv[1:])
    shutil.copy(os.path.join(out, "variants.vcf"), keep_as)
    mask_date(keep_as)
    shutil.rmtree(out)


def haploid():
    d = os.path.join(HERE, "haploid")
    os.makedirs(d, exist_ok=True)
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2"], [220000, 150000, 180000], 36, 2.4e4, 4101, sv_per_event=8e-3,
                            split_fraction=0.5, sv_max=1500)
    rb = synth.make_haploid(cfg)
    bamio.write_bam(os.path.join(d, "h.bam"), rb, level=6)
    bamio.write_fasta(os.path.join(d, "ref.fa"), synth.random_reference(cfg), cfg.contig_names)
    run_cli(d, ["haploid", os.path.join(d, "h.bam"), os.path.join(d, "ref.fa")], os.path.join(d, "variants.vcf"))
    d2 = os.path.join(HERE, "haploid_opts")
    os.makedirs(d2, exist_ok=True)
    run_cli(d, ["haploid", os.path.join(d, "h.bam"), os.path.join(d, "ref.fa"), "--min_sv_size", "50", "--symbolic_alleles",
                "--query_names", "--tandem_duplications_as_insertions", "--sample", "NA12878"], os.path.join(d2, "variants.vcf"))
    opts = refrun.parse_options(["haploid", "/tmp/x", os.path.join(d, "h.bam"), os.path.join(d, "ref.fa")])
    cands, _ = refrun.collect(os.path.join(d, "h.bam"), opts)
    json.dump(canon_json(cands), open(os.path.join(d, "candidates.json"), "w"))
    print("haploid:", len(cands), "candidates")


def diploid():
    d = os.path.join(HERE, "diploid")
    os.makedirs(d, exist_ok=True)
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2"], [220000, 150000, 180000], 40, 2.6e4, 4202, sv_per_event=8e-3,
                            split_fraction=0.5, sv_max=1500)
    rb1, rb2 = synth.make_diploid(cfg)
    bamio.write_bam(os.path.join(d, "h1.bam"), rb1, level=6)
    bamio.write_bam(os.path.join(d, "h2.bam"), rb2, level=6)
    bamio.write_fasta(os.path.join(d, "ref.fa"), synth.random_reference(cfg), cfg.contig_names)
    argv = ["diploid", os.path.join(d, "h1.bam"), os.path.join(d, "h2.bam"), os.path.join(d, "ref.fa"), "--query_names"]
    run_cli(d, argv, os.path.join(d, "variants.vcf"))
    mods = refrun.modules()
    opts = refrun.parse_options(["diploid", "/tmp/x"] + argv[1:])
    c1, bam1 = refrun.collect(os.path.join(d, "h1.bam"), opts)
    c2, _ = refrun.collect(os.path.join(d, "h2.bam"), opts)
    fasta = mods["pysam"].FastaFile(os.path.join(d, "ref.fa"))
    paired = mods["SVIM_COMBINE"].pair_candidates(c1, c2, fasta, bam1, opts)
    json.dump({"hap1": canon_json(c1), "hap2": canon_json(c2), "paired": canon_json(paired)},
              open(os.path.join(d, "candidates.json"), "w"))
    print("diploid:", len(c1), len(c2), "->", len(paired))


def reference_fixtures():
    """The reference's own BAM fixtures -> record images + what the reference finds in them."""
    for stem in ("chimeric_read", "chimeric_read_errors"):
        src = os.path.join(refrun.REFERENCE_SRC, "tests", stem + ".bam")
        host = HostBatch.from_bam(src)
        found = {}
        for min_sv in (40, 2):
            opts = refrun.parse_options(["haploid", "/tmp/x", src, "ref.fa", "--min_sv_size", str(min_sv)])
            cands, _ = refrun.collect(src, opts)
            found[str(min_sv)] = canon_json(cands)
        names = [host.query_name(i) for i in range(host.n_aln)]
        np.savez_compressed(os.path.join(HERE, stem + ".npz"), hdr=host.hdr, cigar=host.cigar, seg=host.seg,
                            sa_count=host.sa_count, seq4=host.seq4, seq_off=host.seq_off,
                            contig_lengths=host.contig_lengths, contig_names=np.array(host.contig_names),
                            query_names=np.array(names), expected=json.dumps(found))
        print(stem, {k: len(v) for k, v in found.items()})


if __name__ == "__main__":
    assert refrun.available(), "needs the reference tree"
    haploid()
    diploid()
    reference_fixtures()
