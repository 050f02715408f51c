"""Golden vectors for the VCF writer: seeded random candidates of every class, written by the UNMODIFIED reference's
write_final_vcf (SVIM_COMBINE.py:379-477, Candidate*.get_vcf_entry* of SVCandidate.py) under several option sets.

    python tests/golden/make_golden_vcf.py        (needs /root/reference; writes tests/golden/vcf_writer.json)

The fixture holds the reference bases, the candidates as the reference's constructors left them (clamped, breakends
normalised) and the record lines of every option set.  tests/test_vcf_writer.py checks the python writer, the host
build of the device plan (csrc/vcf_core.cuh) and -- on a GPU -- svb_vcf_body against it.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refrun                                   # noqa: E402

OPTION_SETS = {
    "default": dict(types="DEL,INS,INV,DUP:TANDEM,DUP:INT,BND"),
    "symbolic": dict(types="DEL,INS,INV,DUP:TANDEM,DUP:INT,BND", symbolic_alleles=True),
    "dups_as_ins": dict(types="DEL,INS,INV,DUP:TANDEM,DUP:INT,BND", tandem_duplications_as_insertions=True,
                        interspersed_duplications_as_insertions=True),
    "dups_as_ins_symbolic": dict(types="DEL,INS,INV,DUP:TANDEM,DUP:INT,BND", tandem_duplications_as_insertions=True,
                                 interspersed_duplications_as_insertions=True, symbolic_alleles=True),
    "subset": dict(types="INS,BND,DUP:INT", interspersed_duplications_as_insertions=False),
    "no_ins": dict(types="DEL,DUP:TANDEM,DUP:INT", tandem_duplications_as_insertions=True,
                   interspersed_duplications_as_insertions=True),
}


class Fasta(object):
    def __init__(self, bases):
        self.bases = bases

    def fetch(self, contig, start, end):
        return self.bases[contig][start:end]

    def close(self):
        pass


class Bam(object):
    def __init__(self, lengths):
        self.lengths = lengths

    def get_reference_length(self, contig):
        return self.lengths[contig]


def main():
    assert refrun.available(), "needs the reference tree"
    mods = refrun.modules()
    C = mods["SVCandidate"]
    rng = np.random.default_rng(777)
    names = ["chr2", "chr10", "chr1", "chrX", "1", "scaffold_12", "scaffold_3", "chr01"]
    lengths = {n: int(rng.integers(2500, 6000)) for n in names}
    bases = {n: "".join(rng.choice(list("ACGTacgtNn"), lengths[n], p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .02]).tolist())
             for n in names}
    bam, gts = Bam(lengths), ["1/1", "1/0", "0/1"]

    def span(n, longest=400):
        L = lengths[n]
        mode = int(rng.integers(0, 8))
        if mode == 0:
            return 0, int(rng.integers(1, longest))
        if mode == 1:
            return L - int(rng.integers(1, longest)), L + int(rng.integers(0, 50))      # clamped by the constructor
        if mode == 2:
            return -int(rng.integers(1, 30)), int(rng.integers(1, longest))
        s = int(rng.integers(0, L - longest))
        return s, s + int(rng.integers(0, longest))

    def gt():
        return gts[int(rng.integers(0, 3))]

    def reads(k):
        return ["read_%d" % k] + (["mate_%d" % k] if rng.random() < 0.4 else [])
    cands, k = [], 0
    for _ in range(30):
        n = names[int(rng.integers(0, len(names)))]
        s, e = span(n)
        cands.append(C.CandidateDeletion(n, s, e, reads(k), bam, gt())); k += 1
        s, e = span(n)
        cands.append(C.CandidateInversion(n, s, e, reads(k), bool(rng.integers(0, 2)), bam, gt())); k += 1
        s, e = span(n)
        seq = "".join(rng.choice(list("ACGTNRY="), int(rng.integers(0, 300)), p=[.23, .23, .23, .23, .02, .02, .02, .02]).tolist())
        cands.append(C.CandidateInsertion(n, s, e, reads(k), seq, bam, gt())); k += 1
        s, e = span(n, 150)
        cands.append(C.CandidateDuplicationTandem(n, s, e, int(rng.integers(1, 5)), bool(rng.integers(0, 2)), reads(k), bam, gt())); k += 1
        m = names[int(rng.integers(0, len(names)))]
        s, e = span(n, 200)
        ds, de = span(m, 200)
        cands.append(C.CandidateDuplicationInterspersed(n, s, e, m, ds, de, reads(k), bam, bool(rng.integers(0, 2)), gt())); k += 1
        p1 = int(rng.integers(-5, lengths[n] + 5))
        p2 = int(rng.integers(-5, lengths[m] + 5))
        d1, d2 = ("fwd", "rev")[int(rng.integers(0, 2))], ("fwd", "rev")[int(rng.integers(0, 2))]
        cands.append(C.CandidateBreakend(n, p1, d1, m, p2, d2, reads(k), bam, gt())); k += 1
    # equal sort keys across classes and inside one class: the stable order of the append sequence decides
    cands.append(C.CandidateDeletion("chr1", 100, 160, reads(k), bam, "1/1")); k += 1
    cands.append(C.CandidateDeletion("chr1", 100, 160, reads(k), bam, "0/1")); k += 1
    cands.append(C.CandidateInversion("chr1", 99, 160, reads(k), True, bam, "1/0")); k += 1
    cands.append(C.CandidateInsertion("chr1", 100, 160, reads(k), "ACGT", bam, "1/1")); k += 1
    cands.append(C.CandidateInsertion("chr01", 0, 40, reads(k), "TTTT", bam, "1/1")); k += 1      # no anchor base at position 0
    order = rng.permutation(len(cands))
    cands = [cands[i] for i in order]

    def plain(c):
        d = {"type": c.type, "genotype": c.genotype, "reads": list(c.reads)}
        for f in ("source_contig", "source_start", "source_end", "dest_contig", "dest_start", "dest_end", "sequence", "copies",
                  "fully_covered", "complete", "cutpaste", "source_direction", "dest_direction"):
            if hasattr(c, f):
                v = getattr(c, f)
                d[f] = v if isinstance(v, str) else (bool(v) if isinstance(v, (bool, np.bool_)) else int(v))
        return d
    outputs = {}
    for key, o in OPTION_SETS.items():
        tmp = tempfile.mkdtemp()
        options = argparse.Namespace(working_dir=tmp, sample="Sample", query_names=False,
                                     symbolic_alleles=o.get("symbolic_alleles", False),
                                     tandem_duplications_as_insertions=o.get("tandem_duplications_as_insertions", False),
                                     interspersed_duplications_as_insertions=o.get("interspersed_duplications_as_insertions", False))
        by = {t: [c for c in cands if c.type == t] for t in ("DEL", "INS", "INV", "DUP_TAN", "BND", "DUP_INT")}
        mods["SVIM_COMBINE"].write_final_vcf(by["DUP_INT"], by["INV"], by["DUP_TAN"], by["DEL"], by["INS"], by["BND"], "1.0.3",
                                             names, [lengths[n] for n in names], [t.strip() for t in o["types"].split(",")],
                                             Fasta(bases), options)
        text = open(os.path.join(tmp, "variants.vcf")).read()
        outputs[key] = "".join(ln + "\n" for ln in text.split("\n") if ln and not ln.startswith("#"))
        header = [ln for ln in text.split("\n") if ln.startswith("#") and not ln.startswith("##fileDate")]
        outputs[key + "/header"] = header
    json.dump({"contigs": [[n, lengths[n]] for n in names], "bases": bases, "candidates": [plain(c) for c in cands],
               "option_sets": OPTION_SETS, "outputs": outputs}, open(os.path.join(HERE, "vcf_writer.json"), "w"))
    print(len(cands), "candidates;", {k: len(v) for k, v in outputs.items() if not k.endswith("/header")})


if __name__ == "__main__":
    main()
