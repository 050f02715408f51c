"""Golden vectors for the pairing seams of SVIM_COMBINE.py: seeded candidates of every class with haplotypes, and what the
UNMODIFIED reference's form_partitions (:15-32), pair_haplotypes (:120-140), pair_haplotypes_breakends (:143-161),
compute_distance (:35-102) and span_position_distance_breakends (:105-117) return for them.

    python tests/golden/make_golden_seams.py      (needs /root/reference; writes tests/golden/combine_seams.json)

Outputs are stored as index structures over the candidate list (partitions / clusters = lists of candidate indices), so the
fixture does not depend on object identity.  Includes the SURVEY App. D pairing vectors (label order A2, B, A, A3; the
200 / 200.5 cut).  tests/test_combine_seams.py replays them through svim_asm_b200.SVIM_COMBINE on the GPU.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refrun                                   # noqa: E402


class Fasta(object):
    def __init__(self, bases):
        self.bases = bases

    def fetch(self, contig, start, end):
        return self.bases[contig][start:end]

    def get_reference_length(self, contig):
        return len(self.bases[contig])

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
    C, COMB = mods["SVCandidate"], mods["SVIM_COMBINE"]
    rng = np.random.default_rng(4711)
    names = ["chr2", "chr10", "chr1"]
    lengths = {"chr2": 90000, "chr10": 70000, "chr1": 110000}
    bases = {n: "".join(rng.choice(list("ACGTacgtN"), lengths[n], p=[.22, .22, .22, .22, .03, .03, .02, .02, .02]).tolist())
             for n in names}
    bam, fasta = Bam(lengths), Fasta(bases)
    cands = []          # (haplotype, candidate)

    def seq(n):
        return "".join(rng.choice(list("ACGT"), n).tolist())

    def mutate(s, k):
        b = list(s)
        for _ in range(k):
            i = int(rng.integers(0, max(1, len(b))))
            kind = int(rng.integers(0, 3))
            if kind == 0 and b:
                b[i] = "ACGT"[int(rng.integers(0, 4))]
            elif kind == 1:
                b.insert(i, "ACGT"[int(rng.integers(0, 4))])
            elif b:
                del b[i]
        return "".join(b)
    k = 0
    for contig in names:
        L = lengths[contig]
        pos = 300
        while pos < L - 2500:
            kind = ("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND")[int(rng.integers(0, 6))]
            size = int(rng.integers(40, 260))
            scenario = rng.random()
            # shared (both haplotypes, second shifted / edited a little), private, or a crowd of neighbours
            copies = [(1, 0, 0)]
            if scenario < 0.45:
                copies.append((2, int(rng.integers(-8, 9)), int(rng.integers(0, 12))))
            elif scenario < 0.6:
                copies = [(2, 0, 0)]
            elif scenario < 0.8:
                copies += [(2, int(rng.integers(-300, 300)), int(rng.integers(0, 80))), (1, int(rng.integers(100, 700)), 0),
                           (2, int(rng.integers(100, 700)), int(rng.integers(0, 30)))]
            elif scenario < 0.85:
                copies = [(1 + (j % 2), 60 * j, j) for j in range(12)]                 # > 10 members: dropped
            ins = seq(size)
            other = names[int(rng.integers(0, 3))]
            opos = int(rng.integers(300, lengths[other] - 600))
            for hap, shift, edits in copies:
                s = max(0, pos + shift)
                reads = ["ctg_h%d_%d" % (hap, k)]
                k += 1
                if kind == "DEL":
                    c = C.CandidateDeletion(contig, s, s + size + edits % 5, reads, bam)
                elif kind == "INS":
                    body = mutate(ins, edits)
                    c = C.CandidateInsertion(contig, s, s + len(body), reads, body, bam)
                elif kind == "INV":
                    c = C.CandidateInversion(contig, s, s + size, reads, bool(edits % 2), bam)
                elif kind == "DUP_TAN":
                    c = C.CandidateDuplicationTandem(contig, s, s + size, 1 + edits % 3, bool(edits % 2), reads, bam)
                elif kind == "DUP_INT":
                    c = C.CandidateDuplicationInterspersed(other, opos + shift // 4, opos + shift // 4 + size, contig, s, s + size, reads, bam,
                                                           bool(edits % 2))
                else:
                    d1, d2 = ("fwd", "rev")[(edits // 2) % 2 if scenario >= 0.6 else 0], ("fwd", "rev")[edits % 2 if scenario >= 0.6 else 1]
                    c = C.CandidateBreakend(contig, s, d1, other, opos + shift, d2, reads, bam)
                cands.append((hap, c))
            pos += int(rng.integers(1200, 3500))
    # label order A2, B, A, A3 (four singletons) and a tight shared pair
    for hap, s, e in ((1, 105000, 105100), (1, 105300, 105800), (1, 105050, 105400), (2, 105900, 106400), (1, 107000, 107100), (2, 107002, 107102)):
        cands.append((hap, C.CandidateDeletion("chr1", s, e, ["appD_%d_%d" % (hap, s)], bam)))
    order = rng.permutation(len(cands))
    cands = [cands[i] for i in order]
    index = {id(c): i for i, (_h, c) in enumerate(cands)}

    def plain(c):
        d = {"type": c.type, "reads": list(c.reads)}
        for f in ("source_contig", "source_start", "source_end", "dest_contig", "dest_start", "dest_end", "sequence", "copies",
                  "fully_covered", "complete", "cutpaste", "source_direction", "dest_direction"):
            if hasattr(c, f):
                v = getattr(c, f)
                d[f] = v if isinstance(v, str) else (bool(v) if isinstance(v, (bool, np.bool_)) else int(v))
        return d

    out = {"contigs": [[n, lengths[n]] for n in names], "bases": bases,
           "candidates": [[hap, plain(c)] for hap, c in cands], "types": {}}
    for t in ("DEL", "INV", "INS", "DUP_TAN", "DUP_INT", "BND"):
        items = [(hap, c) for hap, c in cands if c.type == t]
        entry = {}
        for max_distance in (1000, 150):
            parts = COMB.form_partitions(items, max_distance)
            entry["partitions_%d" % max_distance] = [[index[id(c)] for _h, c in part] for part in parts]
        parts = COMB.form_partitions(items, 1000)
        if t == "BND":
            for thr in (0.3, 0.05):
                clusters = COMB.pair_haplotypes_breakends(parts, thr)
                entry["clusters_%g" % thr] = [[index[id(c)] for _h, c in cl] for cl in clusters]
        else:
            for thr in (200, 10):
                clusters = COMB.pair_haplotypes(parts, fasta, thr)
                entry["clusters_%d" % thr] = [[index[id(c)] for _h, c in cl] for cl in clusters]
            dist = []
            for part in parts:
                if 2 <= len(part) <= 4:
                    for i in range(len(part) - 1):
                        for j in range(i + 1, len(part)):
                            dist.append([index[id(part[i][1])], index[id(part[j][1])], int(COMB.compute_distance(part[i], part[j], fasta))])
            entry["distances"] = dist
        out["types"][t] = entry
    spd = []
    for _ in range(40):
        a = [int(rng.integers(1, 3)), int(rng.integers(0, 5000)), int(rng.integers(0, 2)), int(rng.integers(0, 5000)), int(rng.integers(0, 2))]
        b = [int(rng.integers(1, 3)), a[1] + int(rng.integers(-900, 900)), int(rng.integers(0, 2)), a[3] + int(rng.integers(-900, 900)),
             int(rng.integers(0, 2))]
        spd.append([a, b, float(COMB.span_position_distance_breakends(np.array(a), np.array(b)))])
    out["span_position"] = spd
    json.dump(out, open(os.path.join(HERE, "combine_seams.json"), "w"))
    print(len(cands), "candidates;", {t: (len(e["partitions_1000"]), len(e.get("clusters_200", e.get("clusters_0.3")))) for t, e in out["types"].items()})


if __name__ == "__main__":
    main()
