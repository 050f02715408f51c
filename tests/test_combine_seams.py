"""The pairing seams of SVIM_COMBINE.py -- form_partitions (:15-32), compute_distance (:35-102),
span_position_distance_breakends (:105-117), pair_haplotypes (:120-140), pair_haplotypes_breakends (:143-161) --
replayed on the GPU against what the UNMODIFIED reference returned for the same candidates
(tests/golden/combine_seams.json, written by tests/golden/make_golden_seams.py in the build container).
Includes the SURVEY App. D vectors: label order A2, B, A, A3 and the tight shared pair."""
import json
import os

import pytest

from svim_asm_b200 import SVCandidate as C

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "combine_seams.json")


class _Fasta(object):
    def __init__(self, bases):
        self.bases = bases

    def fetch(self, contig, start, end):
        return self.bases[contig][start:end]

    def get_reference_length(self, contig):
        return len(self.bases[contig])

    def close(self):
        pass


@pytest.fixture(scope="module")
def fx():
    d = json.load(open(GOLDEN))
    names, lengths = [n for n, _ in d["contigs"]], [l for _, l in d["contigs"]]
    bam = C._Lengths(names, lengths)
    items = []
    for hap, c in d["candidates"]:
        t, reads = c["type"], c["reads"]
        if t == "DEL":
            o = C.CandidateDeletion(c["source_contig"], c["source_start"], c["source_end"], reads, bam)
        elif t == "INV":
            o = C.CandidateInversion(c["source_contig"], c["source_start"], c["source_end"], reads, c["complete"], bam)
        elif t == "INS":
            o = C.CandidateInsertion(c["dest_contig"], c["dest_start"], c["dest_end"], reads, c["sequence"], bam)
        elif t == "DUP_TAN":
            o = C.CandidateDuplicationTandem(c["source_contig"], c["source_start"], c["source_end"], c["copies"], c["fully_covered"], reads, bam)
        elif t == "DUP_INT":
            o = C.CandidateDuplicationInterspersed(c["source_contig"], c["source_start"], c["source_end"], c["dest_contig"], c["dest_start"],
                                                   c["dest_end"], reads, bam, c["cutpaste"])
        else:
            o = C.CandidateBreakend(c["source_contig"], c["source_start"], c["source_direction"], c["dest_contig"], c["dest_start"],
                                    c["dest_direction"], reads, bam)
        items.append((hap, o))
    d["items"], d["fasta"] = items, _Fasta(d["bases"])
    d["index"] = {id(o): i for i, (_h, o) in enumerate(items)}
    return d


def _ids(groups, fx):
    return [[fx["index"][id(c)] for _h, c in g] for g in groups]


def test_span_position_distance_breakends_cpu(fx):
    from svim_asm_b200 import SVIM_COMBINE as combine
    for a, b, want in fx["span_position"]:
        assert combine.span_position_distance_breakends(a, b) == want


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["DEL", "INV", "INS", "DUP_TAN", "DUP_INT", "BND"])
def test_form_partitions(fx, engine, kind):
    from svim_asm_b200 import SVIM_COMBINE as combine
    items = [it for it in fx["items"] if it[1].type == kind]
    for max_distance in (1000, 150):
        got = combine.form_partitions(items, max_distance)
        assert _ids(got, fx) == fx["types"][kind]["partitions_%d" % max_distance]
    assert combine.form_partitions([], 1000) == []


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["DEL", "INV", "INS", "DUP_TAN", "DUP_INT"])
def test_pair_haplotypes(fx, engine, kind):
    from svim_asm_b200 import SVIM_COMBINE as combine
    items = [it for it in fx["items"] if it[1].type == kind]
    parts = combine.form_partitions(items, 1000)
    for thr in (200, 10):
        got = combine.pair_haplotypes(parts, fx["fasta"], thr)
        assert _ids(got, fx) == fx["types"][kind]["clusters_%d" % thr], (kind, thr)
    for i, j, want in fx["types"][kind]["distances"]:
        assert combine.compute_distance(fx["items"][i], fx["items"][j], fx["fasta"]) == want


@pytest.mark.gpu
def test_pair_haplotypes_breakends(fx, engine):
    from svim_asm_b200 import SVIM_COMBINE as combine
    items = [it for it in fx["items"] if it[1].type == "BND"]
    parts = combine.form_partitions(items, 1000)
    for thr in (0.3, 0.05):
        got = combine.pair_haplotypes_breakends(parts, thr)
        assert _ids(got, fx) == fx["types"]["BND"]["clusters_%g" % thr]


@pytest.mark.gpu
def test_app_d_label_order(fx, engine):
    """hap1 DEL A 105000-105100, A2 105300-105800, A3 105050-105400 + hap2 B 105900-106400: singletons in scipy's
    label order A2, B, A, A3; the neighbouring shared pair (107000 / 107002) merges."""
    from svim_asm_b200 import SVIM_COMBINE as combine
    items = [it for it in fx["items"] if it[1].type == "DEL" and it[1].source_contig == "chr1" and it[1].source_start >= 105000]
    clusters = combine.pair_haplotypes(combine.form_partitions(items, 1000), fx["fasta"], 200)
    starts = [[c.source_start for _h, c in cl] for cl in clusters]
    assert starts == [[107000, 107002], [105300], [105900], [105000], [105050]]
