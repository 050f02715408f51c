"""The CPU oracle (oracle/port.py) against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and against the reference's own BAM fixtures.
These are the tests that pin the oracle on boxes where /root/reference does not exist."""
import json
import os

import numpy as np
import pytest

from oracle import port
from tests import util

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _as_tuples(items):
    return [tuple(x[:-1]) + (tuple(x[-1]),) for x in items]


def _canon(rows, hosts, names):
    return [util.canon_row(r, hosts, names) for r in rows]


def test_haploid_collect_matches_reference(built_library):
    from svim_asm_b200.engine import HostBatch
    host = HostBatch.from_bam(os.path.join(GOLDEN, "haploid", "h.bam"))
    want = _as_tuples(json.load(open(os.path.join(GOLDEN, "haploid", "candidates.json"))))
    got = _canon(port.collect(host, port.Params()), {0: host}, host.contig_names)
    assert got == want
    assert len({w[0] for w in want}) >= 5


def test_diploid_pair_matches_reference(built_library, oracle_clib):
    from svim_asm_b200.engine import HostBatch
    from svim_asm_b200.fasta import FastaFile
    d = os.path.join(GOLDEN, "diploid")
    h1, h2 = HostBatch.from_bam(os.path.join(d, "h1.bam")), HostBatch.from_bam(os.path.join(d, "h2.bam"))
    want = json.load(open(os.path.join(d, "candidates.json")))
    p = port.Params()
    r1, r2 = port.collect(h1, p, hap=1), port.collect(h2, p, hap=2)
    assert _canon(r1, {1: h1, 2: h2}, h1.contig_names) == _as_tuples(want["hap1"])
    assert _canon(r2, {1: h1, 2: h2}, h1.contig_names) == _as_tuples(want["hap2"])
    fasta = FastaFile(os.path.join(d, "ref.fa"))

    def fetch(tid, s, e):
        return fasta.fetch(h1.contig_names[tid], s, e).upper().encode()
    paired = port.pair(r1, r2, h1, h2, fetch, p)
    assert _canon(paired, {1: h1, 2: h2}, h1.contig_names) == _as_tuples(want["paired"])


class _NpzHost(object):
    """Record image stored by make_golden.reference_fixtures()."""

    def __init__(self, path):
        z = np.load(path)
        self.hdr, self.cigar, self.seg, self.sa_count = z["hdr"], z["cigar"], z["seg"], z["sa_count"]
        self.seq4, self.seq_off = z["seq4"], z["seq_off"]
        self.contig_lengths, self.contig_names = z["contig_lengths"], [str(x) for x in z["contig_names"]]
        self._names = [str(x) for x in z["query_names"]]
        self.expected = json.loads(str(z["expected"]))

    n_aln = property(lambda self: self.hdr.shape[0])

    def query_name(self, i):
        return self._names[int(i)]

    def sequence_slice(self, i, start, length):
        from svim_asm_b200.engine import HostBatch
        return HostBatch.sequence_slice(self, i, start, length)


@pytest.mark.parametrize("stem,n_default,n_small", [("chimeric_read", 1, 31), ("chimeric_read_errors", 1, 39)])
def test_reference_bam_fixtures(built_library, stem, n_default, n_small):
    host = _NpzHost(os.path.join(GOLDEN, stem + ".npz"))
    for min_sv, n in (("40", n_default), ("2", n_small)):
        want = _as_tuples(host.expected[min_sv])
        got = _canon(port.collect(host, port.Params(min_sv_size=int(min_sv))), {0: host}, host.contig_names)
        assert len(want) == n and got == want
    if stem == "chimeric_read":
        # DUP_TAN chr21:35351848-35352405 copies=3 fully_covered=True
        assert _as_tuples(host.expected["40"])[0][:6] == ("DUP_TAN", "chr21", 35351848, 35352405, 3, True)
