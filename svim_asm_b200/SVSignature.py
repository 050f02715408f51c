"""Compatibility names of the sibling project SVIM (read-based caller) that BASELINE.json's north_star mentions.

The svim-asm reference tree does NOT contain `SVSignature`: only its stale test does
(src/tests/test_Signature.py:3), and that test is the whole specification available (SURVEY.md section 0 and
8f row 3): constructor `(contig, start, end, signature, read)`, `get_source`, `get_key`, `position_distance_to`,
`as_string`.  svim-asm's own object surface is `SVCandidate.Candidate*` (svim_asm_b200/SVCandidate.py).
`SVIM_clustering.partition_and_cluster` has no definition anywhere in the reference tree (its test pins only the
number of clusters and a score range that depends on SVIM's unpublished scoring), so it is not provided; the
svim-asm equivalent is `SVIM_COMBINE.form_partitions` + `pair_haplotypes` (SVIM_COMBINE.py:15-32,120-161).
"""


class Signature(object):
    """src/tests/test_Signature.py:8-27."""
    type = None

    def __init__(self, contig, start, end, signature, read):
        self.contig = contig
        self.start = start
        self.end = end
        self.signature = signature
        self.read = read
        if self.end < self.start:
            raise ValueError("Signature end is smaller than its start")

    def get_source(self):
        return (self.contig, self.start, self.end)                          # test_Signature.py:11

    def get_key(self):
        return (self.type, self.contig, (self.start + self.end) // 2)       # test_Signature.py:12 ("DEL", "chr1", 200)

    def position_distance_to(self, other):
        """Distance of the midpoints; infinite across contigs or types (test_Signature.py:20-22)."""
        if self.type != other.type or self.contig != other.contig:
            return float("Inf")
        return abs((self.start + self.end) // 2 - (other.start + other.end) // 2)

    def as_string(self, sep="\t"):
        return sep.join(["{0}", "{1}", "{2}", "{3}", "{4}"]).format(
            self.contig, self.start, self.end, "{0};{1}".format(self.type, self.signature), self.read)   # :27-28


class SignatureDeletion(Signature):
    type = "DEL"


class SignatureInsertion(Signature):
    type = "INS"

    def __init__(self, contig, start, end, signature, read, sequence):
        Signature.__init__(self, contig, start, end, signature, read)
        self.sequence = sequence


class SignatureInversion(Signature):
    type = "INV"


class SignatureDuplicationTandem(Signature):
    type = "DUP_TAN"

    def __init__(self, contig, start, end, copies, signature, read):
        Signature.__init__(self, contig, start, end, signature, read)
        self.copies = copies
