"""The SVIM compatibility names (north_star's "Signature" surface): the asserts of the reference's stale
src/tests/test_Signature.py:8-27, which is the only specification of that surface in the reference tree."""
from svim_asm_b200.SVSignature import SignatureDeletion, SignatureInsertion


def test_accessors():
    deletion = SignatureDeletion("chr1", 100, 300, "cigar", "read1")
    assert deletion.get_source() == ("chr1", 100, 300)
    assert deletion.get_key() == ("DEL", "chr1", 200)


def test_position_distance_to():
    deletion1 = SignatureDeletion("chr1", 100, 300, "cigar", "read1")
    deletion2 = SignatureDeletion("chr1", 150, 200, "cigar", "read2")
    deletion3 = SignatureDeletion("chr2", 150, 200, "cigar", "read2")
    insertion = SignatureInsertion("chr1", 150, 200, "cigar", "read2", "ACGTAGTAGCTAGCTTTGCTAGCATTAGCGACTGCTTACGCAGCTCCCTA")
    assert deletion1.position_distance_to(deletion2) == 25
    assert deletion1.position_distance_to(deletion3) == float("Inf")
    assert deletion1.position_distance_to(insertion) == float("Inf")


def test_as_string():
    deletion1 = SignatureDeletion("chr1", 100, 300, "cigar", "read1")
    assert deletion1.as_string() == "chr1\t100\t300\tDEL;cigar\tread1"
    assert deletion1.as_string(":") == "chr1:100:300:DEL;cigar:read1"
