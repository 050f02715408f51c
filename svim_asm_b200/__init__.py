"""svim_asm_b200 -- B200-native (sm_100a) implementation of SVIM-asm's alignment-scan and haplotype-pairing
hot path behind the reference's python seams.  See DESIGN.md and INTEGRATION.md."""
__version__ = "1.0.3+b200.1"
