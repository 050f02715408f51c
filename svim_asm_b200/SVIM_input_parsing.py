"""Command-line surface of SVIM-asm (reference src/svim_asm/SVIM_input_parsing.py:7-264): the sub-commands
`haploid` and `diploid` with the same positionals, flags and defaults, built from one option table."""
import argparse
import os
import sys

_COLLECT = (  # flag, default, help                                        (SVIM_input_parsing.py:45-95)
    ("min_mapq", 20, "Minimum mapping quality of alignments to consider. Alignments with a lower mapping quality are ignored."),
    ("min_sv_size", 40, "Minimum SV size to detect."),
    ("max_sv_size", 100000, "Maximum SV size to detect; larger events are reported as translocation breakpoints."),
    ("query_gap_tolerance", 50, "Maximum tolerated gap between adjacent alignment segments on the query."),
    ("query_overlap_tolerance", 50, "Maximum tolerated overlap between adjacent alignment segments on the query."),
    ("reference_gap_tolerance", 50, "Maximum tolerated gap between adjacent alignment segments on the reference."),
    ("reference_overlap_tolerance", 50, "Maximum tolerated overlap between adjacent alignment segments on the reference."),
)
_PAIR = (     #                                                             (SVIM_input_parsing.py:221,229)
    ("partition_max_distance", 1000, "Maximum distance in bp between SVs in a partition."),
    ("max_edit_distance", 200, "Maximum edit distance between both alleles to be paired."),
)
_OUTPUT_FLAGS = (
    ("symbolic_alleles", "Use symbolic alleles, such as <DEL> or <INV> in the VCF output."),
    ("tandem_duplications_as_insertions", "Represent tandem duplications as insertions in output VCF."),
    ("interspersed_duplications_as_insertions", "Represent interspersed duplications as insertions in output VCF."),
    ("query_names", "Output names of supporting query sequences in INFO tag of VCF."),
)


def _add_common(sub, diploid):
    sub.add_argument("working_dir", type=os.path.abspath,
                     help="Working and output directory. Existing files are overwritten; it is created if missing.")
    if diploid:
        sub.add_argument("bam_file1", type=str, help="Coordinate-sorted, indexed BAM of the first haplotype")
        sub.add_argument("bam_file2", type=str, help="Coordinate-sorted, indexed BAM of the second haplotype")
    else:
        sub.add_argument("bam_file", type=str, help="Coordinate-sorted, indexed BAM of query assembly vs reference")
    sub.add_argument("genome", type=str, help="Reference genome the assembly was aligned to (FASTA with .fai)")
    sub.add_argument("--verbose", action="store_true", help="Enable more verbose logging (default: %(default)s)")
    grp = sub.add_argument_group("COLLECT")
    for flag, default, text in _COLLECT:
        grp.add_argument("--" + flag, type=int, default=default, help=text + " (default: %(default)s)")
    if diploid:
        grp = sub.add_argument_group("PAIR")
        for flag, default, text in _PAIR:
            grp.add_argument("--" + flag, type=int, default=default, help=text + " (default: %(default)s)")
    grp = sub.add_argument_group("OUTPUT")
    grp.add_argument("--sample", type=str, default="Sample", help="Sample ID to include in output vcf file (default: %(default)s)")
    grp.add_argument("--types", type=str, default="DEL,INS,INV,DUP:TANDEM,DUP:INT,BND",
                     help="Comma-separated SV types to include in the output VCF (default: %(default)s)")
    for flag, text in _OUTPUT_FLAGS:
        grp.add_argument("--" + flag, action="store_true", help=text + " (default: %(default)s)")


def parse_arguments(program_version, arguments=None):
    if arguments is None:
        arguments = sys.argv[1:]
    parser = argparse.ArgumentParser(
        formatter_class=argparse.RawDescriptionHelpFormatter,
        description="SVIM-asm hot path on B200: structural variant calling from genome-genome alignments.\n"
                    "COLLECT (GPU) detects SVs from BAM alignments; PAIR (GPU, diploid only) merges the calls of the two\n"
                    "haplotypes; OUTPUT writes the VCF.")
    subparsers = parser.add_subparsers(help="modes", dest="sub")
    parser.add_argument("--version", "-v", action="version", version="%(prog)s {0}".format(program_version))
    _add_common(subparsers.add_parser("haploid", help="Detect SVs from the alignment of a haploid query assembly"), False)
    _add_common(subparsers.add_parser("diploid", help="Detect SVs from the alignments of both haplotypes of a diploid assembly"), True)
    return parser.parse_args(arguments)
