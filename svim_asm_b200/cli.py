"""`svim-asm haploid|diploid ...` (reference src/svim_asm/svim-asm:23-185) on top of the GPU hot path:
same sub-commands, flags, log lines, checks and output files (variants.vcf, SVIM_<date>.log)."""
import logging
import os
import sys
from time import localtime, strftime

__version__ = "1.0.3"

def _open_sorted_bam(path, which):
    """svim-asm:63-80 / 85-120: header must say SO:coordinate and an index must sit next to the file."""
    from .bamfile import AlignmentFile
    from .runtime import get_engine
    bam = AlignmentFile(path, engine=get_engine())      # device ingest: the records stay in HBM for COLLECT
    label = "" if which is None else ("first " if which == 1 else "second ")
    try:
        sorted_ok = bam.header["HD"]["SO"] == "coordinate"
    except KeyError:
        # (svim-asm:79,100,119: only the first-BAM message of the diploid mode lacks the trailing " Exiting..")
        logging.error("Is the given {0}input BAM file coordinate-sorted? It does not contain a sorting order in its "
                      "header line.{1}".format(label, "" if which == 1 else " Exiting.."))
        return None
    if not sorted_ok:
        logging.error("{0} BAM file needs to be coordinate-sorted. Exiting..".format(
            "Input" if which is None else "The " + label + "input"))
        return None
    try:
        bam.check_index()
    except ValueError:
        logging.error("{0} BAM file is missing an index. Please generate with 'samtools index'. Exiting..".format(
            "Input" if which is None else "The " + label + "input"))
        return None
    return bam


def main(argv=None):
    from .SVIM_input_parsing import parse_arguments
    options = parse_arguments(program_version=__version__, arguments=argv)
    if not options.sub:
        print("Please choose one of the two modes ('haploid' or 'diploid'). See --help for more information.")
        return
    formatter = logging.Formatter("%(asctime)s [%(levelname)-7.7s]  %(message)s")
    root = logging.getLogger()
    root.setLevel(logging.DEBUG if options.verbose else logging.INFO)
    os.makedirs(options.working_dir, exist_ok=True)
    handlers = [logging.FileHandler("{0}/SVIM_{1}.log".format(options.working_dir, strftime("%y%m%d_%H%M%S", localtime())), mode="w"),
                logging.StreamHandler()]
    for h in handlers:
        h.setFormatter(formatter)
        root.addHandler(h)
    try:
        return _run(options)
    finally:
        from .SVIM_COMBINE import drop_prefetches
        drop_prefetches()                  # a run that stopped early (bad second BAM, ...) must not leave the loader running
        for h in handlers:
            root.removeHandler(h)
            h.close()


def _prefetching():
    return os.environ.get("SVIM_ASM_B200_PREFETCH", "1") != "0" and os.environ.get("SVIM_ASM_B200_INGEST", "device") != "host"


def _prefetch_reference(options, contig_names):
    """Start the FASTA -> HBM load in the background (needed by PAIR and by the VCF alleles) once the contigs are known:
    before the first ingest when the header's contig names can be peeked at, else right after it."""
    if not _prefetching() or not contig_names:
        return
    if os.path.exists(options.genome) and os.path.exists(options.genome + ".fai"):
        from .SVIM_COMBINE import ReferencePrefetch, prefetch_pending
        if not prefetch_pending(options.genome, contig_names):
            ReferencePrefetch(options.genome, contig_names)


def _run(options):
    from .fasta import FastaFile
    from .SVIM_COLLECT import analyze_alignment_file_coordsorted
    from .SVIM_COMBINE import pair_candidates, write_final_vcf
    from .SVIM_plot import plot_sv_lengths
    logging.info("****************** Start SVIM-asm, version {0} ******************".format(__version__))
    logging.info("CMD: python3 {0}".format(" ".join(sys.argv)))
    logging.info("WORKING DIR: {0}".format(os.path.abspath(options.working_dir)))
    for arg in vars(options):
        logging.info("PARAMETER: {0}, VALUE: {1}".format(arg, getattr(options, arg)))
    logging.info("****************** STEP 1: COLLECT ******************")
    if options.sub == "haploid":
        logging.info("MODE: haploid")
        logging.info("INPUT: {0}".format(os.path.abspath(options.bam_file)))
        if _prefetching():
            from .bamio import read_reference_names
            _prefetch_reference(options, read_reference_names(options.bam_file))
        aln_file1 = _open_sorted_bam(options.bam_file, None)
        if aln_file1 is None:
            return
        if aln_file1.records is not None:
            _prefetch_reference(options, aln_file1.references)
        options._haplotype = 0
        sv_candidates = analyze_alignment_file_coordsorted(aln_file1, options)
    else:
        logging.info("MODE: diploid")
        logging.info("INPUT1: {0}".format(os.path.abspath(options.bam_file1)))
        logging.info("INPUT2: {0}".format(os.path.abspath(options.bam_file2)))
        if _prefetching():                 # the FASTA starts moving now, next to the first ingest
            from .bamio import read_reference_names
            _prefetch_reference(options, read_reference_names(options.bam_file1))
        aln_file1 = _open_sorted_bam(options.bam_file1, 1)
        if aln_file1 is None:
            return
        if aln_file1.records is not None:
            _prefetch_reference(options, aln_file1.references)
        options._haplotype = 1
        sv_candidates1 = analyze_alignment_file_coordsorted(aln_file1, options)
        aln_file2 = _open_sorted_bam(options.bam_file2, 2)
        if aln_file2 is None:
            return
        options._haplotype = 2
        sv_candidates2 = analyze_alignment_file_coordsorted(aln_file2, options)
    try:
        reference = FastaFile(options.genome)
    except ValueError:
        logging.error("The given reference genome is missing an index file ({0}.fai). Sequence alleles cannot be "
                      "retrieved.".format(options.genome))
        return
    except IOError:
        logging.error("The given reference genome is missing ({0}). Sequence alleles cannot be retrieved.".format(options.genome))
        return
    if options.sub == "haploid":
        final = sv_candidates
        logging.info("****************** STEP 2: OUTPUT ******************")
    else:
        logging.info("****************** STEP 2: PAIR ******************")
        final = pair_candidates(sv_candidates1, sv_candidates2, reference, aln_file1, options)
        logging.info("****************** STEP 3: OUTPUT ******************")
    split = getattr(final, "of_type", None)        # device-backed list: per-class views, no python objects built
    by_type = {t: (split(t) if split else [c for c in final if c.type == t]) for t in ("DEL", "INS", "INV", "DUP_TAN", "BND", "DUP_INT")}
    for label, key in (("deletion", "DEL"), ("inversion", "INV"), ("insertion", "INS"), ("tandem duplication", "DUP_TAN"),
                       ("interspersed duplication", "DUP_INT"), ("breakend", "BND")):
        logging.info("Found {0} {1} candidates.".format(len(by_type[key]), label))
    logging.info("Write SV candidates..")
    types_to_output = [entry.strip() for entry in options.types.split(",")]
    write_final_vcf(by_type["DUP_INT"], by_type["INV"], by_type["DUP_TAN"], by_type["DEL"], by_type["INS"], by_type["BND"],
                    __version__, aln_file1.references, aln_file1.lengths, types_to_output, reference, options)
    logging.info("Draw plots..")
    plot_sv_lengths(by_type["DEL"], by_type["INV"], by_type["DUP_INT"], by_type["DUP_TAN"], by_type["INS"], options)
    logging.info("Done.")


def entry():
    try:
        sys.exit(main())
    except Exception as exc:           # same top-level behaviour as svim-asm:182-185
        logging.error(exc, exc_info=True)
