"""TEST INFRASTRUCTURE -- run the UNMODIFIED reference (/root/reference/src/svim_asm) in this container.

The reference needs modules *named* pysam / edlib / matplotlib; `oracle/shims`
provides stand-ins (real packages win if they are importable).  This file is
only usable where /root/reference exists (the build container); it generates
the golden fixtures under tests/golden/ and validates the travelling CPU
restatement (`oracle/port.py`).  Nothing in the product imports it.
"""
import importlib
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = os.environ.get("SVIM_REFERENCE_SRC", "/root/reference/src")


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "svim_asm"))


def _prepare_path():
    shims = os.path.join(HERE, "shims")
    for mod in ("pysam", "edlib", "matplotlib"):
        try:
            if mod not in sys.modules:
                importlib.import_module(mod)
        except ImportError:
            if shims not in sys.path:
                sys.path.insert(0, shims)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


def modules():
    """dict of the reference's modules, imported in place."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_SRC)
    _prepare_path()
    names = ("SVCandidate", "SVIM_intra", "SVIM_inter", "SVIM_COLLECT", "SVIM_COMBINE", "SVIM_input_parsing")
    mods = {name: importlib.import_module("svim_asm." + name) for name in names}
    mods["pysam"] = importlib.import_module("pysam")
    return mods


def run_cli(argv):
    """Execute the reference's `svim-asm` script with `argv` (list of str, without the program name)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_SRC)
    _prepare_path()
    import logging
    script = os.path.join(REFERENCE_SRC, "svim_asm", "svim-asm")
    saved = sys.argv
    root = logging.getLogger()
    before = list(root.handlers)
    try:
        sys.argv = [script] + list(argv)
        # parse_arguments() binds sys.argv[1:] as a default argument when its module is imported
        # (SVIM_input_parsing.py:7), so a second CLI run in this process needs a fresh import
        sys.modules.pop("svim_asm.SVIM_input_parsing", None)
        try:
            runpy.run_path(script, run_name="__main__")
        except SystemExit:
            pass
    finally:
        sys.argv = saved
        for h in list(root.handlers):
            if h not in before:
                root.removeHandler(h)
                h.close()


def parse_options(argv):
    return modules()["SVIM_input_parsing"].parse_arguments("1.0.3", list(argv))


def canon(cand):
    """Flatten one reference Candidate object into a plain comparable tuple."""
    t = cand.type
    g = cand.genotype
    reads = tuple(cand.reads)
    if t == "DEL":
        return (t, cand.source_contig, cand.source_start, cand.source_end, g, reads)
    if t == "INV":
        return (t, cand.source_contig, cand.source_start, cand.source_end, bool(cand.complete), g, reads)
    if t == "INS":
        return (t, cand.dest_contig, cand.dest_start, cand.dest_end, cand.sequence, g, reads)
    if t == "DUP_TAN":
        return (t, cand.source_contig, cand.source_start, cand.source_end, int(cand.copies),
                bool(cand.fully_covered), g, reads)
    if t == "DUP_INT":
        return (t, cand.source_contig, cand.source_start, cand.source_end, cand.dest_contig, cand.dest_start,
                cand.dest_end, bool(cand.cutpaste), g, reads)
    if t == "BND":
        return (t, cand.source_contig, cand.source_start, cand.source_direction, cand.dest_contig,
                cand.dest_start, cand.dest_direction, g, reads)
    raise ValueError(t)


def collect(bam_path, options):
    mods = modules()
    bam = mods["pysam"].AlignmentFile(bam_path)
    return mods["SVIM_COLLECT"].analyze_alignment_file_coordsorted(bam, options), bam
