"""Seams of reference src/svim_asm/SVIM_COMBINE.py.

pair_candidates (SVIM_COMBINE.py:164-366) runs on the GPU (svb_pair: radix sort by key, partition split,
bit-parallel edit distances, complete-linkage clustering, genotype rules).  write_final_vcf / sorted_nicely
(:369-477) are host text formatting, byte-identical to the reference apart from the wall-clock
##fileDate line.
"""
import logging
import re
import time
from collections import defaultdict

import numpy as np

from . import _lib
from . import synth
from .engine import HostBatch, make_params
from .runtime import get_engine
from .SVCandidate import (F_COMPLETE, F_CUTPASTE, F_DST_FWD, F_FULLY_COVERED, F_SRC_FWD, GENOTYPES, NO_MATE, TYPE_NAMES,
                          candidates_from_rows)
from .SVIM_COLLECT import CandidateList


def _device_reference(reference, contig_names):
    """Upload the FASTA once per (FastaFile, contig list): upper-cased bases in BAM header order."""
    cache = getattr(reference, "_svb_ref", None)
    key = tuple(contig_names)
    if cache is None or cache[0] != key:
        if hasattr(reference, "fai_rows"):         # our FastaFile: the bases go from the file to HBM without a host pass
            cache = (key, get_engine().load_reference_fasta(reference.filename, reference.fai_rows(contig_names)))
        else:
            bases, offsets = reference.load_upper(contig_names)
            cache = (key, get_engine().load_reference(bases, offsets))
        reference._svb_ref = cache
    return cache[1]


def _rows_from_objects(cands, hap, contig_names):
    """Candidate objects -> (rows, HostBatch holding the inserted sequences and read names)."""
    tid = {n: i for i, n in enumerate(contig_names)}
    rows = np.zeros(len(cands), dtype=_lib.ROW_DTYPE)
    seqs = []
    for i, c in enumerate(cands):
        r = rows[i]
        r["type"] = TYPE_NAMES.index(c.type)
        r["hap"], r["aln_idx"], r["mate_aln"], r["ordinal"] = hap, i, NO_MATE, i
        r["src_tid"] = r["dst_tid"] = -1
        r["genotype"] = GENOTYPES.index(c.genotype)
        if c.type in ("DEL", "INV", "DUP_TAN", "DUP_INT"):
            r["src_tid"], r["src_start"], r["src_end"] = tid[c.source_contig], c.source_start, c.source_end
        if c.type in ("INS", "DUP_INT"):
            r["dst_tid"], r["dst_start"], r["dst_end"] = tid[c.dest_contig], c.dest_start, c.dest_end
        if c.type == "INV":
            r["flags"] = F_COMPLETE if c.complete else 0
        elif c.type == "DUP_TAN":
            r["flags"], r["copies"] = (F_FULLY_COVERED if c.fully_covered else 0), c.copies
        elif c.type == "DUP_INT":
            r["flags"] = F_CUTPASTE if c.cutpaste else 0
        elif c.type == "BND":
            r["src_tid"], r["src_start"], r["dst_tid"], r["dst_start"] = (tid[c.source_contig], c.source_start,
                                                                          tid[c.dest_contig], c.dest_start)
            r["flags"] = (F_SRC_FWD if c.source_direction == "fwd" else 0) | (F_DST_FWD if c.dest_direction == "fwd" else 0)
        seq = c.sequence if c.type == "INS" else ""
        r["seq_len"] = len(seq)
        seqs.append(seq)
    # one pseudo record per candidate: its "query sequence" is the inserted sequence
    lut = {ch: k for k, ch in enumerate(synth.NT16)}
    l_seq = np.array([len(s) for s in seqs], dtype=np.uint32)
    seq_off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    seq_off[1:] = np.cumsum((l_seq.astype(np.int64) + 1) // 2)
    seq4 = np.zeros(int(seq_off[-1]), dtype=np.uint8)
    for i, s in enumerate(seqs):
        if s:
            codes = np.array([lut.get(ch, 15) for ch in s] + ([0] if len(s) % 2 else []), dtype=np.uint8)
            seq4[int(seq_off[i]):int(seq_off[i + 1])] = (codes[0::2] << 4) | codes[1::2]
    n = len(cands)
    rb = synth.RecordBatch(list(contig_names), None, np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint16),
                           np.zeros(n, np.uint8), np.zeros(n, np.uint32), np.zeros(n + 1, np.uint64), l_seq, seq_off,
                           np.zeros(0, np.uint32), seq4, [",".join(c.reads) for c in cands], {})
    return rows, rb


def pair_candidates(sv_candidates1, sv_candidates2, reference, bam, options):
    """SVIM_COMBINE.py:164-366 on the GPU.  Accepts the CandidateList objects of
    analyze_alignment_file_coordsorted (device tables are reused) or plain lists of Candidate objects."""
    eng = get_engine()
    names, lengths = list(bam.references), list(bam.lengths)
    counts = defaultdict(int)
    for c in list(sv_candidates1) + list(sv_candidates2):
        counts[c.type] += 1
    for label, kind in (("deletions", "DEL"), ("inversions", "INV"), ("insertions", "INS"),
                        ("tandem duplications", "DUP_TAN"), ("interspersed duplications", "DUP_INT"), ("breakends", "BND")):
        logging.info("Pairing {0} {1}...".format(counts[kind], label))
    sides = []
    for hap, cands in ((1, sv_candidates1), (2, sv_candidates2)):
        if getattr(cands, "table", None) is not None:
            table, records, host = cands.table, cands.records, cands.host
            if not records.has_sequences:
                eng.set_sequences(records)
        else:
            rows, rb = _rows_from_objects(cands, hap, names)
            rb.contig_lengths = np.asarray(lengths, dtype=np.int32)
            host = HostBatch.from_record_batch(rb)
            records = eng.load_records(host, with_sequences=True)
            table = eng.table_from_numpy(rows)
        sides.append((table, records, host))
    ref = _device_reference(reference, names)
    paired = eng.pair(sides[0][0], sides[1][0], sides[0][1], sides[1][1], ref, make_params(options))
    known = {hap: getattr(c, "sequences", None) for hap, c in ((1, sv_candidates1), (2, sv_candidates2))}
    out = CandidateList(candidates_from_rows(paired.to_numpy(), {1: sides[0][2], 2: sides[1][2]}, names, lengths,
                                             sequences={h: d for h, d in known.items() if d is not None}))
    out.table = paired
    return out


def compute_distance(candidate_with_haplotype1, candidate_with_haplotype2, reference):
    """SVIM_COMBINE.py:35-102 for one pair: the edit distance comes from the GPU kernel (svb_edit_distance)."""
    (hap1, c1), (hap2, c2) = candidate_with_haplotype1, candidate_with_haplotype2
    if hap1 == hap2:
        return 1000000000
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}

    def fetch(contig, s, e):
        return reference.fetch(contig, s, e).upper()

    def build(c, lo, hi):
        if c.type in ("DEL", "INV", "DUP_TAN"):
            left, right = fetch(c.source_contig, lo, c.source_start), fetch(c.source_contig, c.source_end, hi)
            inner = fetch(c.source_contig, c.source_start, c.source_end)
            mid = "" if c.type == "DEL" else ("".join(comp.get(b, b) for b in reversed(inner)) if c.type == "INV"
                                               else inner * (c.copies + 1))
            return left + mid + right
        mid = c.sequence if c.type == "INS" else fetch(c.source_contig, c.source_start, c.source_end)
        return fetch(c.dest_contig, lo, c.dest_start) + mid + fetch(c.dest_contig, c.dest_start, hi)

    if c1.type in ("DEL", "INV", "DUP_TAN"):
        length = reference.get_reference_length(c1.source_contig)
        lo = max(0, min(c1.source_start, c2.source_start) - 100)
        hi = min(length, max(c1.source_end, c2.source_end) + 100)
    else:
        length = reference.get_reference_length(c1.dest_contig)
        lo = max(0, min(c1.dest_start, c2.dest_start) - 100)
        hi = min(length, max(c1.dest_start, c2.dest_start) + 100)
    a, b = build(c1, lo, hi).encode("latin-1"), build(c2, lo, hi).encode("latin-1")
    return int(get_engine().edit_distance([(a, b)])[0])


def sorted_nicely(vcf_entries):
    """SVIM_COMBINE.py:369-376: natural ("human") order of contig names, then start, then end; stable."""
    def natural(name):
        return [int(tok) if tok.isdigit() else tok for tok in re.split("([0-9]+)", str(name))]
    return sorted(vcf_entries, key=lambda entry: (natural(entry[0][0]), entry[0][1], entry[0][2]))


_HEADER_ALT = (("DEL", "Deletion"), ("INV", "Inversion"))


def write_final_vcf(int_duplication_candidates, inversion_candidates, tandem_duplication_candidates, deletion_candidates,
                    insertion_candidates, breakend_candidates, version, contig_names, contig_lengths, types_to_output,
                    reference, options):
    """SVIM_COMBINE.py:379-477, same 12 positional arguments."""
    tan_as_ins = options.tandem_duplications_as_insertions
    int_as_ins = options.interspersed_duplications_as_insertions
    want_tan = (not tan_as_ins) and "DUP:TANDEM" in types_to_output
    want_int = (not int_as_ins) and "DUP:INT" in types_to_output
    lines = ["##fileformat=VCFv4.2",
             "##fileDate={0}".format(time.strftime("%Y-%m-%d|%I:%M:%S%p|%Z|%z")),
             "##source=SVIM-asm-v{0}".format(version)]
    lines += ["##contig=<ID={0},length={1}>".format(n, l) for n, l in zip(contig_names, contig_lengths)]
    alts = [("DEL", "Deletion", "DEL" in types_to_output), ("INV", "Inversion", "INV" in types_to_output),
            ("DUP", "Duplication", want_tan or want_int), ("DUP:TANDEM", "Tandem Duplication", want_tan),
            ("DUP:INT", "Interspersed Duplication", want_int), ("INS", "Insertion", "INS" in types_to_output),
            ("BND", "Breakend", "BND" in types_to_output)]
    lines += ['##ALT=<ID={0},Description="{1}">'.format(i, d) for i, d, on in alts if on]
    lines += ['##INFO=<ID=SVTYPE,Number=1,Type=String,Description="Type of structural variant">',
              '##INFO=<ID=CUTPASTE,Number=0,Type=Flag,Description="Genomic origin of interspersed duplication seems to be deleted">',
              '##INFO=<ID=END,Number=1,Type=Integer,Description="End position of the variant described in this record">',
              '##INFO=<ID=SVLEN,Number=1,Type=Integer,Description="Difference in length between REF and ALT alleles">']
    if options.query_names:
        lines.append('##INFO=<ID=READS,Number=.,Type=String,Description="Names of all supporting reads">')
    lines += ['##FILTER=<ID=not_fully_covered,Description="Tandem duplication is not fully covered by a contig">',
              '##FILTER=<ID=incomplete_inversion,Description="Only one inversion breakpoint is supported">',
              '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">']
    if want_tan:
        lines.append('##FORMAT=<ID=CN,Number=1,Type=Integer,Description="Copy number of tandem duplication (e.g. 2 for one additional copy)">')
    lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + options.sample)

    alleles = not options.symbolic_alleles
    names = options.query_names
    entries = []                       # ((contig, start, end), line, id label) in the reference's append order (:431-464)
    if "DEL" in types_to_output:
        for c in deletion_candidates:
            contig, start, end = c.get_source()
            entries.append(((contig, max(1, start), end), c.get_vcf_entry(alleles, reference, names), "DEL"))
    if "INV" in types_to_output:
        for c in inversion_candidates:
            contig, start, end = c.get_source()
            entries.append(((contig, start + 1, end), c.get_vcf_entry(alleles, reference, names), "INV"))
    if "INS" in types_to_output:
        for c in insertion_candidates:
            contig, start, end = c.get_destination()
            entries.append(((contig, max(1, start), end), c.get_vcf_entry(alleles, reference, names), "INS"))
    if tan_as_ins:
        if "INS" in types_to_output:
            for c in tandem_duplication_candidates:
                entries.append(((c.source_contig, c.source_start + 1, c.source_end),
                                c.get_vcf_entry_as_ins(alleles, reference, names), "INS"))
    elif "DUP:TANDEM" in types_to_output:
        for c in tandem_duplication_candidates:
            entries.append(((c.source_contig, c.source_start + 1, c.source_end), c.get_vcf_entry_as_dup(names), "DUP_TANDEM"))
    if int_as_ins:
        if "INS" in types_to_output:
            for c in int_duplication_candidates:
                contig, start, end = c.get_destination()
                entries.append(((contig, max(1, start), end), c.get_vcf_entry_as_ins(alleles, reference, names), "INS"))
    elif "DUP:INT" in types_to_output:
        for c in int_duplication_candidates:
            contig, start, end = c.get_source()
            entries.append(((contig, start + 1, end), c.get_vcf_entry_as_dup(names), "DUP_INT"))
    if "BND" in types_to_output:
        for c in breakend_candidates:
            (sc, sp), (dc, dp) = c.get_source(), c.get_destination()
            entries.append(((sc, sp + 1, sp + 2), c.get_vcf_entry(names), "BND"))
            entries.append(((dc, dp + 1, dp + 2), c.get_vcf_entry_reverse(names), "BND"))
    if alleles:
        reference.close()

    numbering = defaultdict(int)
    for _key, line, label in sorted_nicely(entries):
        numbering[label] += 1
        lines.append(line.replace("PLACEHOLDERFORID", "svim_asm.{0}.{1}".format(label, numbering[label]), 1))
    with open(options.working_dir + "/variants.vcf", "w") as out:
        out.write("\n".join(lines) + "\n")
