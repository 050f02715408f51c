"""Seams of reference src/svim_asm/SVIM_COMBINE.py.

pair_candidates (SVIM_COMBINE.py:164-366) runs on the GPU (svb_pair: radix sort by key, partition split,
bit-parallel edit distances, complete-linkage clustering, genotype rules).  write_final_vcf / sorted_nicely
(:369-477) are host text formatting, byte-identical to the reference apart from the wall-clock
##fileDate line.
"""
import logging
import os
import re
import threading
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


_PREFETCH = {}        # (absolute FASTA path, contig names) -> ReferencePrefetch


class ReferencePrefetch(threading.Thread):
    """FASTA -> HBM on a context (stream, staging buffers) of its own while the main thread ingests the BAM files: the two
    only share the PCIe link, and the ingest is mostly the inflate kernel.  The CLI starts one as soon as the first BAM
    header has named the contigs; _device_reference picks the result up.  A failure is kept quiet here: the synchronous
    load that follows raises it where the reference would."""

    def __init__(self, fasta_path, contig_names):
        threading.Thread.__init__(self, daemon=True)
        self.key = (os.path.abspath(fasta_path), tuple(contig_names))
        self.path, self.names, self.result = fasta_path, list(contig_names), None
        self.device = get_engine().device            # (the main context is created here, on the caller's thread)
        _PREFETCH[self.key] = self
        self.start()

    def run(self):
        try:
            from .fasta import FastaFile
            fasta = FastaFile(self.path)
            rows = fasta.fai_rows(self.names)
            fasta.close()
            engine = get_engine(self.device, role="loader")
            self.result = engine.load_reference_fasta(self.path, rows)
            self.result.loader = engine                  # (the context lives as long as the process)
        except Exception:
            self.result = None


def prefetch_pending(fasta_path, contig_names):
    return (os.path.abspath(fasta_path), tuple(contig_names)) in _PREFETCH


def drop_prefetches():
    """Wait for background loads nobody picked up and release what they produced."""
    while _PREFETCH:
        _key, pending = _PREFETCH.popitem()
        pending.join()
        if pending.result is not None:
            pending.result.free()
            pending.result = None


def _device_reference(reference, contig_names):
    """Upload the FASTA once per (FastaFile, contig list): upper-cased bases in BAM header order."""
    cache = getattr(reference, "_svb_ref", None)
    key = tuple(contig_names)
    if cache is None or cache[0] != key:
        pending = _PREFETCH.pop((os.path.abspath(getattr(reference, "filename", "") or ""), key), None)
        if pending is not None:
            pending.join()
        if pending is not None and pending.result is not None:
            cache = (key, pending.result)
        elif hasattr(reference, "fai_rows"):       # our FastaFile: the bases go from the file to HBM without a host pass
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


def _onto_first_header(eng, side, names):
    """The reference pairs by contig NAME and takes every length from the first BAM (SVIM_COMBINE.py:164-366 pass `bam` =
    aln_file1 to the constructors).  The device rows carry tids of their own file: when the second BAM lists its contigs in
    another order (or lists other contigs), its rows are renumbered onto the first header before svb_pair sees them."""
    table, records, host = side
    other = list(getattr(host, "contig_names", names))
    if other == list(names):
        return side
    first = {n: t for t, n in enumerate(names)}
    remap = np.array([first.get(n, -1) for n in other] + [-1], dtype=np.int32)        # last slot: tid -1 stays -1
    rows = table.to_numpy()
    for field in ("src_tid", "dst_tid"):
        tid = rows[field]
        used = tid >= 0
        bad = used & (remap[np.where(used, tid, -1)] < 0)
        if np.any(bad):
            raise KeyError(other[int(tid[np.nonzero(bad)[0][0]])])       # bam.get_reference_length(name) of the first BAM would raise
        rows[field] = np.where(used, remap[np.where(used, tid, -1)], tid)
    renumbered = eng.table_from_numpy(rows)
    if not records.has_sequences:
        eng.set_sequences(records)
    renumbered.gather_sequences(records)          # the inserted bases travel with the table: svb_pair no longer needs its image
    return renumbered, records, host


def _require_fasta_contigs(reference, names, tables):
    """reference.fetch(contig, ...) raises KeyError for a contig the FASTA does not hold (SVIM_COMBINE.py:48-99); the device
    copy would silently read it as empty.  Every contig that carries a candidate with sequence context must be there."""
    have = getattr(reference, "references", None)
    if have is None:
        return
    have = set(have)
    missing = [t for t, n in enumerate(names) if n not in have]
    if not missing:
        return
    for table in tables:
        rows = table.to_numpy()
        rows = rows[rows["type"] != 5]
        for field in ("src_tid", "dst_tid"):
            hit = np.isin(rows[field], missing)
            if np.any(hit):
                raise KeyError(names[int(rows[field][np.nonzero(hit)[0][0]])])


def pair_candidates(sv_candidates1, sv_candidates2, reference, bam, options):
    """SVIM_COMBINE.py:164-366 on the GPU.  Accepts the CandidateList objects of
    analyze_alignment_file_coordsorted (device tables are reused) or plain lists of Candidate objects."""
    eng = get_engine()
    names, lengths = list(bam.references), list(bam.lengths)
    counts = defaultdict(int)
    for cands in (sv_candidates1, sv_candidates2):
        if getattr(cands, "rows", None) is not None:           # device-backed list: count without building objects
            for kind, k in zip(TYPE_NAMES, np.bincount(cands.rows["type"], minlength=len(TYPE_NAMES))):
                counts[kind] += int(k)
        else:
            for c in cands:
                counts[c.type] += 1
    for label, kind in (("deletions", "DEL"), ("inversions", "INV"), ("insertions", "INS"),
                        ("tandem duplications", "DUP_TAN"), ("interspersed duplications", "DUP_INT"), ("breakends", "BND")):
        logging.info("Pairing {0} {1}...".format(counts[kind], label))
    sides = []
    for hap, cands in ((1, sv_candidates1), (2, sv_candidates2)):
        if (getattr(cands, "table", None) is not None and getattr(cands, "records", None) is not None
                and cands.rows.shape[0] == len(cands.table)):          # the untouched result of a collect: reuse its device table
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
    sides[1] = _onto_first_header(eng, sides[1], names)
    _require_fasta_contigs(reference, names, [side[0] for side in sides])
    ref = _device_reference(reference, names)
    paired = eng.pair(sides[0][0], sides[1][0], sides[0][1], sides[1][1], ref, make_params(options))
    def known():
        found = {hap: getattr(c, "sequences", None) for hap, c in ((1, sv_candidates1), (2, sv_candidates2))}
        return {h: d for h, d in found.items() if d is not None}
    out = CandidateList.from_rows(paired.to_numpy(), {1: sides[0][2], 2: sides[1][2]}, names, lengths, known, paired,
                                  {1: sides[0][1], 2: sides[1][1]})
    return out


def _haplotype_strings(c1, c2, reference):
    """The two haplotype strings compute_distance aligns (SVIM_COMBINE.py:43-100), as bytes."""
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
    return build(c1, lo, hi).encode("latin-1"), build(c2, lo, hi).encode("latin-1")


def compute_distance(candidate_with_haplotype1, candidate_with_haplotype2, reference):
    """SVIM_COMBINE.py:35-102 for one pair: the edit distance comes from the GPU kernel (svb_edit_distance)."""
    (hap1, c1), (hap2, c2) = candidate_with_haplotype1, candidate_with_haplotype2
    if hap1 == hap2:
        return 1000000000
    return int(get_engine().edit_distance([_haplotype_strings(c1, c2, reference)])[0])


def form_partitions(sv_candidates_with_haplotype, max_distance):
    """SVIM_COMBINE.py:15-32: stable sort of (haplotype, candidate) items by candidate.get_key(), split where type or contig
    change or consecutive key positions are more than max_distance apart.  The sort and the split run on the device
    (svb_form_partitions: LSD radix sort + head flags); the keys are ranked here the way python orders the tuples."""
    items = list(sv_candidates_with_haplotype)
    if not items:
        return []
    keys = [c.get_key() for _hap, c in items]
    groups = {g: k for k, g in enumerate(sorted(set((key[0], key[1]) for key in keys)))}
    packed = np.zeros(len(items), dtype=np.uint64)
    for i, key in enumerate(keys):
        pos = int(key[2])
        if not 0 <= pos < (1 << 31):
            raise ValueError("form_partitions: key position %d outside [0, 2^31)" % pos)
        packed[i] = (groups[(key[0], key[1])] << 32) | pos
    order, part_start = get_engine().form_partitions(packed, max_distance)
    return [[items[int(k)] for k in order[part_start[p]:part_start[p + 1]]] for p in range(part_start.shape[0] - 1)]


def span_position_distance_breakends(candidate1, candidate2):
    """SVIM_COMBINE.py:105-117 on (haplotype, pos1, dir1, pos2, dir2) rows (the destination contig is never compared)."""
    hap1, pos1a, dir1a, pos2a, dir2a = candidate1
    hap2, pos1b, dir1b, pos2b, dir2b = candidate2
    if hap1 != hap2 and dir1a == dir1b and dir2a == dir2b:
        return (abs(pos1a - pos1b) + abs(pos2a - pos2b)) / 3000
    return 99999


def _clusters_by_label(partitions, condensed_of, threshold):
    """pair_haplotypes' frame (SVIM_COMBINE.py:120-140 / :143-161): singletons pass, partitions above 10 are dropped,
    the rest is clustered by complete linkage cut at `threshold` (svb_cluster_labels reproduces scipy's labels, hence the
    order of the clusters)."""
    todo = [k for k, part in enumerate(partitions) if 2 <= len(part) <= 10]
    labels = dict(zip(todo, get_engine().cluster_labels([condensed_of(partitions[k]) for k in todo], threshold))) if todo else {}
    out = []
    for k, part in enumerate(partitions):
        if len(part) < 2:
            out.append(part)
        elif len(part) > 10:
            if part and hasattr(part[0][1], "get_key"):
                logging.debug("Ignored partition of size {0} and type {1}: {2}".format(
                    len(part), part[0][1].get_key()[0], ",".join("{0}:{1}".format(c.get_key()[1], c.get_key()[2]) for _h, c in part)))
            continue
        else:
            lab = labels[k]
            clusters = [[] for _ in range(max(lab))]
            for item, which in zip(part, lab):
                clusters[which - 1].append(item)
            out.extend(clusters)
    return out


def pair_haplotypes(partitions, reference, edit_distance_threshold=10):
    """SVIM_COMBINE.py:120-140.  Every cross-haplotype edit distance of every partition goes to the GPU in ONE batch
    (svb_edit_distance), then every partition's linkage in one (svb_cluster_labels)."""
    partitions = list(partitions)
    jobs, slots = [], {}
    for k, part in enumerate(partitions):
        if 2 <= len(part) <= 10:
            for i in range(len(part) - 1):
                for j in range(i + 1, len(part)):
                    if part[i][0] != part[j][0]:
                        slots[(k, i, j)] = len(jobs)
                        jobs.append(_haplotype_strings(part[i][1], part[j][1], reference))
    dist = get_engine().edit_distance(jobs) if jobs else []
    index = {id(part): k for k, part in enumerate(partitions)}

    def condensed(part):
        k = index[id(part)]
        return [1000000000 if part[i][0] == part[j][0] else int(dist[slots[(k, i, j)]])
                for i in range(len(part) - 1) for j in range(i + 1, len(part))]
    return _clusters_by_label(partitions, condensed, edit_distance_threshold)


def pair_haplotypes_breakends(partitions, span_position_distance_threshold=0.3):
    """SVIM_COMBINE.py:143-161: clusters from the span-position distance of breakends."""
    def condensed(part):
        data = [(hap, c.get_source()[1], 1 if c.source_direction == "fwd" else 0, c.get_destination()[1],
                 1 if c.dest_direction == "fwd" else 0) for hap, c in part]
        return [float(span_position_distance_breakends(data[i], data[j])) for i in range(len(data) - 1) for j in range(i + 1, len(data))]
    return _clusters_by_label(list(partitions), condensed, span_position_distance_threshold)


def sorted_nicely(vcf_entries):
    """SVIM_COMBINE.py:369-376: natural ("human") order of contig names, then start, then end; stable."""
    return sorted(vcf_entries, key=lambda entry: (_natural_key(entry[0][0]), entry[0][1], entry[0][2]))


def _natural_key(name):
    return [int(tok) if tok.isdigit() else tok for tok in re.split("([0-9]+)", str(name))]


def natural_ranks(contig_names):
    """Dense rank of every contig under sorted_nicely's natural order (names with equal keys share a rank)."""
    keys = [_natural_key(n) for n in contig_names]
    rank = np.zeros(len(keys), dtype=np.int64)
    r, prev = -1, None
    for i in sorted(range(len(keys)), key=lambda k: keys[k]):
        if keys[i] != prev:
            r, prev = r + 1, keys[i]
        rank[i] = r
    return rank


VCF_DEL, VCF_INV, VCF_INS, VCF_TAN_AS_INS, VCF_TAN_AS_DUP, VCF_INT_AS_INS, VCF_INT_AS_DUP, VCF_BND, VCF_BND_MATE = range(9)
_LABEL_OF_MODE = np.array([0, 1, 2, 2, 3, 2, 4, 5, 5])        # DEL, INV, INS, DUP_TANDEM, DUP_INT, BND share a counter each


def vcf_entries(rows, row_index, contig_names, types_to_output, tan_as_ins, int_as_ins):
    """write_final_vcf's record list (SVIM_COMBINE.py:428-475) for table rows, without objects: which rows are written,
    through which get_vcf_entry* (mode), in sorted_nicely's order (natural contig order, start, end; stable over the
    append order of :431-464), with the running number of every ID label.  Returns svb_vcf_entry rows."""
    kind = rows["type"]
    rank = natural_ranks(contig_names)
    blocks = []                                         # (indices into rows, mode, tid, key start, key end) in append order

    def add(sel, mode, tid, start, end):
        blocks.append((sel, np.full(sel.shape[0], mode, dtype=np.uint32), tid[sel], start[sel], end[sel]))
    s, e, ds, de = (rows[n].astype(np.int64) for n in ("src_start", "src_end", "dst_start", "dst_end"))
    st, dt = rows["src_tid"].astype(np.int64), rows["dst_tid"].astype(np.int64)
    if "DEL" in types_to_output:
        add(np.nonzero(kind == 0)[0], VCF_DEL, st, np.maximum(1, s), e)
    if "INV" in types_to_output:
        add(np.nonzero(kind == 1)[0], VCF_INV, st, s + 1, e)
    if "INS" in types_to_output:
        add(np.nonzero(kind == 2)[0], VCF_INS, dt, np.maximum(1, ds), de)
    if tan_as_ins:
        if "INS" in types_to_output:
            add(np.nonzero(kind == 3)[0], VCF_TAN_AS_INS, st, s + 1, e)
    elif "DUP:TANDEM" in types_to_output:
        add(np.nonzero(kind == 3)[0], VCF_TAN_AS_DUP, st, s + 1, e)
    if int_as_ins:
        if "INS" in types_to_output:
            add(np.nonzero(kind == 4)[0], VCF_INT_AS_INS, dt, np.maximum(1, ds), de)
    elif "DUP:INT" in types_to_output:
        add(np.nonzero(kind == 4)[0], VCF_INT_AS_DUP, st, s + 1, e)
    if "BND" in types_to_output:
        sel = np.repeat(np.nonzero(kind == 5)[0], 2)                     # the record and its mate, adjacent
        mate = np.arange(sel.shape[0]) % 2 == 1
        blocks.append((sel, np.where(mate, VCF_BND_MATE, VCF_BND).astype(np.uint32), np.where(mate, dt[sel], st[sel]),
                       np.where(mate, ds[sel], s[sel]) + 1, np.where(mate, ds[sel], s[sel]) + 2))
    out = np.zeros(sum(b[0].shape[0] for b in blocks), dtype=_lib.VCF_ENTRY_DTYPE)
    if out.shape[0] == 0:
        return out
    sel, mode, tid, k1, k2 = (np.concatenate([b[k] for b in blocks]) for k in range(5))
    order = np.lexsort((k2, k1, rank[tid]))                             # stable: ties keep the append order
    sel, mode = sel[order], mode[order]
    out["row"], out["mode"] = row_index[sel], mode
    label = _LABEL_OF_MODE[mode]
    for k in range(6):
        hit = np.nonzero(label == k)[0]
        out["id"][hit] = np.arange(1, hit.shape[0] + 1)
    return out


def _device_lists(lists):
    """The six per-class lists are untouched views of ONE device table (CandidateList.of_type): (table, rows, row_index,
    records_by_hap), else None."""
    table, parts = None, []
    for kind, cands in lists:
        if getattr(cands, "table", None) is None or getattr(cands, "rows", None) is None:
            if len(cands) == 0:
                continue
            return None
        if table is not None and cands.table is not table:
            return None
        if cands.rows.shape[0] and not np.all(cands.rows["type"] == TYPE_NAMES.index(kind)):
            return None
        table = cands.table
        parts.append(cands)
    if table is None:
        return None
    rows = np.concatenate([c.rows for c in parts])
    index = np.concatenate([c.row_index for c in parts]).astype(np.uint32)
    return table, rows, index, parts[0].records_by_hap


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

    device = None if options.query_names or os.environ.get("SVIM_ASM_B200_VCF", "device") == "host" else _device_lists(
        (("DUP_INT", int_duplication_candidates), ("INV", inversion_candidates), ("DUP_TAN", tandem_duplication_candidates),
         ("DEL", deletion_candidates), ("INS", insertion_candidates), ("BND", breakend_candidates)))
    if device is not None:
        # the candidates never left the GPU as objects: the record lines are assembled there too (svb_vcf_body), the REF / ALT
        # bases gathered from the resident reference and query sequences
        table, rows, row_index, records_by_hap = device
        eng = table.engine
        entries = vcf_entries(rows, row_index, contig_names, types_to_output, tan_as_ins, int_as_ins)
        symbolic = bool(options.symbolic_alleles)
        records_by_hap = dict(records_by_hap or {})
        if not symbolic:
            for hap in np.unique(rows["hap"][(rows["type"] == 2) & (rows["seq_len"] > 0)]):
                records = records_by_hap.get(int(hap))
                if records is not None and not records.has_sequences:
                    eng.set_sequences(records)
        body = eng.vcf_body(table, records_by_hap, _device_reference(reference, list(contig_names)), contig_names, entries, symbolic)
        if not symbolic:
            reference.close()
        with open(options.working_dir + "/variants.vcf", "wb") as out:
            out.write(("\n".join(lines) + "\n").encode())
            out.write(body)
        return

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
