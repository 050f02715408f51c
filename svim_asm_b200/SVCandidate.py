"""Candidate object surface of SVIM-asm (reference src/svim_asm/SVCandidate.py:1-443), kept so that code
written against the reference's classes (attributes, get_key, get_vcf_entry*) keeps working.

In this package the objects are only a VIEW: candidates are computed on the GPU as 64-byte table rows
(include/svimasm_b200.h, svb_row) and materialised lazily by `candidates_from_rows` for the VCF writer or
for callers that want python objects.  The constructors repeat the reference's clamps and assertions so
that hand-built objects behave identically.
"""
import numpy as np

_PLACEHOLDER = "PLACEHOLDERFORID"


def _vcf_record(chrom, pos, ref, alt, filters, info, fmt, sample):
    """One VCF body line; the ID is filled in by write_final_vcf (SVIM_COMBINE.py:472-475)."""
    return "\t".join((chrom, str(pos), _PLACEHOLDER, ref, alt, ".", ";".join(filters) if filters else "PASS", info, fmt,
                      sample))


def _with_reads(info, reads, read_names):
    return info + ";READS=" + ",".join(reads) if read_names else info


class Candidate(object):
    """Common behaviour (SVCandidate.py:1-35)."""
    type = None

    def __init__(self, source_contig, source_start, source_end, genotype="1/1"):
        self.source_contig, self.source_start, self.source_end = source_contig, source_start, source_end
        self.genotype = genotype

    def get_source(self):
        return (self.source_contig, self.source_start, self.source_end)

    def get_key(self):
        contig, start, end = self.get_source()
        return (self.type, contig, (start + end) // 2)

    def position_distance_to(self, other):
        c1, s1, e1 = self.get_source()
        c2, s2, e2 = other.get_source()
        if self.type != other.type or c1 != c2:
            return float("inf")
        return min(abs(s1 - s2), abs(e1 - e2), abs((s1 + e1) // 2 - (s2 + e2) // 2))

    def get_vcf_entry(self):
        raise NotImplementedError

    def _set_span(self, prefix, contig, start, end, bam, what, reads):
        # assert + clamp to [0, contig length] (SVCandidate.py:40-46, 83-89, 130-136, 181-187, 266-279)
        assert end >= start, "{3} end ({0}:{1}) is smaller than its start ({0}:{2}). From read {4}".format(
            contig, end, start, what, reads)
        setattr(self, prefix + "_contig", contig)
        setattr(self, prefix + "_start", max(0, start))
        setattr(self, prefix + "_end", min(bam.get_reference_length(contig), end))


class CandidateDeletion(Candidate):
    type = "DEL"

    def __init__(self, source_contig, source_start, source_end, reads, bam, genotype="1/1"):
        self._set_span("source", source_contig, source_start, source_end, bam, "Deletion", reads)
        self.reads, self.genotype = reads, genotype

    def get_vcf_entry(self, sequence_alleles=False, reference=None, read_names=False):
        contig, start, end = self.get_source()
        if sequence_alleles:
            anchor = max(0, start - 1)
            ref_allele = reference.fetch(contig, anchor, end).upper()
            alt_allele = reference.fetch(contig, anchor, start).upper()
        else:
            ref_allele, alt_allele = "N", "<DEL>"
        info = _with_reads("SVTYPE=DEL;END=%d;SVLEN=%d" % (end, start - end), self.reads, read_names)
        return _vcf_record(contig, max(1, start), ref_allele, alt_allele, [], info, "GT", self.genotype)


_REVCOMP = str.maketrans("ACGT", "TGCA")


class CandidateInversion(Candidate):
    type = "INV"
    complement = {"A": "T", "C": "G", "G": "C", "T": "A"}

    def __init__(self, source_contig, source_start, source_end, reads, complete, bam, genotype="1/1"):
        self._set_span("source", source_contig, source_start, source_end, bam, "Inversion", reads)
        self.reads, self.complete, self.genotype = reads, complete, genotype

    def get_vcf_entry(self, sequence_alleles=False, reference=None, read_names=False):
        contig, start, end = self.get_source()
        if sequence_alleles:
            ref_allele = reference.fetch(contig, start, end).upper()
            alt_allele = ref_allele[::-1].translate(_REVCOMP)      # complement.get(b, b) per base (SVCandidate.py:106)
        else:
            ref_allele, alt_allele = "N", "<INV>"
        info = _with_reads("SVTYPE=INV;END=%d" % end, self.reads, read_names)
        filters = [] if self.complete else ["incomplete_inversion"]
        return _vcf_record(contig, start + 1, ref_allele, alt_allele, filters, info, "GT", self.genotype)


class CandidateInsertion(Candidate):
    type = "INS"

    def __init__(self, dest_contig, dest_start, dest_end, reads, sequence, bam, genotype="1/1"):
        self._set_span("dest", dest_contig, dest_start, dest_end, bam, "Insertion", reads)
        self.reads, self.sequence, self.genotype = reads, sequence, genotype

    def get_destination(self):
        return (self.dest_contig, self.dest_start, self.dest_end)

    def get_key(self):
        return (self.type, self.dest_contig, self.dest_start)

    def get_vcf_entry(self, sequence_alleles=False, reference=None, read_names=False):
        contig, start, end = self.get_destination()
        if sequence_alleles:
            ref_allele = reference.fetch(contig, max(0, start - 1), start).upper()
            alt_allele = ref_allele + self.sequence
        else:
            ref_allele, alt_allele = "N", "<INS>"
        info = _with_reads("SVTYPE=INS;END=%d;SVLEN=%d" % (start, end - start), self.reads, read_names)   # END is the start
        return _vcf_record(contig, max(1, start), ref_allele, alt_allele, [], info, "GT", self.genotype)


class CandidateDuplicationTandem(Candidate):
    type = "DUP_TAN"

    def __init__(self, source_contig, source_start, source_end, copies, fully_covered, reads, bam, genotype="1/1"):
        self._set_span("source", source_contig, source_start, source_end, bam, "Tandem duplication", reads)
        self.copies, self.fully_covered, self.reads, self.genotype = copies, fully_covered, reads, genotype

    def get_destination(self):
        contig, start, end = self.get_source()
        return (contig, end, end + self.copies * (end - start))

    def _filters(self):
        return [] if self.fully_covered else ["not_fully_covered"]

    def get_vcf_entry_as_ins(self, sequence_alleles=False, reference=None, read_names=False):
        contig, start, end = self.get_source()
        if sequence_alleles:
            ref_allele = reference.fetch(contig, start, end).upper()
            alt_allele = ref_allele * (self.copies + 1)
        else:
            ref_allele, alt_allele = "N", "<INS>"
        info = _with_reads("SVTYPE=INS;END=%d;SVLEN=%d" % (end, (end - start) * self.copies), self.reads, read_names)
        return _vcf_record(contig, start + 1, ref_allele, alt_allele, self._filters(), info, "GT", self.genotype)

    def get_vcf_entry_as_dup(self, read_names=False):
        contig, start, end = self.get_source()
        info = _with_reads("SVTYPE=DUP:TANDEM;END=%d;SVLEN=%d" % (end, end - start), self.reads, read_names)
        return _vcf_record(contig, start + 1, "N", "<DUP:TANDEM>", self._filters(), info, "GT:CN",
                           "%s:%d" % (self.genotype, self.copies + 1))


class CandidateDuplicationInterspersed(Candidate):
    type = "DUP_INT"

    def __init__(self, source_contig, source_start, source_end, dest_contig, dest_start, dest_end, reads, bam,
                 cutpaste=False, genotype="1/1"):
        self._set_span("source", source_contig, source_start, source_end, bam, "Interspersed duplication source", reads)
        self._set_span("dest", dest_contig, dest_start, dest_end, bam, "Interspersed duplication destination", reads)
        self.cutpaste, self.reads, self.genotype = cutpaste, reads, genotype

    def get_destination(self):
        return (self.dest_contig, self.dest_start, self.dest_end)

    def get_key(self):
        return (self.type, self.dest_contig, self.dest_start)

    def _info(self, svtype, end, length, read_names):
        return _with_reads("SVTYPE=%s;%sEND=%d;SVLEN=%d" % (svtype, "CUTPASTE;" if self.cutpaste else "", end, length),
                           self.reads, read_names)

    def get_vcf_entry_as_ins(self, sequence_alleles=False, reference=None, read_names=False):
        contig, start, end = self.get_destination()
        if sequence_alleles:
            ref_allele = reference.fetch(contig, max(0, start - 1), start).upper()
            alt_allele = ref_allele + reference.fetch(self.source_contig, self.source_start, self.source_end).upper()
        else:
            ref_allele, alt_allele = "N", "<INS>"
        return _vcf_record(contig, max(1, start), ref_allele, alt_allele, [], self._info("INS", start, end - start, read_names),
                           "GT", self.genotype)

    def get_vcf_entry_as_dup(self, read_names=False):
        contig, start, end = self.get_source()
        return _vcf_record(contig, start + 1, "N", "<DUP:INT>", [], self._info("DUP:INT", end, end - start, read_names), "GT",
                           self.genotype)


class CandidateBreakend(Candidate):
    type = "BND"
    # ALT of the forward record, by (source_direction, dest_direction)        (SVCandidate.py:392-399)
    _ALT = {("fwd", "fwd"): "N[%s:%d[", ("fwd", "rev"): "N]%s:%d]", ("rev", "rev"): "]%s:%d]N", ("rev", "fwd"): "[%s:%d[N"}
    # ALT of the mate record                                                   (SVCandidate.py:420-427)
    _ALT_MATE = {("rev", "rev"): "N[%s:%d[", ("fwd", "rev"): "N]%s:%d]", ("fwd", "fwd"): "]%s:%d]N", ("rev", "fwd"): "[%s:%d[N"}

    def __init__(self, source_contig, source_start, source_direction, dest_contig, dest_start, dest_direction, reads, bam,
                 genotype="1/1"):
        # (contig, position) of the source sorts first -- contigs compare as python strings (SVCandidate.py:352)
        if not (source_contig < dest_contig or (source_contig == dest_contig and source_start < dest_start)):
            flip = {"fwd": "rev", "rev": "fwd"}
            source_contig, source_start, source_direction, dest_contig, dest_start, dest_direction = (
                dest_contig, dest_start, flip[dest_direction], source_contig, source_start, flip[source_direction])
        self.source_contig, self.source_direction = source_contig, source_direction
        self.source_start = min(bam.get_reference_length(source_contig), max(0, source_start))
        self.dest_contig, self.dest_direction = dest_contig, dest_direction
        self.dest_start = min(bam.get_reference_length(dest_contig), max(0, dest_start))
        self.reads, self.genotype = reads, genotype

    def get_source(self):
        return (self.source_contig, self.source_start)

    def get_destination(self):
        return (self.dest_contig, self.dest_start)

    def get_key(self):
        return (self.type, self.source_contig, self.source_start)

    def _entry(self, here, there, table, read_names):
        alt = table[(self.source_direction, self.dest_direction)] % (there[0], there[1] + 1)
        return _vcf_record(here[0], here[1] + 1, "N", alt, [], _with_reads("SVTYPE=BND", self.reads, read_names), "GT",
                           self.genotype)

    def get_vcf_entry(self, read_names=False):
        return self._entry(self.get_source(), self.get_destination(), self._ALT, read_names)

    def get_vcf_entry_reverse(self, read_names=False):
        return self._entry(self.get_destination(), self.get_source(), self._ALT_MATE, read_names)


# ---------------------------------------------------------------------------------------------------------
# device rows -> objects

TYPE_NAMES = ("DEL", "INV", "INS", "DUP_TAN", "DUP_INT", "BND")
GENOTYPES = ("1/1", "1/0", "0/1")
F_COMPLETE, F_FULLY_COVERED, F_CUTPASTE, F_SRC_FWD, F_DST_FWD = 1, 2, 4, 8, 16
NO_MATE = 0xFFFFFFFF


class _Lengths(object):
    """The only thing the constructors ask of `bam` (get_reference_length)."""

    def __init__(self, names, lengths):
        self._len = dict(zip(names, (int(x) for x in lengths)))

    def get_reference_length(self, name):
        return self._len[name]


def decode_pool(rows, pool, starts):
    """Inserted sequences of the INS rows of a table whose sequence pool was gathered on the device (4-bit packed, one
    byte-aligned run per row): dict (aln_idx, seq_pos, seq_len) -> str."""
    lut = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)
    out = {}
    for i in np.nonzero((rows["type"] == 2) & (rows["seq_len"] > 0))[0]:
        n = int(rows["seq_len"][i])
        lo = int(starts[i])
        raw = pool[lo:lo + (n + 1) // 2]
        nib = np.empty(raw.shape[0] * 2, dtype=np.uint8)
        nib[0::2] = raw >> 4
        nib[1::2] = raw & 15
        out[(int(rows["aln_idx"][i]), int(rows["seq_pos"][i]), n)] = lut[nib[:n]].tobytes().decode("ascii")
    return out


def candidates_from_rows(rows, hosts, contig_names, contig_lengths, sequences=None):
    """Materialise table rows (numpy structured array, svb_row layout) as Candidate objects.

    hosts: dict haplotype -> HostBatch (key 0 for a haploid run); they supply query names and the
    inserted sequences (query_sequence slices, SVIM_intra.py:42, SVIM_inter.py:117,120).
    sequences: optional dict haplotype -> decode_pool() result; rows found there do not touch the host's query bases
    (after a device ingest those would have to be downloaded first)."""
    bam = _Lengths(contig_names, contig_lengths)
    sequences = sequences or {}
    out = []
    for r in rows:
        hap = int(r["hap"])
        host = hosts[hap]
        reads = [host.query_name(int(r["aln_idx"]))]
        if int(r["mate_aln"]) != NO_MATE:
            reads.append(hosts[3 - hap].query_name(int(r["mate_aln"])))
        gt = GENOTYPES[int(r["genotype"])]
        kind = int(r["type"])
        flags = int(r["flags"])
        if kind == 0:
            c = CandidateDeletion(contig_names[r["src_tid"]], int(r["src_start"]), int(r["src_end"]), reads, bam, gt)
        elif kind == 1:
            c = CandidateInversion(contig_names[r["src_tid"]], int(r["src_start"]), int(r["src_end"]), reads,
                                   bool(flags & F_COMPLETE), bam, gt)
        elif kind == 2:
            key = (int(r["aln_idx"]), int(r["seq_pos"]), int(r["seq_len"]))
            seq = sequences.get(hap, {}).get(key) if key[2] > 0 else ""
            if seq is None:
                seq = host.sequence_slice(*key)
            c = CandidateInsertion(contig_names[r["dst_tid"]], int(r["dst_start"]), int(r["dst_end"]), reads, seq, bam, gt)
        elif kind == 3:
            c = CandidateDuplicationTandem(contig_names[r["src_tid"]], int(r["src_start"]), int(r["src_end"]), int(r["copies"]),
                                           bool(flags & F_FULLY_COVERED), reads, bam, gt)
        elif kind == 4:
            c = CandidateDuplicationInterspersed(contig_names[r["src_tid"]], int(r["src_start"]), int(r["src_end"]),
                                                 contig_names[r["dst_tid"]], int(r["dst_start"]), int(r["dst_end"]), reads, bam,
                                                 bool(flags & F_CUTPASTE), gt)
        else:
            # rows hold the already-normalised breakend; set the fields directly so that the constructor's
            # swap (which is not idempotent when both ends coincide) is not applied a second time
            c = CandidateBreakend.__new__(CandidateBreakend)
            c.source_contig, c.source_start = contig_names[r["src_tid"]], int(r["src_start"])
            c.source_direction = "fwd" if flags & F_SRC_FWD else "rev"
            c.dest_contig, c.dest_start = contig_names[r["dst_tid"]], int(r["dst_start"])
            c.dest_direction = "fwd" if flags & F_DST_FWD else "rev"
            c.reads, c.genotype = reads, gt
        out.append(c)
    return out
