"""TEST INFRASTRUCTURE -- a pure-Python stand-in for the `pysam` module.

pysam/htslib is not installed in this image and there is no network.  The
reference (`/root/reference/src/svim_asm/*.py`) imports `pysam` by name
(SVIM_COLLECT.py:2, svim-asm:12,15), so the oracle puts *this* directory on
`sys.path` and lets the UNMODIFIED reference run on top of it.  Only the
surface listed in SURVEY.md App. C is provided; semantics follow the SAM/BAM
specification v1 and htslib behaviour as documented there [ext].

Nothing in the product (`svim_asm_b200/`) may import this file.
"""
import gzip
import os
import re
import struct

__version__ = "0.0-shim"

_CIGAR_LETTERS = "MIDNSHP=XB"
_NT16 = "=ACMGRSVTWYHKDBN"
_NT16_PAIRS = [_NT16[b >> 4] + _NT16[b & 15] for b in range(256)]
# op classes (SAM spec table): consumes query / consumes reference
_Q_OPS = (0, 1, 4, 7, 8)
_R_OPS = (0, 2, 3, 7, 8)


_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=XB])")


def _parse_cigar_string(text):
    # pysam: CIGAR_REGEX.findall(text) -- anything that is not "<digits><op>" is silently ignored;
    # a length that does not fit 28 bits overflows the packed uint32 -> OverflowError
    ops = []
    for num, letter in _CIGAR_RE.findall(text):
        n = int(num)
        if n >= (1 << 28):
            raise OverflowError("value too large to convert to uint32_t")
        ops.append((_CIGAR_LETTERS.index(letter), n))
    return ops


class AlignedSegment(object):
    """One alignment record (the attribute set of SURVEY.md App. C)."""

    def __init__(self, header=None):
        self.query_name = None
        self.flag = 0
        self.reference_id = -1
        self.reference_start = -1
        self._mapq = 0
        self._cigar = []
        self._seq = None          # decoded str or None
        self._seq4 = None         # raw packed bytes (lazy decode)
        self._l_seq = 0
        self.next_reference_id = -1
        self.next_reference_start = -1
        self.template_length = 0
        self.query_qualities = None
        self._tags = {}

    # ---- flags
    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    @property
    def is_reverse(self):
        return bool(self.flag & 0x10)

    @property
    def is_secondary(self):
        return bool(self.flag & 0x100)

    @property
    def is_supplementary(self):
        return bool(self.flag & 0x800)

    # ---- mapq (uint8 in htslib -> OverflowError outside 0..255)
    @property
    def mapping_quality(self):
        return self._mapq

    @mapping_quality.setter
    def mapping_quality(self, value):
        if not 0 <= value <= 255:
            raise OverflowError("value too large to convert to uint8_t")
        self._mapq = value

    # ---- cigar
    @property
    def cigartuples(self):
        return list(self._cigar) if self._cigar else None

    @property
    def cigarstring(self):
        if not self._cigar:
            return None
        return "".join("%d%s" % (n, _CIGAR_LETTERS[op]) for op, n in self._cigar)

    @cigarstring.setter
    def cigarstring(self, text):
        self._cigar = _parse_cigar_string(text) if text and text != "*" else []

    def get_cigar_stats(self):
        base_counts = [0] * 11
        block_counts = [0] * 11
        for op, n in self._cigar:
            base_counts[op] += n
            block_counts[op] += 1
        nm = self._tags.get("NM")
        if isinstance(nm, int):
            base_counts[10] = nm
        return base_counts, block_counts

    # ---- sequence
    @property
    def query_sequence(self):
        if self._seq is None and self._seq4 is not None and self._l_seq > 0:
            self._seq = "".join(map(_NT16_PAIRS.__getitem__, self._seq4))[: self._l_seq]
        return self._seq

    @query_sequence.setter
    def query_sequence(self, value):
        if not value:
            self._seq = None
            self._seq4 = None
            self._l_seq = 0
        else:
            self._seq = value
            self._l_seq = len(value)

    @property
    def query_length(self):
        return self._l_seq

    # ---- CIGAR derived coordinates (htslib bam_endpos / pysam getQueryStart/End)
    @property
    def reference_end(self):
        if self.is_unmapped or not self._cigar:
            return None
        span = sum(n for op, n in self._cigar if op in _R_OPS)
        return self.reference_start + (span if span > 0 else 1)

    def infer_read_length(self):
        if not self._cigar:
            return None
        return sum(n for op, n in self._cigar if op in (0, 1, 4, 7, 8, 5))

    @property
    def query_alignment_start(self):
        offset = 0
        for op, n in self._cigar:
            if op == 5:
                continue
            if op == 4:
                offset += n
            else:
                break
        return offset

    @property
    def query_alignment_end(self):
        end = self._l_seq
        if end == 0:
            for op, n in self._cigar:
                if op in (0, 1, 7, 8) or (op == 4 and end == 0):
                    end += n
            return end
        for op, n in reversed(self._cigar[1:]):
            if op == 5:
                continue
            if op == 4:
                end -= n
            else:
                break
        return end

    # ---- tags
    def get_tag(self, name):
        return self._tags[name]          # KeyError if absent, like pysam

    def has_tag(self, name):
        return name in self._tags

    def set_tags(self, tags):
        self._tags = {}
        for entry in tags:
            self._tags[entry[0]] = entry[1]

    def set_tag(self, name, value, value_type=None):
        self._tags[name] = value


def _parse_tags(buf, pos, end):
    """BAM auxiliary fields -> dict (only the value kinds the spec defines)."""
    tags = {}
    scalar = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}
    while pos + 3 <= end:
        name = buf[pos:pos + 2].decode("ascii")
        kind = chr(buf[pos + 2])
        pos += 3
        if kind in scalar:
            fmt = scalar[kind]
            size = struct.calcsize(fmt)
            tags[name] = struct.unpack_from(fmt, buf, pos)[0]
            pos += size
        elif kind == "A":
            tags[name] = chr(buf[pos])
            pos += 1
        elif kind in "ZH":
            stop = buf.index(b"\0", pos)
            tags[name] = buf[pos:stop].decode("ascii")
            pos = stop + 1
        elif kind == "B":
            sub = chr(buf[pos])
            count = struct.unpack_from("<i", buf, pos + 1)[0]
            pos += 5
            fmt = scalar[sub]
            size = struct.calcsize(fmt)
            tags[name] = ("B" + sub, struct.unpack_from("<%d%s" % (count, fmt[1]), buf, pos))
            pos += size * count
        else:
            raise ValueError("unknown BAM tag type %r" % kind)
    return tags


class AlignmentFile(object):
    """Whole-file in-memory BAM reader (fixtures and synthetic files are small)."""

    def __init__(self, path, mode="rb", **kwargs):
        self.filename = path
        with gzip.open(path, "rb") as handle:      # BGZF == concatenated gzip members
            data = handle.read()
        if data[:4] != b"BAM\1":
            raise ValueError("not a BAM file: %s" % path)
        l_text = struct.unpack_from("<i", data, 4)[0]
        text = data[8:8 + l_text].split(b"\0", 1)[0].decode("ascii", "replace")
        pos = 8 + l_text
        n_ref = struct.unpack_from("<i", data, pos)[0]
        pos += 4
        names, lengths = [], []
        for _ in range(n_ref):
            l_name = struct.unpack_from("<i", data, pos)[0]
            names.append(data[pos + 4:pos + 4 + l_name - 1].decode("ascii"))
            lengths.append(struct.unpack_from("<i", data, pos + 4 + l_name)[0])
            pos += 8 + l_name
        self.references = tuple(names)
        self.lengths = tuple(lengths)
        self.nreferences = n_ref
        self._tid = {name: i for i, name in enumerate(names)}
        self.header = self._parse_header(text)
        self.text = text
        self._records = []
        size = len(data)
        while pos + 4 <= size:
            block = struct.unpack_from("<i", data, pos)[0]
            self._records.append(self._decode(data, pos + 4, pos + 4 + block))
            pos += 4 + block

    @staticmethod
    def _parse_header(text):
        header = {}
        for line in text.splitlines():
            if not line.startswith("@") or len(line) < 3:
                continue
            kind = line[1:3]
            fields = {}
            for item in line.split("\t")[1:]:
                if ":" in item:
                    key, value = item.split(":", 1)
                    fields[key] = value
            if kind == "HD":
                header["HD"] = fields
            else:
                header.setdefault(kind, []).append(fields)
        return header

    def _decode(self, data, pos, end):
        (ref_id, start, l_name, mapq, _bin, n_cigar, flag, l_seq, nref, npos, tlen) = struct.unpack_from(
            "<iiBBHHHiiii", data, pos)
        rec = AlignedSegment()
        cursor = pos + 32
        rec.query_name = data[cursor:cursor + l_name - 1].decode("ascii")
        cursor += l_name
        cigar = [(v & 15, v >> 4) for v in struct.unpack_from("<%dI" % n_cigar, data, cursor)]
        cursor += 4 * n_cigar
        rec._seq4 = data[cursor:cursor + (l_seq + 1) // 2] if l_seq > 0 else None
        rec._l_seq = l_seq
        cursor += (l_seq + 1) // 2 + l_seq
        rec._tags = _parse_tags(data, cursor, end)
        # long CIGAR convention: "<l_seq>S<ref_len>N" placeholder + CG:B,I tag
        if (n_cigar == 2 and cigar[0] == (4, l_seq) and cigar[1][0] == 3 and "CG" in rec._tags
                and rec._tags["CG"][0] == "BI"):
            cigar = [(v & 15, v >> 4) for v in rec._tags["CG"][1]]
            del rec._tags["CG"]
        # B-arrays other than CG are returned as plain tuples
        for key, val in list(rec._tags.items()):
            if isinstance(val, tuple) and len(val) == 2 and isinstance(val[0], str) and val[0].startswith("B"):
                rec._tags[key] = val[1]
        rec._cigar = cigar
        rec.flag = flag
        rec.reference_id = ref_id
        rec.reference_start = start
        rec._mapq = mapq
        rec.next_reference_id = nref
        rec.next_reference_start = npos
        rec.template_length = tlen
        return rec

    # ---- header access
    def get_tid(self, name):
        return self._tid.get(name, -1)

    def get_reference_name(self, tid):
        if not 0 <= tid < self.nreferences:
            raise ValueError("reference_id %i out of range 0<=tid<%i" % (tid, self.nreferences))
        return self.references[tid]

    getrname = get_reference_name

    def get_reference_length(self, name):
        if name not in self._tid:
            raise KeyError("unknown reference %s" % name)
        return self.lengths[self._tid[name]]

    def check_index(self):
        for suffix in (".bai", ".csi"):
            if os.path.exists(self.filename + suffix):
                return True
        stem = os.path.splitext(self.filename)[0]
        if os.path.exists(stem + ".bai"):
            return True
        raise ValueError("mapping information not recorded in index or index not available")

    def fetch(self, contig=None, start=None, stop=None, until_eof=False, **kwargs):
        if contig is None:
            return iter(list(self._records))
        tid = self._tid[contig]
        # htslib returns every record placed on the contig (placed-unmapped mates included)
        return iter([rec for rec in self._records if rec.reference_id == tid])

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class FastaFile(object):
    """faidx-style random access; `ValueError` without .fai, `IOError` without file."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise IOError("file `%s` not found" % path)
        if not os.path.exists(path + ".fai"):
            raise ValueError("could not locate index file for %s" % path)
        self._index = {}
        self.references = []
        with open(path + ".fai") as fai:
            for line in fai:
                cols = line.rstrip("\n").split("\t")
                if len(cols) < 5:
                    continue
                self._index[cols[0]] = tuple(int(c) for c in cols[1:5])
                self.references.append(cols[0])
        self._fh = open(path, "rb")

    def get_reference_length(self, contig):
        return self._index[contig][0]

    def fetch(self, reference=None, start=None, end=None, region=None):
        length, offset, linebases, linewidth = self._index[reference]
        start = 0 if start is None else max(0, start)
        end = length if end is None else min(length, end)
        if start >= end:
            return ""
        first = offset + (start // linebases) * linewidth + start % linebases
        last = offset + (end // linebases) * linewidth + end % linebases
        self._fh.seek(first)
        chunk = self._fh.read(last - first)
        return chunk.replace(b"\n", b"").replace(b"\r", b"").decode("ascii")[: end - start]

    def close(self):
        self._fh.close()
