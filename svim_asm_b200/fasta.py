"""Indexed FASTA access: the part of pysam.FastaFile the hot path uses (svim-asm:124; SVCandidate.py:57-58,105,
155,210,301-302; SVIM_COMBINE.py:45-99).  Semantics per SURVEY.md App. C: needs `<path>.fai`
(ValueError without it, IOError without the FASTA itself); fetch() is 0-based half-open, clipped at the contig
end, '' when start >= end."""
import os

import numpy as np


class FastaFile(object):
    def __init__(self, path):
        if not os.path.exists(path):
            raise IOError("file `%s` not found" % path)
        if not os.path.exists(path + ".fai"):
            raise ValueError("could not locate index file for %s" % path)
        self.filename = path
        self.references = []
        self._index = {}
        with open(path + ".fai") as fai:
            for line in fai:
                cols = line.rstrip("\n").split("\t")
                if len(cols) >= 5:
                    self.references.append(cols[0])
                    self._index[cols[0]] = tuple(int(c) for c in cols[1:5])
        self.lengths = [self._index[n][0] for n in self.references]
        self._fh = open(path, "rb")

    def get_reference_length(self, contig):
        return self._index[contig][0]

    def _read(self, contig, start, end):
        length, offset, linebases, linewidth = self._index[contig]
        first = offset + (start // linebases) * linewidth + start % linebases
        last = offset + (end // linebases) * linewidth + end % linebases
        self._fh.seek(first)
        raw = self._fh.read(last - first)
        if linewidth != linebases:
            raw = raw.replace(b"\n", b"").replace(b"\r", b"")
        return raw[: end - start]

    def fetch(self, reference=None, start=None, end=None, region=None):
        length = self._index[reference][0]
        start = 0 if start is None else max(0, start)
        end = length if end is None else min(length, end)
        if start >= end:
            return ""
        return self._read(reference, start, end).decode("ascii")

    def load_upper(self, contig_names):
        """Upper-cased bases of `contig_names` (BAM header order), concatenated, for svb_ref_load.
        Returns (uint8 array, uint64 offsets[n + 1])."""
        offsets = np.zeros(len(contig_names) + 1, dtype=np.uint64)
        parts = []
        for i, name in enumerate(contig_names):
            if name in self._index:
                raw = np.frombuffer(self._read(name, 0, self._index[name][0]), dtype=np.uint8)
                lower = (raw >= 97) & (raw <= 122)
                parts.append(np.where(lower, raw - 32, raw).astype(np.uint8))
            else:
                parts.append(np.zeros(0, dtype=np.uint8))     # contig absent from the FASTA: fetch() would raise KeyError
            offsets[i + 1] = offsets[i] + np.uint64(parts[-1].shape[0])
        return (np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)), offsets

    def fai_rows(self, contig_names):
        """(length, offset, linebases, linewidth) per contig of `contig_names`, zeros for contigs the FASTA lacks
        (svb_ref_load_fasta builds the device copy from these)."""
        return [self._index.get(name, (0, 0, 0, 0)) for name in contig_names]

    def close(self):
        self._fh.close()
