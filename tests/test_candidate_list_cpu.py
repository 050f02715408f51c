"""Host logic around the device tables: the lazy CandidateList of analyze_alignment_file_coordsorted / pair_candidates
(objects only when somebody looks), its per-class views, and vcf_entries' order and numbering against the python writer on
randomised tables (contig names that sort differently as strings and naturally, coinciding keys, every class)."""
import argparse
import os

import numpy as np
import pytest

from svim_asm_b200 import SVCandidate as C
from svim_asm_b200 import _lib
from svim_asm_b200 import SVIM_COMBINE as combine
from svim_asm_b200.SVIM_COLLECT import CandidateList
from tests import hostcheck


class _Host(object):
    def query_name(self, i):
        return "read_%d" % i

    def sequence_slice(self, aln, pos, n):
        return "ACGT"[aln % 4] * n


def _random_rows(rng, n, names, lengths):
    rows = np.zeros(n, dtype=_lib.ROW_DTYPE)
    rows["type"] = rng.integers(0, 6, n)
    rows["hap"] = rng.integers(1, 3, n)
    rows["genotype"] = rng.integers(0, 3, n)
    rows["aln_idx"] = np.arange(n)
    rows["mate_aln"] = np.where(rng.random(n) < 0.3, rng.integers(0, n, n), C.NO_MATE)
    rows["ordinal"] = np.arange(n)
    for i in range(n):
        t = rows["type"][i]
        a, b = int(rng.integers(0, len(names))), int(rng.integers(0, len(names)))
        s = int(rng.integers(0, lengths[a] - 50) // 7 * 7)                 # coarse grid: equal keys do occur
        ds = int(rng.integers(0, lengths[b] - 50) // 7 * 7)
        rows["src_tid"][i], rows["dst_tid"][i] = (a if t != 2 else -1), (b if t in (2, 4, 5) else -1)
        rows["src_start"][i], rows["src_end"][i] = s, s + int(rng.integers(0, 40))
        rows["dst_start"][i], rows["dst_end"][i] = ds, ds + int(rng.integers(0, 40))
        rows["copies"][i] = rng.integers(1, 4)
        rows["flags"][i] = rng.integers(0, 32)
        if t == 2:
            rows["seq_len"][i] = rng.integers(0, 30)
    return rows


@pytest.fixture(scope="module")
def table():
    rng = np.random.default_rng(99)
    names = ["chr10", "chr2", "chr1", "chrM", "2", "10", "scaf_9", "scaf_10", "chr01"]
    lengths = [int(x) for x in rng.integers(400, 900, len(names))]
    rows = _random_rows(rng, 600, names, lengths)
    bases = {n: "".join(rng.choice(list("ACGTn"), L).tolist()) for n, L in zip(names, lengths)}
    return names, lengths, rows, bases


def _list(table):
    names, lengths, rows, _ = table
    return CandidateList.from_rows(rows, {1: _Host(), 2: _Host()}, names, lengths, None, object(), {1: None, 2: None})


def test_lazy_until_looked_at(table):
    lst = _list(table)
    n = table[2].shape[0]
    assert len(lst) == n and bool(lst) and lst._pending is not None            # counting builds nothing
    views = {t: lst.of_type(t) for t in C.TYPE_NAMES}
    assert sum(len(v) for v in views.values()) == n and lst._pending is not None
    assert all(v._pending is not None and v.table is lst.table for v in views.values())
    assert np.array_equal(np.sort(np.concatenate([v.row_index for v in views.values()])), np.arange(n))
    first = lst[0]                                                             # indexing does
    assert lst._pending is None and first.type == C.TYPE_NAMES[int(table[2]["type"][0])]
    assert [c.type for c in lst] == [C.TYPE_NAMES[int(t)] for t in table[2]["type"]]
    dels = lst.of_type("DEL")                                                  # views of a built list share its objects
    assert dels._pending is None and all(a is b for a, b in zip(dels, [c for c in lst if c.type == "DEL"]))
    assert dels.table is lst.table


def test_mutation_cuts_the_device_link(table):
    lst = _list(table)
    extra = C.CandidateDeletion("chr1", 5, 9, ["r"], C._Lengths(table[0], table[1]))
    lst.append(extra)
    assert lst.table is None and lst.rows is None and len(lst) == table[2].shape[0] + 1 and lst[-1] is extra
    plain = _list(table) + [extra]                                             # list + list: a plain list of objects
    assert type(plain) is list and len(plain) == table[2].shape[0] + 1
    assert combine._device_lists((("DEL", lst),)) is None                      # the writer then takes the python path


class _Fasta(object):
    def __init__(self, bases):
        self.bases = bases

    def fetch(self, contig, start, end):
        return self.bases[contig][start:end]

    def close(self):
        pass


@pytest.mark.parametrize("opts", [dict(), dict(symbolic_alleles=True), dict(tandem_duplications_as_insertions=True,
                                                                          interspersed_duplications_as_insertions=True),
                                  dict(types="DEL,BND,DUP:TANDEM")])
def test_entries_and_plan_match_python_writer(table, tmp_path, opts):
    """600 random rows: vcf_entries + the device plan (host build) give the python writer's record lines byte for byte."""
    names, lengths, rows, bases = table
    objs = _list(table)
    types = [t.strip() for t in opts.get("types", "DEL,INS,INV,DUP:TANDEM,DUP:INT,BND").split(",")]
    o = argparse.Namespace(working_dir=str(tmp_path), sample="S1", query_names=False, symbolic_alleles=opts.get("symbolic_alleles", False),
                           tandem_duplications_as_insertions=opts.get("tandem_duplications_as_insertions", False),
                           interspersed_duplications_as_insertions=opts.get("interspersed_duplications_as_insertions", False))
    by = {t: [c for c in objs if c.type == t] for t in C.TYPE_NAMES}
    combine.write_final_vcf(by["DUP_INT"], by["INV"], by["DUP_TAN"], by["DEL"], by["INS"], by["BND"], "1.0.3", names, lengths, types,
                            _Fasta(bases), o)
    want = "".join(ln + "\n" for ln in open(os.path.join(str(tmp_path), "variants.vcf")).read().split("\n")
                   if ln and not ln.startswith("#")).encode()
    entries = combine.vcf_entries(rows, np.arange(rows.shape[0], dtype=np.uint32), names, types, o.tandem_duplications_as_insertions,
                                  o.interspersed_duplications_as_insertions)
    # the pseudo query sequences of _Host.sequence_slice, 4-bit packed, one record per row
    lut = {"A": 1, "C": 2, "G": 4, "T": 8}
    seq_off = np.zeros(rows.shape[0] + 1, dtype=np.uint64)
    seq_off[1:] = np.cumsum((rows["seq_pos"].astype(np.int64) + rows["seq_len"] + 1) // 2)
    seq4 = np.zeros(int(seq_off[-1]) + 1, dtype=np.uint8)
    for i in np.nonzero(rows["seq_len"] > 0)[0]:
        code = lut["ACGT"[int(rows["aln_idx"][i]) % 4]]
        seq4[int(seq_off[i]):int(seq_off[i + 1])] = (code << 4) | code
    upper = [bases[n].upper().encode() for n in names]
    off = np.zeros(len(names) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in upper])
    flat = np.frombuffer(b"".join(upper), dtype=np.uint8)
    blob = b"".join(n.encode() for n in names)
    name_off = np.zeros(len(names) + 1, dtype=np.uint32)
    name_off[1:] = np.cumsum([len(n) for n in names])
    out = np.zeros(len(want) + 64, dtype=np.uint8)
    lib = hostcheck.load()
    n = lib.hc_vcf_body(rows.ctypes.data, entries.ctypes.data, entries.shape[0], flat.ctypes.data, off.ctypes.data, len(names), blob,
                        name_off.ctypes.data, None, None, seq4.ctypes.data, seq_off.ctypes.data, seq4.ctypes.data, seq_off.ctypes.data,
                        1 if o.symbolic_alleles else 0, out.ctypes.data, out.shape[0])
    assert out[:n].tobytes() == want
