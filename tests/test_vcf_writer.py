"""The VCF writer against the reference's own output (tests/golden/vcf_writer.json, written by the unmodified
write_final_vcf / get_vcf_entry* of the reference for seeded candidates of every class under six option sets):

  * the python writer (svim_asm_b200.SVIM_COMBINE.write_final_vcf over Candidate objects),
  * the record plan of the device writer compiled for the host (csrc/vcf_core.cuh through tests/hostcheck) fed by
    vcf_entries (order, ID numbering),
  * on a GPU: svb_vcf_body itself, called directly and through write_final_vcf over device-backed lists.
"""
import argparse
import json
import os

import numpy as np
import pytest

from svim_asm_b200 import SVCandidate as C
from svim_asm_b200 import _lib
from svim_asm_b200 import SVIM_COMBINE as combine
from tests import hostcheck

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vcf_writer.json")


class _Fasta(object):
    def __init__(self, bases):
        self.bases = bases

    def fetch(self, contig, start, end):
        return self.bases[contig][start:end]

    def close(self):
        pass


@pytest.fixture(scope="module")
def fx():
    d = json.load(open(GOLDEN))
    d["names"] = [n for n, _ in d["contigs"]]
    d["lengths"] = [l for _, l in d["contigs"]]
    bam = C._Lengths(d["names"], d["lengths"])
    objs = []
    for c in d["candidates"]:
        t, gt, reads = c["type"], c["genotype"], c["reads"]
        if t == "DEL":
            o = C.CandidateDeletion(c["source_contig"], c["source_start"], c["source_end"], reads, bam, gt)
        elif t == "INV":
            o = C.CandidateInversion(c["source_contig"], c["source_start"], c["source_end"], reads, c["complete"], bam, gt)
        elif t == "INS":
            o = C.CandidateInsertion(c["dest_contig"], c["dest_start"], c["dest_end"], reads, c["sequence"], bam, gt)
        elif t == "DUP_TAN":
            o = C.CandidateDuplicationTandem(c["source_contig"], c["source_start"], c["source_end"], c["copies"], c["fully_covered"],
                                             reads, bam, gt)
        elif t == "DUP_INT":
            o = C.CandidateDuplicationInterspersed(c["source_contig"], c["source_start"], c["source_end"], c["dest_contig"],
                                                   c["dest_start"], c["dest_end"], reads, bam, c["cutpaste"], gt)
        else:
            o = C.CandidateBreakend.__new__(C.CandidateBreakend)       # already normalised by the reference's constructor
            o.source_contig, o.source_start, o.source_direction = c["source_contig"], c["source_start"], c["source_direction"]
            o.dest_contig, o.dest_start, o.dest_direction = c["dest_contig"], c["dest_start"], c["dest_direction"]
            o.reads, o.genotype = reads, gt
        objs.append(o)
    d["objects"] = objs
    return d


def _options(tmp, o):
    return argparse.Namespace(working_dir=str(tmp), sample="Sample", query_names=False,
                              symbolic_alleles=o.get("symbolic_alleles", False),
                              tandem_duplications_as_insertions=o.get("tandem_duplications_as_insertions", False),
                              interspersed_duplications_as_insertions=o.get("interspersed_duplications_as_insertions", False))


def _split(text):
    lines = text.split("\n")
    header = [ln for ln in lines if ln.startswith("#") and not ln.startswith("##fileDate")]
    return header, "".join(ln + "\n" for ln in lines if ln and not ln.startswith("#"))


def _write(fx, tmp, o, lists):
    combine.write_final_vcf(lists["DUP_INT"], lists["INV"], lists["DUP_TAN"], lists["DEL"], lists["INS"], lists["BND"], "1.0.3",
                            fx["names"], fx["lengths"], [t.strip() for t in o["types"].split(",")], _Fasta(fx["bases"]),
                            _options(tmp, o))
    return _split(open(os.path.join(str(tmp), "variants.vcf")).read())


@pytest.mark.parametrize("key", ["default", "symbolic", "dups_as_ins", "dups_as_ins_symbolic", "subset", "no_ins"])
def test_python_writer_matches_reference(fx, tmp_path, key):
    lists = {t: [c for c in fx["objects"] if c.type == t] for t in C.TYPE_NAMES}
    header, body = _write(fx, tmp_path, fx["option_sets"][key], lists)
    assert header == fx["outputs"][key + "/header"]
    assert body == fx["outputs"][key]


def _tables(fx):
    """Rows in candidate order, alternating between two haplotype record images that hold the inserted sequences."""
    names = fx["names"]
    objs = fx["objects"]
    rows = np.zeros(len(objs), dtype=_lib.ROW_DTYPE)
    batches = {}
    for hap in (1, 2):
        part, rb = combine._rows_from_objects(objs[hap - 1::2], hap, names)
        rb.contig_lengths = np.asarray(fx["lengths"], dtype=np.int32)
        rows[hap - 1::2] = part
        batches[hap] = rb
    upper = [fx["bases"][n].upper().encode() for n in names]
    off = np.zeros(len(names) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in upper])
    return rows, batches, np.frombuffer(b"".join(upper), dtype=np.uint8), off


def _entries(fx, rows, o):
    return combine.vcf_entries(rows, np.arange(rows.shape[0], dtype=np.uint32), fx["names"], [t.strip() for t in o["types"].split(",")],
                               o.get("tandem_duplications_as_insertions", False), o.get("interspersed_duplications_as_insertions", False))


@pytest.mark.parametrize("key", ["default", "symbolic", "dups_as_ins", "dups_as_ins_symbolic", "subset", "no_ins"])
def test_device_plan_on_host_matches_reference(fx, key):
    lib = hostcheck.load()
    o = fx["option_sets"][key]
    rows, batches, bases, off = _tables(fx)
    entries = _entries(fx, rows, o)
    blob = b"".join(n.encode() for n in fx["names"])
    name_off = np.zeros(len(fx["names"]) + 1, dtype=np.uint32)
    name_off[1:] = np.cumsum([len(n) for n in fx["names"]])
    seq = {h: (np.ascontiguousarray(batches[h].seq4), np.ascontiguousarray(batches[h].seq_off, dtype=np.uint64)) for h in (1, 2)}
    want = fx["outputs"][key].encode()
    out = np.zeros(len(want) + 64, dtype=np.uint8)
    n = lib.hc_vcf_body(rows.ctypes.data, entries.ctypes.data, entries.shape[0], bases.ctypes.data, off.ctypes.data, len(fx["names"]),
                        blob, name_off.ctypes.data, None, None, seq[1][0].ctypes.data, seq[1][1].ctypes.data,
                        seq[2][0].ctypes.data, seq[2][1].ctypes.data, 1 if o.get("symbolic_alleles") else 0, out.ctypes.data,
                        out.shape[0])
    assert out[:n].tobytes() == want


def test_entry_order_and_numbering(fx):
    """sorted_nicely (SVIM_COMBINE.py:369-376): natural contig order ('1' < 'chr1' = 'chr01' < 'chr2' < 'chr10'), then the
    key start and end, ties in append order; every ID label numbered on its own."""
    rows, _, _, _ = _tables(fx)
    entries = _entries(fx, rows, fx["option_sets"]["default"])
    lines = fx["outputs"]["default"].split("\n")[:-1]
    assert entries.shape[0] == len(lines)
    labels = ("DEL", "INV", "INS", "INS", "DUP_TANDEM", "INS", "DUP_INT", "BND", "BND")
    assert [ln.split("\t")[2] for ln in lines] == ["svim_asm.%s.%d" % (labels[m], i) for m, i in zip(entries["mode"], entries["id"])]


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["default", "symbolic", "dups_as_ins", "dups_as_ins_symbolic", "subset", "no_ins"])
def test_device_writer_matches_reference(fx, engine, tmp_path, key):
    from svim_asm_b200.engine import HostBatch
    from svim_asm_b200.SVIM_COLLECT import CandidateList
    o = fx["option_sets"][key]
    rows, batches, bases, off = _tables(fx)
    hosts = {h: HostBatch.from_record_batch(batches[h]) for h in (1, 2)}
    records = {h: engine.load_records(hosts[h], with_sequences=True) for h in (1, 2)}
    table = engine.table_from_numpy(rows)
    ref = engine.load_reference(bases, off)
    body = engine.vcf_body(table, records, ref, fx["names"], _entries(fx, rows, o), bool(o.get("symbolic_alleles")))
    assert body == fx["outputs"][key].encode()
    # the same through write_final_vcf: device-backed per-class lists, no Candidate object is built
    whole = CandidateList.from_rows(rows, hosts, fx["names"], fx["lengths"], None, table, records)
    lists = {t: whole.of_type(t) for t in C.TYPE_NAMES}
    fasta = _Fasta(fx["bases"])
    fasta._svb_ref = (tuple(fx["names"]), ref)
    combine.write_final_vcf(lists["DUP_INT"], lists["INV"], lists["DUP_TAN"], lists["DEL"], lists["INS"], lists["BND"], "1.0.3",
                            fx["names"], fx["lengths"], [t.strip() for t in o["types"].split(",")], fasta, _options(tmp_path, o))
    header, text = _split(open(os.path.join(str(tmp_path), "variants.vcf")).read())
    assert header == fx["outputs"][key + "/header"]
    assert text == fx["outputs"][key]
    assert all(lst._pending is not None for lst in lists.values())
