"""Parity at (a quarter of) whole-genome size, where the per-record oracle would take minutes:
* every indel row of svb_collect against the vectorised oracle (numpy prefix sums over the flat op array);
* size-independent properties of the full table: strictly increasing ordinal (the reference's append order),
  idempotence (the same launch twice gives the same bytes), both load paths (TMA ring / LDG) agree bit for bit,
  and the walk rows equal the per-record oracle restricted to the primaries that carry SA tags."""
import numpy as np
import pytest

from oracle import port
from svim_asm_b200 import synth
from svim_asm_b200.engine import HostBatch, make_params
from tests import util

pytestmark = pytest.mark.gpu


def test_quarter_genome_collect(engine):
    cfg = synth.config_c3(seed=1003, scale=0.25)
    cfg.with_sequence = False
    rb = synth.make_haploid(cfg)
    host = HostBatch.from_record_batch(rb)
    assert host.n_ops > 4.5e7 and int(host.hdr["n_cigar"].max()) > 65535
    rec = engine.load_records(host)
    params = make_params()
    tables = []
    for variant in (1, 0, 1):
        engine.set_scan_variant(variant)
        tables.append(engine.collect(rec, params).to_numpy())
    engine.set_scan_variant(0)
    assert tables[0].tobytes() == tables[1].tobytes() == tables[2].tobytes()
    rows = tables[0]
    assert np.all(np.diff(rows["ordinal"].astype(np.uint64).view(np.int64)) > 0)
    is_walk = (rows["ordinal"] & np.uint64(0x80000000)) != 0
    want = port.collect_indels_vectorised(host, port.Params())
    assert util.rows_equal(rows[~is_walk], want) is None, util.rows_equal(rows[~is_walk], want)
    assert np.array_equal(rows[~is_walk]["ordinal"], want["ordinal"])
    # walk rows: per-record oracle on the primaries with SA only (a few hundred records)
    with_sa = np.nonzero(host.sa_count > 0)[0]
    sub = HostBatch.from_record_batch(rb.subset(with_sa))
    want_walk = port.collect(sub, port.Params())
    want_walk = want_walk[(want_walk["ordinal"] & np.uint64(0x80000000)) != 0]
    got_walk = rows[is_walk].copy()
    got_walk["aln_idx"] = np.searchsorted(with_sa, got_walk["aln_idx"])          # global -> index in the subset
    assert util.rows_equal(got_walk, want_walk) is None, util.rows_equal(got_walk, want_walk)
    assert want.shape[0] > 2000 and want_walk.shape[0] > 100


def test_full_diploid_pair(engine, oracle_clib, tmp_path):
    """BASELINE.json configs[3] at FULL size (scale 1.0: 2 x 42,366 alignments, 2 x 2.04e8 ops, 24 contigs): collect x2 + pair
    on the GPU, then every paired row keyed on chr21, chr22, chrX, chrY against the oracle run on the closed sample of
    those contigs (partitions never span contigs, SVIM_COMBINE.py:24-26; oracle/hostimage.py), and the VCF record lines
    of those rows assembled on the device (svb_vcf_body) against the python writer fed with the ORACLE's rows.
    Covers what the small cases cannot: radix sort over > 18k keys, 24 contigs in lexrank order, the long-pair paths of
    the edit-distance stage, pool offsets of a whole genome."""
    import argparse
    import bench
    from oracle import hostimage
    from svim_asm_b200 import SVIM_COMBINE as combine
    from svim_asm_b200.SVCandidate import TYPE_NAMES, candidates_from_rows
    cfg, rb1, rb2, bases, off = bench.build_workload(1.0)
    h1, h2 = HostBatch.from_record_batch(rb1), HostBatch.from_record_batch(rb2)
    params = make_params()
    r1, r2 = engine.load_records(h1, with_sequences=True), engine.load_records(h2, with_sequences=True)
    ref = engine.load_reference(bases, off)
    t1, t2 = engine.collect(r1, params, hap=1), engine.collect(r2, params, hap=2)
    paired = engine.pair(t1, t2, r1, r2, ref, params)
    got = paired.to_numpy()
    assert got.shape[0] > 10000

    s1, s2, tids, idx1, idx2 = bench.cpu_sample(rb1, rb2, cfg)
    _dt, _n_aln, _n_ops, want = bench.run_cpu_pipeline(s1, s2, bases, off)
    n, diff = hostimage.compare_on_contigs(got, want, tids, idx1, idx2)
    assert diff is None, diff
    assert n > 800

    # VCF lines of those rows: device body over the full table vs the python writer over the oracle's rows
    names, lengths = list(cfg.contig_names), [int(x) for x in cfg.contig_lengths]
    types = ["DEL", "INS", "INV", "DUP:TANDEM", "DUP:INT", "BND"]
    sel = np.nonzero(np.isin(hostimage.key_contig(got), tids))[0]
    entries = combine.vcf_entries(got[sel], sel.astype(np.uint32), names, types, False, False)
    body = engine.vcf_body(paired, {1: r1, 2: r2}, ref, names, entries, False)

    class Fasta(object):
        def fetch(self, contig, start, end):
            t = names.index(contig)
            return bases[int(off[t]) + start:int(off[t]) + end].tobytes().decode("ascii")

        def close(self):
            pass
    want_sel = want[np.isin(hostimage.key_contig(want), tids)]
    sub_hosts = {1: HostBatch.from_record_batch(s1), 2: HostBatch.from_record_batch(s2)}
    objs = candidates_from_rows(want_sel, sub_hosts, names, lengths)
    by = {t: [c for c in objs if c.type == t] for t in TYPE_NAMES}
    opts = argparse.Namespace(working_dir=str(tmp_path), sample="Sample", query_names=False, symbolic_alleles=False,
                              tandem_duplications_as_insertions=False, interspersed_duplications_as_insertions=False)
    import os
    os.environ["SVIM_ASM_B200_VCF"] = "host"
    try:
        combine.write_final_vcf(by["DUP_INT"], by["INV"], by["DUP_TAN"], by["DEL"], by["INS"], by["BND"], "1.0.3", names, lengths,
                                types, Fasta(), opts)
    finally:
        del os.environ["SVIM_ASM_B200_VCF"]
    text = open(os.path.join(str(tmp_path), "variants.vcf")).read()
    lines = "".join(ln + "\n" for ln in text.split("\n") if ln and not ln.startswith("#")).encode()
    assert len(lines) > 100000
    assert body == lines
