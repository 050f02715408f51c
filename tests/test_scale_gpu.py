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
