"""Device-resident multi-GPU exchange (csrc/exchange.cu) on ONE GPU with virtual ranks.

Records are sharded by contig exactly as bench.py does under torchrun; every virtual rank collects its shard, packs its
tables, the packed buffers are laid side by side (what the NCCL all-gather produces), and every rank unpacks its share.
Checked against (a) the numpy statement of the same merge/select (svim_asm_b200/sharded.py, also exercised over gloo in
test_sharded_cpu.py) and (b) the unsharded single-GPU result: sharding must not change one bit of the paired table
(reference: one process sees every candidate, SVIM_COLLECT.py:67-91 -> SVIM_COMBINE.py:164)."""
import numpy as np
import pytest

from svim_asm_b200 import _lib, sharded, synth
from svim_asm_b200.engine import HostBatch, lexrank, make_params
from tests import util
from tests.test_pair_gpu import _reference_arrays

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,seed", [(2, 41), (3, 42), (5, 43)])
def test_exchange_matches_host_logic_and_unsharded(engine, world, seed):
    names = ["chr1", "chr10", "chr2", "chr3", "chrX"]
    cfg = synth.SynthConfig(names, [300000, 200000, 250000, 150000, 100000], 120, 5e4, seed, sv_per_event=6e-3,
                            split_fraction=0.6, sv_max=2500)
    rb = synth.make_diploid(cfg)
    bases, off = _reference_arrays(cfg)
    ref = engine.load_reference(bases, off)
    params = make_params()
    ranks = lexrank(names)

    # unsharded truth
    full_hosts = [HostBatch.from_record_batch(b) for b in rb]
    full_recs = [engine.load_records(h, with_sequences=True) for h in full_hosts]
    full_tabs = [engine.collect(r, params, hap=k + 1) for k, r in enumerate(full_recs)]
    want = engine.pair(full_tabs[0], full_tabs[1], full_recs[0], full_recs[1], ref, params).to_numpy()
    assert want.shape[0] > 50

    owner = sharded.lpt_assign(sharded.contig_weights(rb[0], len(names)) + sharded.contig_weights(rb[1], len(names)), world)
    recs, host_parts, sizes, tables = [], [], [], []
    for r in range(world):
        recs_r, parts_r, tabs_r = [], [], []
        for k in range(2):
            sub, gidx = sharded.shard_records(rb[k], owner, r)
            host = HostBatch.from_record_batch(sub)
            rec = engine.load_records(host, with_sequences=True)
            engine.set_global_index(rec, gidx)
            t = engine.collect(rec, params, hap=k + 1)
            t.gather_sequences(rec)
            t.remap_records(rec)
            pool, starts = t.pool_to_numpy()
            parts_r.append((t.to_numpy(), pool, starts))
            recs_r.append(rec)
            tabs_r.append(t)
        recs.append(recs_r)
        host_parts.append(parts_r)
        tables.append(tabs_r)
        sizes.append(engine.exchange_sizes(tabs_r[0], tabs_r[1]))
    sizes = np.stack(sizes)
    assert any(s[0] and s[2] for s in sizes)
    stride = max(engine.exchange_bytes(s) for s in sizes)
    gathered = engine.device_alloc(world * stride)
    for r in range(world):
        engine.exchange_pack(tables[r][0], tables[r][1], gathered.ptr + r * stride, stride)

    paired = []
    for r in range(world):
        unpacked = []
        for hap in (1, 2):
            u = engine.exchange_unpack(gathered.ptr, stride, sizes, hap, owner, r)
            rows, pool, starts = sharded.select_owned(*sharded.merge_gathered(host_parts, hap - 1), owner, r)
            got_rows = u.to_numpy()
            assert got_rows.tobytes() == rows.tobytes(), "rank %d hap %d rows" % (r, hap)
            got_pool, got_starts = u.pool_to_numpy()
            assert got_pool.tobytes() == pool.tobytes()
            assert got_starts.tolist() == starts.tolist()
            unpacked.append(u)
        paired.append(engine.pair(unpacked[0], unpacked[1], recs[r][0], recs[r][1], ref, params).to_numpy())
    got = sharded.order_paired(paired, ranks)
    diff = util.rows_equal(got, want)
    assert diff is None, diff
    gathered.free()


def test_pool_with_explicit_offsets(engine):
    """A pool attached with per-row offsets in arbitrary order (what a gathered pool looks like) pairs like the original."""
    cfg = synth.SynthConfig(["chr1", "chr2"], [300000, 200000], 60, 5e4, 44, sv_per_event=6e-3, split_fraction=0.3, sv_max=2000)
    rb = synth.make_diploid(cfg)
    bases, off = _reference_arrays(cfg)
    ref = engine.load_reference(bases, off)
    params = make_params()
    hosts = [HostBatch.from_record_batch(b) for b in rb]
    recs = [engine.load_records(h, with_sequences=True) for h in hosts]
    tabs = [engine.collect(r, params, hap=k + 1) for k, r in enumerate(recs)]
    want = engine.pair(tabs[0], tabs[1], recs[0], recs[1], ref, params).to_numpy()
    rng = np.random.default_rng(5)
    shuffled = []
    for t, rec in zip(tabs, recs):
        t.gather_sequences(rec)
        rows = t.to_numpy()
        pool, starts = t.pool_to_numpy()
        lens = np.where(rows["type"] == 2, (rows["seq_len"].astype(np.int64) + 1) // 2, 0)
        order = rng.permutation(rows.shape[0])
        new_pool, new_starts, at = [np.zeros(7, np.uint8)], np.zeros(rows.shape[0], dtype=np.uint64), 7
        for i in order:
            new_starts[i] = at
            new_pool.append(pool[int(starts[i]):int(starts[i]) + int(lens[i])])
            at += int(lens[i]) + 3
            new_pool.append(np.full(3, 0xEE, np.uint8))
        t2 = engine.table_from_numpy(rows)
        t2.set_pool(np.concatenate(new_pool), new_starts)
        shuffled.append(t2)
    got = engine.pair(shuffled[0], shuffled[1], recs[0], recs[1], ref, params).to_numpy()
    assert util.rows_equal(got, want) is None
