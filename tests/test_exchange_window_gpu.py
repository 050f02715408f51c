"""The multi-GPU step over the peer-memory exchange window (csrc/exchange.cu: svb_exchange_share / svb_exchange_gather_paired)
with REAL ranks: one process and one CUDA context per rank, window handles over CUDA IPC, NVLink / peer stores, device-side
flags.  With two or more GPUs every rank gets its own; on a one-GPU box the ranks share the device (IPC works across
processes on one device), so the path bench.py --gpus N takes is exercised wherever the tests run.
The gathered, ordered table of EVERY rank and EVERY step must equal the unsharded single-GPU result bit for bit
(reference: one process sees every candidate, SVIM_COLLECT.py:67-91 -> SVIM_COMBINE.py:164)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from svim_asm_b200 import synth
from svim_asm_b200.engine import HostBatch, make_params
from tests import util
from tests.test_pair_gpu import _reference_arrays

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,seed", [(2, 41), (3, 42)])
def test_window_exchange_equals_unsharded(engine, tmp_path, world, seed):
    import torch
    names = ["chr1", "chr10", "chr2", "chr3", "chrX"]
    cfg = synth.SynthConfig(names, [300000, 200000, 250000, 150000, 100000], 120, 5e4, seed, sv_per_event=6e-3,
                            split_fraction=0.6, sv_max=2500)
    rb = synth.make_diploid(cfg)
    bases, off = _reference_arrays(cfg)
    ref = engine.load_reference(bases, off)
    params = make_params()
    hosts = [HostBatch.from_record_batch(b) for b in rb]
    recs = [engine.load_records(h, with_sequences=True) for h in hosts]
    tabs = [engine.collect(r, params, hap=k + 1) for k, r in enumerate(recs)]
    want = engine.pair(tabs[0], tabs[1], recs[0], recs[1], ref, params).to_numpy()
    assert want.shape[0] > 50

    n_dev = torch.cuda.device_count()
    env = dict(os.environ, SVB_EXCHANGE_TIMEOUT_S="120")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "window_worker.py"), str(r), str(world), str(r % n_dev),
                               str(tmp_path), str(seed)], env=env) for r in range(world)]
    try:
        codes = [p.wait(timeout=600) for p in procs]
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert codes == [0] * world
    for r in range(world):
        for step in range(3):
            got = np.load(tmp_path / ("table_r%d_s%d.npy" % (r, step)))
            diff = util.rows_equal(got, want)
            assert diff is None, (r, step, diff)
            assert np.array_equal(got["ordinal"], np.arange(got.shape[0], dtype=np.uint64))
