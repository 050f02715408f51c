"""Host logic of the multi-GPU path (svim_asm_b200/sharded.py) with world_size 2 on CPU (gloo).
The GPU stages are replaced by the CPU oracle, so what is checked is the sharding, the all-gatherv payloads, the
ordinal merge, the ownership filter and the final ordering: the 2-rank result must equal the 1-process result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_sharding_equals_single_process(tmp_path, built_library, oracle_clib):
    import subprocess
    from tests import util
    port = _free_port()
    script = os.path.join(ROOT, "tests", "sharded_worker.py")
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", str(port), str(tmp_path)]) for r in range(2)]
    assert [p.wait(timeout=600) for p in procs] == [0, 0]
    got, want = np.load(tmp_path / "got.npy"), np.load(tmp_path / "want.npy")
    assert want.shape[0] > 50
    assert util.rows_equal(got, want) is None, util.rows_equal(got, want)


def test_lpt_assignment_balances():
    from svim_asm_b200 import sharded, synth
    w = np.array(synth.HG38_LENGTHS, dtype=np.float64)
    owner = sharded.lpt_assign(w, 8)
    load = np.bincount(owner, weights=w, minlength=8)
    assert load.max() / load.mean() < 1.05
