"""One rank of the peer-memory exchange test (started by tests/test_exchange_window_gpu.py): its own process, its own CUDA
context, the window handles exchanged through files.  Runs the REAL multi-GPU step of bench.py --gpus N
(svim_asm_b200/sharded.sharded_step_window) on the GPU it is given (two ranks may share one GPU: CUDA IPC works across
processes on one device too) and leaves the gathered table of every step in `out_dir`."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(rank, world, device, out_dir, seed):
    sys.path.insert(0, ROOT)
    from svim_asm_b200 import sharded, synth
    from svim_asm_b200.bench_util import pinned_host
    from svim_asm_b200.engine import Engine, HostBatch, lexrank, make_params
    names = ["chr1", "chr10", "chr2", "chr3", "chrX"]
    cfg = synth.SynthConfig(names, [300000, 200000, 250000, 150000, 100000], 120, 5e4, seed, sv_per_event=6e-3,
                            split_fraction=0.6, sv_max=2500)
    rb = synth.make_diploid(cfg)
    ref_dict = synth.random_reference(cfg)
    off = np.zeros(len(names) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([ref_dict[n].shape[0] for n in names])
    bases = np.concatenate([ref_dict[n] for n in names])
    eng = Engine(device)
    ref = eng.load_reference(bases, off)
    params = make_params()
    owner = sharded.lpt_assign(sharded.contig_weights(rb[0], len(names)) + sharded.contig_weights(rb[1], len(names)), world)
    hosts, gidx = [], []
    for k in range(2):
        sub, g = sharded.shard_records(rb[k], owner, rank)
        hosts.append(pinned_host(HostBatch.from_record_batch(sub)))        # the upload-per-step leg reads the sequences in place
        gidx.append(g)
    resident = [eng.load_records(h, with_sequences=True) for h in hosts]
    for rec, g in zip(resident, gidx):
        eng.set_global_index(rec, g)
    window = eng.exchange_window(world, rank, 8 << 20, 2 << 20)
    with open(os.path.join(out_dir, "handle_%d.tmp" % rank), "wb") as f:
        f.write(window.handle)
    os.rename(os.path.join(out_dir, "handle_%d.tmp" % rank), os.path.join(out_dir, "handle_%d.bin" % rank))
    handles, deadline = [], time.time() + 120
    for r in range(world):
        path = os.path.join(out_dir, "handle_%d.bin" % r)
        while not os.path.exists(path):
            if time.time() > deadline:
                raise RuntimeError("rank %d never published its window handle" % r)
            time.sleep(0.01)
        handles.append(open(path, "rb").read())
    window.open(handles)
    ranks = lexrank(names)
    for step in range(3):              # several steps: epochs advance, slots are reused
        stage = sharded.DeviceStage(eng, hosts, gidx, ref, params, resident if step != 1 else None)      # step 1: the upload-per-step leg
        table = sharded.sharded_step_window(stage, window, owner, ranks)
        np.save(os.path.join(out_dir, "table_r%d_s%d.npy" % (rank, step)), table)
    window.close()
    eng.close()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]))
