"""One rank of the world_size-2 gloo test of svim_asm_b200/sharded.py (started by tests/test_sharded_cpu.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import port as oracle
    from svim_asm_b200 import sharded, synth
    from svim_asm_b200.engine import HostBatch, lexrank
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.SynthConfig(["chr1", "chr10", "chr2", "chr3"], [300000, 200000, 250000, 150000], 70, 4.5e4, 515,
                            sv_per_event=8e-3, split_fraction=0.5, sv_max=1200)
    rb1, rb2 = synth.make_diploid(cfg)
    ref = synth.random_reference(cfg)
    names = cfg.contig_names
    clen = [int(x) for x in cfg.contig_lengths]
    p = oracle.Params()

    def fetch(tid, s, e):
        return ref[names[tid]][s:e].tobytes()
    owner = sharded.lpt_assign(sharded.contig_weights(rb1, 4) + sharded.contig_weights(rb2, 4), world)
    shards = [sharded.shard_records(rb, owner, rank) for rb in (rb1, rb2)]
    hosts = [HostBatch.from_record_batch(s[0]) for s in shards]
    full = [HostBatch.from_record_batch(rb1), HostBatch.from_record_batch(rb2)]

    class OracleStage(object):
        def collect(self, hap):
            rows = oracle.collect(hosts[hap - 1], p, hap=hap)
            # sequence pool of the INS rows, 4-bit packed like the device gather
            off = np.zeros(rows.shape[0] + 1, dtype=np.uint64)
            chunks = [np.zeros(3, np.uint8)]          # pools need not start at offset 0
            lut = {c: i for i, c in enumerate(synth.NT16)}
            for i, r in enumerate(rows):
                seq = hosts[hap - 1].sequence_slice(int(r["aln_idx"]), int(r["seq_pos"]), int(r["seq_len"])) if r["type"] == 2 else ""
                codes = np.array([lut[c] for c in seq] + ([0] if len(seq) % 2 else []), dtype=np.uint8)
                packed = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8) if codes.shape[0] else np.zeros(0, np.uint8)
                chunks.append(packed)
                off[i + 1] = off[i] + np.uint64(packed.shape[0])
            pool = np.concatenate(chunks)
            return sharded.remap_to_global(rows, shards[hap - 1][1]), pool, off[:-1] + np.uint64(3)

        def pair(self, part1, part2):
            class PoolHost(object):
                def __init__(self, part, host):
                    self.rows, self.pool, self.starts = part
                    self.contig_names, self.contig_lengths = host.contig_names, host.contig_lengths
                    self.index = {int(r["ordinal"]): i for i, r in enumerate(self.rows)}

                def sequence_slice(self, aln, pos, length):
                    raise AssertionError("pairing must read sequences from the pool")
            hosts_pool = [PoolHost(part1, full[0]), PoolHost(part2, full[1])]

            # oracle.pair reads INS sequences through host.sequence_slice(aln_idx, seq_pos, seq_len): serve them from the pool
            def make(hp):
                def slicer(aln, pos, length, _hp=hp):
                    hits = np.nonzero((_hp.rows["aln_idx"] == aln) & (_hp.rows["seq_pos"] == pos) & (_hp.rows["type"] == 2))[0]
                    i = int(hits[0])
                    raw = _hp.pool[int(_hp.starts[i]):int(_hp.starts[i]) + (length + 1) // 2]
                    nib = np.empty(raw.shape[0] * 2, dtype=np.uint8)
                    nib[0::2], nib[1::2] = raw >> 4, raw & 15
                    return "".join(synth.NT16[c] for c in nib[:length])
                _hp = hp
                hp.sequence_slice = slicer
                return hp
            return oracle.pair(part1[0], part2[0], make(hosts_pool[0]), make(hosts_pool[1]), fetch, p)

    table = sharded.sharded_step(OracleStage(), rank, owner, lexrank(names), torch.device("cpu"), oracle.ROW_DTYPE)
    if rank == 0:
        want = oracle.pair(oracle.collect(full[0], p, hap=1), oracle.collect(full[1], p, hap=2), full[0], full[1], fetch, p)
        np.save(os.path.join(out_dir, "got.npy"), table)
        np.save(os.path.join(out_dir, "want.npy"), want)
    dist.destroy_process_group()



if __name__ == "__main__":
    _worker(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
