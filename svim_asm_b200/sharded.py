"""Multi-GPU layout of the hot path: one process per GPU, records sharded by reference contig.

COLLECT needs no communication: every BAM record lives on one contig and the split-alignment walk of a primary
only reads that record's own SA tag.  The one exchange step is an all-gatherv of the
candidate tables (rows + the sequence pools of their INS rows): walk-derived candidates can land on a contig owned
by another rank, both haplotypes' rows of a key contig must meet for pairing, and the writer needs the global order.
After the exchange every rank pairs the key contigs it owns; the paired rows are gathered and put into the
reference's order (type, then contig by python string order, then partition order).

Everything here is host logic on numpy arrays plus torch.distributed collectives (NCCL on the GPU box, gloo in the
CPU tests); the compute stages are passed in, so the same code is exercised with world_size 2 on a CPU-only box.
"""
import os
import sys
import time

import numpy as np

HEADER_WORDS = 4


def lpt_assign(weights, n_ranks):
    """Longest-processing-time bin packing of contigs onto ranks by CIGAR-op count.  Returns owner[contig]."""
    order = np.argsort(-np.asarray(weights, dtype=np.float64), kind="stable")
    load = np.zeros(n_ranks)
    owner = np.zeros(len(weights), dtype=np.int32)
    for c in order:
        r = int(np.argmin(load))
        owner[c] = r
        load[r] += weights[c]
    return owner


def contig_weights(rb, n_contigs):
    return np.bincount(rb.tid[rb.tid >= 0], weights=rb.n_cigar[rb.tid >= 0].astype(np.float64), minlength=n_contigs)


def shard_records(rb, owner, rank):
    """(sub-batch of the records on this rank's contigs, their indices in the full batch)."""
    mine = np.nonzero((rb.tid >= 0) & (owner[np.maximum(rb.tid, 0)] == rank))[0]
    return rb.subset(mine), mine.astype(np.uint32)


def key_contig(rows):
    """tid of Candidate.get_key() (SVCandidate.py:17-19,147-148,292-293,386-387)."""
    by_dest = (rows["type"] == 2) | (rows["type"] == 4)
    return np.where(by_dest, rows["dst_tid"], rows["src_tid"])


def remap_to_global(rows, global_idx):
    """Local record indices -> indices in the full (unsharded) batch, in aln_idx and in the ordering key."""
    rows = rows.copy()
    g = global_idx[rows["aln_idx"]].astype(np.uint64)
    rows["aln_idx"] = g
    rows["ordinal"] = (g << np.uint64(32)) | (rows["ordinal"] & np.uint64(0xFFFFFFFF))
    return rows


def pack_payload(parts):
    """[(rows, pool, starts), ...] -> one uint8 buffer (header: counts per part).  starts[i] = byte offset of row i's
    sequence run inside `pool`."""
    header = np.zeros(HEADER_WORDS * len(parts), dtype=np.uint64)
    chunks = []
    for k, (rows, pool, starts) in enumerate(parts):
        header[HEADER_WORDS * k:HEADER_WORDS * k + 2] = (rows.shape[0], pool.shape[0])
        chunks += [rows.view(np.uint8).reshape(-1), np.ascontiguousarray(starts, dtype=np.uint64).view(np.uint8).reshape(-1), pool]
    return np.concatenate([header.view(np.uint8)] + chunks)


def unpack_payload(buf, n_parts, row_dtype):
    header = buf[:8 * HEADER_WORDS * n_parts].view(np.uint64)
    pos = 8 * HEADER_WORDS * n_parts
    out = []
    for k in range(n_parts):
        n_rows, n_pool = int(header[HEADER_WORDS * k]), int(header[HEADER_WORDS * k + 1])
        rows = buf[pos:pos + n_rows * row_dtype.itemsize].view(row_dtype)
        pos += n_rows * row_dtype.itemsize
        starts = buf[pos:pos + 8 * n_rows].view(np.uint64)
        pos += 8 * n_rows
        pool = buf[pos:pos + n_pool]
        pos += n_pool
        out.append((rows, pool, starts))
    return out


def all_gather_bytes(buf, device):
    """Variable-length all-gather of a uint8 numpy buffer: sizes first, then one padded all_gather."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    size = torch.tensor([buf.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    sizes = [int(s.item()) for s in sizes]
    width = max(max(sizes), 1)
    mine = torch.zeros(width, dtype=torch.uint8, device=device)
    if buf.shape[0]:
        mine[:buf.shape[0]] = torch.from_numpy(np.ascontiguousarray(buf)).to(device)
    outs = [torch.empty(width, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [o[:s].cpu().numpy() for o, s in zip(outs, sizes)]


def merge_gathered(parts_by_rank, part):
    """One haplotype's (rows, pool, starts) of all ranks -> the global append order (ordinal).  The pools are only
    concatenated; the per-row start offsets move with their rows."""
    rows = np.concatenate([p[part][0] for p in parts_by_rank])
    pools, starts, base = [], [], 0
    for p in parts_by_rank:
        pools.append(p[part][1])
        starts.append(p[part][2].astype(np.uint64) + np.uint64(base))
        base += int(p[part][1].shape[0])
    pool = np.concatenate(pools) if pools else np.zeros(0, dtype=np.uint8)
    starts = np.concatenate(starts) if starts else np.zeros(0, dtype=np.uint64)
    order = np.argsort(rows["ordinal"], kind="stable")
    return rows[order], pool, starts[order]


def select_owned(rows, pool, starts, owner, rank):
    keep = np.nonzero(owner[key_contig(rows)] == rank)[0]
    return rows[keep], pool, starts[keep]


def order_paired(rows_by_rank, lexrank):
    """Final order of pair_candidates' output: type, then key contig in python string order; inside one
    (type, contig) every row comes from the owning rank, already in partition / label order."""
    dtype = rows_by_rank[0].dtype
    # the 64-byte rows are handled as 8 machine words each: concatenating / fancy-indexing the structured dtype is an
    # order of magnitude slower
    words = np.concatenate([np.ascontiguousarray(r).view(np.uint64).reshape(r.shape[0], dtype.itemsize // 8) for r in rows_by_rank])
    rows = words.reshape(-1).view(dtype)
    if rows.shape[0] == 0:
        return rows
    lexrank = np.asarray(lexrank, dtype=np.int64)
    key = rows["type"].astype(np.int64) * lexrank.shape[0] + lexrank[key_contig(rows)]
    if int(key.max()) < 65536:
        key = key.astype(np.uint16)                    # numpy sorts 16-bit keys with a (stable) radix sort
    out = np.take(words, np.argsort(key, kind="stable"), axis=0).reshape(-1).view(dtype)
    out["ordinal"] = np.arange(out.shape[0], dtype=np.uint64)
    return out


def sharded_step(stage, rank, owner, lexrank, device, row_dtype):
    """One step on one rank.  `stage` supplies the compute:
         stage.collect(hap) -> (rows, pool, off) with GLOBAL record indices
         stage.pair(part1, part2) -> paired rows (numpy)
       Returns the complete paired table (identical on every rank)."""
    mine = [stage.collect(1), stage.collect(2)]
    gathered = [unpack_payload(b, 2, row_dtype) for b in all_gather_bytes(pack_payload(mine), device)]
    parts = []
    for hap in (0, 1):
        rows, pool, off = merge_gathered(gathered, hap)
        parts.append(select_owned(rows, pool, off, owner, rank))
    paired = stage.pair(parts[0], parts[1])
    empty = (np.zeros(0, dtype=np.uint8), np.zeros(paired.shape[0], dtype=np.uint64))
    back = [unpack_payload(b, 1, row_dtype)[0][0] for b in all_gather_bytes(pack_payload([(paired, empty[0], empty[1])]), device)]
    return order_paired(back, lexrank)


class EngineStage(object):
    """The compute stages of one rank on its GPU (libsvimasm_b200 through the ctypes engine)."""

    def __init__(self, eng, hosts, global_idx, ref, params, resident):
        self.eng, self.hosts, self.global_idx, self.ref, self.params = eng, hosts, global_idx, ref, params
        self.resident = resident                       # record images (with sequences) kept in HBM, or None: upload per step
        self.records = list(resident) if resident else [None, None]
        self.h2d = 0

    def collect(self, hap):
        k = hap - 1
        if self.resident is None:
            self.records[k] = self.eng.load_records(self.hosts[k])
            h = self.hosts[k]
            self.h2d += sum(getattr(h, n).nbytes for n in ("hdr", "cigar", "seg", "sa_count"))
        table = self.eng.collect(self.records[k], self.params, hap=hap)
        if self.resident is None:
            table.attach_sequences_host(self.hosts[k])
        else:
            table.gather_sequences(self.records[k])
        rows = remap_to_global(table.to_numpy(), self.global_idx[k])
        pool, off = table.pool_to_numpy()
        table.free()
        return rows, pool, off

    def pair(self, part1, part2):
        tables = []
        for rows, pool, starts in (part1, part2):
            t = self.eng.table_from_numpy(rows)
            t.set_pool(pool, starts)
            tables.append(t)
        paired = self.eng.pair(tables[0], tables[1], self.records[0], self.records[1], self.ref, self.params)
        out = paired.to_numpy()
        for t in tables + [paired]:
            t.free()
        if self.resident is None:
            for r in self.records:
                r.free()
        return out


class DeviceStage(object):
    """One rank's compute on its GPU with the exchange kept on the device (csrc/exchange.cu): the tables never visit
    the host between COLLECT and the paired result."""

    def __init__(self, eng, hosts, global_idx, ref, params, resident):
        self.eng, self.hosts, self.global_idx, self.ref, self.params = eng, hosts, global_idx, ref, params
        self.resident = resident                       # record images (sequences + global index) kept in HBM, or None
        self.records = list(resident) if resident else [None, None]
        self.h2d = 0

    def collect(self, hap):
        k = hap - 1
        h = self.hosts[k]
        if self.resident is None:                      # end-to-end leg: this step's records come from pinned host memory
            self.records[k] = self.eng.load_records(h)
            self.eng.set_global_index(self.records[k], self.global_idx[k])
            self.h2d += sum(getattr(h, n).nbytes for n in ("hdr", "cigar", "seg", "sa_count")) + 4 * h.n_aln
        table = self.eng.collect(self.records[k], self.params, hap=hap)
        if self.resident is None:
            table.attach_sequences_host(h)
        else:
            table.gather_sequences(self.records[k])
        table.remap_records(self.records[k])
        return table

    def collect_both(self):
        """Both haplotypes with one host synchronisation (svb_collect2), their sequence pools and global record indices."""
        if self.resident is None:                      # end-to-end leg: this step's records come from pinned host memory
            for k in range(2):
                h = self.hosts[k]
                self.records[k] = self.eng.load_records(h)
                self.eng.set_global_index(self.records[k], self.global_idx[k])
                self.eng.map_sequences_host(self.records[k])
                self.h2d += sum(getattr(h, n).nbytes for n in ("hdr", "cigar", "seg", "sa_count")) + 4 * h.n_aln
        t1, t2 = self.eng.collect2(self.records[0], self.records[1], self.params, with_pools=True)
        t1.remap_records(self.records[0])
        t2.remap_records(self.records[1])
        return t1, t2

    def done(self):
        if self.resident is None:
            for r in self.records:
                r.free()


def _gather_sizes(values, world, device):
    import torch
    import torch.distributed as dist
    mine = torch.from_numpy(np.asarray(values, dtype=np.int64)).to(device)
    out = torch.empty(world * mine.shape[0], dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy().reshape(world, -1)


def sharded_step_device(stage, rank, world, owner, lexrank, device, row_dtype):
    """One step on one rank, NCCL collectives straight on the library's device buffers.  Everything is enqueued on the
    library's own stream (torch sees it as an ExternalStream), so kernels and collectives are stream ordered."""
    import torch
    import torch.distributed as dist
    from .engine import DeviceView
    eng = stage.eng
    prof = _PROFILE if os.environ.get("SVB_SHARD_PROFILE") else None

    def tick(name):
        if prof is not None:              # profiling aid: synchronised phase times of rank 0 (tools/perf_sharded.py)
            eng.synchronize()
            now = time.perf_counter()
            prof.setdefault(name, 0.0)
            prof[name] += now - prof["_t"]
            prof["_t"] = now
    if prof is not None:
        eng.synchronize()
        prof["_t"] = time.perf_counter()
    with torch.cuda.stream(torch.cuda.ExternalStream(eng.stream_handle(), device=device)):
        t1, t2 = stage.collect(1), stage.collect(2)
        tick("collect x2")
        sizes = _gather_sizes(eng.exchange_sizes(t1, t2).astype(np.int64), world, device).astype(np.uint64)
        tick("gather sizes")
        stride = max(eng.exchange_bytes(sizes[r]) for r in range(world))
        mine = torch.empty(stride, dtype=torch.uint8, device=device)
        eng.exchange_pack(t1, t2, mine.data_ptr(), stride)
        tick("pack")
        gathered = torch.empty(world * stride, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(gathered, mine)
        tick("all-gather tables")
        t1.free()
        t2.free()
        u1 = eng.exchange_unpack(gathered.data_ptr(), stride, sizes, 1, owner, rank)
        u2 = eng.exchange_unpack(gathered.data_ptr(), stride, sizes, 2, owner, rank)
        tick("unpack x2")
        paired = eng.pair(u1, u2, stage.records[0], stage.records[1], stage.ref, stage.params)
        tick("pair")
        counts = _gather_sizes([len(paired)], world, device)[:, 0]
        width = max(int(counts.max()), 1) * row_dtype.itemsize
        rows = torch.zeros(width, dtype=torch.uint8, device=device)
        ptr, nbytes = paired.device_rows()
        if nbytes:
            rows[:nbytes].copy_(torch.as_tensor(DeviceView(ptr, nbytes), device=device))
        back = torch.empty(world * width, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(back, rows)
        host = back.cpu().numpy().reshape(world, width)
        tick("gather paired rows")
        for t in (u1, u2, paired):
            t.free()
        stage.done()
    parts = [host[r, :int(counts[r]) * row_dtype.itemsize].view(row_dtype) for r in range(world)]
    out = order_paired(parts, lexrank)
    tick("order on host")
    return out


def sharded_step_window(stage, window, owner, lexrank, download=True):
    """One step on one rank over the peer-memory exchange window (csrc/exchange.cu: svb_exchange_share /
    svb_exchange_gather_paired): no collective library and no torch in the step.  The tables go from this rank's HBM
    straight into every peer's window (one kernel of NVLink stores + a flag), the consumer waits on the device, the paired
    rows travel the same way and are put into pair_candidates' order on the device; one download at the end.
    Returns the complete paired table (identical on every rank); with download=False the table stays in HBM, like the
    result of the one-GPU resident step, and the number of its rows is returned."""
    eng = stage.eng
    prof = _PROFILE if os.environ.get("SVB_SHARD_PROFILE") else None

    def tick(name):
        if prof is not None:              # profiling aid: synchronised phase times (host clock)
            eng.synchronize()
            now = time.perf_counter()
            prof[name] = prof.get(name, 0.0) + now - prof["_t"]
            prof["_t"] = now
    if prof is not None:
        eng.synchronize()
        prof["_t"] = time.perf_counter()
    t1, t2 = stage.collect_both()
    tick("collect x2 (+ pool, remap)")
    u1, u2 = window.share(t1, t2, owner)
    tick("share (put, wait, unpack x2)")
    t1.free()
    t2.free()
    paired = eng.pair(u1, u2, stage.records[0], stage.records[1], stage.ref, stage.params)
    tick("pair")
    everything = window.gather_paired(paired, lexrank)
    tick("gather paired (put, wait, order)")
    out = everything.to_numpy() if download else len(everything)
    for t in (u1, u2, paired, everything):
        t.free()
    stage.done()
    tick("download + free" if download else "free")
    return out


def open_window(eng, rank, world, slot_bytes=None, result_bytes=None):
    """Create this rank's exchange window and map every peer's (handles travel through torch.distributed once, at set-up).
    Returns None when CUDA IPC is not available between the ranks (all ranks agree): the caller falls back to NCCL."""
    import torch
    import torch.distributed as dist
    slot_bytes = slot_bytes or int(os.environ.get("SVB_EXCHANGE_SLOT_MB", "64")) << 20
    result_bytes = result_bytes or int(os.environ.get("SVB_EXCHANGE_RESULT_MB", "16")) << 20
    window, ok = None, 1
    try:
        window = eng.exchange_window(world, rank, slot_bytes, result_bytes)
        handles = [None] * world
        dist.all_gather_object(handles, window.handle)
    except RuntimeError:
        handles, ok = None, 0
        junk = [None] * world
        dist.all_gather_object(junk, b"")
    if ok:
        try:
            window.open(handles)
        except RuntimeError as exc:
            print("[sharded] peer-memory window unavailable on rank %d (%s): falling back to NCCL" % (rank, exc), file=sys.stderr)
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", eng.device))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        if window is not None:
            window.close()
        return None
    return window


_PROFILE = {}


def bench_sharded(args, build_workload, workload_config, peak_hbm, ClockSampler, cpu_baseline_block, build_reference, parity_tools=None):
    """bench.py for WORLD_SIZE > 1 (launched by torchrun): strong scaling, max over ranks."""
    import json
    import torch
    import torch.distributed as dist
    from . import _lib
    from .engine import Engine, HostBatch, lexrank, make_params
    from .bench_util import pinned_host, pin
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    # host memory: every rank generates the whole (deterministic) workload but keeps only its shard; the reference
    # genome is generated afterwards, uploaded from pageable memory and dropped (8 ranks x 3 GB pinned would not fit)
    cfg, rb1, rb2, _b, _o = build_workload(args.scale, with_reference=False)
    n_contigs = len(cfg.contig_names)
    owner = lpt_assign(contig_weights(rb1, n_contigs) + contig_weights(rb2, n_contigs), world)
    s1, g1 = shard_records(rb1, owner, rank)
    s2, g2 = shard_records(rb2, owner, rank)
    n_aln_total, n_ops_total = rb1.n_aln + rb2.n_aln, rb1.n_ops + rb2.n_ops
    # rank 0 keeps the bounded CPU sample (a few contigs, closed under SA tags) for the parity check of the gathered table
    sample = parity_tools[0](rb1, rb2, cfg) if (parity_tools and rank == 0) else None
    del rb1, rb2
    hosts = [pinned_host(HostBatch.from_record_batch(s1)), pinned_host(HostBatch.from_record_batch(s2))]
    del s1, s2
    ranks = lexrank(cfg.contig_names)
    eng = Engine(local)
    params = make_params()
    for turn in range(world):                              # one rank at a time holds the 3.1 GB host copy
        if turn == rank:
            bases, off = build_reference(cfg)
            ref = eng.load_reference(bases, off)           # every rank keeps the whole reference in HBM (3.1 GB of 180 GB)
            if sample is None:
                del bases
        dist.barrier()
    resident = [eng.load_records(h, with_sequences=True) for h in hosts]
    for rec, g in zip(resident, (g1, g2)):
        eng.set_global_index(rec, g)

    window = None if os.environ.get("SVB_EXCHANGE", "window") == "nccl" else open_window(eng, rank, world)
    exchange = "nccl all-gather (torch.distributed)" if window is None else "peer-memory window (CUDA IPC, NVLink stores + device flags)"

    def one_step(stage, download):
        if window is not None:
            return sharded_step_window(stage, window, owner, ranks, download=download)
        return sharded_step_device(stage, rank, world, owner, ranks, device, _lib.ROW_DTYPE)

    def timed(stage_factory, steps, warmup, download):
        one = lambda stage: one_step(stage, download)
        for _ in range(warmup):
            table = one(stage_factory())
        dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()
        # timed on the device: every kernel, copy and collective of the step is enqueued on the library's stream, so two
        # events on that stream bracket the whole loop (host gaps between the enqueues included); max over ranks
        _PROFILE.clear()
        eng.mark(2)
        for _ in range(steps):
            table = one(stage_factory())
        eng.mark(3)
        eng.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([eng.elapsed_ms(2, 3) * 1e-3], dtype=torch.float64, device=device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if os.environ.get("SVB_SHARD_PROFILE") and rank == 0:
            for name, v in _PROFILE.items():
                if not name.startswith("_"):
                    print("[shard profile] %-20s %8.3f ms per step" % (name, v / steps * 1e3), file=sys.stderr)
            print("[shard profile] ---", file=sys.stderr)
        return float(dt.item()) / steps, table

    warm = max(args.warmup, 3)
    eng.timing_reset()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    # `value`: inputs resident, and the gathered, ordered table stays in HBM on every rank (what the one-GPU resident step
    # leaves behind as well); `e2e`: host buffers in, the table downloaded
    sec, table = timed(lambda: DeviceStage(eng, hosts, [g1, g2], ref, params, resident), args.steps, warm, window is None)
    n_rows_resident = int(table) if window is not None else int(table.shape[0])
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count() - launches0
    timing = eng.timing()
    scan_ms, scan_launches = timing["cigar_scan"]
    stages = []

    def e2e_factory():
        st = DeviceStage(eng, hosts, [g1, g2], ref, params, None)
        stages.append(st)
        return st
    e2e_sec, table2 = timed(e2e_factory, args.steps, warm, True)
    h2d = torch.tensor([float(stages[-1].h2d)], dtype=torch.float64, device=device)
    dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
    assert n_rows_resident == table2.shape[0]
    parity = None
    if sample is not None:
        # the gathered, ordered table of the sharded run against the oracle on the CPU sample (outside the timed region)
        s1, s2, tids, idx1, idx2 = sample
        _dt, _na, _no, want = parity_tools[1](s1, s2, bases, off)
        parity = parity_tools[2](table2, (want, tids, idx1, idx2))
        del bases
    if rank == 0:
        peak, peak_src = peak_hbm()
        local_ops, local_aln = hosts[0].n_ops + hosts[1].n_ops, hosts[0].n_aln + hosts[1].n_aln
        scan_avg = scan_ms / max(scan_launches, 1)
        alg = (4.0 * local_ops + 32.0 * local_aln) / 2.0
        achieved = alg / (scan_avg * 1e-3) / 1e9 if scan_avg > 0 else 0.0
        # DRAM bytes per launch: the ncu capture of the full-size launch (profiles/) measured 1.016 x the algorithmic bytes;
        # a shard runs the same kernel on fewer units, so the shard's figure is that ratio times its algorithmic bytes
        traffic, traffic_src = None, None
        try:
            cap = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_cigar_scan_traffic.json")))
            ratio = cap["cigar_scan_dram_bytes_per_launch"] / (4.0 * cap["n_ops"] + 32.0 * cap.get("n_aln", 0) + 64.0 * cap.get("n_rows", 0))
            traffic = ratio * alg
            traffic_src = "profiles/r2_cigar_scan_traffic.json: measured DRAM bytes / algorithmic bytes of the full-size launch (%.3f) x this shard's algorithmic bytes" % ratio
        except Exception:
            pass
        print(json.dumps({
            "metric": "alignments_per_sec", "value": n_aln_total / sec, "unit": "alignments/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "cigar_ops_per_sec": n_ops_total / sec,
            "config": workload_config(cfg, args, paired_rows=n_rows_resident,
                                      shard="rank 0 holds %d of %d alignments" % (local_aln, n_aln_total), exchange=exchange),
            "roofline": {"kernel": "cigar_scan (rank 0 shard)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "launch_ms": scan_avg, "algorithmic_bytes_per_launch": alg},
            "e2e": {"value": n_aln_total / e2e_sec, "unit": "alignments/s", "h2d_bytes_per_step": int(h2d.item()),
                    "d2h_bytes_per_step": int(table2.nbytes), "ms_per_step": e2e_sec * 1e3},
            "parity_check": parity, "gpu_launches": int(launches), "clocks": clocks,
        }), flush=True)
    if window is not None:
        window.close()
    dist.destroy_process_group()
