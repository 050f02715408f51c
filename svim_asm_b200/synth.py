"""Seeded synthetic genome-genome alignments.

Produces `RecordBatch` objects: the flat, 16-byte-aligned host layout that the
C-ABI (`include/svimasm_b200.h`, `svb_load_records`) consumes, i.e. exactly
what the BAM ingest produces from a file.  `bamio.write_bam` can turn a batch
into a real BGZF BAM so that file-based tools (and the reference under the
oracle shims) read the very same alignments.

The generator works in "reference space": every alignment covers a fixed
reference interval; small noise events (X / short I / short D) form a renewal
process along it and structural variants (the truth list, shared between the
two haplotypes of a diploid sample) are merged in by coordinate, so that the
same SV gets the same reference position in both haplotypes.
All randomness comes from numpy.random.Generator(PCG64(seed)).
"""
import os
from dataclasses import dataclass, field

import numpy as np

OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X = range(9)
OP_PAD = 15                      # filler used to pad every CIGAR run to 4 ops (16 bytes)
NT16 = "=ACMGRSVTWYHKDBN"
_ACGT_CODE = np.array([1, 2, 4, 8], dtype=np.uint8)
_ACGT_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)

_GAP_LUT = (1 + np.floor(np.log1p(-(np.arange(65536) + 0.5) / 65536.0) / np.log(1.0 - 1.0 / 12.0))).astype(np.uint16)
_KIND_LUT = np.where(np.arange(256) < 154, 8, np.where(np.arange(256) < 205, 1, 2)).astype(np.int8)      # X / I / D
_LEN_LUT = np.array([1 + (bin(b ^ (b + 1)).count("1") - 1 if b != 255 else 8) for b in range(256)], dtype=np.uint8)

HG38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
                138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
                83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
HG38_NAMES = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"]


@dataclass
class RecordBatch:
    """Flat host-side image of one BAM file (one haplotype)."""
    contig_names: list
    contig_lengths: np.ndarray          # int32[n_contig]
    tid: np.ndarray                     # int32[n]
    pos: np.ndarray                     # int32[n]  0-based reference_start
    flag: np.ndarray                    # uint16[n]
    mapq: np.ndarray                    # uint8[n]
    n_cigar: np.ndarray                 # uint32[n] real op count
    cigar_off: np.ndarray               # uint64[n+1] op index of each run (multiple of 4)
    l_seq: np.ndarray                   # uint32[n]
    seq_off: np.ndarray                 # uint64[n+1] byte offset into seq4
    cigar: np.ndarray                   # uint32[cigar_off[-1]] BAM-packed (len<<4|op), padded with op 15
    seq4: np.ndarray                    # uint8 4-bit packed query bases, high nibble first
    names: list                         # query names
    sa: dict = field(default_factory=dict)     # record index -> raw SA:Z text

    @property
    def n_aln(self):
        return int(self.tid.shape[0])

    @property
    def n_ops(self):
        return int(self.n_cigar.sum())

    def cigartuples(self, i):
        lo = int(self.cigar_off[i])
        run = self.cigar[lo:lo + int(self.n_cigar[i])]
        return [(int(v & 15), int(v >> 4)) for v in run]

    def query_sequence(self, i):
        n = int(self.l_seq[i])
        if n == 0:
            return None
        lo = int(self.seq_off[i])
        raw = self.seq4[lo:lo + (n + 1) // 2]
        nib = np.empty(raw.shape[0] * 2, dtype=np.uint8)
        nib[0::2] = raw >> 4
        nib[1::2] = raw & 15
        table = np.frombuffer(NT16.encode(), dtype=np.uint8)
        return table[nib[:n]].tobytes().decode("ascii")

    def subset(self, idx):
        """New batch holding records `idx` (ascending) -- used for bounded CPU samples and rank shards."""
        idx = np.asarray(idx, dtype=np.int64)
        n_c = self.n_cigar[idx].astype(np.int64)
        padded = (n_c + 3) // 4 * 4
        off = np.zeros(idx.shape[0] + 1, dtype=np.uint64)
        off[1:] = np.cumsum(padded)
        cigar = np.full(int(off[-1]), OP_PAD, dtype=np.uint32)
        nbytes = (self.l_seq[idx].astype(np.int64) + 1) // 2
        soff = np.zeros(idx.shape[0] + 1, dtype=np.uint64)
        soff[1:] = np.cumsum(nbytes)
        seq4 = np.zeros(int(soff[-1]), dtype=np.uint8)
        for k, i in enumerate(idx):
            lo = int(self.cigar_off[i])
            cigar[int(off[k]):int(off[k]) + int(n_c[k])] = self.cigar[lo:lo + int(n_c[k])]
            so = int(self.seq_off[i])
            seq4[int(soff[k]):int(soff[k + 1])] = self.seq4[so:so + int(nbytes[k])]
        remap = {int(i): k for k, i in enumerate(idx)}
        return RecordBatch(self.contig_names, self.contig_lengths, self.tid[idx].copy(), self.pos[idx].copy(),
                           self.flag[idx].copy(), self.mapq[idx].copy(), self.n_cigar[idx].copy(), off,
                           self.l_seq[idx].copy(), soff, cigar, seq4, [self.names[int(i)] for i in idx],
                           {remap[i]: s for i, s in self.sa.items() if i in remap})


@dataclass
class SynthConfig:
    contig_names: list
    contig_lengths: list
    n_aln: int
    target_ops: float
    seed: int
    sv_per_event: float = 2.5e-4         # probability that an indel event is an SV (>=40 bp)
    sv_min: int = 40
    sv_max: int = 10000
    split_fraction: float = 0.02         # primaries that carry an SA tag
    low_mapq_fraction: float = 0.01
    secondary_fraction: float = 0.005
    mstyle_fraction: float = 0.2         # alignments written with M instead of =/X
    giant_ops: int = 0                   # force the first alignment to have about this many ops (CG-tag path)
    lognorm_sigma: float = 1.0
    with_sequence: bool = True


def config_c1(seed=1001):
    return SynthConfig(["chr1"], [1_000_000], 40, 5e4, seed, sv_per_event=4e-3, split_fraction=0.25)


def config_c2(seed=1002):
    return SynthConfig(["chr20"], [64_444_167], 2000, 1.0e7, seed, giant_ops=200_000)


def config_c3(seed=1003, scale=1.0):
    return SynthConfig(list(HG38_NAMES), list(HG38_LENGTHS), max(24, int(40000 * scale)), 2.0e8 * scale, seed,
                       giant_ops=int(1_000_000 * min(1.0, scale * 4)))


_ACGT_QUADS = np.array([[_ACGT_ASCII[(b >> s) & 3] for s in (0, 2, 4, 6)] for b in range(256)], dtype=np.uint8).view(np.uint32).reshape(256)


def random_reference(cfg, seed=None):
    """dict name -> uint8 ASCII array (uniform ACGT); one random byte yields four bases."""
    rng = np.random.Generator(np.random.PCG64((cfg.seed if seed is None else seed) ^ 0x5EED))
    ref = {}
    for name, length in zip(cfg.contig_names, cfg.contig_lengths):
        raw = np.frombuffer(rng.bytes((int(length) + 3) // 4), dtype=np.uint8)
        ref[name] = _ACGT_QUADS[raw].view(np.uint8)[:int(length)]
    return ref


# ----------------------------------------------------------------------------------------------
# layout shared by both haplotypes


@dataclass
class Layout:
    tid: np.ndarray            # int32[n] contig of each alignment
    pos: np.ndarray            # int64[n] reference start
    span: np.ndarray           # int64[n] reference length
    bound: np.ndarray          # int64[n+1] start of each alignment in concatenated "alignment space"


def make_layout(cfg):
    rng = np.random.Generator(np.random.PCG64(cfg.seed))
    n = cfg.n_aln
    weights = rng.lognormal(0.0, cfg.lognorm_sigma, n)
    ops_per_ref = 2.0 / 13.3                              # two ops per (match run + event)
    total_ref = cfg.target_ops / ops_per_ref
    if cfg.giant_ops:
        giant_ref = cfg.giant_ops / ops_per_ref
        weights[0] = 0.0
        span = weights / weights.sum() * max(total_ref - giant_ref, total_ref * 0.5)
        span[0] = giant_ref
    else:
        span = weights / weights.sum() * total_ref
    span = np.maximum(span.astype(np.int64), 200)
    lengths = np.asarray(cfg.contig_lengths, dtype=np.int64)
    # the giant alignment goes to the largest contig; the rest are dealt out in proportion to contig length
    tid = rng.choice(len(lengths), size=n, p=lengths / lengths.sum()).astype(np.int32)
    tid[0] = int(np.argmax(lengths))
    span = np.minimum(span, lengths[tid] - 2)
    order = np.lexsort((rng.random(n), tid))
    tid, span = tid[order], span[order]
    pos = np.zeros(n, dtype=np.int64)
    for c in range(len(lengths)):
        sel = np.nonzero(tid == c)[0]
        if sel.size == 0:
            continue
        cum = np.concatenate(([0], np.cumsum(span[sel])[:-1]))
        room = lengths[c] - span[sel].max() - 1
        scale = min(1.0, room / max(1, cum[-1])) if cum[-1] > 0 else 1.0
        slack = max(0, room - int(cum[-1] * scale))
        jitter = np.sort(rng.integers(0, slack + 1, sel.size))
        pos[sel] = np.minimum((cum * scale).astype(np.int64) + jitter, lengths[c] - span[sel])
    order = np.lexsort((pos, tid))
    tid, span, pos = tid[order], span[order], pos[order]
    bound = np.zeros(n + 1, dtype=np.int64)
    bound[1:] = np.cumsum(span)
    return Layout(tid, pos, span, bound)


@dataclass
class SVTruth:
    gpos: np.ndarray      # int64 position in alignment space
    kind: np.ndarray      # OP_I / OP_D
    length: np.ndarray    # int64
    sv_id: np.ndarray     # int64 identity (drives the inserted sequence)
    edit: np.ndarray      # float fraction of substituted bases of the inserted sequence


def make_truth(cfg, layout, seed, n_sv=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    total = int(layout.bound[-1])
    if n_sv is None:
        n_sv = int(total / 13.3 * 0.4 * cfg.sv_per_event) + 1
    gpos = np.sort(rng.integers(0, total, n_sv))
    kind = np.where(rng.random(n_sv) < 0.5, OP_I, OP_D)
    length = np.exp(rng.uniform(np.log(cfg.sv_min), np.log(cfg.sv_max), n_sv)).astype(np.int64)
    return SVTruth(gpos, kind.astype(np.int8), length, rng.integers(1, 2 ** 40, n_sv), np.zeros(n_sv))


def diploid_truth(cfg, layout):
    """Truth for the two haplotypes: 60 % identical, 10 % shifted/edited (still pair), 30 % private each."""
    base = make_truth(cfg, layout, cfg.seed + 11)
    rng = np.random.Generator(np.random.PCG64(cfg.seed + 12))
    u = rng.random(base.gpos.shape[0])
    shared = u < 0.6
    near = (u >= 0.6) & (u < 0.7)
    only1 = (u >= 0.7) & (u < 0.85)
    only2 = u >= 0.85
    h1 = shared | near | only1
    h2 = shared | near | only2

    def take(mask, jitter):
        t = SVTruth(base.gpos[mask].copy(), base.kind[mask].copy(), base.length[mask].copy(),
                    base.sv_id[mask].copy(), np.zeros(int(mask.sum())))
        if jitter:
            sel = near[mask]
            t.gpos[sel] += rng.integers(-20, 21, int(sel.sum()))
            t.edit[sel] = rng.uniform(0.0, 0.05, int(sel.sum()))
            order = np.argsort(t.gpos, kind="stable")
            t = SVTruth(t.gpos[order], t.kind[order], t.length[order], t.sv_id[order], t.edit[order])
        return t

    return take(h1, False), take(h2, True)


def sv_sequence(sv_id, length, edit=0.0):
    """Deterministic inserted sequence (nt16 codes) of an SV; `edit` substitutes that fraction of bases."""
    rng = np.random.Generator(np.random.PCG64(int(sv_id)))
    codes = _ACGT_CODE[rng.integers(0, 4, int(length))]
    if edit > 0.0:
        erng = np.random.Generator(np.random.PCG64(int(sv_id) ^ 0xED17))
        hit = erng.random(int(length)) < edit
        codes = np.where(hit, _ACGT_CODE[erng.integers(0, 4, int(length))], codes)
    return codes


def _write_nibbles(seq4, nib_off, codes):
    """Overwrite nibbles [nib_off, nib_off+len) of a 4-bit packed array (high nibble first)."""
    n = codes.shape[0]
    if n == 0:
        return
    b0, b1 = nib_off // 2, (nib_off + n + 1) // 2
    chunk = seq4[b0:b1]
    nib = np.empty(chunk.shape[0] * 2, dtype=np.uint8)
    nib[0::2] = chunk >> 4
    nib[1::2] = chunk & 15
    s = nib_off - 2 * b0
    nib[s:s + n] = codes
    seq4[b0:b1] = (nib[0::2] << 4) | nib[1::2]


# ----------------------------------------------------------------------------------------------
# one haplotype


N_GROUPS = 16          # fixed, so that the data do not depend on the number of worker processes


def make_haplotype(cfg, layout, truth, hap_seed, name_prefix="ctg", workers=None):
    """One haplotype.  Large configurations are generated as N_GROUPS independent slices of the alignment
    list (own PCG64 stream each) in a process pool and concatenated."""
    n = layout.tid.shape[0]
    if cfg.target_ops < 2e7 or n < 16 * N_GROUPS:
        batch, supp, rng = _make_group((cfg, layout, truth, hap_seed, name_prefix, 0))
        return _append_records(batch, supp, rng) if supp else batch
    cuts = np.searchsorted(layout.bound, np.linspace(0, layout.bound[-1], N_GROUPS + 1)[1:-1])
    cuts = np.unique(np.concatenate(([0], cuts, [n])))
    jobs = []
    for g in range(cuts.shape[0] - 1):
        a, b = int(cuts[g]), int(cuts[g + 1])
        base = layout.bound[a]
        sub = Layout(layout.tid[a:b], layout.pos[a:b], layout.span[a:b], layout.bound[a:b + 1] - base)
        sel = (truth.gpos >= base) & (truth.gpos < layout.bound[b])
        sub_truth = SVTruth(truth.gpos[sel] - base, truth.kind[sel], truth.length[sel], truth.sv_id[sel], truth.edit[sel])
        jobs.append((cfg, sub, sub_truth, hap_seed + 1000003 * (g + 1), name_prefix, a))
    if workers is None:
        workers = min(len(jobs), int(os.environ.get("SVIM_SYNTH_WORKERS", os.cpu_count() or 1)))
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_make_group, jobs)
    else:
        parts = [_make_group(j) for j in jobs]
    batches = [p[0] for p in parts]
    supp = [r for p in parts for r in p[1]]
    n_c = np.concatenate([b.n_cigar for b in batches])
    l_seq = np.concatenate([b.l_seq for b in batches])
    c_off = np.zeros(n + 1, dtype=np.uint64)
    c_off[1:] = np.cumsum((n_c.astype(np.int64) + 3) // 4 * 4)
    s_off = np.zeros(n + 1, dtype=np.uint64)
    s_off[1:] = np.cumsum((l_seq.astype(np.int64) + 1) // 2)
    sa = {}
    names = []
    for (cfg_, sub, tr, seed, pre, first), b in zip(jobs, batches):
        sa.update({i + first: t for i, t in b.sa.items()})
        names.extend(b.names)
    whole = RecordBatch(list(cfg.contig_names), np.asarray(cfg.contig_lengths, dtype=np.int32),
                        np.concatenate([b.tid for b in batches]), np.concatenate([b.pos for b in batches]),
                        np.concatenate([b.flag for b in batches]), np.concatenate([b.mapq for b in batches]), n_c, c_off,
                        l_seq, s_off, np.concatenate([b.cigar for b in batches]),
                        np.concatenate([b.seq4 for b in batches]), names, sa)
    rng = np.random.Generator(np.random.PCG64(hap_seed ^ 0xABCDEF))
    return _append_records(whole, supp, rng) if supp else whole


def _make_group(job):
    cfg, layout, truth, hap_seed, name_prefix, first_index = job
    rng = np.random.Generator(np.random.PCG64(hap_seed))
    n = layout.tid.shape[0]
    bound = layout.bound
    total = int(bound[-1])

    # ---- noise events as a renewal process over alignment space.  One 32-bit draw per event feeds three
    # inverse-CDF tables: gap ~ geometric(1/12) (16 bits), kind X/I/D = 60/20/20 % (8 bits), short indel
    # length ~ geometric(1/2) (8 bits)
    n_guess = int(total / 13.0 * 1.02) + 1024
    raw = rng.integers(0, 2 ** 32, n_guess, dtype=np.uint32)
    gaps = _GAP_LUT[raw & 0xFFFF].astype(np.int64)
    kind = _KIND_LUT[(raw >> 16) & 0xFF]
    length = _LEN_LUT[raw >> 24].astype(np.int64)
    del raw
    length[kind == OP_X] = 1
    rspan = np.where(kind == OP_I, 0, length)
    start = np.cumsum(gaps + rspan) - rspan           # event start = previous end + gap
    del gaps
    keep = start < total - 2
    # ---- structural variants: clear the noise around each one
    tpos = truth.gpos.astype(np.int64)
    tspan = np.where(truth.kind == OP_I, 0, truth.length)
    ok = np.ones(tpos.shape[0], dtype=bool)
    if tpos.shape[0] > 1:                              # SVs must not touch each other
        prev_end = np.maximum.accumulate(tpos + tspan)
        ok[1:] = tpos[1:] > prev_end[:-1] + 60
    a_sv = np.searchsorted(bound, tpos, side="right") - 1
    ok &= (tpos > bound[a_sv] + 50) & (tpos + tspan < bound[np.minimum(a_sv + 1, n)] - 50)
    tpos, tspan = tpos[ok], tspan[ok]
    tkind, tlen = truth.kind[ok], truth.length[ok]
    tid_sv, tedit = truth.sv_id[ok], truth.edit[ok]
    lo = np.searchsorted(start, tpos - 42, side="left")
    hi = np.searchsorted(start, tpos + tspan + 2, side="right")
    marks = (np.bincount(lo, minlength=start.shape[0] + 1) - np.bincount(hi, minlength=start.shape[0] + 1)).astype(np.int32)
    keep &= np.cumsum(marks[:-1]) == 0
    del marks
    start, kind, length, rspan = start[keep], kind[keep], length[keep], rspan[keep]
    del keep
    # ---- alignment boundaries: an event needs a match run on both sides inside its alignment
    aln = (np.searchsorted(bound, start, side="right") - 1).astype(np.int64)
    inside = (start > bound[aln]) & (start + rspan < bound[aln + 1])
    start, kind, length, rspan, aln = start[inside], kind[inside], length[inside], rspan[inside], aln[inside]
    del inside
    # ---- merge the SVs in by coordinate
    where = np.searchsorted(start, tpos)
    is_sv = np.insert(np.zeros(start.shape[0], dtype=bool), where, True)
    kind = np.insert(kind, where, tkind)
    length = np.insert(length, where, tlen)
    rspan = np.insert(rspan, where, tspan)
    aln = np.insert(aln, where, np.searchsorted(bound, tpos, side="right") - 1)
    start = np.insert(start, where, tpos)
    n_ev = start.shape[0]

    # ---- match run before every event, trailing run per alignment
    end = start + rspan
    first = np.ones(n_ev, dtype=bool)
    first[1:] = aln[1:] != aln[:-1]
    prev_end = np.empty(n_ev, dtype=np.int64)
    prev_end[1:] = end[:-1]
    prev_end[first] = bound[aln[first]]
    match = start - prev_end
    assert match.min() >= 1, "generator produced an empty match run"
    count = np.bincount(aln, minlength=n).astype(np.int64)
    last_end = bound[:-1].copy()
    last_idx = np.nonzero(np.append(first[1:], True))[0]
    last_end[aln[last_idx]] = end[last_idx]
    tail = bound[1:] - last_end
    assert tail.min() >= 1

    # ---- record attributes
    flag = np.where(rng.random(n) < 0.5, 16, 0).astype(np.uint16)
    mapq = np.full(n, 60, dtype=np.uint8)
    low = rng.random(n) < cfg.low_mapq_fraction
    mapq[low] = rng.integers(0, 20, int(low.sum()))
    flag[rng.random(n) < cfg.secondary_fraction] |= 256
    mstyle = rng.random(n) < cfg.mstyle_fraction
    lead = np.where(rng.random(n) < 0.3, rng.integers(1, 500, n), 0).astype(np.int64)
    trail = np.where(rng.random(n) < 0.3, rng.integers(1, 500, n), 0).astype(np.int64)
    hard = np.zeros(n, dtype=bool)

    # ---- split reads: extra segments + SA text (python loop, a few hundred reads)
    read_adv = np.where(kind == OP_D, 0, length)
    q_body = (np.bincount(aln, weights=(match + read_adv).astype(np.float64), minlength=n).astype(np.int64) + tail)
    eligible = np.nonzero((flag & 256) == 0)[0]
    n_split = int(round(cfg.split_fraction * n))
    split_idx = np.sort(rng.choice(eligible, size=min(n_split, eligible.size), replace=False)) if n_split else []
    sa_text, supp = {}, []
    lengths = np.asarray(cfg.contig_lengths, dtype=np.int64)
    for i in split_idx:
        i = int(i)
        segs, li, ti, is_hard = _split_read(rng, cfg, lengths, int(layout.tid[i]), int(layout.pos[i]),
                                            int(layout.span[i]), bool(flag[i] & 16), int(q_body[i]))
        lead[i], trail[i], hard[i] = li, ti, is_hard
        sa_text[i] = segs["sa"]
        supp.extend(segs["records"])

    # ---- assemble the flat CIGAR array
    has_lead = lead > 0
    has_trail = trail > 0
    n_cigar = has_lead.astype(np.int64) + 2 * count + 1 + has_trail
    padded = (n_cigar + 3) // 4 * 4
    cigar_off = np.zeros(n + 1, dtype=np.int64)
    cigar_off[1:] = np.cumsum(padded)
    cigar = np.full(int(cigar_off[-1]), OP_PAD, dtype=np.uint32)
    first_ev = np.zeros(n + 1, dtype=np.int64)
    first_ev[1:] = np.cumsum(count)
    rank = np.arange(n_ev, dtype=np.int64) - first_ev[aln]
    slot = cigar_off[aln] + has_lead[aln] + 2 * rank
    mop = np.where(mstyle[aln], OP_M, OP_EQ).astype(np.uint32)
    cigar[slot] = (match.astype(np.uint32) << 4) | mop
    eop = np.where((kind == OP_X) & mstyle[aln], OP_M, kind).astype(np.uint32)
    cigar[slot + 1] = (length.astype(np.uint32) << 4) | eop
    tslot = cigar_off[:-1] + has_lead + 2 * count
    cigar[tslot] = (tail.astype(np.uint32) << 4) | np.where(mstyle, OP_M, OP_EQ).astype(np.uint32)
    clip = np.where(hard, OP_H, OP_S).astype(np.uint32)
    cigar[cigar_off[:-1][has_lead]] = (lead[has_lead].astype(np.uint32) << 4) | clip[has_lead]
    cigar[(tslot + 1)[has_trail]] = (trail[has_trail].astype(np.uint32) << 4) | clip[has_trail]

    # ---- query sequences (4-bit) with the SV insertions written at their read offsets
    l_seq = q_body + np.where(hard, 0, lead + trail)
    nbytes = (l_seq + 1) // 2
    seq_off = np.zeros(n + 1, dtype=np.int64)
    seq_off[1:] = np.cumsum(nbytes)
    if cfg.with_sequence:
        raw = np.frombuffer(rng.bytes(int(seq_off[-1])), dtype=np.uint8)
        seq4 = (_ACGT_CODE[raw & 3] << 4) | _ACGT_CODE[(raw >> 2) & 3]
        del raw
        adv = match + read_adv
        cum = np.cumsum(adv)
        base = np.zeros(n + 1, dtype=np.int64)
        base[1:] = np.where(first_ev[1:] > 0, cum[np.maximum(first_ev[1:] - 1, 0)], 0)
        qpos = cum - read_adv - base[aln] + np.where(hard[aln], 0, lead[aln])   # read offset of each event
        sv_ins = np.nonzero(is_sv & (kind == OP_I))[0]
        sv_order = np.cumsum(is_sv) - 1
        for e in sv_ins:
            k = int(sv_order[e])
            codes = sv_sequence(tid_sv[k], tlen[k], float(tedit[k]))
            _write_nibbles(seq4, 2 * int(seq_off[aln[e]]) + int(qpos[e]), codes)
    else:
        seq4 = np.zeros(int(seq_off[-1]), dtype=np.uint8)

    names = ["%s%06d" % (name_prefix, first_index + i) for i in range(n)]
    batch = RecordBatch(list(cfg.contig_names), np.asarray(cfg.contig_lengths, dtype=np.int32),
                        layout.tid.astype(np.int32), layout.pos.astype(np.int32), flag, mapq,
                        n_cigar.astype(np.uint32), cigar_off.astype(np.uint64), l_seq.astype(np.uint32),
                        seq_off.astype(np.uint64), cigar, seq4, names, sa_text)
    return batch, supp, rng


# ----------------------------------------------------------------------------------------------
# split reads


_RELATIONS = ("ins", "del", "bigdel", "tandem", "tandem_far", "tandem_huge", "inv", "inv_far", "interchr",
              "interchr_flip", "overlap_read", "gap_read", "dupint", "random")


def _split_read(rng, cfg, lengths, tid, pos, span, rev, q_body):
    """Segments around one primary.  Returns SA text, supplementary records and the primary's clips."""
    n_contig = lengths.shape[0]
    k_right = int(rng.integers(0, 4))
    k_left = int(rng.integers(0, 3)) if k_right else int(rng.integers(1, 3))
    # segments in read order of the FORWARD read: dicts with q0,q1,tid,r0,r1,rev
    prim = dict(q0=0, q1=q_body, tid=tid, r0=pos, r1=pos + span, rev=rev, prim=True)

    def neighbour(cur, side):
        rel = _RELATIONS[int(rng.integers(0, len(_RELATIONS)))]
        qlen = int(rng.integers(300, 6000))
        dr = int(rng.choice([0, 0, 0, 5, -5, 30, -30, 49, 50, 51, -49, -50, -51, 200]))
        new = dict(tid=cur["tid"], rev=cur["rev"], prim=False)
        dref = int(rng.choice([0, 0, 3, -3, 40, -40, 50, 51, -50, -51, 60]))
        big = int(rng.integers(cfg.sv_min, 3000))
        huge = int(rng.integers(100001, 400000))
        if rel == "ins":
            dr = big + dref if rng.random() < 0.8 else dr
        elif rel == "del":
            dref = big
            dr = int(rng.choice([0, 10, 50, 51]))
        elif rel == "bigdel":
            dref = huge
            dr = int(rng.choice([0, 10, 50]))
        elif rel == "tandem":
            dref = -int(rng.integers(60, 5000))
        elif rel == "tandem_far":
            dref = -int(rng.integers(8000, 90000))
        elif rel == "tandem_huge":
            dref = -huge
        elif rel in ("inv", "inv_far"):
            new["rev"] = not cur["rev"]
            dref = big if rel == "inv" else huge
            if rng.random() < 0.5:
                dref = -dref
        elif rel in ("interchr", "interchr_flip", "dupint") and n_contig > 1:
            new["tid"] = int((cur["tid"] + 1 + rng.integers(0, n_contig - 1)) % n_contig)
            if rel == "interchr_flip":
                new["rev"] = not cur["rev"]
        elif rel == "overlap_read":
            dr = -int(rng.integers(51, 300))
        elif rel == "gap_read":
            dr = int(rng.integers(51, 2000))
        elif rel == "random":
            new["rev"] = bool(rng.random() < 0.5)
            dref = int(rng.integers(-200000, 200000))
            dr = int(rng.integers(-80, 120))
        # place on the read
        if side > 0:
            new["q0"] = cur["q1"] + dr
            new["q1"] = new["q0"] + qlen
        else:
            new["q1"] = cur["q0"] - dr
            new["q0"] = new["q1"] - qlen
        rlen = qlen + int(rng.choice([0, 0, 0, 7, -7]))
        clen = int(lengths[new["tid"]])
        if new["tid"] != cur["tid"]:
            r0 = int(rng.integers(0, max(1, clen - rlen)))
        else:
            # forward reads walk up the reference, reverse reads walk down it (side flips that again)
            forward_like = (not cur["rev"]) == (side > 0)
            if new["rev"] != cur["rev"]:
                r0 = (cur["r1"] + dref) if rng.random() < 0.5 else (cur["r0"] - dref - rlen)
            elif forward_like:
                r0 = cur["r1"] + dref
            else:
                r0 = cur["r0"] - dref - rlen
        r0 = int(min(max(0, r0), max(0, clen - rlen - 1)))
        new["r0"], new["r1"] = r0, r0 + rlen
        return new

    chain = [prim]
    cur = prim
    for _ in range(k_right):
        cur = neighbour(cur, +1)
        chain.append(cur)
    cur = prim
    for _ in range(k_left):
        cur = neighbour(cur, -1)
        chain.insert(0, cur)
    # occasionally replay an earlier segment position to provoke tandem runs / interspersed duplications
    if len(chain) >= 3 and rng.random() < 0.5:
        src = chain[0] if chain[0] is not prim else chain[-1]
        extra = dict(src)
        extra["prim"] = False
        shift = int(rng.integers(-15, 16))
        extra["r0"], extra["r1"] = max(0, src["r0"] + shift), max(1, src["r1"] + shift)
        qlen = src["q1"] - src["q0"]
        extra["q0"] = chain[-1]["q1"] + int(rng.integers(0, 40))
        extra["q1"] = extra["q0"] + qlen
        chain.append(extra)
    q_min = min(s["q0"] for s in chain)
    for s in chain:
        s["q0"] -= q_min
        s["q1"] -= q_min
    read_len = max(s["q1"] for s in chain) + int(rng.integers(0, 30))
    p = next(s for s in chain if s["prim"])
    # clips of the primary in ITS OWN orientation (reverse records store the reversed read)
    p_lead, p_trail = (p["q0"], read_len - p["q1"]) if not rev else (read_len - p["q1"], p["q0"])
    is_hard = bool(rng.random() < 0.06)
    entries, records = [], []
    for s in chain:
        if s["prim"]:
            continue
        qlen, rlen = s["q1"] - s["q0"], s["r1"] - s["r0"]
        a_lead, a_trail = (s["q0"], read_len - s["q1"]) if not s["rev"] else (read_len - s["q1"], s["q0"])
        body = []
        common = min(qlen, rlen)
        half = common // 2
        body.append((OP_M, half))
        if rlen > qlen:
            body.append((OP_D, rlen - qlen))
        elif qlen > rlen:
            body.append((OP_I, qlen - rlen))
        body.append((OP_M, common - half))
        ops = ([(OP_S, a_lead)] if a_lead else []) + body + ([(OP_S, a_trail)] if a_trail else [])
        text = "".join("%d%s" % (ln, "MIDNSHP=X"[op]) for op, ln in ops)
        mq = 60 if rng.random() < 0.85 else int(rng.integers(0, 20))
        roll = rng.random()
        name = cfg.contig_names[s["tid"]]
        if roll < 0.03:
            entries.append("%s,%d,%s,%s,%d,%d,extra" % (name, s["r0"] + 1, "-" if s["rev"] else "+", text, mq, 7))
            continue                                   # 7 fields: the reference skips it (SVIM_COLLECT.py:22)
        if roll < 0.06:
            mq_txt = str(int(rng.choice([-400, 300, 256])))   # out of uint8 range -> 0 (SVIM_COLLECT.py:42-45)
        else:
            mq_txt = str(mq)
        entries.append("%s,%d,%s,%s,%s,%d" % (name, s["r0"] + 1, "-" if s["rev"] else "+", text, mq_txt, 7))
        hops = [(OP_H if op == OP_S else op, ln) for op, ln in ops]
        records.append(dict(tid=s["tid"], pos=s["r0"], flag=2048 | (16 if s["rev"] else 0), mapq=mq, ops=hops,
                            l_seq=qlen))
    sa = ";".join(entries) + (";" if entries else "")
    return dict(sa=sa, records=records), p_lead, p_trail, is_hard


def _append_records(batch, extra, rng):
    """Insert supplementary records at their coordinate-sorted places (stable on (tid, pos))."""
    n0 = batch.n_aln
    m = len(extra)
    e_tid = np.array([r["tid"] for r in extra], dtype=np.int64)
    e_pos = np.array([r["pos"] for r in extra], dtype=np.int64)
    e_order = np.lexsort((e_pos, e_tid))
    extra = [extra[int(i)] for i in e_order]
    e_key = e_tid[e_order] * (1 << 32) + e_pos[e_order]
    o_key = batch.tid.astype(np.int64) * (1 << 32) + batch.pos.astype(np.int64)
    where = np.searchsorted(o_key, e_key, side="right")          # originals first on ties
    # per-record tables
    def ins(arr, vals, dtype):
        return np.insert(arr, where, np.asarray(vals, dtype=dtype))
    tid = ins(batch.tid, [r["tid"] for r in extra], np.int32)
    pos = ins(batch.pos, [r["pos"] for r in extra], np.int32)
    flag = ins(batch.flag, [r["flag"] for r in extra], np.uint16)
    mapq = ins(batch.mapq, [r["mapq"] for r in extra], np.uint8)
    n_c = ins(batch.n_cigar, [len(r["ops"]) for r in extra], np.uint32)
    l_seq = ins(batch.l_seq, [r["l_seq"] for r in extra], np.uint32)
    # flat arrays: slices of the original interleaved with the new runs
    c_parts, s_parts = [], []
    prev = 0
    for k, r in enumerate(extra):
        w = int(where[k])
        if w > prev:
            c_parts.append(batch.cigar[int(batch.cigar_off[prev]):int(batch.cigar_off[w])])
            s_parts.append(batch.seq4[int(batch.seq_off[prev]):int(batch.seq_off[w])])
            prev = w
        run = np.full((len(r["ops"]) + 3) // 4 * 4, OP_PAD, dtype=np.uint32)
        run[:len(r["ops"])] = [(ln << 4) | op for op, ln in r["ops"]]
        c_parts.append(run)
        raw = np.frombuffer(rng.bytes((r["l_seq"] + 1) // 2), dtype=np.uint8)
        s_parts.append((_ACGT_CODE[raw & 3] << 4) | _ACGT_CODE[(raw >> 2) & 3])
    c_parts.append(batch.cigar[int(batch.cigar_off[prev]):])
    s_parts.append(batch.seq4[int(batch.seq_off[prev]):])
    cigar = np.concatenate(c_parts)
    seq4 = np.concatenate(s_parts)
    padded = (n_c.astype(np.int64) + 3) // 4 * 4
    new_off = np.zeros(n0 + m + 1, dtype=np.uint64)
    new_off[1:] = np.cumsum(padded)
    new_soff = np.zeros(n0 + m + 1, dtype=np.uint64)
    new_soff[1:] = np.cumsum((l_seq.astype(np.int64) + 1) // 2)
    # new index of every original record = old index + number of inserts before it
    shift = np.searchsorted(where, np.arange(n0), side="right")
    new_index = np.arange(n0) + shift
    names = [None] * (n0 + m)
    for i, nm in enumerate(batch.names):
        names[int(new_index[i])] = nm
    k = 0
    for i in range(n0 + m):
        if names[i] is None:
            names[i] = "supp%06d" % k
            k += 1
    sa = {int(new_index[i]): t for i, t in batch.sa.items()}
    return RecordBatch(batch.contig_names, batch.contig_lengths, tid, pos, flag, mapq, n_c, new_off, l_seq, new_soff,
                       cigar, seq4, names, sa)


# ----------------------------------------------------------------------------------------------
# convenience front-ends


def make_haploid(cfg):
    layout = make_layout(cfg)
    truth = make_truth(cfg, layout, cfg.seed + 11)
    return make_haplotype(cfg, layout, truth, cfg.seed + 101)


def make_diploid(cfg):
    layout = make_layout(cfg)
    t1, t2 = diploid_truth(cfg, layout)
    return (make_haplotype(cfg, layout, t1, cfg.seed + 101, "h1ctg"),
            make_haplotype(cfg, layout, t2, cfg.seed + 202, "h2ctg"))
