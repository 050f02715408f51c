"""Shared helpers of the parity tests: canonical tuples for candidate rows."""
import numpy as np

TYPE_NAMES = ("DEL", "INV", "INS", "DUP_TAN", "DUP_INT", "BND")
GT = ("1/1", "1/0", "0/1")
F_COMPLETE, F_FULLY, F_CUTPASTE, F_SRC_FWD, F_DST_FWD = 1, 2, 4, 8, 16
COMPARE_FIELDS = ("type", "flags", "genotype", "hap", "src_tid", "src_start", "src_end", "dst_tid", "dst_start",
                  "dst_end", "copies", "aln_idx", "seq_pos", "seq_len", "mate_aln")


def rows_equal(a, b, fields=COMPARE_FIELDS):
    """First differing (index, field, a, b) or None.  `ordinal` is an ordering device and is not compared."""
    if a.shape[0] != b.shape[0]:
        return ("length", None, a.shape[0], b.shape[0])
    for f in fields:
        if f in ("src_end", "dst_end"):
            # BND rows carry no end; INS carries no source at all
            pass
        bad = np.nonzero(a[f] != b[f])[0]
        if bad.size:
            i = int(bad[0])
            return (i, f, a[i], b[i])
    return None


def canon_row(r, hosts, names):
    """The tuple oracle/refrun.canon() makes of the reference's Candidate object, from a table row.
    hosts: {hap: HostBatch} (hap 0 for haploid)."""
    t = TYPE_NAMES[int(r["type"])]
    g = GT[int(r["genotype"])]
    hap = int(r["hap"])
    host = hosts[hap]
    reads = [host.query_name(int(r["aln_idx"]))]
    if int(r["mate_aln"]) != 0xFFFFFFFF:
        other = hosts[3 - hap]
        reads.append(other.query_name(int(r["mate_aln"])))
    reads = tuple(reads)
    if t == "DEL":
        return (t, names[r["src_tid"]], int(r["src_start"]), int(r["src_end"]), g, reads)
    if t == "INV":
        return (t, names[r["src_tid"]], int(r["src_start"]), int(r["src_end"]), bool(r["flags"] & F_COMPLETE), g, reads)
    if t == "INS":
        seq = host.sequence_slice(int(r["aln_idx"]), int(r["seq_pos"]), int(r["seq_len"]))
        return (t, names[r["dst_tid"]], int(r["dst_start"]), int(r["dst_end"]), seq, g, reads)
    if t == "DUP_TAN":
        return (t, names[r["src_tid"]], int(r["src_start"]), int(r["src_end"]), int(r["copies"]),
                bool(r["flags"] & F_FULLY), g, reads)
    if t == "DUP_INT":
        return (t, names[r["src_tid"]], int(r["src_start"]), int(r["src_end"]), names[r["dst_tid"]],
                int(r["dst_start"]), int(r["dst_end"]), bool(r["flags"] & F_CUTPASTE), g, reads)
    return (t, names[r["src_tid"]], int(r["src_start"]), "fwd" if r["flags"] & F_SRC_FWD else "rev",
            names[r["dst_tid"]], int(r["dst_start"]), "fwd" if r["flags"] & F_DST_FWD else "rev", g, reads)


def batch_from_records(contig_names, contig_lengths, records, rng=None):
    """RecordBatch from explicit records: dicts with tid, pos, flag, mapq, cigar [(op,len)...], sa (optional)."""
    from svim_asm_b200 import synth
    rng = rng or np.random.default_rng(0)
    n = len(records)
    n_c = np.array([len(r["cigar"]) for r in records], dtype=np.uint32)
    padded = (n_c.astype(np.int64) + 3) // 4 * 4
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(padded)
    cigar = np.full(int(off[-1]), synth.OP_PAD, dtype=np.uint32)
    l_seq = np.zeros(n, dtype=np.uint32)
    for i, r in enumerate(records):
        vals = np.array([(int(ln) << 4) | int(op) for op, ln in r["cigar"]], dtype=np.uint32)
        cigar[int(off[i]):int(off[i]) + vals.shape[0]] = vals
        l_seq[i] = r.get("l_seq", sum(ln for op, ln in r["cigar"] if op in (0, 1, 4, 7, 8)))
    nb = (l_seq.astype(np.int64) + 1) // 2
    soff = np.zeros(n + 1, dtype=np.uint64)
    soff[1:] = np.cumsum(nb)
    raw = rng.integers(0, 256, int(soff[-1]), dtype=np.uint8)
    code = np.array([1, 2, 4, 8], dtype=np.uint8)
    seq4 = (code[raw & 3] << 4) | code[(raw >> 2) & 3]
    return synth.RecordBatch(list(contig_names), np.asarray(contig_lengths, dtype=np.int32),
                             np.array([r["tid"] for r in records], dtype=np.int32),
                             np.array([r["pos"] for r in records], dtype=np.int32),
                             np.array([r.get("flag", 0) for r in records], dtype=np.uint16),
                             np.array([r.get("mapq", 60) for r in records], dtype=np.uint8),
                             n_c, off, l_seq, soff, cigar, seq4,
                             [r.get("name", "rec%05d" % i) for i, r in enumerate(records)],
                             {i: r["sa"] for i, r in enumerate(records) if r.get("sa")})
