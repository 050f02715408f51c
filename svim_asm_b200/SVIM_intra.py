"""Seams of reference src/svim_asm/SVIM_intra.py, computed by the cigar_scan kernel (csrc/cigar_scan.cu)."""
import numpy as np

from . import synth
from .engine import HostBatch, make_params
from .runtime import get_engine
from .SVCandidate import candidates_from_rows


def analyze_cigar_indel(tuples, min_length):
    """SVIM_intra.py:8-30: [(pos_ref, pos_read, length, "INS"|"DEL")] for I/D ops with length >= min_length."""
    return get_engine().cigar_indel(list(tuples), min_length)


def single_record_batch(alignment, bam, supplementaries=()):
    """HostBatch holding one alignment (and, optionally, pre-built SA pseudo alignments as its segments)."""
    cigar = alignment.cigartuples or []
    seq = alignment.query_sequence
    l_seq = len(seq) if seq else 0
    ops = np.full((len(cigar) + 3) // 4 * 4, synth.OP_PAD, dtype=np.uint32)
    ops[:len(cigar)] = [(int(n) << 4) | int(op) for op, n in cigar]
    lut = {c: i for i, c in enumerate(synth.NT16)}
    codes = np.array([lut.get(c, 15) for c in (seq or "")], dtype=np.uint8)
    if codes.shape[0] % 2:
        codes = np.append(codes, 0)
    seq4 = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8) if codes.shape[0] else np.zeros(0, dtype=np.uint8)
    rb = synth.RecordBatch(list(bam.references), np.asarray(bam.lengths, dtype=np.int32),
                           np.array([alignment.reference_id], dtype=np.int32),
                           np.array([alignment.reference_start], dtype=np.int32),
                           np.array([alignment.flag], dtype=np.uint16), np.array([alignment.mapping_quality], dtype=np.uint8),
                           np.array([len(cigar)], dtype=np.uint32), np.array([0, ops.shape[0]], dtype=np.uint64),
                           np.array([l_seq], dtype=np.uint32), np.array([0, seq4.shape[0]], dtype=np.uint64), ops, seq4,
                           [alignment.query_name], {})
    host = HostBatch.from_record_batch(rb)
    if supplementaries:
        from ._lib import SEG_DTYPE
        seg = np.zeros(len(supplementaries), dtype=SEG_DTYPE)
        for k, s in enumerate(supplementaries):
            seg[k]["tid"], seg[k]["pos"] = s.reference_id, s.reference_start
            seg[k]["is_reverse"], seg[k]["mapq"] = int(s.is_reverse), s.mapping_quality
            seg[k]["ref_end"], seg[k]["q_astart"] = s.reference_end, s.query_alignment_start
            seg[k]["q_aend"], seg[k]["read_len"] = s.query_alignment_end, s.infer_read_length()
        host.seg = seg
        host.sa_count = np.array([len(supplementaries)], dtype=np.uint32)
    return host


def _collect_single(alignment, bam, options, supplementaries=(), keep_flag=False):
    host = single_record_batch(alignment, bam, supplementaries)
    if not keep_flag:
        # the seam is called AFTER the record filter (SVIM_COLLECT.py:71): always analyse this alignment
        host.hdr["flag"] &= np.uint16(~(0x4 | 0x100) & 0xFFFF)
        host.hdr["mapq"] = 255
    eng = get_engine()
    rec = eng.load_records(host)
    params = make_params(options)
    if not keep_flag:
        params["min_mapq"] = min(int(params["min_mapq"][0]), 255)
    rows = eng.collect(rec, params).to_numpy()
    rec.free()
    return rows, host


def analyze_alignment_indel(alignment, bam, query_name, options):
    """SVIM_intra.py:33-44: CandidateDeletion / CandidateInsertion objects of one alignment."""
    rows, host = _collect_single(alignment, bam, options)
    rows = rows[(rows["ordinal"] & np.uint64(0x80000000)) == 0]
    host._names = [query_name]
    return candidates_from_rows(rows, {0: host}, list(bam.references), list(bam.lengths))
