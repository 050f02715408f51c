"""Writers for the on-disk formats around the hot path (SAM/BAM spec v1).

`write_bam` serialises a `synth.RecordBatch` as a real BGZF BAM plus a `.bai`
index, `write_fasta` a FASTA plus `.fai`.  Reading BAM files is NOT done here:
the product path reads them with the device ingest / the multi-threaded C++ ingest inside
`libsvimasm_b200.so` (`svb_bam_open_device`, `svb_bam_open`); `read_reference_names` only peeks at
the contig names of a header so that the CLI can start loading the FASTA before the ingest is through.
"""
import os
import struct
import zlib

import numpy as np

_BLOCK_PAYLOAD = 0xFF00
_EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(payload, level):
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = comp.compress(payload) + comp.flush()
    size = 18 + len(body) + 8
    head = struct.pack("<4BIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 0x42, 0x43, 2, size - 1)
    return head + body + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _ref_span(ops):
    code = ops & 15
    consume = (code == 0) | (code == 2) | (code == 3) | (code == 7) | (code == 8)
    return int((ops[consume] >> 4).sum())


def write_bam(path, batch, level=1, sort_order="coordinate", index=True):
    """Write `batch` to `path` (+ `path.bai`).  Records are written in batch order."""
    text = "@HD\tVN:1.6\tSO:%s\n" % sort_order if sort_order else "@HD\tVN:1.6\n"
    for name, length in zip(batch.contig_names, batch.contig_lengths):
        text += "@SQ\tSN:%s\tLN:%d\n" % (name, int(length))
    text += "@PG\tID:synth\tPN:svim_asm_b200.synth\n"
    head = bytearray(b"BAM\1")
    tb = text.encode("ascii")
    head += struct.pack("<i", len(tb)) + tb + struct.pack("<i", len(batch.contig_names))
    for name, length in zip(batch.contig_names, batch.contig_lengths):
        nb = name.encode("ascii") + b"\0"
        head += struct.pack("<i", len(nb)) + nb + struct.pack("<i", int(length))
    stream = bytearray(head)
    rec_start = np.zeros(batch.n_aln + 1, dtype=np.int64)
    spans = np.zeros(batch.n_aln, dtype=np.int64)
    for i in range(batch.n_aln):
        rec_start[i] = len(stream)
        lo = int(batch.cigar_off[i])
        n_c = int(batch.n_cigar[i])
        ops = batch.cigar[lo:lo + n_c]
        l_seq = int(batch.l_seq[i])
        span = _ref_span(ops)
        spans[i] = span
        name = batch.names[i].encode("ascii") + b"\0"
        tags = b""
        if n_c > 65535:                   # long-CIGAR convention: placeholder + CG:B,I
            tags += b"CGBI" + struct.pack("<i", n_c) + ops.astype("<u4").tobytes()
            cig_bytes = struct.pack("<II", (l_seq << 4) | 4, (span << 4) | 3)
            n_field = 2
        else:
            cig_bytes = ops.astype("<u4").tobytes()
            n_field = n_c
        if i in batch.sa:
            tags += b"SAZ" + batch.sa[i].encode("ascii") + b"\0"
        so = int(batch.seq_off[i])
        seq = batch.seq4[so:so + (l_seq + 1) // 2].tobytes()
        pos = int(batch.pos[i])
        core = struct.pack("<iiBBHHHiiii", int(batch.tid[i]), pos, len(name), int(batch.mapq[i]),
                           reg2bin(pos, pos + max(span, 1)), n_field, int(batch.flag[i]), l_seq, -1, -1, 0)
        body = core + name + cig_bytes + seq + b"\xff" * l_seq + tags
        stream += struct.pack("<i", len(body)) + body
    rec_start[-1] = len(stream)
    # ---- BGZF
    # (zlib releases the GIL: the members are compressed by a thread pool, a whole-genome file in seconds instead of half a minute)
    block_file_off = []
    view = memoryview(stream)
    starts = list(range(0, len(stream), _BLOCK_PAYLOAD))
    workers = max(1, min(32, int(os.environ.get("SVIM_BAM_WRITE_THREADS", os.cpu_count() or 1))))
    with open(path, "wb") as out:
        if workers > 1 and len(starts) > 64:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(workers) as pool:
                for lo0 in range(0, len(starts), 4096):                 # bounded batches keep the compressed copies small
                    for block in pool.map(lambda lo: _bgzf_block(bytes(view[lo:lo + _BLOCK_PAYLOAD]), level), starts[lo0:lo0 + 4096]):
                        block_file_off.append(out.tell())
                        out.write(block)
        else:
            for lo in starts:
                block_file_off.append(out.tell())
                out.write(_bgzf_block(bytes(view[lo:lo + _BLOCK_PAYLOAD]), level))
        block_file_off.append(out.tell())
        out.write(_EOF_BLOCK)
    del view
    if index:
        _write_bai(path + ".bai", batch, rec_start, spans, np.asarray(block_file_off, dtype=np.int64))


def _voffset(raw, block_file_off):
    blk = raw // _BLOCK_PAYLOAD
    return (int(block_file_off[blk]) << 16) | int(raw - blk * _BLOCK_PAYLOAD)


def _write_bai(path, batch, rec_start, spans, block_file_off):
    n_ref = len(batch.contig_names)
    bins = [dict() for _ in range(n_ref)]
    linear = [dict() for _ in range(n_ref)]
    for i in range(batch.n_aln):
        tid = int(batch.tid[i])
        if tid < 0:
            continue
        beg = int(batch.pos[i])
        end = beg + max(int(spans[i]), 1)
        v0 = _voffset(int(rec_start[i]), block_file_off)
        v1 = _voffset(int(rec_start[i + 1]), block_file_off)
        chunks = bins[tid].setdefault(reg2bin(beg, end), [])
        if chunks and chunks[-1][1] == v0:
            chunks[-1][1] = v1
        else:
            chunks.append([v0, v1])
        for win in range(beg >> 14, ((end - 1) >> 14) + 1):
            if win not in linear[tid] or v0 < linear[tid][win]:
                linear[tid][win] = v0
    with open(path, "wb") as out:
        out.write(b"BAI\1" + struct.pack("<i", n_ref))
        for tid in range(n_ref):
            out.write(struct.pack("<i", len(bins[tid])))
            for b in sorted(bins[tid]):
                out.write(struct.pack("<Ii", b, len(bins[tid][b])))
                for v0, v1 in bins[tid][b]:
                    out.write(struct.pack("<QQ", v0, v1))
            n_win = (max(linear[tid]) + 1) if linear[tid] else 0
            out.write(struct.pack("<i", n_win))
            last = 0
            for win in range(n_win):
                last = linear[tid].get(win, last)
                out.write(struct.pack("<Q", last))


def write_fasta(path, reference, names, line_width=60):
    """`reference`: dict name -> uint8 ASCII array.  Writes `path` and `path.fai`."""
    with open(path, "wb") as out, open(path + ".fai", "w") as fai:
        for name in names:
            seq = np.asarray(reference[name], dtype=np.uint8)
            out.write(b">" + name.encode("ascii") + b"\n")
            offset = out.tell()
            n = seq.shape[0]
            full = n // line_width
            if full:
                block = np.empty((full, line_width + 1), dtype=np.uint8)
                block[:, :line_width] = seq[:full * line_width].reshape(full, line_width)
                block[:, line_width] = 10
                out.write(block.tobytes())
            if n % line_width:
                out.write(seq[full * line_width:].tobytes() + b"\n")
            fai.write("%s\t%d\t%d\t%d\t%d\n" % (name, n, offset, line_width, line_width + 1))


def read_reference_names(path, limit=64 << 20):
    """Contig names of a BAM header, in header order, from the first BGZF members of the file (a few kilobytes for a
    human genome); None when the file does not look like a BAM.  The CLI uses it to start the FASTA -> HBM load while
    the first BAM is still being ingested; the ingest itself reads the header again, on its own."""
    try:
        with open(path, "rb") as f:
            data = bytearray()

            def need(n):
                while len(data) < n:
                    head = f.read(18)
                    if len(head) < 18 or head[:4] != b"\x1f\x8b\x08\x04" or head[12:14] != b"BC" or len(data) > limit:
                        return False
                    xlen = struct.unpack("<H", head[10:12])[0]
                    bsize = struct.unpack("<H", head[16:18])[0] + 1
                    rest = f.read(bsize - 18)
                    if len(rest) != bsize - 18:
                        return False
                    data.extend(zlib.decompress(rest[xlen - 6:-8], -15))
                return True
            if not need(12) or data[:4] != b"BAM\x01":
                return None
            l_text = struct.unpack_from("<i", data, 4)[0]
            at = 8 + l_text
            if l_text < 0 or not need(at + 4):
                return None
            n_ref = struct.unpack_from("<i", data, at)[0]
            at += 4
            names = []
            for _ in range(n_ref):
                if not need(at + 4):
                    return None
                l_name = struct.unpack_from("<i", data, at)[0]
                if l_name <= 0 or not need(at + 4 + l_name + 4):
                    return None
                names.append(bytes(data[at + 4:at + 4 + l_name - 1]).decode("ascii"))
                at += 4 + l_name + 4
            return names
    except (OSError, ValueError, zlib.error, struct.error, UnicodeDecodeError):
        return None
