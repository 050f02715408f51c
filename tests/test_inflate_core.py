"""Host check of the per-member DEFLATE decoder (csrc/inflate_core.cuh, one GPU thread per BGZF member in
bgzf_inflate.cu) against zlib: stored, fixed-Huffman and dynamic-Huffman blocks, every compression level, data from
incompressible to highly repetitive, and truncated / corrupt streams (which must fail, not hang or overrun)."""
import zlib

import numpy as np
import pytest

from tests import hostcheck


def _raw_deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def _inflate(lib, comp, n):
    """Both decoders: the plain one (inflate_member) and the device flow (direct tables, word refill, short matches inline)
    at every alignment of the compressed bytes; they must agree."""
    out = np.zeros(max(n, 1), dtype=np.uint8)
    rc = lib.hc_inflate(comp, len(comp), out.ctypes.data, n)
    got = out[:n].tobytes()
    for misalign in range(4):
        shift = (5 * misalign + 1) % 16                                            # the member's place in the output: any alignment
        out2 = np.zeros(n + 8 + 16, dtype=np.uint8)
        out2[:] = 0xA5
        rc2 = lib.hc_inflate_fast(comp, len(comp), out2.ctypes.data + shift, n, misalign)
        assert (rc2 == 0) == (rc == 0), (rc, rc2, misalign)
        assert bytes(out2[:shift]) == b"\xa5" * shift and bytes(out2[shift + n:]) == b"\xa5" * (24 - shift)   # nothing outside the member
        if rc == 0:
            assert out2[shift:shift + n].tobytes() == got, misalign
    return rc, got


def _samples():
    rng = np.random.default_rng(3)
    yield b""
    yield b"A"
    yield bytes(rng.integers(0, 256, 65280, dtype=np.uint8))                       # incompressible: stored blocks
    yield bytes(rng.choice(list(b"ACGT"), 65280).astype(np.uint8))                 # 2 bits of entropy per byte
    yield b"ACGT" * 16000                                                          # long matches, distance 4
    yield bytes(rng.integers(0, 4, 30000, dtype=np.uint8)) + b"\xff" * 30000       # mixed
    ops = ((rng.geometric(1 / 12.0, 16000).astype(np.uint32) << 4) | rng.choice([7, 8, 1, 2], 16000).astype(np.uint32))
    yield ops.tobytes()                                                            # BAM-packed CIGAR ops
    yield bytes((rng.geometric(0.04, 60000) % 256).astype(np.uint8))               # skewed: Huffman codes longer than the direct tables
    far = bytes(rng.integers(0, 256, 3000, dtype=np.uint8))
    yield far + bytes(rng.integers(0, 8, 28000, dtype=np.uint8)) + far + b"\xff" * 3000 + far[:40]   # distances near 32 KB, long runs


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_matches_zlib(level):
    lib = hostcheck.load()
    for data in _samples():
        comp = _raw_deflate(data, level)
        rc, got = _inflate(lib, comp, len(data))
        assert rc == 0 and got == data, (level, len(data), rc)


@pytest.mark.parametrize("strategy", [zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FILTERED])
def test_other_encoder_strategies(strategy):
    """Streams no BGZF writer produces by default but the format allows: runs only (distance 1), literals only, filtered;
    with memLevel 1 the encoder also emits many short dynamic blocks (the tables are rebuilt every few hundred symbols)."""
    lib = hostcheck.load()
    for data in _samples():
        for mem_level in (1, 9):
            c = zlib.compressobj(6, zlib.DEFLATED, -15, mem_level, strategy)
            comp = c.compress(data) + c.flush()
            rc, got = _inflate(lib, comp, len(data))
            assert rc == 0 and got == data, (strategy, mem_level, len(data), rc)


def test_flush_markers_inside_a_stream():
    """Z_SYNC_FLUSH / Z_FULL_FLUSH leave empty stored blocks between Huffman blocks (bgzip never does, other writers may)."""
    lib = hostcheck.load()
    rng = np.random.default_rng(5)
    data = bytes(rng.choice(list(b"ACGTN"), 50000).astype(np.uint8)) + b"\xff" * 5000
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = b""
    for k, lo in enumerate(range(0, len(data), 7000)):
        comp += c.compress(data[lo:lo + 7000]) + c.flush(zlib.Z_SYNC_FLUSH if k % 2 else zlib.Z_FULL_FLUSH)
    comp += c.flush()
    rc, got = _inflate(lib, comp, len(data))
    assert rc == 0 and got == data


def test_fixed_huffman_blocks():
    lib = hostcheck.load()
    for data in _samples():
        comp = _raw_deflate(data, 6, zlib.Z_FIXED)
        rc, got = _inflate(lib, comp, len(data))
        assert rc == 0 and got == data


def test_corrupt_streams_fail_cleanly():
    lib = hostcheck.load()
    rng = np.random.default_rng(4)
    data = bytes(rng.choice(list(b"ACGTN"), 20000).astype(np.uint8))
    comp = _raw_deflate(data, 6)
    assert _inflate(lib, comp[:len(comp) // 2], len(data))[0] != 0                 # truncated input
    assert _inflate(lib, comp, len(data) - 7)[0] != 0                              # ISIZE too small
    assert _inflate(lib, comp, len(data) + 7)[0] != 0                              # ISIZE too large
    bad = 0
    for _ in range(200):                                                           # random bit flips: error or different bytes, never a crash
        c = bytearray(comp)
        c[int(rng.integers(0, len(c)))] ^= 1 << int(rng.integers(0, 8))
        rc, got = _inflate(lib, bytes(c), len(data))
        bad += rc != 0 or got != data
    assert bad > 150
    assert _inflate(lib, bytes(rng.integers(0, 256, 500, dtype=np.uint8)), 4000)[0] != 0   # garbage
