"""Python face of the C ABI: device context, uploaded BAM images, candidate tables.

`HostBatch` is the host-side record image (what pysam.AlignmentFile + fetch() are to the
reference, svim-asm:63, SVIM_COLLECT.py:62-65): built either from a BAM file by the C++ ingest or
from a synthetic `synth.RecordBatch`.  `Engine` owns one GPU context.  Every failure of the
library surfaces as `RuntimeError(svb_last_error())`, which lands in the same top-level handler as
the reference's exceptions (svim-asm:182-185).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import lib

NT16 = "=ACMGRSVTWYHKDBN"
_NT16_LUT = np.frombuffer(NT16.encode("ascii"), dtype=np.uint8)

DEFAULT_PARAMS = dict(min_mapq=20, min_sv_size=40, max_sv_size=100000, query_gap_tolerance=50,
                      query_overlap_tolerance=50, reference_gap_tolerance=50, reference_overlap_tolerance=50,
                      partition_max_distance=1000, max_edit_distance=200)


def make_params(options=None, **overrides):
    """svb_params from an argparse namespace (SVIM_input_parsing.py) and/or keyword overrides."""
    p = np.zeros(1, dtype=_lib.PARAMS_DTYPE)
    for name, default in DEFAULT_PARAMS.items():
        value = overrides.get(name, getattr(options, name, default) if options is not None else default)
        p[name] = int(value)
    return p


def lexrank(names):
    """Rank of each contig name under python str ordering (SVCandidate.py:352, SVIM_COMBINE.py:17)."""
    order = sorted(range(len(names)), key=lambda i: names[i])
    rank = np.zeros(len(names), dtype=np.int32)
    rank[order] = np.arange(len(names), dtype=np.int32)
    return rank


class HostBatch(object):
    """Flat host image of one BAM file."""

    def __init__(self):
        self.contig_names = []
        self.contig_lengths = np.zeros(0, dtype=np.int32)
        self.sort_order = ""
        self.hdr = np.zeros(0, dtype=_lib.HDR_DTYPE)
        self.cigar = np.zeros(0, dtype=np.uint32)
        self.seg = np.zeros(0, dtype=_lib.SEG_DTYPE)
        self.sa_count = np.zeros(0, dtype=np.uint32)
        self.seq4 = np.zeros(0, dtype=np.uint8)
        self.seq_off = np.zeros(1, dtype=np.uint64)
        self._names = None
        self._bam = None                 # svb_bam* keeping the C++ buffers alive
        self.path = None

    # ---- construction
    @classmethod
    def from_bam(cls, path, threads=0):
        handle = ctypes.c_void_p()
        err = ctypes.create_string_buffer(512)
        rc = lib.svb_bam_open(str(path).encode(), int(threads), ctypes.byref(handle), err, len(err))
        if rc != 0:
            if rc == -4:
                raise IOError(err.value.decode() or "cannot read %s" % path)
            raise RuntimeError(err.value.decode() or "svb_bam_open failed (%d)" % rc)
        return cls._from_handle(handle, path)

    @classmethod
    def _from_handle(cls, handle, path):
        self = cls()
        self._bam = handle
        self.path = str(path)
        n = lib.svb_bam_n_records(handle)
        n_contig = lib.svb_bam_n_contigs(handle)
        self.contig_names = [lib.svb_bam_contig_name(handle, t).decode() for t in range(n_contig)]

        def view(address, dtype, count):
            if count == 0 or not address:
                return np.zeros(0, dtype=dtype)
            buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(address)
            return np.frombuffer(buf, dtype=dtype, count=count)

        self.contig_lengths = view(lib.svb_bam_contig_lengths(handle), np.int32, n_contig)
        self.sort_order = lib.svb_bam_sort_order(handle).decode()
        self.hdr = view(lib.svb_bam_headers(handle), _lib.HDR_DTYPE, n)
        self.cigar = view(lib.svb_bam_cigar(handle), np.uint32, lib.svb_bam_n_ops_padded(handle))
        self.seg = view(lib.svb_bam_segments(handle), _lib.SEG_DTYPE, lib.svb_bam_n_segments(handle))
        self.sa_count = view(lib.svb_bam_sa_count(handle), np.uint32, n)
        self.seq_off = view(lib.svb_bam_seq_offsets(handle), np.uint64, n + 1)
        self.seq4 = view(lib.svb_bam_seq4(handle), np.uint8, int(self.seq_off[-1]) if n else 0)
        return self

    @classmethod
    def from_bam_device(cls, engine, path, keep_sequences=True):
        """Device ingest (svb_bam_open_device): BGZF inflate + record split on the GPU.  Returns (host, records): the
        record image is already resident; `host` carries the small per-record data, its CIGAR / sequence arrays are
        downloaded on first use (only the per-alignment seams and the VCF writer's INS alleles read them)."""
        handle, rec = ctypes.c_void_p(), ctypes.c_void_p()
        err = ctypes.create_string_buffer(512)
        rc = lib.svb_bam_open_device(engine.handle, str(path).encode(), int(bool(keep_sequences)), None, ctypes.byref(handle),
                                     ctypes.byref(rec), err, len(err))
        if rc != 0:
            if rc == -4:
                raise IOError(err.value.decode() or "cannot read %s" % path)
            raise RuntimeError(err.value.decode() or "svb_bam_open_device failed (%d)" % rc)
        records = Records(engine, rec, None)
        records.has_sequences = bool(keep_sequences)
        self = DeviceBackedHostBatch._from_handle(handle, path)
        records.host = self
        self._engine, self._records = engine, records
        return self, records

    @classmethod
    def from_record_batch(cls, rb):
        """From a synth.RecordBatch (SA text goes through the same parser as the ingest)."""
        self = cls()
        n = rb.n_aln
        self.contig_names = list(rb.contig_names)
        self.contig_lengths = np.ascontiguousarray(rb.contig_lengths, dtype=np.int32)
        self.sort_order = "coordinate"
        hdr = np.zeros(n, dtype=_lib.HDR_DTYPE)
        hdr["tid"], hdr["pos"], hdr["flag"], hdr["mapq"] = rb.tid, rb.pos, rb.flag, rb.mapq
        hdr["n_cigar"], hdr["cigar_off"], hdr["l_seq"] = rb.n_cigar, rb.cigar_off[:-1], rb.l_seq
        self.cigar = np.ascontiguousarray(rb.cigar, dtype=np.uint32)
        self.seq4 = np.ascontiguousarray(rb.seq4, dtype=np.uint8)
        self.seq_off = np.ascontiguousarray(rb.seq_off, dtype=np.uint64)
        self.sa_count = np.zeros(n, dtype=np.uint32)
        names_c = (ctypes.c_char_p * len(self.contig_names))(*[s.encode() for s in self.contig_names])
        segs = []
        total = 0
        tmp = np.zeros(64, dtype=_lib.SEG_DTYPE)
        first = np.zeros(n, dtype=np.uint32)
        for i in sorted(rb.sa):
            text = rb.sa[i].encode()
            cnt = lib.svb_parse_sa(text, names_c, len(self.contig_names), tmp.ctypes.data, tmp.shape[0])
            if cnt > tmp.shape[0]:
                tmp = np.zeros(cnt, dtype=_lib.SEG_DTYPE)
                cnt = lib.svb_parse_sa(text, names_c, len(self.contig_names), tmp.ctypes.data, tmp.shape[0])
            if cnt < 0:
                raise ValueError("invalid literal in SA tag: %r" % rb.sa[i])
            first[i] = total
            self.sa_count[i] = cnt
            segs.append(tmp[:cnt].copy())
            total += cnt
        # sa_first of records without SA: running total (never dereferenced)
        running = np.concatenate(([0], np.cumsum(self.sa_count)[:-1])).astype(np.uint32) if n else first
        hdr["sa_first"] = running
        self.hdr = hdr
        self.seg = np.concatenate(segs) if segs else np.zeros(0, dtype=_lib.SEG_DTYPE)
        self._names = list(rb.names)
        self._sa_text = dict(rb.sa)
        return self

    # ---- access
    @property
    def n_aln(self):
        return int(self.hdr.shape[0])

    @property
    def n_ops(self):
        return int(self.hdr["n_cigar"].sum(dtype=np.uint64))

    def query_name(self, i):
        if self._names is not None:
            return self._names[int(i)]
        return lib.svb_bam_query_name(self._bam, int(i)).decode()

    def sa_text(self, i):
        """Raw SA:Z value of record i or None."""
        if self._bam is not None:
            raw = lib.svb_bam_sa_text(self._bam, int(i))
            return raw.decode() if raw is not None else None
        return getattr(self, "_sa_text", {}).get(int(i))

    def sequence_slice(self, i, start, length):
        """query_sequence[start:start+length] of record i (start/length already python-slice normalised)."""
        if length <= 0:
            return ""
        nib0 = 2 * int(self.seq_off[i]) + int(start)
        b0, b1 = nib0 // 2, (nib0 + int(length) + 1) // 2
        raw = self.seq4[b0:b1]
        nib = np.empty(raw.shape[0] * 2, dtype=np.uint8)
        nib[0::2] = raw >> 4
        nib[1::2] = raw & 15
        s = nib0 - 2 * b0
        return _NT16_LUT[nib[s:s + int(length)]].tobytes().decode("ascii")

    def close(self):
        if self._bam is not None:
            for name in ("hdr", "cigar", "seg", "sa_count", "seq4", "seq_off", "contig_lengths"):
                setattr(self, name, np.zeros(0, dtype=getattr(self, name).dtype))
            lib.svb_bam_close(self._bam)
            self._bam = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Table(object):
    """Device-resident candidate table (svb_table*)."""

    def __init__(self, engine, handle):
        self.engine = engine
        self.handle = handle

    def __len__(self):
        return int(lib.svb_table_size(self.handle))

    def to_numpy(self):
        n = len(self)
        rows = np.empty(n, dtype=_lib.ROW_DTYPE)
        got = ctypes.c_uint64()
        self.engine._check(lib.svb_table_to_host(self.engine.handle, self.handle, _lib.ptr(rows), n, ctypes.byref(got)))
        return rows

    # ---- sequence pool (inserted bases of the INS rows travelling with the table)
    def gather_sequences(self, rec):
        """Device-side gather from the record image's resident query sequences."""
        self.engine._check(lib.svb_table_gather_sequences(self.engine.handle, self.handle, rec.handle))

    def attach_sequences_host(self, host):
        """Host-side gather from a HostBatch: only the inserted bases cross PCIe, not the whole assembly."""
        self.engine._check(lib.svb_table_attach_sequences_host(self.engine.handle, self.handle, _lib.ptr(host.seq4),
                                                               _lib.ptr(host.seq_off)))

    def pool_to_numpy(self):
        """(pool bytes, start offset of every row's run)."""
        n = len(self)
        off = np.zeros(n + 1, dtype=np.uint64)
        size = ctypes.c_uint64()
        self.engine._check(lib.svb_table_pool_to_host(self.engine.handle, self.handle, None, 0, _lib.ptr(off), ctypes.byref(size)))
        pool = np.zeros(int(size.value), dtype=np.uint8)
        if pool.shape[0]:
            self.engine._check(lib.svb_table_pool_to_host(self.engine.handle, self.handle, _lib.ptr(pool), pool.shape[0], None,
                                                          ctypes.byref(size)))
        return pool, off[:-1].copy()

    def remap_records(self, rec):
        """aln_idx / ordinal -> indices in the unsharded batch (Records.set_global_index)."""
        self.engine._check(lib.svb_table_remap_records(self.engine.handle, self.handle, rec.handle))

    def device_rows(self):
        """(device pointer, bytes) of the rows."""
        return int(lib.svb_table_device_rows(self.handle) or 0), len(self) * _lib.ROW_DTYPE.itemsize

    def set_pool(self, pool, starts):
        """Attach a pool given the start offset of every row's run (any order, gaps allowed)."""
        pool = np.ascontiguousarray(pool, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        assert starts.shape[0] == len(self)
        self.engine._check(lib.svb_table_set_pool_from_host(self.engine.handle, self.handle, _lib.ptr(pool), pool.shape[0],
                                                            _lib.ptr(starts)))

    def free(self):
        if self.handle:
            if getattr(self.engine, "handle", None):
                lib.svb_table_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceBackedHostBatch(HostBatch):
    """HostBatch of a device ingest: `cigar` and `seq4` live in HBM and come to the host only when somebody reads them
    (each on its own: the VCF writer's INS alleles need the bases, the per-alignment seams the CIGAR ops)."""
    _engine = None
    _records = None
    _have = 0

    def _materialize(self, name):
        bit = 1 if name == "cigar" else 2
        if not (self._have & bit) and self._bam is not None and self._records is not None and self._records.handle:
            self._engine._check(lib.svb_bam_materialize_host(self._engine.handle, self._bam, self._records.handle, bit))
            n = lib.svb_bam_n_records(self._bam)

            def view(address, dtype, count):
                if count == 0 or not address:
                    return np.zeros(0, dtype=dtype)
                buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(address)
                return np.frombuffer(buf, dtype=dtype, count=count)
            if bit == 1:
                self.__dict__["cigar"] = view(lib.svb_bam_cigar(self._bam), np.uint32, lib.svb_bam_n_ops_padded(self._bam))
            else:
                self.__dict__["seq4"] = view(lib.svb_bam_seq4(self._bam), np.uint8, int(self.seq_off[-1]) if n else 0)
            self._have |= bit

    def __getattribute__(self, name):
        if name == "cigar" or name == "seq4":
            object.__getattribute__(self, "_materialize")(name)
        return object.__getattribute__(self, name)


class Records(object):
    """Device-resident BAM image (svb_records*)."""

    def __init__(self, engine, handle, host):
        self.engine = engine
        self.handle = handle
        self.host = host
        self.has_sequences = False

    def free(self):
        if self.handle:
            if getattr(self.engine, "handle", None):      # a closed context has already released its stream and pools
                lib.svb_records_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Reference(object):
    def __init__(self, engine, handle):
        self.engine = engine
        self.handle = handle

    def to_numpy(self):
        """(bases, class map) of the resident copy (tests)."""
        n = ctypes.c_uint64()
        self.engine._check(lib.svb_ref_to_host(self.engine.handle, self.handle, None, 0, ctypes.byref(n), None))
        bases, cmap = np.zeros(int(n.value), dtype=np.uint8), np.zeros(256, dtype=np.uint8)
        self.engine._check(lib.svb_ref_to_host(self.engine.handle, self.handle, _lib.ptr(bases), bases.shape[0], ctypes.byref(n),
                                               cmap.ctypes.data))
        return bases, cmap

    def free(self):
        if self.handle:
            if getattr(self.engine, "handle", None):
                lib.svb_ref_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceView(object):
    """A span of device memory owned by someone else, visible to torch (`torch.as_tensor(view, device=...)`) through
    __cuda_array_interface__: the collectives of the multi-GPU exchange read and write the library's buffers in place."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = int(ptr), int(nbytes)

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}


class DeviceBuffer(DeviceView):
    """Stream-ordered scratch of the engine's context."""

    def __init__(self, engine, nbytes):
        out = ctypes.c_void_p()
        engine._check(lib.svb_device_alloc(engine.handle, int(nbytes), ctypes.byref(out)))
        DeviceView.__init__(self, out.value, nbytes)
        self.engine = engine

    def free(self):
        if self.ptr:
            lib.svb_device_free(self.engine.handle, ctypes.c_void_p(self.ptr))
            self.ptr = 0


class ExchangeWindow(object):
    """One rank's window of the peer-memory exchange (csrc/exchange.cu): `handle` is the 64-byte CUDA IPC handle the other
    processes need; after open(handles of all ranks) share() / gather_paired() are the two collective calls of a step."""

    def __init__(self, engine, world, rank, slot_bytes, result_bytes):
        self.engine, self.world, self.rank = engine, int(world), int(rank)
        out = ctypes.c_void_p()
        engine._check(lib.svb_exchange_create(engine.handle, self.world, self.rank, int(slot_bytes), int(result_bytes), ctypes.byref(out)))
        self.ptr = out
        buf = np.zeros(64, dtype=np.uint8)
        engine._check(lib.svb_exchange_handle(engine.handle, self.ptr, _lib.ptr(buf)))
        self.handle = buf.tobytes()

    def open(self, handles):
        blob = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
        assert blob.shape[0] == 64 * self.world
        self.engine._check(lib.svb_exchange_open(self.engine.handle, self.ptr, _lib.ptr(blob)))

    def share(self, table1, table2, owner):
        owner = np.ascontiguousarray(owner, dtype=np.int32)
        u1, u2 = ctypes.c_void_p(), ctypes.c_void_p()
        self.engine._check(lib.svb_exchange_share(self.engine.handle, self.ptr, table1.handle, table2.handle, _lib.ptr(owner),
                                                  owner.shape[0], ctypes.byref(u1), ctypes.byref(u2)))
        return Table(self.engine, u1), Table(self.engine, u2)

    def gather_paired(self, paired, contig_lexrank):
        ranks = np.ascontiguousarray(contig_lexrank, dtype=np.int32)
        out = ctypes.c_void_p()
        self.engine._check(lib.svb_exchange_gather_paired(self.engine.handle, self.ptr, paired.handle, _lib.ptr(ranks), ranks.shape[0],
                                                          ctypes.byref(out)))
        return Table(self.engine, out)

    def close(self):
        if self.ptr:
            if getattr(self.engine, "handle", None):
                lib.svb_exchange_destroy(self.engine.handle, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine(object):
    """One GPU context (svb_ctx*).  Not thread-safe, like the reference."""

    def __init__(self, device=0):
        handle = ctypes.c_void_p()
        rc = lib.svb_create(int(device), ctypes.byref(handle))
        if rc != 0:
            raise RuntimeError("svb_create(device=%d) failed (%d): no usable CUDA device -- this package has no "
                               "CPU fallback" % (device, rc))
        self.handle = handle
        self.device = int(device)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("libsvimasm_b200 error %d: %s" % (rc, lib.svb_last_error(self.handle).decode()))

    def close(self):
        if self.handle:
            lib.svb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads
    def load_records(self, host, with_sequences=False):
        out = ctypes.c_void_p()
        rank = lexrank(host.contig_names)
        self._check(lib.svb_load_records(self.handle, _lib.ptr(host.hdr), host.n_aln, _lib.ptr(host.cigar),
                                         int(host.cigar.shape[0]), _lib.ptr(host.seg), _lib.ptr(host.sa_count),
                                         int(host.seg.shape[0]), _lib.ptr(host.contig_lengths), _lib.ptr(rank),
                                         len(host.contig_names), ctypes.byref(out)))
        rec = Records(self, out, host)
        if with_sequences:
            self.set_sequences(rec)
        return rec

    def set_sequences(self, rec):
        host = rec.host
        self._check(lib.svb_records_set_sequences(self.handle, rec.handle, _lib.ptr(host.seq4), _lib.ptr(host.seq_off)))
        rec.has_sequences = True

    def load_reference_fasta(self, path, fai_rows):
        """Reference genome from the FASTA file itself (svb_ref_load_fasta).  fai_rows: one (length, offset, linebases,
        linewidth) per contig in BAM header order, (0, 0, 0, 0) for a contig the FASTA lacks."""
        out = ctypes.c_void_p()
        fai = np.ascontiguousarray(fai_rows, dtype=np.uint64).reshape(-1, 4)
        self._check(lib.svb_ref_load_fasta(self.handle, str(path).encode(), _lib.ptr(fai) if fai.size else None, fai.shape[0],
                                           ctypes.byref(out)))
        return Reference(self, out)

    def ingest_timings(self):
        """ms of the last device ingest: file read, H2D, inflate, record chase, fields + copy, host parse, total; inflated bytes."""
        v = np.zeros(12, dtype=np.float64)
        self._check(lib.svb_bam_device_timings(self.handle, _lib.ptr(v)))
        keys = ("read_file", "h2d", "inflate", "chase", "fields_copy", "host_parse", "total", "inflated_bytes",
                "inflate_ctas_per_sm", "inflate_cycles_per_member", "members", "indexed_chase_segments")
        return dict(zip(keys, (float(x) for x in v)))

    def load_reference(self, bases, contig_off):
        """bases: uint8 upper-cased ASCII, contigs concatenated in BAM header order; contig_off: uint64[n+1]."""
        out = ctypes.c_void_p()
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        contig_off = np.ascontiguousarray(contig_off, dtype=np.uint64)
        self._check(lib.svb_ref_load(self.handle, _lib.ptr(bases), _lib.ptr(contig_off), contig_off.shape[0] - 1,
                                     ctypes.byref(out)))
        return Reference(self, out)

    # ---- hot path
    def collect(self, rec, params, hap=0):
        out = ctypes.c_void_p()
        self._check(lib.svb_collect(self.handle, rec.handle, _lib.ptr(params), int(hap), ctypes.byref(out)))
        return Table(self, out)

    def collect2(self, rec1, rec2, params, with_pools=False):
        """Both haplotypes of a diploid run with one host synchronisation (svb_collect2); with_pools also copies the inserted
        bases of the INS rows next to each table."""
        o1, o2 = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(lib.svb_collect2(self.handle, rec1.handle, rec2.handle, _lib.ptr(params), 1 if with_pools else 0,
                                     ctypes.byref(o1), ctypes.byref(o2)))
        return Table(self, o1), Table(self, o2)

    def map_sequences_host(self, rec):
        """The record image reads its query sequences in place from the host batch's PINNED buffers (no upload)."""
        host = rec.host
        self._check(lib.svb_records_map_sequences_host(self.handle, rec.handle, _lib.ptr(host.seq4), _lib.ptr(host.seq_off)))
        rec.has_sequences = True

    def pair(self, table1, table2, rec1, rec2, reference, params):
        out = ctypes.c_void_p()
        self._check(lib.svb_pair(self.handle, table1.handle, table2.handle, rec1.handle, rec2.handle,
                                 reference.handle, _lib.ptr(params), ctypes.byref(out)))
        return Table(self, out)

    def vcf_body(self, table, records_by_hap, reference, contig_names, entries, symbolic=False):
        """Record lines of variants.vcf for `entries` (svb_vcf_entry rows, already in output order) as bytes; the alleles
        are gathered on the device from the resident reference and query sequences (svb_vcf_body)."""
        entries = np.ascontiguousarray(entries, dtype=_lib.VCF_ENTRY_DTYPE)
        encoded = [str(n).encode() for n in contig_names]
        name_off = np.zeros(len(encoded) + 1, dtype=np.uint32)
        name_off[1:] = np.cumsum([len(b) for b in encoded])
        recs = (ctypes.c_void_p * 3)(*[(records_by_hap[h].handle if records_by_hap.get(h) is not None else None) for h in range(3)])
        text, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._check(lib.svb_vcf_body(self.handle, table.handle, recs, reference.handle, b"".join(encoded) + b"\0", _lib.ptr(name_off),
                                     len(encoded), _lib.ptr(entries) if entries.shape[0] else None, entries.shape[0],
                                     1 if symbolic else 0, ctypes.byref(text), ctypes.byref(n)))
        return ctypes.string_at(text.value, int(n.value)) if n.value else b""

    def table_from_numpy(self, rows):
        out = ctypes.c_void_p()
        rows = np.ascontiguousarray(rows, dtype=_lib.ROW_DTYPE)
        self._check(lib.svb_table_from_host(self.handle, _lib.ptr(rows), rows.shape[0], ctypes.byref(out)))
        return Table(self, out)

    def cigar_indel(self, tuples, min_length):
        """analyze_cigar_indel (SVIM_intra.py:8-30) on the GPU."""
        ops = np.array([(int(n) << 4) | int(op) for op, n in tuples], dtype=np.uint32)
        cap = max(1, ops.shape[0])
        out = np.zeros((cap, 4), dtype=np.int64)
        n = ctypes.c_uint32()
        self._check(lib.svb_cigar_indel(self.handle, _lib.ptr(ops), ops.shape[0], int(min_length), _lib.ptr(out), cap,
                                        ctypes.byref(n)))
        return [(int(r[0]), int(r[1]), int(r[2]), "DEL" if r[3] else "INS") for r in out[:n.value]]

    def edit_distance(self, pairs, max_distance=None):
        """Unit-cost global edit distance of each (a, b) byte-string pair (edlib.align(a, b)['editDistance']); with
        max_distance = k: edlib.align(a, b, k=k), i.e. -1 where the distance exceeds k (thresholded wavefront kernel)."""
        a = b"".join(p[0] for p in pairs)
        b = b"".join(p[1] for p in pairs)
        a_off = np.cumsum([0] + [len(p[0]) for p in pairs]).astype(np.uint64)
        b_off = np.cumsum([0] + [len(p[1]) for p in pairs]).astype(np.uint64)
        av = np.frombuffer(a, dtype=np.uint8) if a else np.zeros(0, dtype=np.uint8)
        bv = np.frombuffer(b, dtype=np.uint8) if b else np.zeros(0, dtype=np.uint8)
        out = np.zeros(len(pairs), dtype=np.int64)
        if max_distance is None:
            self._check(lib.svb_edit_distance(self.handle, _lib.ptr(av), _lib.ptr(a_off), _lib.ptr(bv), _lib.ptr(b_off),
                                              len(pairs), _lib.ptr(out)))
        else:
            self._check(lib.svb_edit_distance_bounded(self.handle, _lib.ptr(av), _lib.ptr(a_off), _lib.ptr(bv), _lib.ptr(b_off),
                                                      len(pairs), int(max_distance), _lib.ptr(out)))
        return out

    def form_partitions(self, keys, max_distance):
        """svb_form_partitions: (order, part_start) of the stable key sort + the consecutive-gap split (SVIM_COMBINE.py:15-32)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        n = keys.shape[0]
        order, part_start = np.zeros(n, dtype=np.uint32), np.zeros(n + 1, dtype=np.uint32)
        n_parts = ctypes.c_uint32()
        self._check(lib.svb_form_partitions(self.handle, _lib.ptr(keys), n, int(max_distance), _lib.ptr(order), _lib.ptr(part_start),
                                            ctypes.byref(n_parts)))
        return order, part_start[:n_parts.value + 1].copy()

    def cluster_labels(self, condensed_list, threshold):
        n_points = []
        for d in condensed_list:
            m = len(d)
            n = int(round((1 + (1 + 8 * m) ** 0.5) / 2))
            assert n * (n - 1) // 2 == m
            n_points.append(n)
        flat = np.concatenate([np.asarray(d, dtype=np.float64) for d in condensed_list]) if condensed_list else np.zeros(0)
        npts = np.asarray(n_points, dtype=np.uint32)
        out = np.zeros((len(n_points), 32), dtype=np.int32)
        self._check(lib.svb_cluster_labels(self.handle, _lib.ptr(flat), _lib.ptr(npts), len(n_points), float(threshold),
                                           _lib.ptr(out)))
        return [out[i, :n].tolist() for i, n in enumerate(n_points)]

    # ---- multi-GPU exchange (device resident, see csrc/exchange.cu)
    def stream_handle(self):
        return int(lib.svb_stream(self.handle) or 0)

    def set_global_index(self, rec, global_idx):
        idx = np.ascontiguousarray(global_idx, dtype=np.uint32)
        assert idx.shape[0] == rec.host.n_aln
        self._check(lib.svb_records_set_global_index(self.handle, rec.handle, _lib.ptr(idx)))

    def exchange_sizes(self, table1, table2):
        sizes = np.zeros(4, dtype=np.uint64)
        self._check(lib.svb_exchange_sizes(table1.handle, table2.handle, _lib.ptr(sizes)))
        return sizes

    @staticmethod
    def exchange_bytes(sizes):
        sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        return int(lib.svb_exchange_bytes(_lib.ptr(sizes)))

    def exchange_pack(self, table1, table2, device_ptr, cap_bytes):
        self._check(lib.svb_exchange_pack(self.handle, table1.handle, table2.handle, ctypes.c_void_p(int(device_ptr)), int(cap_bytes)))

    def exchange_unpack(self, device_ptr, stride, sizes, hap, owner, rank):
        sizes = np.ascontiguousarray(sizes, dtype=np.uint64).reshape(-1, 4)
        owner = np.ascontiguousarray(owner, dtype=np.int32)
        out = ctypes.c_void_p()
        self._check(lib.svb_exchange_unpack(self.handle, ctypes.c_void_p(int(device_ptr)), int(stride), _lib.ptr(sizes),
                                            sizes.shape[0], int(hap), _lib.ptr(owner), owner.shape[0], int(rank), ctypes.byref(out)))
        return Table(self, out)

    def exchange_window(self, world, rank, slot_bytes, result_bytes):
        """Peer-memory exchange window of this rank (svb_exchange_create); see ExchangeWindow."""
        return ExchangeWindow(self, world, rank, slot_bytes, result_bytes)

    def device_alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    # ---- timing
    def timing_reset(self):
        self._check(lib.svb_timing_reset(self.handle))

    def timing(self):
        buf = np.zeros(2 * _lib.SVB_K_COUNT, dtype=np.uint64)
        self._check(lib.svb_timing_get(self.handle, _lib.ptr(buf)))
        ms = buf[:_lib.SVB_K_COUNT].view(np.float64)
        launches = buf[_lib.SVB_K_COUNT:]
        return {name: (float(ms[i]), int(launches[i])) for i, name in enumerate(_lib.KERNEL_NAMES)}

    def pair_stats(self):
        """The last pair(): partitions, cross-haplotype pairs (edit-distance jobs), pairs that needed the exact kernel."""
        v = np.zeros(4, dtype=np.uint64)
        self._check(lib.svb_pair_stats(self.handle, _lib.ptr(v)))
        return {"partitions": int(v[0]), "pairs": int(v[1]), "exact_pairs": int(v[2]), "table_cells": int(v[3])}

    def launch_count(self):
        n = ctypes.c_uint64()
        self._check(lib.svb_launch_count(self.handle, ctypes.byref(n)))
        return n.value

    def mark(self, slot):
        """Record a timing marker on the library's stream."""
        self._check(lib.svb_mark(self.handle, int(slot)))

    def elapsed_ms(self, slot_begin, slot_end):
        ms = ctypes.c_double()
        self._check(lib.svb_elapsed_ms(self.handle, int(slot_begin), int(slot_end), ctypes.byref(ms)))
        return ms.value

    def set_scan_variant(self, variant):
        self._check(lib.svb_set_scan_variant(self.handle, int(variant)))

    def synchronize(self):
        self._check(lib.svb_synchronize(self.handle))
