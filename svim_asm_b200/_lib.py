"""ctypes binding of libsvimasm_b200.so (the C ABI declared in include/svimasm_b200.h).

There is no fallback: if the shared library has not been built, importing this module raises.
Build it with `python -m svim_asm_b200.build` (nvcc, sm_100a).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVIM_ASM_B200_LIB", os.path.join(_HERE, "libsvimasm_b200.so"))   # override: tuning builds only

if not os.path.exists(LIB_PATH):
    raise ImportError("libsvimasm_b200.so is missing at %s -- build it with `python -m svim_asm_b200.build`; "
                      "this package has no CPU fallback" % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

c_i32, c_u32, c_i64, c_u64 = ctypes.c_int32, ctypes.c_uint32, ctypes.c_int64, ctypes.c_uint64
c_void_p, c_char_p, c_int, c_double = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_double
P = ctypes.POINTER

SVB_K_COUNT = 8
KERNEL_NAMES = ("cigar_scan", "segment_walk", "merge", "sort", "edit_distance", "cluster", "cigar_scan_finalize", "vcf_body")

VCF_ENTRY_DTYPE = np.dtype([("row", "<u4"), ("id", "<u4"), ("mode", "<u4")])      # svb_vcf_entry

PARAMS_DTYPE = np.dtype([(n, "<i4") for n in (
    "min_mapq", "min_sv_size", "max_sv_size", "query_gap_tolerance", "query_overlap_tolerance",
    "reference_gap_tolerance", "reference_overlap_tolerance", "partition_max_distance", "max_edit_distance")])

HDR_DTYPE = np.dtype([("tid", "<i4"), ("pos", "<i4"), ("flag", "<u2"), ("mapq", "u1"), ("reserved0", "u1"),
                      ("n_cigar", "<u4"), ("cigar_off", "<u8"), ("l_seq", "<u4"), ("sa_first", "<u4")])
SEG_DTYPE = np.dtype([("tid", "<i4"), ("pos", "<i4"), ("is_reverse", "u1"), ("mapq", "u1"), ("reserved0", "<u2"),
                      ("ref_end", "<i4"), ("q_astart", "<i4"), ("q_aend", "<i4"), ("read_len", "<i4"),
                      ("reserved1", "<i4")])
ROW_DTYPE = np.dtype([("type", "u1"), ("flags", "u1"), ("genotype", "u1"), ("hap", "u1"),
                      ("src_tid", "<i4"), ("src_start", "<i4"), ("src_end", "<i4"),
                      ("dst_tid", "<i4"), ("dst_start", "<i4"), ("dst_end", "<i4"),
                      ("copies", "<i4"), ("aln_idx", "<u4"), ("seq_pos", "<u4"), ("seq_len", "<u4"),
                      ("mate_aln", "<u4"), ("ordinal", "<u8"), ("reserved0", "<u8")])
assert HDR_DTYPE.itemsize == 32 and SEG_DTYPE.itemsize == 32 and ROW_DTYPE.itemsize == 64

# every symbol include/svimasm_b200.h and include/svimasm_b200_debug.h declare: name -> (restype, argtypes)
SIGNATURES = {
    "svb_abi_version": (c_int, []),
    "svb_create": (c_int, [c_int, P(c_void_p)]),
    "svb_destroy": (None, [c_void_p]),
    "svb_last_error": (c_char_p, [c_void_p]),
    "svb_synchronize": (c_int, [c_void_p]),
    "svb_timing_reset": (c_int, [c_void_p]),
    "svb_timing_get": (c_int, [c_void_p, c_void_p]),
    "svb_set_scan_variant": (c_int, [c_void_p, c_int]),
    "svb_launch_count": (c_int, [c_void_p, P(c_u64)]),
    "svb_mark": (c_int, [c_void_p, c_int]),
    "svb_elapsed_ms": (c_int, [c_void_p, c_int, c_int, P(c_double)]),
    "svb_bam_open": (c_int, [c_char_p, c_int, P(c_void_p), c_char_p, c_int]),
    "svb_bam_close": (None, [c_void_p]),
    "svb_bam_n_records": (c_i64, [c_void_p]),
    "svb_bam_n_ops_padded": (c_i64, [c_void_p]),
    "svb_bam_n_segments": (c_i64, [c_void_p]),
    "svb_bam_n_contigs": (c_i32, [c_void_p]),
    "svb_bam_contig_name": (c_char_p, [c_void_p, c_i32]),
    "svb_bam_contig_lengths": (c_void_p, [c_void_p]),
    "svb_bam_sort_order": (c_char_p, [c_void_p]),
    "svb_bam_headers": (c_void_p, [c_void_p]),
    "svb_bam_cigar": (c_void_p, [c_void_p]),
    "svb_bam_segments": (c_void_p, [c_void_p]),
    "svb_bam_sa_count": (c_void_p, [c_void_p]),
    "svb_bam_seq4": (c_void_p, [c_void_p]),
    "svb_bam_seq_offsets": (c_void_p, [c_void_p]),
    "svb_bam_query_name": (c_char_p, [c_void_p, c_i64]),
    "svb_bam_sa_text": (c_char_p, [c_void_p, c_i64]),
    "svb_parse_sa": (c_int, [c_char_p, P(c_char_p), c_i32, c_void_p, c_i32]),
    "svb_bam_open_device": (c_int, [c_void_p, c_char_p, c_int, c_void_p, P(c_void_p), P(c_void_p), c_char_p, c_int]),
    "svb_bam_materialize_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "svb_bam_device_timings": (c_int, [c_void_p, c_void_p]),
    "svb_load_records": (c_int, [c_void_p, c_void_p, c_u32, c_void_p, c_u64, c_void_p, c_void_p, c_u32, c_void_p,
                                 c_void_p, c_i32, P(c_void_p)]),
    "svb_records_free": (None, [c_void_p]),
    "svb_records_set_sequences": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "svb_collect2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, P(c_void_p), P(c_void_p)]),
    "svb_records_map_sequences_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "svb_collect": (c_int, [c_void_p, c_void_p, c_void_p, c_int, P(c_void_p)]),
    "svb_cigar_indel": (c_int, [c_void_p, c_void_p, c_u32, c_i32, c_void_p, c_u32, P(c_u32)]),
    "svb_ref_load": (c_int, [c_void_p, c_void_p, c_void_p, c_i32, P(c_void_p)]),
    "svb_ref_load_fasta": (c_int, [c_void_p, c_char_p, c_void_p, c_i32, P(c_void_p)]),
    "svb_ref_to_host": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, P(c_u64), c_void_p]),
    "svb_ref_free": (None, [c_void_p]),
    "svb_pair": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, P(c_void_p)]),
    "svb_edit_distance": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_void_p]),
    "svb_edit_distance_bounded": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_i64, c_void_p]),
    "svb_form_partitions": (c_int, [c_void_p, c_void_p, c_u32, c_i64, c_void_p, c_void_p, P(c_u32)]),
    "svb_pair_stats": (c_int, [c_void_p, c_void_p]),
    "svb_bgzf_member_table_check": (c_int, [c_char_p, c_int, c_void_p]),
    "svb_cluster_labels": (c_int, [c_void_p, c_void_p, c_void_p, c_u32, c_double, c_void_p]),
    "svb_table_size": (c_i64, [c_void_p]),
    "svb_table_to_host": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, P(c_u64)]),
    "svb_table_from_host": (c_int, [c_void_p, c_void_p, c_u64, P(c_void_p)]),
    "svb_table_export": (c_int, [c_void_p, c_void_p, c_void_p, c_u64]),
    "svb_table_import": (c_int, [c_void_p, c_void_p, c_u64, P(c_void_p)]),
    "svb_table_free": (None, [c_void_p]),
    "svb_table_gather_sequences": (c_int, [c_void_p, c_void_p, c_void_p]),
    "svb_vcf_body": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_char_p, c_void_p, c_i32, c_void_p, c_u64, c_u32, P(c_void_p), P(c_u64)]),
    "svb_table_attach_sequences_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "svb_table_pool_to_host": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, c_void_p, P(c_u64)]),
    "svb_stream": (c_void_p, [c_void_p]),
    "svb_device_alloc": (c_int, [c_void_p, c_u64, c_void_p]),
    "svb_device_free": (None, [c_void_p, c_void_p]),
    "svb_records_set_global_index": (c_int, [c_void_p, c_void_p, c_void_p]),
    "svb_table_remap_records": (c_int, [c_void_p, c_void_p, c_void_p]),
    "svb_exchange_sizes": (c_int, [c_void_p, c_void_p, c_void_p]),
    "svb_exchange_bytes": (c_u64, [c_void_p]),
    "svb_exchange_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_u64]),
    "svb_exchange_unpack": (c_int, [c_void_p, c_void_p, c_u64, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "svb_table_device_rows": (c_void_p, [c_void_p]),
    "svb_exchange_create": (c_int, [c_void_p, c_int, c_int, c_u64, c_u64, P(c_void_p)]),
    "svb_exchange_handle": (c_int, [c_void_p, c_void_p, c_void_p]),
    "svb_exchange_open": (c_int, [c_void_p, c_void_p, c_void_p]),
    "svb_exchange_destroy": (None, [c_void_p, c_void_p]),
    "svb_exchange_share": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, P(c_void_p), P(c_void_p)]),
    "svb_exchange_gather_paired": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, P(c_void_p)]),
    "svb_table_set_pool_from_host": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, c_void_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


def ptr(array):
    """Raw pointer of a C-contiguous numpy array (None for empty / None)."""
    if array is None or array.size == 0:
        return None
    assert array.flags["C_CONTIGUOUS"]
    return array.ctypes.data
