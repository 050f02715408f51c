"""TEST INFRASTRUCTURE: builds tests/hostcheck/hostcheck.cpp (host compile of the kernels' cores)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "hostcheck.so")


def load():
    src = os.path.join(_HERE, "hostcheck.cpp")
    deps = [src] + [os.path.join(_HERE, "..", "..", "svim_asm_b200", "csrc", f) for f in ("linkage.cuh", "walk.cuh", "edit_core.cuh", "inflate_core.cuh", "vcf_core.cuh", "wfa_core.cuh")]
    if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", _SO, src])
    lib = ctypes.CDLL(_SO)
    vp = ctypes.c_void_p
    lib.hc_cluster_labels.restype = ctypes.c_int
    lib.hc_cluster_labels.argtypes = [vp, ctypes.c_int, ctypes.c_double, vp]
    lib.hc_labels_determined.restype = ctypes.c_int
    lib.hc_labels_determined.argtypes = [vp, vp, ctypes.c_int, ctypes.c_double, vp]
    lib.hc_walk.restype = ctypes.c_int
    lib.hc_walk.argtypes = [vp, ctypes.c_int, ctypes.c_int32, ctypes.c_uint32, vp, vp, vp, ctypes.c_int32,
                            ctypes.c_uint32, ctypes.c_uint32, vp, ctypes.c_int]
    lib.hc_myers.restype = ctypes.c_longlong
    lib.hc_myers.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_longlong, ctypes.c_int]
    lib.hc_myers_window.restype = ctypes.c_longlong
    lib.hc_myers_window.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_longlong, ctypes.c_uint, ctypes.c_uint]
    lib.hc_win_kmax.restype = ctypes.c_uint
    lib.hc_win_kmax.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint]
    lib.hc_myers_window_split.restype = ctypes.c_longlong
    lib.hc_myers_window_split.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_longlong, ctypes.c_uint]
    lib.hc_wfa.restype = ctypes.c_longlong
    lib.hc_wfa.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_longlong, ctypes.c_int]
    lib.hc_wfa_bidir.restype = ctypes.c_longlong
    lib.hc_wfa_bidir.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_longlong, ctypes.c_int]
    lib.hc_wfa_rounds.restype = ctypes.c_longlong
    lib.hc_wfa_rounds.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_longlong, ctypes.c_int]
    lib.hc_inflate.restype = ctypes.c_int
    lib.hc_inflate.argtypes = [ctypes.c_char_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint]
    lib.hc_inflate_fast.restype = ctypes.c_int
    lib.hc_inflate_fast.argtypes = [ctypes.c_char_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint]
    lib.hc_vcf_body.restype = ctypes.c_longlong
    lib.hc_vcf_body.argtypes = [vp, vp, ctypes.c_longlong, vp, vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_uint, vp, ctypes.c_longlong]
    return lib
