"""CPU checks: the library builds for sm_100a, loads, and exports every symbol the header declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_library):
    from svim_asm_b200 import _lib
    declared = set()
    for name in ("svimasm_b200.h", "svimasm_b200_debug.h"):          # the drop-in surface + the measurement / test hooks
        declared |= set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", open(os.path.join(ROOT, "include", name)).read()))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), "library does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib.svb_abi_version() == 1


def test_no_cpu_fallback(built_library):
    import ctypes
    import torch
    from svim_asm_b200 import _lib
    if torch.cuda.is_available():
        return
    handle = ctypes.c_void_p()
    assert _lib.lib.svb_create(0, ctypes.byref(handle)) != 0       # fails loudly without a device
