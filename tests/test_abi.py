"""The C-ABI library loads on a CPU-only box and exports exactly what include/mpformer_b200.h
declares; the ctypes table in mp_former_b200/_lib.py covers every declared symbol."""
import ctypes
import os
import re

from mp_former_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpformer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "mpf_msda_forward_f32" in syms and "mpf_msda_backward_f32" in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported by the .so"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_abi_version_and_error_channel():
    lib = _lib.load()
    assert lib.mpf_abi_version() == 1
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = lib.mpf_msda_forward_f32(None, None, None, None, None, 1, 1, 1, 4, 1, 1, 4, None, None)
    assert rc == -1
    assert b"null pointer" in lib.mpf_last_error()
    rc = lib.mpf_msda_forward_f32(None, None, None, None, None, 0, 1, 1, 4, 1, 1, 4, None, None)
    assert rc == -1 and b"positive" in lib.mpf_last_error()
