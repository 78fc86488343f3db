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


def test_matcher_and_sampler_entry_points_validate_arguments_without_a_gpu():
    """mpf_match_cost_* / mpf_lsap_f32 / mpf_point_sample_rows*: size and pointer checks precede any CUDA call."""
    lib = _lib.load()
    # workspace: S point splits x (3 sums per (query, target) + one per (image, query) + one per target)
    B, Q, ntot, nmax, P = 16, 100, 156, 20, 12544
    ws = lib.mpf_match_cost_workspace_bytes(B, Q, ntot, nmax, P)
    qtiles, ttiles, nchunks = 2, 1, 196                       # 64 queries x 32 targets per CTA, chunks of 64 points
    S = min(nchunks, max(1, 3 * 148 // (B * qtiles * ttiles)))  # one wave of 148 SMs x 3 resident CTAs
    assert ws == (S * Q * ntot * 3 + S * B * Q + S * ntot + 4) * 4
    assert lib.mpf_match_cost_workspace_bytes(0, Q, ntot, nmax, P) == -1
    args = [None, 0, 0, 81, None, 0, 0, 256, 256, None, 0, 1024, 1024, None, None, ntot, nmax, None, B, Q, P,
            1.0, 1.0, 1.0, None, 0, None, None]
    assert lib.mpf_match_cost_f32(*args) == -1 and b"null pointer" in lib.mpf_last_error()
    bad = list(args)
    bad[16] = ntot + 1                                          # max_targets > total_targets
    assert lib.mpf_match_cost_f32(*bad) == -1 and b"target counts" in lib.mpf_last_error()
    none = list(args)
    none[15] = none[16] = 0                                     # nothing to match: success, no launch
    assert lib.mpf_match_cost_f32(*none) == 0
    assert lib.mpf_lsap_f32(None, None, B, Q, nmax, None, None, None, None) == -1
    assert lib.mpf_lsap_f32(None, None, B, Q, 0, None, None, None, None) == 0
    assert lib.mpf_lsap_f32(None, None, B, 5000, nmax, None, None, None, None) == -1
    assert lib.mpf_point_sample_rows(None, 0, 64, 64, None, 0, 100, 0, None, None) == 0      # no rows
    assert lib.mpf_point_sample_rows(None, 0, 64, 64, None, 3, 100, 0, None, None) == -1
    assert lib.mpf_point_sample_rows_bwd_f32(None, 64, 64, None, 3, 100, None, None) == -1
    assert lib.mpf_point_sample_rows_bwd_f32(None, 0, 64, None, 3, 100, None, None) == -1 and \
        b"bad sizes" in lib.mpf_last_error()
