"""ctypes binding of the C-ABI shared library ``libmpformer_b200.so`` (see include/mpformer_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc -shared``; no torch headers).
There is NO fallback: if the library is missing or a call fails, a ``RuntimeError`` is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpformer_b200.so")

_c_int = ctypes.c_int
_c_vp = ctypes.c_void_p
_c_ll = ctypes.c_longlong

# name -> (restype, argtypes).  Must list every symbol declared in include/mpformer_b200.h;
# tests/test_abi.py parses the header and checks this table and the .so against it.
_MSDA_FWD = [_c_vp] * 5 + [_c_int] * 7 + [_c_vp, _c_vp]
_MSDA_FWD_EX = [_c_vp] * 5 + [_c_int] * 7 + [_c_vp, _c_vp, _c_vp]
_MSDA_BWD = [_c_vp] * 6 + [_c_int] * 7 + [_c_vp] * 3 + [_c_vp]
_MSDA_BWD_EX = [_c_vp] * 6 + [_c_int] * 7 + [_c_vp] * 3 + [_c_vp, _c_vp]

SIGNATURES = {
    "mpf_abi_version": (_c_int, []),
    "mpf_last_error": (ctypes.c_char_p, []),
    "mpf_launch_count": (ctypes.c_uint64, []),
    "mpf_msda_forward_f32": (_c_int, _MSDA_FWD),
    "mpf_msda_forward_f32_ex": (_c_int, _MSDA_FWD_EX),
    "mpf_msda_forward_f64": (_c_int, _MSDA_FWD),
    "mpf_msda_backward_f32": (_c_int, _MSDA_BWD),
    "mpf_msda_backward_f32_ex": (_c_int, _MSDA_BWD_EX),
    "mpf_msda_backward_f64": (_c_int, _MSDA_BWD),
    "mpf_msda_enc_forward_f32": (_c_int, [_c_vp] * 5 + [_c_ll] + [_c_int] * 7 + [_c_vp, _c_vp, _c_vp]),
    "mpf_msda_enc_backward_f32": (_c_int, [_c_vp] * 6 + [_c_ll] + [_c_int] * 7 + [_c_vp, _c_vp, _c_vp, _c_vp]),
    "mpf_msda_set_staged": (_c_int, [_c_int]),
    "mpf_gemm_bf16x3_set_pair_mode": (_c_int, [_c_int]),
    "mpf_self_attn_fwd_f32": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    "mpf_self_attn_bwd_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "mpf_topk_gather_rows_f32": (_c_int, [_c_vp, _c_int, _c_int, _c_int, _c_vp, _c_int, _c_vp, _c_vp]),
    "mpf_mask_loss_rows_fwd_f32": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp]),
    "mpf_mask_loss_rows_bwd_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp]),
    "mpf_split_tf32": (_c_int, [_c_vp, _c_vp, _c_vp, ctypes.c_longlong, _c_vp]),
    "mpf_gemm_tf32x3": (_c_int, [_c_vp, ctypes.c_longlong, ctypes.c_longlong, _c_vp, _c_vp,
                                 ctypes.c_longlong, ctypes.c_longlong, _c_vp, _c_vp, ctypes.c_longlong,
                                 ctypes.c_longlong, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "mpf_gemm_tf32x3_ex": (_c_int, [_c_vp, _c_ll, _c_ll, _c_vp, _c_vp, _c_ll, _c_ll, _c_vp, _c_vp, _c_vp,
                                    _c_ll, _c_ll, _c_vp, _c_ll, _c_int, _c_int, ctypes.c_float,
                                    _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "mpf_gemm_tf32x3_general": (_c_int, [_c_vp, _c_int, _c_ll, _c_ll, _c_vp, _c_vp, _c_int, _c_ll, _c_ll, _c_vp,
                                         _c_vp, _c_vp, _c_ll, _c_ll, _c_vp, _c_ll, _c_int, _c_int, _c_vp, _c_ll,
                                         ctypes.c_float,
                                         _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "mpf_split_bf16": (_c_int, [_c_vp, _c_vp, _c_vp, _c_ll, _c_vp]),
    "mpf_split_weights_f32": (_c_int, [_c_vp, _c_int, _c_int, _c_vp]),
    "mpf_gemm_bf16x3": (_c_int, [_c_vp, _c_ll, _c_ll, _c_vp, _c_vp, _c_ll, _c_ll, _c_vp, _c_vp, _c_vp, _c_ll, _c_ll,
                                 _c_vp, _c_ll, _c_int, _c_int, _c_vp, _c_ll, ctypes.c_float,
                                 _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "mpf_gemm_bf16x3_relubits": (_c_int, [_c_vp, _c_ll, _c_vp, _c_vp, _c_ll, _c_vp, _c_vp, _c_ll, _c_vp, _c_ll,
                                          _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    "mpf_transpose_split_bf16": (_c_int, [_c_vp] * 3 + [_c_int, _c_ll, _c_int, _c_vp]),
    "mpf_gemm_bf16x3_tn": (_c_int, [_c_vp, _c_ll, _c_ll, _c_vp, _c_ll, _c_ll, _c_vp, _c_ll, _c_ll] + [_c_int] * 5 + [_c_vp]),
    "mpf_gemm_bf16x3_tn_ex": (_c_int, [_c_vp, _c_ll, _c_ll, _c_vp, _c_ll, _c_ll, _c_vp, _c_ll, _c_ll] + [_c_int] * 6 + [_c_vp]),
    "mpf_gemm_bf16x3_tn_colsum": (_c_int, [_c_vp, _c_ll, _c_ll, _c_vp, _c_ll, _c_ll, _c_vp, _c_ll, _c_ll] + [_c_int] * 5
                                  + [_c_vp, _c_ll, _c_vp]),
    "mpf_add_layernorm_partials": (_c_int, [_c_ll]),
    "mpf_add_layernorm_fwd_f32": (_c_int, [_c_vp] * 4 + [ctypes.c_float, _c_ll, _c_int] + [_c_vp] * 4),
    "mpf_add_layernorm_bwd_f32": (_c_int, [_c_vp] * 6 + [_c_ll, _c_int] + [_c_vp] * 3),
    "mpf_colsum_f32": (_c_int, [_c_vp, _c_ll, _c_int, _c_ll, _c_vp, _c_vp]),
    "mpf_groupnorm_cl_fwd_f32": (_c_int, [_c_vp] * 3 + [ctypes.c_float, _c_int, _c_ll, _c_int, _c_int, _c_int] + [_c_vp] * 5),
    "mpf_groupnorm_cl_bwd_f32": (_c_int, [_c_vp] * 6 + [_c_int, _c_ll, _c_int, _c_int, _c_int] + [_c_vp] * 4),
    "mpf_groupnorm_nchw2cl_fwd_f32": (_c_int, [_c_vp] * 3 + [ctypes.c_float, _c_int, _c_ll, _c_int, _c_int, _c_int] + [_c_vp] * 5),
    "mpf_groupnorm_nchw2cl_bwd_f32": (_c_int, [_c_vp] * 6 + [_c_int, _c_ll, _c_int, _c_int, _c_int] + [_c_vp] * 4),
    "mpf_conv3x3_cl_bf16x3": (_c_int, [_c_vp] * 5 + [_c_int] * 6 + [_c_vp]),
    "mpf_conv3x3_cl_wgrad_bf16x3": (_c_int, [_c_vp] * 3 + [_c_int] * 6 + [_c_vp]),
    "mpf_upsample2x_add_cl_fwd_f32": (_c_int, [_c_vp] * 2 + [_c_int] * 4 + [_c_vp] * 2),
    "mpf_upsample2x_cl_bwd_f32": (_c_int, [_c_vp] + [_c_int] * 4 + [_c_vp] * 2),
    "mpf_upsample2x_add_nchw_fwd_f32": (_c_int, [_c_vp] * 2 + [_c_int] * 4 + [_c_vp] * 2),
    "mpf_upsample2x_add_nchw_bwd_f32": (_c_int, [_c_vp] + [_c_int] * 4 + [_c_vp] * 3),
    "mpf_attn_mask_bits_f32": (_c_int, [_c_vp, _c_ll] + [_c_int] * 5 + [_c_vp, _c_int, _c_vp]),
    "mpf_pack_bool_bits": (_c_int, [_c_vp, _c_int, _c_int, _c_vp, _c_int, _c_vp]),
    "mpf_gt_mask_area_bits": (_c_int, [_c_vp] + [_c_int] * 5 + [_c_vp, _c_int, _c_vp]),
    "mpf_masked_xattn_fwd_f32": (_c_int, [_c_vp] * 10 + [_c_int] * 6 + [_c_vp]),
    "mpf_masked_xattn_bwd_f32": (_c_int, [_c_vp] * 21 + [_c_int] * 7 + [_c_vp]),
    "mpf_masked_xattn_fwd_f32_ex": (_c_int, [_c_vp] * 10 + [_c_int] * 7 + [_c_vp] * 3),
    "mpf_masked_xattn_bwd_f32_ex": (_c_int, [_c_vp] * 21 + [_c_int] * 8 + [_c_vp] * 2),
    "mpf_match_cost_workspace_bytes": (_c_ll, [_c_int] * 5),
    "mpf_match_cost_f32": (_c_int, [_c_vp, _c_ll, _c_ll, _c_int, _c_vp, _c_ll, _c_ll, _c_int, _c_int, _c_vp, _c_int,
                                    _c_int, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_int, _c_int, _c_int,
                                    ctypes.c_float, ctypes.c_float, ctypes.c_float, _c_vp, _c_ll, _c_vp, _c_vp]),
    "mpf_sample_shared_points_f32": (_c_int, [_c_vp, _c_ll, _c_ll, _c_int, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_int,
                                              _c_int, _c_int, _c_vp, _c_vp]),
    "mpf_match_cost_presampled_f32": (_c_int, [_c_vp, _c_ll, _c_ll, _c_int, _c_vp, _c_int, _c_int, _c_vp, _c_int,
                                               _c_int, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_int, _c_int,
                                               _c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, _c_vp, _c_ll,
                                               _c_vp, _c_vp]),
    "mpf_lsap_f32": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp]),
    "mpf_point_sample_rows": (_c_int, [_c_vp, _c_int, _c_int, _c_int, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "mpf_point_sample_rows_bwd_f32": (_c_int, [_c_vp, _c_int, _c_int, _c_vp, _c_int, _c_int, _c_vp, _c_vp]),
    "mpf_instance_masks_blocks": (_c_int, [_c_int, _c_int]),
    "mpf_instance_masks_f32": (_c_int, [_c_vp, _c_ll, _c_int, _c_int, _c_vp] + [_c_int] * 7 + [_c_vp, _c_int, _c_vp, _c_vp]),
}

_lib = None


def load():
    """Loads (once) and returns the ctypes handle; raises RuntimeError when the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"mp_former_b200: native library not found at {LIB_PATH}. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "There is no CPU / PyTorch fallback for this path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => stale .so; let it surface
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().mpf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"mp_former_b200.{what} failed (code {rc}): {msg}")


def launch_count():
    return int(load().mpf_launch_count())


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(
            f"mp_former_b200: `{name}` must be a CUDA tensor (got {t.device}); this path has no "
            "CPU implementation (the reference op has none either: "
            "ops/src/cpu/ms_deform_attn_cpu.cpp:22-45)")
