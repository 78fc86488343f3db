"""Drop-in for the reference's compiled extension module ``MultiScaleDeformableAttention``.

The reference builds a pybind module of this name exporting ``ms_deform_attn_forward`` and
``ms_deform_attn_backward`` (ref: mask2former/modeling/pixel_decoder/ops/src/vision.cpp:18-21,
src/ms_deform_attn.h:25-66, setup.py:60).  This module exports the same two functions with the same
positional signatures, argument meaning and error behaviour (RuntimeError on CPU tensors, on
non-contiguous inputs and on batch % im2col_step != 0), implemented by the hand-written sm_100a
kernels behind the C ABI in ``include/mpformer_b200.h``.

    sys.modules["MultiScaleDeformableAttention"] = mp_former_b200.MultiScaleDeformableAttention

makes the reference's own ``ops/functions/ms_deform_attn_func.py`` run on these kernels unchanged
(see INTEGRATION.md).

Extension to the reference signature: ``spatial_shapes`` may carry a Python attribute
``_mpf_host_shapes`` (tuple of (H, W)); when present the launcher tiles queries spatially for L1
locality.  Results do not depend on it.
"""
import torch

from . import _lib

_FLOATS = (torch.float32, torch.float64)

# Optional per-launch device timing (bench.py's roofline leg): CUDA events recorded on the launching
# stream right around the C-ABI call.  Off by default.
_PROFILE = None


def profile_begin():
    global _PROFILE
    _PROFILE = {"fwd": [], "bwd": []}


def profile_end():
    """Returns {"fwd_ms": [...], "bwd_ms": [...]} and switches profiling off (synchronises)."""
    global _PROFILE
    p, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    return {k + "_ms": [a.elapsed_time(b) for a, b in v] for k, v in (p or {"fwd": [], "bwd": []}).items()}


class _Timed:
    def __init__(self, kind):
        self.kind = kind

    def __enter__(self):
        if _PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.b.record()
            _PROFILE[self.kind].append((self.a, self.b))
        return False


def _host_shapes_array(spatial_shapes, host_shapes):
    import ctypes
    hs = host_shapes if host_shapes is not None else getattr(spatial_shapes, "_mpf_host_shapes", None)
    if hs is None:
        return None
    flat = [int(v) for hw in hs for v in hw]
    return (ctypes.c_int64 * len(flat))(*flat)


def _check_inputs(named):
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    for name, t in named:
        _lib.require_cuda(t, name)


def _dims(value, spatial_shapes, sampling_loc, attn_weight, im2col_step):
    if value.dim() != 4 or sampling_loc.dim() != 6 or attn_weight.dim() != 5:
        raise RuntimeError("ms_deform_attn: expected value [N,S,M,D], sampling_loc [N,Lq,M,L,P,2], "
                           "attn_weight [N,Lq,M,L,P]")
    batch, spatial_size, num_heads, channels = value.shape
    num_levels = spatial_shapes.shape[0]
    num_query, num_point = sampling_loc.shape[1], sampling_loc.shape[4]
    if tuple(sampling_loc.shape) != (batch, num_query, num_heads, num_levels, num_point, 2) or \
            tuple(attn_weight.shape) != (batch, num_query, num_heads, num_levels, num_point):
        raise RuntimeError("ms_deform_attn: inconsistent sampling_loc / attn_weight shapes")
    if spatial_shapes.dtype != torch.int64:
        raise RuntimeError("spatial_shapes / level_start_index must be int64 (as in the reference)")
    step = min(batch, int(im2col_step))
    if step <= 0 or batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")
    if value.dtype not in _FLOATS or sampling_loc.dtype != value.dtype or attn_weight.dtype != value.dtype:
        raise RuntimeError("ms_deform_attn supports float32/float64 with matching dtypes "
                           f"(got {value.dtype}, {sampling_loc.dtype}, {attn_weight.dtype})")
    return batch, spatial_size, num_heads, channels, num_levels, num_query, num_point


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                           im2col_step, host_shapes=None):
    """ref: ops/src/ms_deform_attn.h:25-45 -> cuda/ms_deform_attn_cuda.cu:25-85.
    Returns ``[N, Lq, M*D]`` (freshly allocated, same dtype/device as ``value``)."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes),
                   ("level_start_index", level_start_index), ("sampling_loc", sampling_loc),
                   ("attn_weight", attn_weight)])
    B, S, M, D, L, Lq, P = _dims(value, spatial_shapes, sampling_loc, attn_weight, im2col_step)
    lib = _lib.load()
    out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device), _Timed("fwd"):
        stream = torch.cuda.current_stream().cuda_stream
        if value.dtype == torch.float32:
            hs = _host_shapes_array(spatial_shapes, host_shapes)
            rc = lib.mpf_msda_forward_f32_ex(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                sampling_loc.data_ptr(), attn_weight.data_ptr(), B, S, M, D, L, Lq, P,
                out.data_ptr(), hs, stream)
        else:
            rc = lib.mpf_msda_forward_f64(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                sampling_loc.data_ptr(), attn_weight.data_ptr(), B, S, M, D, L, Lq, P,
                out.data_ptr(), stream)
    _lib.check(rc, "ms_deform_attn_forward")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                            grad_output, im2col_step, host_shapes=None):
    """ref: ops/src/ms_deform_attn.h:47-66 -> cuda/ms_deform_attn_cuda.cu:88-158.
    Returns ``[grad_value, grad_sampling_loc, grad_attn_weight]``."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes),
                   ("level_start_index", level_start_index), ("sampling_loc", sampling_loc),
                   ("attn_weight", attn_weight), ("grad_output", grad_output)])
    B, S, M, D, L, Lq, P = _dims(value, spatial_shapes, sampling_loc, attn_weight, im2col_step)
    if grad_output.dtype != value.dtype or grad_output.numel() != B * Lq * M * D:
        raise RuntimeError("ms_deform_attn_backward: grad_output must be [N, Lq, M*D] of value's dtype")
    lib = _lib.load()
    grad_value = torch.empty_like(value)
    grad_loc = torch.empty_like(sampling_loc)
    grad_aw = torch.empty_like(attn_weight)
    with torch.cuda.device(value.device), _Timed("bwd"):
        stream = torch.cuda.current_stream().cuda_stream
        if value.dtype == torch.float32:
            hs = _host_shapes_array(spatial_shapes, host_shapes)
            rc = lib.mpf_msda_backward_f32_ex(
                grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(),
                level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                B, S, M, D, L, Lq, P, grad_value.data_ptr(), grad_loc.data_ptr(),
                grad_aw.data_ptr(), hs, stream)
        else:
            rc = lib.mpf_msda_backward_f64(
                grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(),
                level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                B, S, M, D, L, Lq, P, grad_value.data_ptr(), grad_loc.data_ptr(),
                grad_aw.data_ptr(), stream)
    _lib.check(rc, "ms_deform_attn_backward")
    return [grad_value, grad_loc, grad_aw]


# ------------------------------------------------------------------------------------------------
# Encoder-fused form (extension to the reference's extension API): consumes the raw projection output
# cat(sampling_offsets(q), attention_weights(q)) and the reference points; softmax and location arithmetic
# run inside the kernel (see include/mpformer_b200.h, mpf_msda_enc_*).
# ------------------------------------------------------------------------------------------------
def enc_supported(value, num_levels, num_points):
    D = value.shape[-1]
    return value.dtype == torch.float32 and num_points == 4 and D in (16, 32, 64) and num_levels <= 4


def _ref_stride(reference_points, batch):
    if reference_points.shape[0] == 1 or reference_points.stride(0) == 0:
        return 0
    if reference_points.shape[0] != batch:
        raise RuntimeError("reference_points batch mismatch")
    return reference_points.stride(0)


def ms_deform_attn_enc_forward(value, spatial_shapes, level_start_index, offsets_logits, reference_points,
                               num_points, host_shapes=None):
    """value [N,S,M,D]; offsets_logits [N,Lq,M*L*P*3]; reference_points [N or 1, Lq, L, 2] -> [N,Lq,M*D]."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("offsets_logits", offsets_logits)])
    B, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq = offsets_logits.shape[1]
    if offsets_logits.shape != (B, Lq, M * L * num_points * 3) or reference_points.shape[1:] != (Lq, L, 2):
        raise RuntimeError("ms_deform_attn_enc: inconsistent shapes")
    ref = reference_points if reference_points[0].is_contiguous() else reference_points.contiguous()
    out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device), _Timed("fwd"):
        rc = _lib.load().mpf_msda_enc_forward_f32(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), offsets_logits.data_ptr(),
            ref.data_ptr(), _ref_stride(ref, B), B, S, M, D, L, Lq, num_points, out.data_ptr(),
            _host_shapes_array(spatial_shapes, host_shapes), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ms_deform_attn_enc_forward")
    return out


def ms_deform_attn_enc_backward(value, spatial_shapes, level_start_index, offsets_logits, reference_points,
                                grad_output, num_points, host_shapes=None):
    """Returns [grad_value, grad_offsets_logits]."""
    _check_inputs([("value", value), ("offsets_logits", offsets_logits), ("grad_output", grad_output)])
    B, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq = offsets_logits.shape[1]
    ref = reference_points if reference_points[0].is_contiguous() else reference_points.contiguous()
    grad_value = torch.empty_like(value)
    grad_ow = torch.empty_like(offsets_logits)
    with torch.cuda.device(value.device), _Timed("bwd"):
        rc = _lib.load().mpf_msda_enc_backward_f32(
            grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            offsets_logits.data_ptr(), ref.data_ptr(), _ref_stride(ref, B), B, S, M, D, L, Lq, num_points,
            grad_value.data_ptr(), grad_ow.data_ptr(), _host_shapes_array(spatial_shapes, host_shapes),
            torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ms_deform_attn_enc_backward")
    return [grad_value, grad_ow]
