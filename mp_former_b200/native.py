"""Python launchers for the C-ABI kernels other than MSDeformAttn (tensor in, tensor out; the caller
owns autograd).  Every function requires CUDA fp32 tensors and raises RuntimeError otherwise."""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t, name):
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    return t


def split_tf32(x):
    """x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi); returns (hi, lo), same shape."""
    x = _f32c(x, "x").contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.load().mpf_split_tf32(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "split_tf32")
    return hi, lo


def gemm_tf32x3(a, b_hi, b_lo, bias=None, relu=False, transpose_c=False):
    """C[b] = A[b] @ B[b]^T (+bias)(ReLU).  a: [M,K] or [batch,M,K]; b_hi/b_lo: [N,K] or [batch,N,K]
    (K contiguous, row strides multiples of 4).  Returns [.., M, N] or, with ``transpose_c``,
    [.., N, M]."""
    a = _f32c(a, "a")
    squeeze = a.dim() == 2
    if squeeze:
        a, b_hi, b_lo = a[None], b_hi[None], b_lo[None]
    if a.stride(2) != 1 or a.stride(1) % 4 or (a.shape[0] > 1 and a.stride(0) % 4):
        a = a.contiguous()
    if not (b_hi.is_contiguous() and b_lo.is_contiguous()):
        b_hi, b_lo = b_hi.contiguous(), b_lo.contiguous()
    batch, M, K = a.shape
    N = b_hi.shape[1]
    if b_hi.shape != (batch, N, K) or b_lo.shape != b_hi.shape:
        raise RuntimeError(f"gemm_tf32x3: shape mismatch a={tuple(a.shape)} b={tuple(b_hi.shape)}")
    if bias is not None:
        bias = _f32c(bias, "bias").contiguous()
    out = torch.empty((batch, N, M) if transpose_c else (batch, M, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.load().mpf_gemm_tf32x3(
            a.data_ptr(), a.stride(1), a.stride(0) if batch > 1 else M * a.stride(1),
            b_hi.data_ptr(), b_lo.data_ptr(), K, N * K,
            None if bias is None else bias.data_ptr(), out.data_ptr(),
            M if transpose_c else N, out.stride(0), batch, M, N, K, int(relu), int(transpose_c), _stream())
    _lib.check(rc, "gemm_tf32x3")
    return out[0] if squeeze else out
