"""Device ops of the hot path, each with exactly one implementation.

Everything here runs forward AND backward on this package's hand-written sm_100a kernels through the C ABI:
``linear`` / ``ffn`` / ``encoder_layer`` / ``mask_logits`` / ``conv1x1_nchw_to_cl`` / ``conv3x3_cl`` on the tcgen05
split-precision GEMMs (bf16x3: K-major "NT" kernel and token-reduction "TN" kernel; 3xTF32 for the remaining
mixed-major products), ``attn_mask_from_logits`` on the bit-packed mask kernels, ``masked_cross_attention`` on the
fused attention forward and the two backward kernels that recompute the probabilities from the saved log-sum-exp,
``add_layer_norm`` / ``group_norm_*`` / ``upsample2x_add_*`` on the row-wise and FPN-stage kernels.
``grad_fanout`` / ``collect_mask_heads`` are autograd plumbing that lets the ten prediction heads share one batched
backward.  ``self_attention`` over the few hundred queries runs its projections on the GEMMs and its [Q x Q] core on
the fp32 kernels of csrc/self_attn.cu; ``conv2d_fp32`` pins the library convolution used for
uncovered geometries to true fp32 in both directions.  Ops whose geometry may leave the native kernels record the
route they took in ``ROUTES``.

Inputs must be CUDA fp32 tensors; there is no CPU path.
"""
import collections
import logging
import math
import os

import torch
import torch.nn.functional as F

from . import _lib, native

NATIVE_OPS = {"ms_deform_attn_forward", "ms_deform_attn_backward", "linear fwd+bwd (bf16x3 tcgen05 GEMM, TMA in/out)",
              "conv3x3 fwd+bwd (bf16x3 tcgen05 GEMM, taps as shifted TMA boxes)",
              "mask_logits fwd+bwd (bf16x3 / 3xTF32 tcgen05 GEMM)", "attn_mask_bits", "gt_mask_area_bits",
              "masked_cross_attention fwd+bwd (tcgen05, key-split for small batches)",
              "self_attention core fwd+bwd (fp32 CUDA cores)", "topk_gather_rows (radix select)",
              "match_cost / lsap (Hungarian matcher)", "point_sample_rows fwd+bwd"}

# fp32 ``sigmoid(x) < 0.5`` as evaluated by the reference (1/(1+exp(-x)) with a correctly rounded exp)
# is EXACTLY ``x <= -0x1.7ffffep-23``: for -1.788e-7 < x < 0 the sigmoid rounds to 0.5 and the key stays
# unmasked.  Verified exhaustively over every float32 in [-2.4e-7, -1.2e-7] against torch CPU
# (tests/test_host_logic_cpu.py::test_mask_threshold_equals_sigmoid_rule).
MASK_LOGIT_THRESHOLD = float.fromhex("-0x1.7ffffep-23")

LOG2E = 1.4426950408889634

# ------------------------------------------------------------------------------------------------
# route accounting: which implementation ran
# ------------------------------------------------------------------------------------------------
# Every op whose geometry may fall outside what the hand-written kernels cover records the route it took:
# ROUTES["<op>:native"] / ROUTES["<op>:library"] count calls; the first library-route call of each (op, reason) is
# logged once; MPF_STRICT_NATIVE=1 turns a library route into an error (CI for the bench geometries: nothing of the
# recipe may leave the native path).  bench.py prints the counters as `impl_notes.routes`.
ROUTES = collections.Counter()
_ROUTE_SEEN = set()
STRICT_NATIVE = bool(os.environ.get("MPF_STRICT_NATIVE"))


def _route(op, native, why=""):
    ROUTES[f"{op}:{'native' if native else 'library'}"] += 1
    if not native:
        if STRICT_NATIVE:
            raise RuntimeError(f"mp_former_b200.ops.{op}: geometry outside the native kernels ({why}) and "
                               "MPF_STRICT_NATIVE is set")
        if (op, why) not in _ROUTE_SEEN:
            _ROUTE_SEEN.add((op, why))
            logging.getLogger("mp_former_b200").warning("%s: library (ATen / cuBLAS / cuDNN) route: %s", op, why)
    return native


def _cuda_only(t, name):
    _lib.require_cuda(t, name)


# ------------------------------------------------------------------------------------------------
# linear
# ------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x2 = x.reshape(-1, x.shape[-1])
        w_hi, w_lo = native.split_b(weight)
        y = native.gemm(x2, w_hi, w_lo, bias, relu=relu)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x2, weight, y if relu else None)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, weight, y = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        if ctx.relu:
            g2 = g2 * (y > 0)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            n_out = weight.shape[0]
            if n_out % 32 == 0:
                wt_hi, wt_lo = native.split_bt(weight)
                gx = native.gemm(g2.contiguous(), wt_hi, wt_lo)
            else:
                # e.g. class_embed (81 outputs): the reduction dimension of the input-gradient GEMM is zero-padded to a
                # multiple of 32 (exact) instead of leaving the path for a library matmul
                pad = (-n_out) % 32
                wt_hi, wt_lo = native.split_b(F.pad(weight, (0, 0, 0, pad)).t().contiguous())
                gx = native.gemm(F.pad(g2, (0, pad)), wt_hi, wt_lo)
            ROUTES["linear.input_grad:native"] += 1
            gx = gx.view(*gy.shape[:-1], weight.shape[1])
        if ctx.needs_input_grad[1]:
            if g2.shape[1] % 4 == 0 and x2.shape[1] % 4 == 0:
                if ctx.has_bias and ctx.needs_input_grad[2]:    # dW = dY^T X and db = dY^T 1 from one pass over dY
                    gw, gb = native.matmul_tn(g2.contiguous(), x2, with_colsum=True)
                else:
                    gw = native.matmul_tn(g2.contiguous(), x2)  # dW = dY^T X on the tensor cores, no transposes
                ROUTES["linear.weight_grad:native"] += 1
            elif x2.shape[1] % 4 == 0:
                pad = (-g2.shape[1]) % 4                         # zero columns of dY -> zero rows of dW, sliced away
                gw = native.matmul_tn(F.pad(g2, (0, pad)), x2)[:g2.shape[1]]
                ROUTES["linear.weight_grad:native"] += 1
            else:
                _route("linear.weight_grad", False, f"in_features {x2.shape[1]} % 4 != 0")
                gw = g2.t() @ x2
        if gb is None and ctx.has_bias and ctx.needs_input_grad[2]:
            gb = native.colsum(g2)
        return gx, gw, gb, None


def linear(x, weight, bias=None, relu=False):
    """y = x W^T + b (optionally ReLU) in fp32-accurate 3xTF32 on the tensor cores.
    Replaces nn.Linear at ref ops/modules/ms_deform_attn.py:98,102-103,124 and
    pixel_decoder/msdeformattn.py:116-120."""
    _cuda_only(x, "x")
    if not _route("linear", x.shape[-1] % 32 == 0, f"in_features {x.shape[-1]} % 32 != 0"):
        y = F.linear(x, weight, bias)
        return F.relu(y) if relu else y
    return _Linear.apply(x, weight, bias, relu)


class fp32_math:
    """Context manager: cuDNN / cuBLAS library calls inside run in true fp32 (TF32 off)."""

    def __enter__(self):
        self._c = torch.backends.cudnn.allow_tf32
        self._m = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cudnn.allow_tf32 = self._c
        torch.backends.cuda.matmul.allow_tf32 = self._m
        return False


class _ConvFp32(torch.autograd.Function):
    """Library convolution (the pixel decoder's 3x3 output conv, ref pixel_decoder/msdeformattn.py:268-275) pinned to
    true fp32 in BOTH directions.  The reference forces fp32 here (`autocast(enabled=False)`, :314); PyTorch's
    default ``cudnn.allow_tf32 = True`` would still run the convolution -- and, outside any forward-time context
    manager, its backward -- on TF32 tensor cores (1e-3 .. 1e-2 relative error, algorithm-dependent)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups):
        with fp32_math():
            y = F.conv2d(x, weight, bias, stride, padding, dilation, groups)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding, dilation, groups, None if bias is None else tuple(bias.shape))
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        stride, padding, dilation, groups, bias_shape = ctx.cfg
        mask = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], bias_shape is not None and ctx.needs_input_grad[2]]
        with fp32_math():
            gx, gw, gb = torch.ops.aten.convolution_backward(
                gy, x, weight, None if bias_shape is None else list(bias_shape), list(stride), list(padding),
                list(dilation), False, [0, 0], groups, mask)
        return gx, gw, gb, None, None, None, None


def conv2d_fp32(x, conv):
    """``conv`` (an nn.Conv2d) applied to ``x`` with fp32 arithmetic forward and backward (library convolution: the
    route of layers the tensor-core kernels do not cover)."""
    _cuda_only(x, "x")
    _route("conv2d", False, f"{tuple(conv.weight.shape)} on {tuple(x.shape)}: cuDNN, TF32 off")
    if conv.padding_mode != "zeros" or isinstance(conv.padding, str):
        with fp32_math():
            return conv._conv_forward(x, conv.weight, conv.bias)
    return _ConvFp32.apply(x, conv.weight, conv.bias, tuple(conv.stride), tuple(conv.padding), tuple(conv.dilation),
                           conv.groups)


class _Conv1x1NCHW(torch.autograd.Function):
    """1x1 convolution of an NCHW-contiguous map (backbone output) straight into channels-last tokens: the map is
    the MN-major operand [B, Cin, HW] of the token-reduction ("TN") GEMM, so no layout copy of the input is made."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        B, Cin, H, W = x.shape
        Cout = weight.shape[0]
        x3 = x.view(B, Cin, H * W)
        wt = weight.view(Cout, Cin).t()[None].expand(B, -1, -1).contiguous()          # [B, Cin, Cout], a few MB
        y = native.gemm_tn(x3, wt)                                                   # [B, HW, Cout]
        if bias is not None:
            y += bias
        ctx.save_for_backward(x3, weight)
        ctx.has_bias, ctx.hw = bias is not None, (H, W)
        return y.view(B, H, W, Cout).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        x3, weight = ctx.saved_tensors
        B, Cin, HW = x3.shape
        Cout = weight.shape[0]
        g = gy.permute(0, 2, 3, 1)
        if not g.is_contiguous():
            g = g.contiguous()
        g3 = g.view(B, HW, Cout)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # dX[b] (Cin x HW) = W^T dY[b]^T: token GEMM with a transposed store -> NCHW directly
            wt_hi, wt_lo = native.split_bt(weight.view(Cout, Cin))
            gx = native.gemm(g3, wt_hi[None].expand(B, -1, -1), wt_lo[None].expand(B, -1, -1), transpose_c=True)
            gx = gx.view(B, Cin, *ctx.hw)
        if ctx.needs_input_grad[1]:
            # dW = sum_b dY[b]^T (Cout x HW) X[b]^T (HW x Cin): dY is MN-major, X K-major; K = HW cut into splits
            tiles = ((Cout + 127) // 128) * ((Cin + 127) // 128) * B
            splits = max(1, min(64, 296 // max(1, tiles), (HW + 1023) // 1024))
            gw = native.gemm_general(g3, x3, a_mn=True, b_mn=False, k_splits=splits).sum(0).view(weight.shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = native.colsum(g3.view(B * HW, Cout))
        return gx, gw, gb


def conv1x1_nchw_to_cl(x, weight, bias=None):
    """1x1 convolution (ref pixel_decoder/msdeformattn.py:216-219 input projections, :262 lateral convs) of an
    NCHW-contiguous map; result: the same logical [B, Cout, H, W] in channels-last memory.  None when the geometry is
    not covered."""
    ok = (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous() and native.GEMM_MODE == "bf16x3"
          and (x.shape[2] * x.shape[3]) % 4 == 0 and x.shape[2] * x.shape[3] >= 128 and weight.shape[0] % 4 == 0
          and x.shape[1] % 4 == 0)
    ROUTES[f"conv1x1_nchw:{'tn_gemm' if ok else 'token_gemm'}"] += 1      # both native; the second copies the layout
    return _Conv1x1NCHW.apply(x, weight, bias) if ok else None


class _AddLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, r, weight, bias, eps):
        y, mean, rstd = native.add_layernorm_fwd(x, r, weight, bias, eps)
        ctx.save_for_backward(x, r, weight, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, r, weight, mean, rstd = ctx.saved_tensors
        dx, dgamma, dbeta = native.add_layernorm_bwd(gy, x, r, weight, mean, rstd)
        return dx, (dx if r is not None else None), dgamma, dbeta, None


def add_layer_norm(x, r, norm):
    """``norm(x + r)`` for an nn.LayerNorm over the last dimension (``r`` may be None) as ONE kernel forward and one
    backward: the sum is never materialised (ref pixel_decoder/msdeformattn.py:125-126,129; decoder :52,:112,:169)."""
    _cuda_only(x, "x")
    C = x.shape[-1]
    ok = (C in (128, 256, 512) and x.dtype == torch.float32 and tuple(norm.normalized_shape) == (C,)
          and norm.weight is not None and norm.bias is not None)
    if not _route("add_layer_norm", ok, f"width {C} / dtype {x.dtype} / affine"):
        return norm(x if r is None else x + r)
    return _AddLayerNorm.apply(x, r, norm.weight, norm.bias, norm.eps)


_NO_GN_KERNEL = bool(os.environ.get("MPF_NO_GN_KERNEL"))     # A/B switches for benchmarks only
NO_FUSED_ENCODER_LAYER = bool(os.environ.get("MPF_NO_FUSED_ENCODER_LAYER"))
NO_CONV3X3_KERNEL = bool(os.environ.get("MPF_NO_CONV3X3_KERNEL"))


class _GroupNormCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, groups, relu):
        # x: logical [B, C, H, W] in channels-last memory == tokens [B, H*W, C]
        B, C, H, W = x.shape
        tokens = x.permute(0, 2, 3, 1).reshape(B, H * W, C)
        y, mean, rstd = native.groupnorm_cl_fwd(tokens, weight, bias, eps, groups, relu)
        ctx.save_for_backward(tokens, weight, bias, mean, rstd)
        ctx.groups, ctx.relu, ctx.hw = groups, relu, (H, W)
        return y.view(B, H, W, C).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        tokens, weight, bias, mean, rstd = ctx.saved_tensors
        B, HW, C = tokens.shape
        g = gy.permute(0, 2, 3, 1)
        if not g.is_contiguous():
            g = g.contiguous()
        dx, dgamma, dbeta = native.groupnorm_cl_bwd(g.view(B, HW, C), tokens, weight, bias, mean, rstd, ctx.groups,
                                                    ctx.relu)
        H, W = ctx.hw
        return dx.view(B, H, W, C).permute(0, 3, 1, 2), dgamma, dbeta, None, None, None


def group_norm_cl(x, gn, relu=False):
    """nn.GroupNorm (optionally followed by ReLU) on a logically-NCHW map held in channels-last memory, as two
    HBM passes forward / two backward (ref pixel_decoder/msdeformattn.py:216-219, :262-275).  Other layouts /
    group sizes go through the library GroupNorm."""
    ok = (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and gn.affine and not _NO_GN_KERNEL
          and native.groupnorm_cl_ok(x.shape[1], gn.num_groups) and x.permute(0, 2, 3, 1).is_contiguous())
    if _route("group_norm", ok, f"{tuple(x.shape)} groups {gn.num_groups} (layout / group size)"):
        return _GroupNormCL.apply(x, gn.weight, gn.bias, gn.eps, gn.num_groups, relu)
    y = gn(x)
    return F.relu(y) if relu else y


class _GroupNormNCHW2CL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, groups, relu):
        # x: NCHW-contiguous [B, C, H, W] (a cuDNN convolution output) -> logical NCHW in channels-last memory
        B, C, H, W = x.shape
        x3 = x.view(B, C, H * W)
        y, mean, rstd = native.groupnorm_nchw2cl_fwd(x3, weight, bias, eps, groups, relu)
        ctx.save_for_backward(x3, weight, bias, mean, rstd)
        ctx.groups, ctx.relu, ctx.hw = groups, relu, (H, W)
        return y.view(B, H, W, C).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        x3, weight, bias, mean, rstd = ctx.saved_tensors
        B, C, HW = x3.shape
        g = gy.permute(0, 2, 3, 1)
        if not g.is_contiguous():
            g = g.contiguous()
        dx, dgamma, dbeta = native.groupnorm_nchw2cl_bwd(g.view(B, HW, C), x3, weight, bias, mean, rstd, ctx.groups,
                                                         ctx.relu)
        H, W = ctx.hw
        return dx.view(B, C, H, W), dgamma, dbeta, None, None, None


def group_norm_nchw_to_cl(x, gn, relu=False):
    """nn.GroupNorm (optionally + ReLU) of an NCHW-contiguous map, returned as the same logical [B,C,H,W] tensor in
    channels-last memory (ready to be viewed as tokens): ref pixel_decoder/msdeformattn.py:262-275 output conv norm +
    activation.  Returns None when the geometry is not covered (caller uses the library GroupNorm)."""
    if (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and gn.affine and not _NO_GN_KERNEL
            and x.is_contiguous()
            and native.groupnorm_nchw2cl_ok(x.shape[1], gn.num_groups, x.shape[2] * x.shape[3])):
        ROUTES["group_norm_nchw_to_cl:native"] += 1
        return _GroupNormNCHW2CL.apply(x, gn.weight, gn.bias, gn.eps, gn.num_groups, relu)
    return None          # (the caller continues with group_norm_cl, which records its own route)


class _Upsample2xAddNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cur, prev):
        # cur, prev: logical NCHW in channels-last memory; result NCHW-contiguous
        out = native.upsample2x_add_nchw_fwd(cur.permute(0, 2, 3, 1), prev.permute(0, 2, 3, 1))
        return out

    @staticmethod
    def backward(ctx, g):
        g_cur, g_prev = native.upsample2x_add_nchw_bwd(g)
        return g_cur.permute(0, 3, 1, 2), g_prev.permute(0, 3, 1, 2)


def upsample2x_add_to_nchw(cur, prev):
    """``cur + F.interpolate(prev, size=cur.shape[-2:], mode="bilinear", align_corners=False)`` for the exact x2 case of
    the FPN stage (ref pixel_decoder/msdeformattn.py:347-349), both inputs channels-last, result NCHW-contiguous for
    the 3x3 convolution.  Returns None when the geometry is not covered."""
    if (cur.is_cuda and cur.dtype == torch.float32 and prev.dtype == torch.float32 and cur.dim() == 4
            and not _NO_GN_KERNEL
            and cur.shape[2] == 2 * prev.shape[2] and cur.shape[3] == 2 * prev.shape[3]
            and native.upsample2x_add_ok(cur.shape[2], cur.shape[3], cur.shape[1])
            and cur.permute(0, 2, 3, 1).is_contiguous() and prev.permute(0, 2, 3, 1).is_contiguous()):
        ROUTES["upsample2x_add_to_nchw:native"] += 1
        return _Upsample2xAddNCHW.apply(cur, prev)
    _route("upsample2x_add_to_nchw", False, f"{tuple(cur.shape)} <- {tuple(prev.shape)}")
    return None


class _Upsample2xAddCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cur, prev):
        # cur, prev: logical NCHW in channels-last memory; result likewise
        out = native.upsample2x_add_cl_fwd(cur.permute(0, 2, 3, 1), prev.permute(0, 2, 3, 1))
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        gc = g.permute(0, 2, 3, 1)
        if not gc.is_contiguous():
            gc = gc.contiguous()
        g_prev = native.upsample2x_cl_bwd(gc) if ctx.needs_input_grad[1] else None
        return gc.permute(0, 3, 1, 2), (None if g_prev is None else g_prev.permute(0, 3, 1, 2))


def upsample2x_add_cl(cur, prev):
    """``cur + F.interpolate(prev, x2, bilinear, align_corners=False)`` with inputs and result channels-last (the FPN
    merge in front of the tensor-core 3x3 convolution).  None when the geometry is not covered."""
    if (cur.is_cuda and cur.dtype == torch.float32 and prev.dtype == torch.float32 and cur.dim() == 4
            and cur.shape[2] == 2 * prev.shape[2] and cur.shape[3] == 2 * prev.shape[3] and cur.shape[1] % 4 == 0
            and cur.permute(0, 2, 3, 1).is_contiguous() and prev.permute(0, 2, 3, 1).is_contiguous()):
        ROUTES["upsample2x_add_cl:native"] += 1
        return _Upsample2xAddCL.apply(cur, prev)
    return None          # (the caller tries upsample2x_add_to_nchw next)


class _Conv3x3CL(torch.autograd.Function):
    """3x3 convolution (stride 1, padding 1) of a channels-last map as ONE tensor-core GEMM per direction: the nine
    taps are the outer part of the reduction dimension (K = 9 * Cin) and only shift the TMA box of the activation
    operand, so no im2col buffer, no padded copy and no layout change exist (ref pixel_decoder/msdeformattn.py:268-275;
    cuDNN's fp32 path for this layer: Winograd on CUDA cores, 6.5 ms forward / 12.7 ms backward at [16,256,256,256])."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        Cout, Cin = weight.shape[0], weight.shape[1]
        xc = x.permute(0, 2, 3, 1)                                            # [B, H, W, Cin] contiguous
        w_hi, w_lo = native.split_bf16(weight.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin))
        y = native.conv3x3_cl(xc, w_hi, w_lo, bias)
        ctx.save_for_backward(xc, weight)
        ctx.has_bias = bias is not None
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xc, weight = ctx.saved_tensors
        Cout, Cin = weight.shape[0], weight.shape[1]
        g = gy.permute(0, 2, 3, 1)
        if not g.is_contiguous():
            g = g.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # dX = conv3x3(dY, W') with W'[ci, co, ky, kx] = W[co, ci, 2-ky, 2-kx]
            wt_hi, wt_lo = native.split_bf16(weight.flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, 9 * Cout))
            gx = native.conv3x3_cl(g, wt_hi, wt_lo).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            gw = native.conv3x3_cl_wgrad(g, xc).view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = native.colsum(g.reshape(-1, Cout))
        return gx, gw, gb


def conv3x3_cl_supported(x, conv):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and conv.kernel_size == (3, 3)
            and conv.stride == (1, 1) and conv.padding == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
            and conv.padding_mode == "zeros" and not NO_CONV3X3_KERNEL
            and native.conv3x3_cl_ok(x.shape[2], x.shape[3], conv.in_channels, conv.out_channels))


def conv3x3_cl(x, conv):
    """``conv`` (nn.Conv2d, 3x3, stride 1, padding 1) on a logically-NCHW map in channels-last memory, result likewise.
    None when the layer / geometry is not covered (caller uses the library convolution)."""
    if _route("conv3x3", conv3x3_cl_supported(x, conv) and x.permute(0, 2, 3, 1).is_contiguous(),
              f"{tuple(x.shape)} -> {conv.out_channels} (channels % 64 / layout)"):
        return _Conv3x3CL.apply(x, conv.weight, conv.bias)
    return None


class _FFN(torch.autograd.Function):
    """y = relu(x W1^T + b1) W2^T + b2 as one autograd node, so the ReLU backward is fused into the epilogue
    of the input-gradient GEMM of the second layer (``gate``) instead of a separate pass over [tokens, d_ffn]."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x2 = x.reshape(-1, x.shape[-1])
        w1_hi, w1_lo = native.split_b(w1)
        hidden = native.gemm(x2, w1_hi, w1_lo, b1, relu=True)
        w2_hi, w2_lo = native.split_b(w2)
        y = native.gemm(hidden, w2_hi, w2_lo, b2)
        ctx.save_for_backward(x2, w1, w2, hidden)
        return y.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w1, w2, hidden = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1]).contiguous()
        w2t_hi, w2t_lo = native.split_bt(w2)
        gh = native.gemm_general(g2, w2t_hi, b_lo=w2t_lo, gate=hidden)       # d(hidden) with the ReLU mask applied
        gw2, gb2 = native.matmul_tn(g2, hidden, with_colsum=True)
        w1t_hi, w1t_lo = native.split_bt(w1)
        gx = native.gemm(gh, w1t_hi, w1t_lo).view(*gy.shape[:-1], w1.shape[1]) if ctx.needs_input_grad[0] else None
        gw1, gb1 = native.matmul_tn(gh, x2, with_colsum=True)
        return gx, gw1, gb1, gw2, gb2


def ffn(x, w1, b1, w2, b2):
    """Linear -> ReLU -> Linear (ref pixel_decoder/msdeformattn.py:116-120, decoder :165-169 with
    dropout 0)."""
    _cuda_only(x, "x")
    if x.shape[-1] % 32 or w1.shape[0] % 32 or w2.shape[0] % 4:
        return linear(linear(x, w1, b1, relu=True), w2, b2)          # (each linear records its own route)
    ROUTES["ffn:native"] += 1
    return _FFN.apply(x, w1, b1, w2, b2)


# ------------------------------------------------------------------------------------------------
# MSDeformAttn encoder layer as one autograd node
# ------------------------------------------------------------------------------------------------
class _EncoderLayer(torch.autograd.Function):
    """One MSDeformAttnTransformerEncoderLayer (ref pixel_decoder/msdeformattn.py:92-131 with dropout inactive):

        value = src Wv^T + bv;  ow = (src + pos) [Woff; Watt]^T + [boff; batt];  a = MSDeformAttn(value, ow, ref)
        src1 = LN1(src + a Wo^T + bo);  out = LN2(src1 + relu(src1 W1^T + b1) W2^T + b2)

    as ONE node, so that the elementwise traffic autograd would add around the kernels disappears:
    ``src + pos`` is never formed (pos [1,S,C] is batch independent: its projection enters the GEMM epilogue as a
    row-periodic residual); in the backward the three gradient contributions to ``src`` and the two to ``src1``
    are chained through the ``resid`` input of the input-gradient GEMMs instead of separate full-size additions,
    and the gradient of ``pos`` is one small GEMM on the batch-summed logit gradients."""

    @staticmethod
    def forward(ctx, src, pos, ref, shapes, lsi, n_heads, n_points,
                wv, bv, woff, boff, watt, batt, wo, bo, g1, be1, eps1, w1, b1, w2, b2, g2, be2, eps2):
        from . import MultiScaleDeformableAttention as MSDA
        B, S, C = src.shape
        M, P = n_heads, n_points
        host_shapes = getattr(shapes, "_mpf_host_shapes", None)
        x2 = src.reshape(B * S, C)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        pos2 = pos.reshape(S, C)
        wv_hi, wv_lo = native.split_b(wv)
        value = native.gemm(x2, wv_hi, wv_lo, bv)
        wow = torch.cat([woff, watt], 0)
        bow = torch.cat([boff, batt], 0)
        wow_hi, wow_lo = native.split_b(wow)
        pos_ow = native.gemm(pos2, wow_hi, wow_lo, bow)                             # [S, M*L*P*3]
        ow = native.gemm(x2, wow_hi, wow_lo, None, resid=pos_ow, resid_rows=S)
        attn = MSDA.ms_deform_attn_enc_forward(value.view(B, S, M, C // M), shapes, lsi, ow.view(B, S, -1), ref, P,
                                               host_shapes=host_shapes)
        wo_hi, wo_lo = native.split_b(wo)
        proj = native.gemm(attn.view(B * S, C), wo_hi, wo_lo, bo)
        src1, mean1, rstd1 = native.add_layernorm_fwd(x2, proj, g1, be1, eps1)
        w1_hi, w1_lo = native.split_b(w1)
        if w1_hi.dtype == torch.bfloat16 and w1.shape[0] % 32 == 0:
            # the ReLU pattern is kept as one bit per element for the backward (1/32 of re-reading `hidden`)
            hidden, hbits = native.gemm_relu_bits(src1, w1_hi, w1_lo, b1, relu_bits_out=True)
        else:
            hidden, hbits = native.gemm(src1, w1_hi, w1_lo, b1, relu=True), None
        w2_hi, w2_lo = native.split_b(w2)
        y = native.gemm(hidden, w2_hi, w2_lo, b2)
        out, mean2, rstd2 = native.add_layernorm_fwd(src1, y, g2, be2, eps2)
        ctx.hbits = hbits
        ctx.save_for_backward(x2, pos2, ref, shapes, lsi, value, ow, attn, proj, mean1, rstd1, src1, hidden, y, mean2,
                              rstd2, wv, wow, wo, g1, w1, w2, g2)
        ctx.geom = (B, S, C, M, P, host_shapes, woff.shape[0])
        return out.view(B, S, C)

    @staticmethod
    def backward(ctx, g_out):
        from . import MultiScaleDeformableAttention as MSDA
        (x2, pos2, ref, shapes, lsi, value, ow, attn, proj, mean1, rstd1, src1, hidden, y, mean2, rstd2,
         wv, wow, wo, g1, w1, w2, g2) = ctx.saved_tensors
        B, S, C, M, P, host_shapes, n_off = ctx.geom

        def t_halves(w):
            return native.split_bt(w)

        # FFN block
        dsum2, dg2, db2, gb2 = native.add_layernorm_bwd(g_out.reshape(B * S, C), src1, y, g2, mean2, rstd2,
                                                        with_colsum=True)          # colsum(dx) = bias gradient of W2
        w2t_hi, w2t_lo = t_halves(w2)
        if ctx.hbits is not None and w2t_hi.dtype == torch.bfloat16:
            gh = native.gemm_relu_bits(dsum2, w2t_hi, w2t_lo, gate_bits=ctx.hbits)   # ReLU mask (bits) in the epilogue
        else:
            gh = native.gemm_general(dsum2, w2t_hi, b_lo=w2t_lo, gate=hidden)        # ReLU mask in the epilogue
        gw2 = native.matmul_tn(dsum2, hidden)
        w1t_hi, w1t_lo = t_halves(w1)
        g_src1 = native.gemm(gh, w1t_hi, w1t_lo, resid=dsum2)                        # + the residual branch
        gw1, gb1 = native.matmul_tn(gh, src1, with_colsum=True)      # weight and bias gradient from one pass over gh
        del gh
        # attention block
        dsum1, dg1, db1, gbo = native.add_layernorm_bwd(g_src1, x2, proj, g1, mean1, rstd1, with_colsum=True)
        wot_hi, wot_lo = t_halves(wo)
        g_attn = native.gemm(dsum1, wot_hi, wot_lo)
        gwo = native.matmul_tn(dsum1, attn.view(B * S, C))
        g_value, g_ow = MSDA.ms_deform_attn_enc_backward(value.view(B, S, M, C // M), shapes, lsi, ow.view(B, S, -1), ref,
                                                         g_attn.view(B, S, C), P, host_shapes=host_shapes)
        gv2, gow2 = g_value.view(B * S, C), g_ow.view(B * S, -1)
        wvt_hi, wvt_lo = t_halves(wv)
        wowt_hi, wowt_lo = t_halves(wow)
        g_src = native.gemm(gv2, wvt_hi, wvt_lo, resid=dsum1)                        # residual + value branch
        g_src = native.gemm(gow2, wowt_hi, wowt_lo, resid=g_src)                     # + query branch
        gwv, gbv = native.matmul_tn(gv2, x2, with_colsum=True)
        gow_b = g_ow.view(B, S, -1).sum(0)                                           # [S, M*L*P*3]
        gwow, gbow = native.matmul_tn(gow2, x2, with_colsum=True)
        gwow = gwow + native.matmul_tn(gow_b, pos2)
        g_pos = native.gemm(gow_b, wowt_hi, wowt_lo).view(1, S, C) if ctx.needs_input_grad[1] else None
        return (g_src.view(B, S, C), g_pos, None, None, None, None, None,
                gwv, gbv, gwow[:n_off], gbow[:n_off], gwow[n_off:], gbow[n_off:], gwo, gbo, dg1, db1, None,
                gw1, gb1, gw2, gb2, dg2, db2, None)


def encoder_layer_supported(src, pos, reference_points, d_ffn, n_heads, n_levels, n_points):
    C = src.shape[-1]
    return (src.is_cuda and src.dtype == torch.float32 and src.dim() == 3 and pos is not None and pos.dim() == 3
            and pos.shape[0] == 1 and pos.shape[1:] == src.shape[1:] and C in (128, 256, 512) and C % n_heads == 0
            and (C // n_heads) in (16, 32, 64) and n_points == 4 and n_levels <= 4 and d_ffn % 32 == 0
            and (n_heads * n_levels * n_points * 3) % 32 == 0
            and reference_points.shape[-1] == 2 and not reference_points.requires_grad)


def encoder_layer(src, pos, reference_points, spatial_shapes, level_start_index, attn, norm1, linear1, linear2, norm2):
    """Fused forward/backward of one encoder layer; ``attn``: the layer's MSDeformAttn module (parameters only)."""
    return _EncoderLayer.apply(
        src, pos, reference_points, spatial_shapes, level_start_index, attn.n_heads, attn.n_points,
        attn.value_proj.weight, attn.value_proj.bias, attn.sampling_offsets.weight, attn.sampling_offsets.bias,
        attn.attention_weights.weight, attn.attention_weights.bias, attn.output_proj.weight, attn.output_proj.bias,
        norm1.weight, norm1.bias, norm1.eps, linear1.weight, linear1.bias, linear2.weight, linear2.bias,
        norm2.weight, norm2.bias, norm2.eps)


# ------------------------------------------------------------------------------------------------
# prediction heads: mask logits + boolean stage
# ------------------------------------------------------------------------------------------------
def _channels_last_tokens(mask_features):
    """[B, C, H, W] (any memory format) -> [B, H*W, C] contiguous view/copy."""
    t = mask_features.permute(0, 2, 3, 1)
    if not t.is_contiguous():
        t = t.contiguous()
    return t.view(t.shape[0], -1, t.shape[-1])


class _SharedGrad:
    """State shared by the prediction heads of one decoder forward: the accumulator of the gradients they send to
    ``mask_features`` (see ``grad_fanout``) and, when the heads' outputs pass through ``collect_mask_heads``, the
    per-head operands / precomputed gradients of the batched backward."""
    __slots__ = ("buf", "embeds", "tokens", "dE")

    def __init__(self):
        self.buf = None          # d(mask_features) as tokens [B, H*W, C]
        self.embeds = []         # mask_embed of every head, in call order (detached)
        self.tokens = None       # mask_features as tokens [B, H*W, C]
        self.dE = None           # [B, heads, Qt, C] once the collector's backward has run


class _GradFanout(torch.autograd.Function):
    """x -> n aliases of x, one per consumer.  Consumers that know the protocol (``_MaskLogits``) add their gradient
    into ``state.buf`` inside their own kernel epilogue instead of returning n full-size tensors for autograd to
    sum; autograd runs this node's backward only after every consumer's, so the buffer is complete here."""

    @staticmethod
    def forward(ctx, x, n, state):
        ctx.state = state
        ctx.xshape = x.shape
        ctx.set_materialize_grads(False)
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *grads):
        buf, ctx.state.buf = ctx.state.buf, None
        ctx.state.dE, ctx.state.embeds, ctx.state.tokens = None, [], None
        rest = [g for g in grads if g is not None and (buf is None or g.data_ptr() != buf.data_ptr())]
        total = None
        if buf is not None:                             # tokens [B, H*W, C] -> logical [B, C, H, W]
            B, C, H, W = ctx.xshape
            total = buf.view(B, H, W, C).permute(0, 3, 1, 2)
        for g in rest:                                  # consumers outside the protocol: plain sum
            total = g if total is None else total + g
        return total, None, None


def grad_fanout(x, n):
    """n aliases of ``x`` for n consumers whose gradients are accumulated in place (the 10 prediction heads all read
    ``mask_features``: nine 1 GB gradient additions per step at the bench geometry otherwise).
    Returns (aliases, state); pass ``state`` to ``mask_logits``."""
    state = _SharedGrad()
    if not (torch.is_grad_enabled() and x.requires_grad):
        return [x] * n, None
    return list(_GradFanout.apply(x, n, state)), state


class _MaskLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask_embed, mask_features, shared):
        B, C, H, W = mask_features.shape
        tokens = _channels_last_tokens(mask_features)                       # [B, HW, C]
        e_hi, e_lo = native.split_b(mask_embed)
        out = native.gemm(tokens, e_hi, e_lo, transpose_c=True)             # [B, Q, HW]
        ctx.save_for_backward(mask_embed, tokens)
        ctx.fshape = (B, C, H, W)
        ctx.shared = shared
        ctx.head = -1
        if shared is not None:
            ctx.head = len(shared.embeds)
            shared.embeds.append(mask_embed.detach())
            shared.tokens = tokens.detach()
        return out.view(B, mask_embed.shape[1], H, W)

    @staticmethod
    def backward(ctx, g):
        mask_embed, tokens = ctx.saved_tensors
        B, C, H, W = ctx.fshape
        sh = ctx.shared
        if sh is not None and sh.dE is not None:
            # collect_mask_heads already computed both gradients for all heads in two batched GEMMs
            ge = sh.dE[:, ctx.head] if ctx.needs_input_grad[0] else None
            gf = sh.buf.view(B, H, W, C).permute(0, 3, 1, 2) if ctx.needs_input_grad[1] else None
            return ge, gf, None
        g2 = g.reshape(B, g.shape[1], H * W)
        ge = gf = None
        g2 = g2.contiguous()
        ok = _route("mask_logits.backward", (H * W) % 4 == 0 and C % 4 == 0, f"H*W {H * W} or C {C} % 4 != 0")
        if ctx.needs_input_grad[0]:
            # dE[b] = dOut[b] (Q x HW) @ F[b] (HW x C): a small [Q x C] result reduced over H*W = 65536, so the
            # reduction is cut into K-splits across CTAs; F is consumed MN-major (no transpose)
            ge = native.gemm_general(g2, tokens, a_mn=False, b_mn=True, k_splits=16) if ok \
                else torch.bmm(g2, tokens)
        if ctx.needs_input_grad[1]:
            # dF[b] = dOut[b]^T (HW x Q) @ E[b] (Q x C): both operands MN-major
            if ok and native.GEMM_MODE == "bf16x3":
                # reduction over the Q queries, both operands as stored; with a shared accumulator the heads after
                # the first add into it in the kernel epilogue (TMA reduce-add)
                sh = ctx.shared
                if sh is None:
                    gf = native.gemm_tn(g2, mask_embed)
                else:
                    sh.buf = gf = native.gemm_tn(g2, mask_embed, accumulate_into=sh.buf)
            elif ok:
                gf = native.gemm_general(g2, mask_embed, a_mn=True, b_mn=True)
            else:
                gf = torch.bmm(g2.transpose(1, 2), mask_embed)
            gf = gf.view(B, H, W, C).permute(0, 3, 1, 2)
        return ge, gf, None


def mask_logits(mask_embed, mask_features, shared=None):
    """einsum('bqc,bchw->bqhw') (ref decoder :1865) as a TMA-fed tcgen05 GEMM over channels-last
    pixel features: M = H*W, N = Q, K = C, transposed store.  ``shared``: state of ``grad_fanout`` when
    ``mask_features`` is one of its aliases."""
    _cuda_only(mask_embed, "mask_embed")
    return _MaskLogits.apply(mask_embed.contiguous(), mask_features, shared)


class _HeadCollector(torch.autograd.Function):
    """Identity on the mask logits of ALL prediction heads (optionally split into the DN and the matching queries),
    placed at the end of the decoder so that its backward sees every head's incoming gradient at once.  Those
    gradients come from the loss alone (the attention masks derived from the logits are detached, ref decoder
    :1875), so the heads' backward GEMMs can be batched:

        dE_all[b] = G_all[b] (heads*Qt x HW) @ F[b] (HW x C)        -- mask_features read once, not once per head
        dF[b]     = G_all[b]^T (HW x heads*Qt) @ E_all[b]           -- written once, not accumulated ten times

    G_all [B, heads, Qt, HW] is assembled here from the (up to two per head) gradient pieces -- the pass that
    ``split_queries`` would otherwise spend on concatenating them."""

    @staticmethod
    def forward(ctx, shared, n_dn, *masks):
        ctx.shared, ctx.n_dn, ctx.n_heads = shared, n_dn, len(masks)
        ctx.mshape = masks[0].shape
        ctx.set_materialize_grads(False)
        outs = []
        for m in masks:
            if n_dn > 0:
                outs += [m[:, :n_dn], m[:, n_dn:]]
            else:
                outs.append(m.view_as(m))
        return tuple(outs)

    @staticmethod
    def _side_by_side(grads, sizes, B, H, W):
        """The gradient pieces as ONE [B, T, H*W] tensor without copying, when they already are slices (along dim 1,
        in any order) of one dense buffer that they tile completely -- what ``native.PointSampleViews`` (the
        criterion's joint evaluation of all heads) hands back.  Returns (G2, first row of every piece) or None."""
        if any(g is None for g in grads):
            return None
        T, HW = sum(sizes), H * W
        store = grads[0].untyped_storage().data_ptr()
        want = (T * HW, HW, W, 1)
        for g, n in zip(grads, sizes):
            if g.dtype != torch.float32 or g.untyped_storage().data_ptr() != store or tuple(g.shape) != (B, n, H, W) or \
                    (n > 0 and g.stride() != want):
                return None
        base = min(g.storage_offset() for g in grads)
        rows, covered = [], 0
        for g, n in zip(grads, sizes):
            d = g.storage_offset() - base
            if d % HW or d // HW + n > T:
                return None
            rows.append(d // HW)
            covered += n
        order = sorted(range(len(rows)), key=lambda i: rows[i])
        at = 0
        for i in order:                       # the pieces must tile [0, T) exactly
            if rows[i] != at:
                return None
            at += sizes[i]
        return torch.as_strided(grads[order[0]], (B, T, HW), (T * HW, HW, 1), base), rows

    @staticmethod
    def backward(ctx, *grads):
        sh, n_dn, nH = ctx.shared, ctx.n_dn, ctx.n_heads
        B, Qt, H, W = ctx.mshape
        dev = sh.tokens.device if sh.tokens is not None else next(g.device for g in grads if g is not None)
        sizes = ([n_dn, Qt - n_dn] * nH) if n_dn > 0 else ([Qt] * nH)
        C = sh.tokens.shape[-1] if sh.tokens is not None else 0
        batched = (sh.tokens is not None and len(sh.embeds) == nH and native.GEMM_MODE == "bf16x3" and (H * W) % 4 == 0
                   and C % 4 == 0)
        side = _HeadCollector._side_by_side(grads, sizes, B, H, W) if batched else None
        ROUTES["mask_heads.backward:" + ("gradients_in_place" if side is not None else "gradients_concatenated")] += 1
        if side is not None:
            # no concatenation: the GEMMs read the buffer in place; its row order (any permutation of the pieces) is
            # applied to the small operand E and undone on the small result dE
            G2, rows = side
            pieces_e, k = [], 0
            for h in range(nH):
                for n, lo in (((n_dn, 0), (Qt - n_dn, n_dn)) if n_dn > 0 else ((Qt, 0),)):
                    pieces_e.append((rows[k], sh.embeds[h][:, lo:lo + n]))
                    k += 1
            E_all = torch.cat([e for _, e in sorted(pieces_e, key=lambda p: p[0])], 1)       # [B, T, C], buffer order
        else:
            # ONE concatenation along the (head, query) axis: contiguous output, vectorised batched copy
            pieces = [g.reshape(B, n, H * W) if g is not None else
                      torch.zeros((B, n, H * W), dtype=torch.float32, device=dev) for g, n in zip(grads, sizes)]
            G = torch.cat(pieces, 1).view(B, nH, Qt, H, W)
            del pieces
            if batched:
                G2 = G.view(B, nH * Qt, H * W)
                E_all = torch.stack(sh.embeds, 1).view(B, nH * Qt, C)
        if batched:
            T = nH * Qt
            # dE: reduction over the H*W pixels -> mask_features is needed K-major: one transposing split per step
            # (instead of ten passes of the 3xTF32 mixed-major kernel), then the bf16x3 GEMM with split-K chosen so
            # that the tile count fills whole waves of the 148 SMs
            tiles = ((T + 127) // 128) * ((C + 255) // 256) * B
            max_splits = max(1, min(16, (H * W) // 2048))
            splits = min(range(1, max_splits + 1), key=lambda k: (-(-tiles * k // 148)) / (tiles * k / 148.0) + 0.01 * k)
            if (H * W) % 8 == 0:
                ft_hi, ft_lo = native.transpose_split_bf16(sh.tokens)              # [B, C, HW]
                dE = native.gemm_bf16x3_splitk(G2, ft_hi, ft_lo, splits)
                del ft_hi, ft_lo
            else:
                dE = native.gemm_general(G2, sh.tokens, a_mn=False, b_mn=True, k_splits=splits)
            if side is not None:                 # back to (head, query) order
                k, back = 0, []
                for h in range(nH):
                    for n in ((n_dn, Qt - n_dn) if n_dn > 0 else (Qt,)):
                        back.append(dE[:, rows[k]:rows[k] + n])
                        k += 1
                dE = torch.cat(back, 1)
            sh.dE = dE.view(B, nH, Qt, C)
            sh.buf = native.gemm_tn(G2, E_all)                                 # [B, HW, C]
        if side is not None:
            # the heads' own backward takes dE / dF from the shared state and ignores its incoming gradient
            skip = torch.zeros(1, dtype=torch.float32, device=dev).expand(B, Qt, H, W)
            return (None, None) + (skip,) * nH
        return (None, None) + tuple(G[:, h] for h in range(nH))


def collect_mask_heads(masks, n_dn, shared):
    """masks: the mask logits [B, Qt, H, W] of every prediction head, in call order.  Returns
    ``(dn_parts, matching_parts)`` (``dn_parts`` is None when ``n_dn == 0``): views of the inputs, with the heads'
    backward batched as described in ``_HeadCollector``.  ``shared``: the state ``grad_fanout`` returned."""
    if shared is None or not torch.is_grad_enabled() or not any(m.requires_grad for m in masks):
        if n_dn > 0:
            return [m[:, :n_dn] for m in masks], [m[:, n_dn:] for m in masks]
        return None, list(masks)
    outs = _HeadCollector.apply(shared, n_dn, *masks)
    if n_dn > 0:
        return list(outs[0::2]), list(outs[1::2])
    return None, list(outs)


class PackedMask:
    """Attention mask of one decoder layer: ``bits`` int32 [B, Q, W] (bit i of word j = key 32j+i,
    1 = not allowed), shared by all heads; ``n_keys`` = h*w."""

    def __init__(self, bits, n_keys):
        self.bits, self.n_keys = bits, n_keys

    def to_bool(self):
        return native.unpack_bits(self.bits, self.n_keys)

    def replace_rows(self, other, n_rows):
        """Rows [0, n_rows) come from ``other`` (the mask-piloted GT masks, ref decoder :1046-1048)."""
        return PackedMask(torch.cat([other.bits[:, :n_rows], self.bits[:, n_rows:]], 1), self.n_keys)

    @staticmethod
    def from_bool(mask):
        return PackedMask(native.pack_bool_bits(mask), mask.shape[-1])


def attn_mask_from_logits(outputs_mask, target_size):
    """Boolean stage of the heads (ref decoder :1869-1875): bilinear resize -> ``sigmoid < 0.5``, as ONE
    bit per (image, query, key).  Returns a PackedMask (detached by construction)."""
    _cuda_only(outputs_mask, "outputs_mask")
    h, w = int(target_size[0]), int(target_size[1])
    return PackedMask(native.attn_mask_bits(outputs_mask.detach(), (h, w)), h * w)


class _SplitQueries(torch.autograd.Function):
    """x [B, Qt, ...] -> (x[:, :n_first], x[:, n_first:]) as views.  Unlike two independent slices, whose
    backward each materialise a zero-filled full-size gradient that autograd then adds (five passes over the
    503 MB mask-logit gradient per prediction head at the bench geometry), the backward here writes both
    incoming gradients into ONE buffer."""

    @staticmethod
    def forward(ctx, x, n_first):
        ctx.set_materialize_grads(False)
        ctx.n_first = n_first
        ctx.meta = (x.shape, x.dtype, x.device)
        return x[:, :n_first], x[:, n_first:]

    @staticmethod
    def backward(ctx, g0, g1):
        if g0 is None and g1 is None:
            return None, None
        shape, dtype, device = ctx.meta
        if g0 is not None and g1 is not None:
            return torch.cat([g0, g1], 1), None
        g = torch.zeros(shape, dtype=dtype, device=device)
        if g0 is not None:
            g[:, :ctx.n_first] = g0
        else:
            g[:, ctx.n_first:] = g1
        return g, None


def split_queries(x, n_first):
    """Splits predictions into the mask-piloted (DN) queries and the matching queries (ref decoder :1697-1703)."""
    return _SplitQueries.apply(x, n_first)


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def _split_heads(x, nhead):
    B, L, E = x.shape
    return x.reshape(B, L, nhead, E // nhead).transpose(1, 2)    # [B,h,L,hd]


class _MaskedCrossAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_in, memory, pos, w_in, b_in, w_out, b_out, nhead, bits, n_keys):
        B, Qt, E = q_in.shape
        HW = memory.shape[1]
        hd = E // nhead
        scale2 = LOG2E / math.sqrt(hd)
        wq_hi, wq_lo = native.split_b(w_in[:E])
        wk_hi, wk_lo = native.split_b(w_in[E:2 * E])
        wv_hi, wv_lo = native.split_b(w_in[2 * E:])
        # Q (scaled into the log2 domain), pre-split for the attention kernel
        q_hi, q_lo = native.gemm(q_in.reshape(B * Qt, E), wq_hi, wq_lo, b_in[:E], alpha=scale2, split_out=True)
        # K = (memory + pos) Wk^T + bk = memory Wk^T + (pos Wk^T + bk): the second term is batch independent
        pos_k = native.gemm(pos.reshape(-1, E), wk_hi, wk_lo, b_in[E:2 * E])                  # [HW, E]
        k_hi, k_lo = native.gemm(memory.reshape(B * HW, E), wk_hi, wk_lo, None, resid=pos_k, resid_rows=HW,
                                 split_out=True)
        # V^T [B, E, HW] (keys contiguous) so that P V is a K-major tensor-core product
        vt_hi, vt_lo = native.gemm(memory, wv_hi[None].expand(B, -1, -1), wv_lo[None].expand(B, -1, -1),
                                   b_in[2 * E:], transpose_c=True, split_out=True)
        row_open = (bits == -1).all(-1)                                                      # ref decoder :1780
        o, lse2 = native.masked_xattn_fwd(q_hi.view(B, Qt, E), q_lo.view(B, Qt, E), k_hi.view(B, HW, E),
                                          k_lo.view(B, HW, E), vt_hi, vt_lo, bits, row_open, nhead)
        wo_hi, wo_lo = native.split_b(w_out)
        y = native.gemm(o.view(B * Qt, E), wo_hi, wo_lo, b_out).view(B, Qt, E)
        ctx.save_for_backward(q_in, memory, pos, w_in, b_in, w_out, bits, row_open, o, lse2,
                              q_hi.view(B, Qt, E), q_lo.view(B, Qt, E), k_hi.view(B, HW, E), k_lo.view(B, HW, E))
        ctx.nhead, ctx.n_keys = nhead, n_keys
        return y

    @staticmethod
    def backward(ctx, gy):
        (q_in, memory, pos, w_in, b_in, w_out, bits, row_open, o, lse2, q_hi, q_lo, k_hi, k_lo) = ctx.saved_tensors
        nhead = ctx.nhead
        B, Qt, E = q_in.shape
        HW = memory.shape[1]
        hd = E // nhead
        gy2 = gy.reshape(B * Qt, E).contiguous()
        o2 = o.reshape(B * Qt, E)
        g_wout, g_bout = native.matmul_tn(gy2, o2, with_colsum=True)
        wot_hi, wot_lo = native.split_bt(w_out)
        go = native.gemm(gy2, wot_hi, wot_lo).view(B, Qt, E)                        # d(attention output)
        delta = (go.view(B, Qt, nhead, hd) * o.view(B, Qt, nhead, hd)).sum(-1).permute(0, 2, 1).contiguous()
        # operands the forward did not keep: V row-major and K^T (both pre-split by the GEMM epilogue)
        wq, wk, wv = w_in[:E], w_in[E:2 * E], w_in[2 * E:]
        wk_hi, wk_lo = native.split_b(wk)
        wv_hi, wv_lo = native.split_b(wv)
        pos2 = pos.reshape(-1, E)
        pos_k = native.gemm(pos2, wk_hi, wk_lo, b_in[E:2 * E])                      # [HW, E]
        v_hi, v_lo = native.gemm(memory.reshape(B * HW, E), wv_hi, wv_lo, b_in[2 * E:], split_out=True)
        kt_hi, kt_lo = native.gemm(memory, wk_hi[None].expand(B, -1, -1), wk_lo[None].expand(B, -1, -1), None,
                                   resid=pos_k, resid_rows=HW, transpose_c=True, split_out=True)   # [B, E, HW]
        dq, dk, dv = native.masked_xattn_bwd(q_hi, q_lo, k_hi, k_lo, kt_hi, kt_lo, v_hi.view(B, HW, E),
                                             v_lo.view(B, HW, E), go.contiguous(), bits, row_open, lse2, delta,
                                             nhead)
        dq2, dk2, dv2 = dq.view(B * Qt, E), dk.view(B * HW, E), dv.view(B * HW, E)
        mem2 = memory.reshape(B * HW, E)
        # in-projection weight gradients: dW = dY^T X on the tensor cores (keys see memory + pos)
        g_wk, g_bk = native.matmul_tn(dk2, mem2, with_colsum=True)
        g_wk = g_wk + native.matmul_tn(dk.sum(0), pos2)
        g_wq, g_bq = native.matmul_tn(dq2, q_in.reshape(B * Qt, E), with_colsum=True)
        g_wv, g_bv = native.matmul_tn(dv2, mem2, with_colsum=True)
        g_win = torch.cat([g_wq, g_wk, g_wv], 0)
        g_bin = torch.cat([g_bq, g_bk, g_bv], 0)
        g_qin = None
        if ctx.needs_input_grad[0]:
            wqt_hi, wqt_lo = native.split_bt(wq)
            g_qin = native.gemm(dq2, wqt_hi, wqt_lo).view(B, Qt, E)
        g_mem = g_pos = None
        if ctx.needs_input_grad[1]:
            wkt_hi, wkt_lo = native.split_bt(wk)
            wvt_hi, wvt_lo = native.split_bt(wv)
            # dK Wk + dV Wv: the second GEMM adds the first one's result in its epilogue (no separate full-size add)
            g_mem = native.gemm(dv2, wvt_hi, wvt_lo, resid=native.gemm(dk2, wkt_hi, wkt_lo)).view(B, HW, E)
        if ctx.needs_input_grad[2]:
            g_pos = (dk.sum(0) @ wk).view(1, HW, E)
        return g_qin, g_mem, g_pos, g_win, g_bin, g_wout, g_bout, None, None, None


_TAIL_BITS = {}


def _tail_bits(n_keys, words, device):
    """int32 [words]: bit k of the row set iff key k >= n_keys (the mask of a row that attends to all real keys)."""
    key = (n_keys, words, str(device))
    if key not in _TAIL_BITS:
        k = torch.arange(words * 32, device=device)
        _TAIL_BITS[key] = native.pack_bool_bits((k >= n_keys)[None, None])[0, 0, :words].contiguous()
    return _TAIL_BITS[key]


def masked_cross_attention(q_in, memory, pos, w_in, b_in, w_out, b_out, nhead, mask):
    """softmax((q Wq)(k Wk)^T / sqrt(hd) + mask) (v Wv) Wo with k = memory + pos, v = memory
    (ref decoder :100-112 through nn.MultiheadAttention); a row that is entirely masked attends to
    every key (ref decoder :1780).  ``mask``: PackedMask, bool [B,Q,HW] (True = not allowed) or None."""
    _cuda_only(q_in, "tgt")
    B, Qt, E = q_in.shape
    HW = memory.shape[1]
    if mask is None:
        bits = torch.zeros((B, Qt, native.mask_words(HW)), dtype=torch.int32, device=q_in.device)
        n_keys = HW
    elif isinstance(mask, PackedMask):
        bits, n_keys = mask.bits, mask.n_keys
    else:
        bits, n_keys = native.pack_bool_bits(mask), mask.shape[-1]
    if n_keys != HW or E // nhead != 32:
        raise RuntimeError(f"masked_cross_attention: unsupported geometry (keys {n_keys} vs memory {HW}, "
                           f"head_dim {E // nhead}; the kernel needs head_dim 32)")
    if pos.dim() == 3 and pos.shape[0] != 1:
        pos = pos[:1]
    if HW % 4 != 0:
        # The TMA row stride of V^T [B, E, HW] must be a multiple of 16 bytes: append up to three zero keys.  Every
        # mask producer already sets the bits of keys >= HW, so ordinary rows never see them; a row whose real keys
        # are ALL masked attends to every real key (ref decoder :1780) -- its mask becomes "padding keys only"
        # here instead of the kernel's row_open flag, which would let the padding in.
        pad = (-HW) % 4
        memory = F.pad(memory, (0, 0, 0, pad))
        pos = F.pad(pos, (0, 0, 0, pad))
        all_masked = (bits == -1).all(-1, keepdim=True)
        bits = torch.where(all_masked, _tail_bits(HW, bits.shape[-1], bits.device), bits)
    return _MaskedCrossAttention.apply(q_in.contiguous(), memory.contiguous(), pos.contiguous(), w_in, b_in,
                                       w_out, b_out, nhead, bits, n_keys)


class _SelfAttnCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, mask_u8, nhead):
        out, lse = native.self_attn_fwd(qkv, mask_u8, nhead)
        ctx.save_for_backward(qkv, lse)
        ctx.mask_u8, ctx.nhead = mask_u8, nhead
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, lse = ctx.saved_tensors
        return native.self_attn_bwd(qkv, ctx.mask_u8, g, lse, ctx.nhead), None, None


_MASK_U8 = {}


def _mask_u8(tgt_mask):
    """bool [Q, Q] -> uint8 copy, cached per mask tensor (the decoder builds one tgt_mask per forward and hands it to
    all nine layers)."""
    key = (tgt_mask.data_ptr(), tuple(tgt_mask.shape), tgt_mask._version)
    hit = _MASK_U8.get(key)
    if hit is None or hit[0] is not tgt_mask:
        _MASK_U8.clear()
        hit = (tgt_mask, tgt_mask.to(torch.uint8).contiguous())
        _MASK_U8[key] = hit
    return hit[1]


def self_attention(qk_in, v_in, w_in, b_in, w_out, b_out, nhead, tgt_mask=None):
    """nn.MultiheadAttention self-attention over the (few hundred) queries (ref decoder :42-52); tgt_mask
    bool [Q,Q], True = not allowed (DN groups, ref decoder :1051-1059).  Projections on the tensor-core GEMM, the
    [Q x Q] core on the fp32 kernels of csrc/self_attn.cu (head dim 32, Q <= 320; other geometries: library SDPA,
    recorded in ROUTES)."""
    _cuda_only(qk_in, "tgt")
    E = qk_in.shape[-1]
    if qk_in is v_in:                                              # no query_pos: one GEMM for q, k and v
        qkv = linear(qk_in, w_in, b_in)
    else:
        qkv = torch.cat([linear(qk_in, w_in[:2 * E], b_in[:2 * E]),      # q and k share the input
                         linear(v_in, w_in[2 * E:], b_in[2 * E:])], -1)
    if _route("self_attention.core", qkv.dim() == 3 and E // nhead == 32 and E % nhead == 0
              and qkv.shape[1] <= native.SELF_ATTN_MAX_Q and qkv.dtype == torch.float32,
              f"head_dim {E // nhead} / {qkv.shape[1]} queries"):
        o = _SelfAttnCore.apply(qkv.contiguous(), None if tgt_mask is None else _mask_u8(tgt_mask), nhead)
    else:
        q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
        o = F.scaled_dot_product_attention(_split_heads(q, nhead), _split_heads(k, nhead), _split_heads(v, nhead),
                                           attn_mask=None if tgt_mask is None else ~tgt_mask)
        o = o.transpose(1, 2).reshape(qk_in.shape)
    return linear(o, w_out, b_out)
