"""Device ops of the hot path, each with exactly one implementation.

``linear``, ``mask_logits``, ``attn_mask_from_logits`` and ``masked_cross_attention`` run forward AND
backward on this package's hand-written sm_100a kernels (tcgen05 split-precision GEMMs: bf16x3 with in-kernel
operand splitting and TMA-store epilogue for K-major products and for the token-reduction "TN" products of the
weight gradients, 3xTF32 for the remaining mixed-major case; bit-packed mask kernels; fused masked-attention
forward and the two backward kernels that recompute the probabilities from the saved log-sum-exp) through the C ABI.  ``self_attention`` and the tiny
query-side layers (a few hundred rows) stay on library ops (launch-latency bound).

Inputs must be CUDA fp32 tensors; there is no CPU path.
"""
import math
import os

import torch
import torch.nn.functional as F

from . import _lib, native

NATIVE_OPS = {"ms_deform_attn_forward", "ms_deform_attn_backward", "linear fwd+bwd (bf16x3 tcgen05 GEMM, TMA in/out)",
              "mask_logits fwd+bwd (bf16x3 / 3xTF32 tcgen05 GEMM)", "attn_mask_bits", "gt_mask_area_bits",
              "masked_cross_attention fwd+bwd (tcgen05)"}

# fp32 ``sigmoid(x) < 0.5`` as evaluated by the reference (1/(1+exp(-x)) with a correctly rounded exp)
# is EXACTLY ``x <= -0x1.7ffffep-23``: for -1.788e-7 < x < 0 the sigmoid rounds to 0.5 and the key stays
# unmasked.  Verified exhaustively over every float32 in [-2.4e-7, -1.2e-7] against torch CPU
# (tests/test_host_logic_cpu.py::test_mask_threshold_equals_sigmoid_rule).
MASK_LOGIT_THRESHOLD = float.fromhex("-0x1.7ffffep-23")

LOG2E = 1.4426950408889634


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
            if weight.shape[0] % 32 == 0:            # reduction dim of the input-gradient GEMM
                wt_hi, wt_lo = native.split_b(weight.t().contiguous())
                gx = native.gemm(g2.contiguous(), wt_hi, wt_lo)
            else:
                gx = g2 @ weight
            gx = gx.view(*gy.shape[:-1], weight.shape[1])
        if ctx.needs_input_grad[1]:
            if g2.shape[1] % 4 == 0 and x2.shape[1] % 4 == 0:
                gw = native.matmul_tn(g2.contiguous(), x2)      # dW = dY^T X on the tensor cores, no transposes
            else:
                gw = g2.t() @ x2
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = native.colsum(g2)
        return gx, gw, gb, None


def linear(x, weight, bias=None, relu=False):
    """y = x W^T + b (optionally ReLU) in fp32-accurate 3xTF32 on the tensor cores.
    Replaces nn.Linear at ref ops/modules/ms_deform_attn.py:98,102-103,124 and
    pixel_decoder/msdeformattn.py:116-120."""
    _cuda_only(x, "x")
    if x.shape[-1] % 32 != 0:
        y = F.linear(x, weight, bias)
        return F.relu(y) if relu else y
    return _Linear.apply(x, weight, bias, relu)


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
    if C not in (128, 256, 512) or x.dtype != torch.float32 or tuple(norm.normalized_shape) != (C,) \
            or norm.weight is None or norm.bias is None:
        return norm(x if r is None else x + r)
    return _AddLayerNorm.apply(x, r, norm.weight, norm.bias, norm.eps)


_NO_GN_KERNEL = bool(os.environ.get("MPF_NO_GN_KERNEL"))     # A/B switch for benchmarks only


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
    if (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and gn.affine and not _NO_GN_KERNEL
            and native.groupnorm_cl_ok(x.shape[1], gn.num_groups)
            and x.permute(0, 2, 3, 1).is_contiguous()):
        return _GroupNormCL.apply(x, gn.weight, gn.bias, gn.eps, gn.num_groups, relu)
    y = gn(x)
    return F.relu(y) if relu else y


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
        w2t_hi, w2t_lo = native.split_b(w2.t().contiguous())
        gh = native.gemm_general(g2, w2t_hi, b_lo=w2t_lo, gate=hidden)       # d(hidden) with the ReLU mask applied
        gw2 = native.matmul_tn(g2, hidden)
        gb2 = native.colsum(g2)
        w1t_hi, w1t_lo = native.split_b(w1.t().contiguous())
        gx = native.gemm(gh, w1t_hi, w1t_lo).view(*gy.shape[:-1], w1.shape[1]) if ctx.needs_input_grad[0] else None
        gw1 = native.matmul_tn(gh, x2)
        gb1 = native.colsum(gh)
        return gx, gw1, gb1, gw2, gb2


def ffn(x, w1, b1, w2, b2):
    """Linear -> ReLU -> Linear (ref pixel_decoder/msdeformattn.py:116-120, decoder :165-169 with
    dropout 0)."""
    _cuda_only(x, "x")
    if x.shape[-1] % 32 or w1.shape[0] % 32 or w2.shape[0] % 4:
        return linear(linear(x, w1, b1, relu=True), w2, b2)
    return _FFN.apply(x, w1, b1, w2, b2)


# ------------------------------------------------------------------------------------------------
# prediction heads: mask logits + boolean stage
# ------------------------------------------------------------------------------------------------
def _channels_last_tokens(mask_features):
    """[B, C, H, W] (any memory format) -> [B, H*W, C] contiguous view/copy."""
    t = mask_features.permute(0, 2, 3, 1)
    if not t.is_contiguous():
        t = t.contiguous()
    return t.view(t.shape[0], -1, t.shape[-1])


class _MaskLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask_embed, mask_features):
        B, C, H, W = mask_features.shape
        tokens = _channels_last_tokens(mask_features)                       # [B, HW, C]
        e_hi, e_lo = native.split_b(mask_embed)
        out = native.gemm(tokens, e_hi, e_lo, transpose_c=True)             # [B, Q, HW]
        ctx.save_for_backward(mask_embed, tokens)
        ctx.fshape = (B, C, H, W)
        return out.view(B, mask_embed.shape[1], H, W)

    @staticmethod
    def backward(ctx, g):
        mask_embed, tokens = ctx.saved_tensors
        B, C, H, W = ctx.fshape
        g2 = g.reshape(B, g.shape[1], H * W)
        ge = gf = None
        g2 = g2.contiguous()
        ok = (H * W) % 4 == 0 and C % 4 == 0
        if ctx.needs_input_grad[0]:
            # dE[b] = dOut[b] (Q x HW) @ F[b] (HW x C): a small [Q x C] result reduced over H*W = 65536, so the
            # reduction is cut into K-splits across CTAs; F is consumed MN-major (no transpose)
            ge = native.gemm_general(g2, tokens, a_mn=False, b_mn=True, k_splits=16) if ok \
                else torch.bmm(g2, tokens)
        if ctx.needs_input_grad[1]:
            # dF[b] = dOut[b]^T (HW x Q) @ E[b] (Q x C): both operands MN-major
            if ok and native.GEMM_MODE == "bf16x3":
                gf = native.gemm_tn(g2, mask_embed)              # reduction over the Q queries, both operands as stored
            elif ok:
                gf = native.gemm_general(g2, mask_embed, a_mn=True, b_mn=True)
            else:
                gf = torch.bmm(g2.transpose(1, 2), mask_embed)
            gf = gf.view(B, H, W, C).permute(0, 3, 1, 2)
        return ge, gf


def mask_logits(mask_embed, mask_features):
    """einsum('bqc,bchw->bqhw') (ref decoder :1865) as a TMA-fed tcgen05 GEMM over channels-last
    pixel features: M = H*W, N = Q, K = C, transposed store."""
    _cuda_only(mask_embed, "mask_embed")
    return _MaskLogits.apply(mask_embed.contiguous(), mask_features)


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
        g_wout = native.matmul_tn(gy2, o2)
        g_bout = gy2.sum(0)
        go = (gy2 @ w_out).view(B, Qt, E)                                           # d(attention output)
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
        g_wk = native.matmul_tn(dk2, mem2) + native.matmul_tn(dk.sum(0), pos2)
        g_win = torch.cat([native.matmul_tn(dq2, q_in.reshape(B * Qt, E)), g_wk, native.matmul_tn(dv2, mem2)], 0)
        g_bin = torch.cat([dq2.sum(0), native.colsum(dk2), native.colsum(dv2)], 0)
        g_qin = (dq2 @ wq).view(B, Qt, E) if ctx.needs_input_grad[0] else None
        g_mem = g_pos = None
        if ctx.needs_input_grad[1]:
            wkt_hi, wkt_lo = native.split_b(wk.t().contiguous())
            wvt_hi, wvt_lo = native.split_b(wv.t().contiguous())
            g_mem = (native.gemm(dk2, wkt_hi, wkt_lo) + native.gemm(dv2, wvt_hi, wvt_lo)).view(B, HW, E)
        if ctx.needs_input_grad[2]:
            g_pos = (dk.sum(0) @ wk).view(1, HW, E)
        return g_qin, g_mem, g_pos, g_win, g_bin, g_wout, g_bout, None, None, None


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
    if n_keys != HW or E // nhead != 32 or HW % 4 != 0:
        raise RuntimeError(f"masked_cross_attention: unsupported geometry (keys {n_keys} vs memory {HW}, "
                           f"head_dim {E // nhead}; the kernel needs head_dim 32 and HW % 4 == 0)")
    if pos.dim() == 3 and pos.shape[0] != 1:
        pos = pos[:1]
    return _MaskedCrossAttention.apply(q_in.contiguous(), memory.contiguous(), pos.contiguous(), w_in, b_in,
                                       w_out, b_out, nhead, bits, n_keys)


def self_attention(qk_in, v_in, w_in, b_in, w_out, b_out, nhead, tgt_mask=None):
    """nn.MultiheadAttention self-attention over the (few hundred) queries (ref decoder :42-52); tgt_mask
    bool [Q,Q], True = not allowed (DN groups, ref decoder :1051-1059).  Projections on the tensor-core
    GEMM; the [Q x Q] attention itself is a library SDPA call (launch-latency bound)."""
    _cuda_only(qk_in, "tgt")
    E = qk_in.shape[-1]
    if qk_in is v_in:                                              # no query_pos: one GEMM for q, k and v
        qkv = linear(qk_in, w_in, b_in)
        q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
    else:
        qk = linear(qk_in, w_in[:2 * E], b_in[:2 * E])             # q and k share the input
        q, k = qk[..., :E], qk[..., E:]
        v = linear(v_in, w_in[2 * E:], b_in[2 * E:])
    allowed = None if tgt_mask is None else ~tgt_mask
    o = F.scaled_dot_product_attention(_split_heads(q, nhead), _split_heads(k, nhead),
                                       _split_heads(v, nhead), attn_mask=allowed)
    o = o.transpose(1, 2).reshape(qk_in.shape)
    return linear(o, w_out, b_out)
