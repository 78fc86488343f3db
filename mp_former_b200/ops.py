"""Device ops of the transformer decoder's hot path.

Each function is the single entry the modules call; there is exactly one implementation per op
(no backend dispatch).  Inputs must be CUDA tensors.

Status (round 1): ``mask_logits`` / ``attn_mask_from_logits`` / ``masked_cross_attention`` are being
moved onto hand-written sm_100a kernels behind the C ABI; an op that is listed in
``NATIVE_OPS`` runs on our kernels, the others still call PyTorch CUDA library ops (cuBLAS /
SDPA) and are counted as library calls in bench.py.
"""
import math

import torch
import torch.nn.functional as F

from . import _lib

NATIVE_OPS = {"ms_deform_attn_forward", "ms_deform_attn_backward"}


def _cuda_only(t, name):
    _lib.require_cuda(t, name)


def mask_logits(mask_embed, mask_features):
    """einsum('bqc,bchw->bqhw') (ref decoder :1865).  mask_embed [B,Q,C], mask_features [B,C,H,W]."""
    _cuda_only(mask_embed, "mask_embed")
    return torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)


def attn_mask_from_logits(outputs_mask, target_size):
    """bool [B,Q,h*w], True = key not allowed: bilinear resize (align_corners=False) -> sigmoid ->
    ``< 0.5`` (ref decoder :1869-1875, without the 8x head repeat).  Detached."""
    _cuda_only(outputs_mask, "outputs_mask")
    a = F.interpolate(outputs_mask.detach(), size=target_size, mode="bilinear", align_corners=False)
    return a.sigmoid().flatten(2) < 0.5


def _split_heads(x, nhead):
    B, L, E = x.shape
    return x.view(B, L, nhead, E // nhead).transpose(1, 2)       # [B,h,L,hd]


def masked_cross_attention(q_in, memory, pos, w_in, b_in, w_out, b_out, nhead, mask):
    """softmax((q Wq)(k Wk)^T / sqrt(hd) + mask) (v Wv) Wo with k = memory + pos, v = memory
    (ref decoder :100-112 through nn.MultiheadAttention), mask bool [B,Q,HW] shared by heads;
    a row that is entirely masked attends to every key (ref decoder :1780)."""
    _cuda_only(q_in, "tgt")
    E = q_in.shape[-1]
    q = F.linear(q_in, w_in[:E], b_in[:E])
    k = F.linear(memory + pos, w_in[E:2 * E], b_in[E:2 * E])
    v = F.linear(memory, w_in[2 * E:], b_in[2 * E:])
    allowed = None
    if mask is not None:
        full = mask.all(-1, keepdim=True)
        allowed = (~mask | full)[:, None]                          # [B,1,Q,HW] True = attend
    o = F.scaled_dot_product_attention(_split_heads(q, nhead), _split_heads(k, nhead),
                                       _split_heads(v, nhead), attn_mask=allowed)
    o = o.transpose(1, 2).reshape(q_in.shape)
    return F.linear(o, w_out, b_out)


def self_attention(qk_in, v_in, w_in, b_in, w_out, b_out, nhead, tgt_mask=None):
    """nn.MultiheadAttention self-attention over the queries (ref decoder :42-52); tgt_mask bool
    [Q,Q], True = not allowed (DN groups, ref decoder :1051-1059)."""
    _cuda_only(qk_in, "tgt")
    E = qk_in.shape[-1]
    q = F.linear(qk_in, w_in[:E], b_in[:E])
    k = F.linear(qk_in, w_in[E:2 * E], b_in[E:2 * E])
    v = F.linear(v_in, w_in[2 * E:], b_in[2 * E:])
    allowed = None if tgt_mask is None else ~tgt_mask
    o = F.scaled_dot_product_attention(_split_heads(q, nhead), _split_heads(k, nhead),
                                       _split_heads(v, nhead), attn_mask=allowed)
    o = o.transpose(1, 2).reshape(qk_in.shape)
    return F.linear(o, w_out, b_out)
