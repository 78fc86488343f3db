"""MSDeformAttn autograd function + module with the reference's Python API.

Mirrors (API only; the implementation is ours):
  * ``MSDeformAttnFunction``  -- ref: ops/functions/ms_deform_attn_func.py:32-49
  * ``MSDeformAttn``          -- ref: ops/modules/ms_deform_attn.py:35-125 (same constructor, same
    parameter names => same state-dict keys, same ``forward`` signature)

Differences by design: no silent ``except:`` fallback to a PyTorch path (ref modules:116-121) --
errors propagate; CPU tensors are rejected.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import MultiScaleDeformableAttention as MSDA
from . import ops


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.host_shapes = getattr(value_spatial_shapes, "_mpf_host_shapes", None)
        output = MSDA.ms_deform_attn_forward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations,
            attention_weights, im2col_step, host_shapes=ctx.host_shapes)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, loc, aw = ctx.saved_tensors
        grad_value, grad_loc, grad_aw = MSDA.ms_deform_attn_backward(
            value, shapes, level_start, loc, aw, grad_output.contiguous(), ctx.im2col_step,
            host_shapes=ctx.host_shapes)
        return grad_value, None, None, grad_loc, grad_aw, None


class MSDeformAttnEncFunction(Function):
    """MSDeformAttn with the softmax over the L*P logits and loc = ref + off / (W_l, H_l) evaluated inside
    the kernels (forward and backward): inputs are the raw projection output and the reference points."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, offsets_logits, reference_points, num_points):
        ctx.num_points = num_points
        ctx.host_shapes = getattr(spatial_shapes, "_mpf_host_shapes", None)
        out = MSDA.ms_deform_attn_enc_forward(value, spatial_shapes, level_start_index, offsets_logits,
                                              reference_points, num_points, host_shapes=ctx.host_shapes)
        ctx.save_for_backward(value, spatial_shapes, level_start_index, offsets_logits, reference_points)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, ow, ref = ctx.saved_tensors
        grad_value, grad_ow = MSDA.ms_deform_attn_enc_backward(
            value, shapes, level_start, ow, ref, grad_output.contiguous(), ctx.num_points,
            host_shapes=ctx.host_shapes)
        return grad_value, None, None, grad_ow, None, None


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    """Multi-scale deformable attention.  Parameters (and therefore checkpoints) are those of the
    reference module: ``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj``."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(
                d_model, n_heads))
        self.im2col_step = 128
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # Same initial state as the reference (ref modules:61-80): zero offset weights, bias = a ring
        # of directions (one per head) scaled by the point index; uniform attention; xavier projections.
        nn.init.constant_(self.sampling_offsets.weight, 0.0)
        theta = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        ring = torch.stack([theta.cos(), theta.sin()], -1)
        ring = ring / ring.abs().max(-1, keepdim=True)[0]
        grid = ring.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        grid = grid * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.reshape(-1))
        nn.init.constant_(self.attention_weights.weight, 0.0)
        nn.init.constant_(self.attention_weights.bias, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.0)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes,
                input_level_start_index, input_padding_mask=None):
        """Same contract as ref modules:82-125.
        query (N, Lq, C); reference_points (N, Lq, L, 2|4); input_flatten (N, S, C);
        input_spatial_shapes (L, 2) int64; input_level_start_index (L,); padding mask (N, S) bool."""
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        host_shapes = getattr(input_spatial_shapes, "_mpf_host_shapes", None)
        if host_shapes is not None:
            assert sum(h * w for h, w in host_shapes) == Len_in
        else:  # reference behaviour: device-side check (implies a sync)
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
        M, L, P = self.n_heads, self.n_levels, self.n_points

        value = ops.linear(input_flatten, self.value_proj.weight, self.value_proj.bias)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, M, self.d_model // M)
        # sampling offsets and attention logits share the input: one GEMM with N = M*L*P*3
        n_off = M * L * P * 2
        ow = ops.linear(query, torch.cat([self.sampling_offsets.weight, self.attention_weights.weight], 0),
                        torch.cat([self.sampling_offsets.bias, self.attention_weights.bias], 0))
        if reference_points.shape[-1] == 2 and MSDA.enc_supported(value, L, P) and not reference_points.requires_grad:
            # fused path: softmax + location arithmetic inside the MSDeformAttn kernels
            output = MSDeformAttnEncFunction.apply(value.contiguous(), input_spatial_shapes,
                                                   input_level_start_index, ow.contiguous(), reference_points, P)
            return ops.linear(output, self.output_proj.weight, self.output_proj.bias)
        offsets = ow[..., :n_off].reshape(N, Len_q, M, L, P, 2)
        weights = F.softmax(ow[..., n_off:].reshape(N, Len_q, M, L * P), -1)
        weights = weights.view(N, Len_q, M, L, P)
        if reference_points.shape[-1] == 2:
            normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            locations = reference_points[:, :, None, :, None, :] \
                + offsets / normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            locations = reference_points[:, :, None, :, None, :2] \
                + offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))
        output = MSDeformAttnFunction.apply(value.contiguous(), input_spatial_shapes,
                                            input_level_start_index, locations.contiguous(),
                                            weights.contiguous(), self.im2col_step)
        return ops.linear(output, self.output_proj.weight, self.output_proj.bias)
