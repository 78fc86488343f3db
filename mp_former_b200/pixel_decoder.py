"""MSDeformAttn pixel decoder with the reference's module API and state-dict keys.

API mirrored (implementation is ours):
  * ``MSDeformAttnPixelDecoder``                -- ref: pixel_decoder/msdeformattn.py:164-358
      constructor kwargs == the reference's ``from_config`` output (:294-312);
      ``forward_features(features) -> (mask_features, transformer_encoder_features,
      multi_scale_features[3])`` (:314-358)
  * ``MSDeformAttnTransformerEncoderOnly`` / ``...Encoder`` / ``...EncoderLayer`` (:23-161)

State-dict keys are the reference's: ``input_proj.{i}.{0,1}.*``,
``transformer.level_embed``, ``transformer.encoder.layers.{i}.{self_attn.*,norm1,norm2,linear1,
linear2}.*``, ``mask_features.*``, ``adapter_1.{weight,norm.*}``, ``layer_1.{weight,norm.*}``.

Data layout: tokens are kept channels-last ``[B, S, C]`` end to end (conv outputs are produced in
``channels_last`` memory format so flatten/transpose are views), shapes / level starts /
reference points / position embeddings are host-known and cached per input geometry -- the
reference's per-call device->host syncs (msdeformattn.py:330-339) do not exist here.
"""
import copy
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .msdeform_attn import MSDeformAttn
from .position_encoding import PositionEmbeddingSine
from .registry import configurable, register_pixel_decoder


fp32_math = ops.fp32_math


class ShapeSpec:
    """Minimal stand-in for detectron2.layers.ShapeSpec (only .channels / .stride are read)."""

    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


class ConvNormAct(nn.Conv2d):
    """conv -> optional norm (sub-module ``norm``) -> optional activation; key layout of
    detectron2.layers.Conv2d so that checkpoints load unchanged."""

    def __init__(self, *args, norm=None, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = ops.conv2d_fp32(x, self)           # library convolution, fp32 pinned; CUDA only like every op of the path
        return self.norm_act(x)

    def norm_act(self, x):
        """norm (+ ReLU) of the convolution output; GroupNorm on channels-last maps runs on the fused kernels."""
        if isinstance(self.norm, nn.GroupNorm) and self.activation in (None, F.relu):
            relu = self.activation is F.relu
            y = ops.group_norm_nchw_to_cl(x, self.norm, relu=relu)      # NCHW convolution output -> tokens
            return y if y is not None else ops.group_norm_cl(x, self.norm, relu=relu)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def conv1x1_tokens(x, conv):
    """1x1 convolution of a (logically NCHW, channels-last in memory) map as a token GEMM on the
    tensor-core kernel: [B*H*W, Cin] x W[Cout, Cin]^T (+ bias).  Returns the map in the same form."""
    B, Cin, H, W = x.shape
    if Cin % 32 != 0 or conv.kernel_size != (1, 1) or conv.stride != (1, 1):
        return ops.conv2d_fp32(x, conv)
    if x.is_contiguous() and not x.permute(0, 2, 3, 1).is_contiguous():
        # NCHW backbone output: consumed as the MN-major operand of the token-reduction GEMM, no layout copy
        y = ops.conv1x1_nchw_to_cl(x, conv.weight, conv.bias)
        if y is not None:
            return y
    tokens = x.permute(0, 2, 3, 1)
    if not tokens.is_contiguous():
        tokens = tokens.contiguous()
    y = ops.linear(tokens.view(B * H * W, Cin), conv.weight.view(conv.out_channels, Cin), conv.bias)
    return y.view(B, H, W, conv.out_channels).permute(0, 3, 1, 2)


def _get_norm(norm, channels):
    if norm is None or (isinstance(norm, str) and norm == ""):
        return None
    if isinstance(norm, str):
        if norm == "GN":
            return nn.GroupNorm(32, channels)
        raise ValueError(f"unsupported norm {norm!r} (the MP-Former configs use 'GN')")
    return norm(channels)


def _c2_xavier_fill(module):
    # fvcore.nn.weight_init.c2_xavier_fill (used by the reference at msdeformattn.py:252,282-283)
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8,
                 n_points=4):
        super().__init__()
        if activation != "relu":
            raise RuntimeError("MSDeformAttn encoder: only relu is used by MP-Former")
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, src):
        if self.dropout2.p == 0.0 or not self.training:
            src2 = ops.ffn(src, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)
        else:
            hidden = ops.linear(src, self.linear1.weight, self.linear1.bias, relu=True)
            src2 = ops.linear(self.dropout2(hidden), self.linear2.weight, self.linear2.bias)
        return ops.add_layer_norm(src, self.dropout3(src2), self.norm2)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        dropout_off = not self.training or (self.dropout1.p == 0.0 and self.dropout2.p == 0.0 and self.dropout3.p == 0.0)
        if (padding_mask is None and dropout_off and not ops.NO_FUSED_ENCODER_LAYER and ops.encoder_layer_supported(
                src, pos, reference_points, self.linear1.out_features, self.self_attn.n_heads,
                self.self_attn.n_levels, self.self_attn.n_points)):
            # the whole layer as one autograd node (no ``src + pos``, no full-size gradient additions)
            return ops.encoder_layer(src, pos, reference_points, spatial_shapes, level_start_index, self.self_attn,
                                     self.norm1, self.linear1, self.linear2, self.norm2)
        src2 = self.self_attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes,
                              level_start_index, padding_mask)
        src = ops.add_layer_norm(src, self.dropout1(src2), self.norm1)
        return self.forward_ffn(src)


class MSDeformAttnTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """ref msdeformattn.py:141-153.  ``spatial_shapes``: list of (H, W) ints."""
        pts = []
        for lvl, (H_, W_) in enumerate(spatial_shapes):
            ref_y, ref_x = torch.meshgrid(
                torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device), indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            pts.append(torch.stack((ref_x, ref_y), -1))
        reference_points = torch.cat(pts, 1)
        return reference_points[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios=None, pos=None,
                padding_mask=None, reference_points=None):
        output = src
        if reference_points is None:
            host = getattr(spatial_shapes, "_mpf_host_shapes", None) or [tuple(s) for s in spatial_shapes.tolist()]
            reference_points = self.get_reference_points(host, valid_ratios, src.device)
        for layer in self.layers:
            output = layer(output, pos, reference_points, spatial_shapes, level_start_index, padding_mask)
        return output


class MSDeformAttnTransformerEncoderOnly(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.1,
                 activation="relu", num_feature_levels=4, enc_n_points=4):
        super().__init__()
        self.d_model, self.nhead = d_model, nhead
        layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation,
                                                    num_feature_levels, nhead, enc_n_points)
        self.encoder = MSDeformAttnTransformerEncoder(layer, num_encoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._geom_cache = {}
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.normal_(self.level_embed)

    def _geometry(self, shapes, device):
        """Per-geometry constants: int64 device tensors (API of the op) carrying their host copy,
        and the pixel-centre reference points (valid ratios are 1: no padding masks on this path)."""
        key = (tuple(shapes), str(device))
        if key not in self._geom_cache:
            st = torch.as_tensor(shapes, dtype=torch.long, device=device)
            st._mpf_host_shapes = tuple(shapes)
            sizes = [h * w for h, w in shapes]
            lsi = torch.as_tensor([0] + list(np.cumsum(sizes)[:-1]), dtype=torch.long, device=device)
            ones = torch.ones(1, len(shapes), 2, dtype=torch.float32, device=device)
            ref = MSDeformAttnTransformerEncoder.get_reference_points(shapes, ones, device)  # [1,S,L,2]
            self._geom_cache[key] = (st, lsi, ref)
        return self._geom_cache[key]

    def forward(self, srcs, pos_embeds):
        """srcs / pos_embeds: lists of [B, C, H, W] maps (low -> high resolution).  Returns
        ``(memory [B,S,C], spatial_shapes, level_start_index)`` like ref :61-89."""
        shapes = [(int(s.shape[2]), int(s.shape[3])) for s in srcs]
        B = srcs[0].shape[0]
        spatial_shapes, level_start_index, ref = self._geometry(shapes, srcs[0].device)
        src_flatten = torch.cat([s.permute(0, 2, 3, 1).flatten(1, 2) for s in srcs], 1)
        # the sine embedding of an unpadded map is batch independent (an expanded [1,...] tensor): keep it [1,S,C]
        pos_embeds = [p[:1] if p.shape[0] > 1 and p.stride(0) == 0 else p for p in pos_embeds]
        lvl_pos = torch.cat([p.permute(0, 2, 3, 1).flatten(1, 2) + self.level_embed[i].view(1, 1, -1)
                             for i, p in enumerate(pos_embeds)], 1)
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, None, lvl_pos, None,
                              reference_points=ref.expand(B, -1, -1, -1))
        return memory, spatial_shapes, level_start_index


@register_pixel_decoder
class MSDeformAttnPixelDecoder(nn.Module):
    @configurable
    def __init__(self, input_shape: Dict[str, ShapeSpec], *, transformer_dropout: float,
                 transformer_nheads: int, transformer_dim_feedforward: int, transformer_enc_layers: int,
                 conv_dim: int, mask_dim: int, norm: Optional[Union[str, Callable]] = None,
                 transformer_in_features: List[str], common_stride: int):
        super().__init__()
        tf_shape = {k: v for k, v in input_shape.items() if k in transformer_in_features}
        ordered = sorted(input_shape.items(), key=lambda x: x[1].stride)
        self.in_features = [k for k, v in ordered]
        self.feature_strides = [v.stride for k, v in ordered]
        self.feature_channels = [v.channels for k, v in ordered]
        tf_ordered = sorted(tf_shape.items(), key=lambda x: x[1].stride)
        self.transformer_in_features = [k for k, v in tf_ordered]
        tf_channels = [v.channels for k, v in tf_ordered]
        self.transformer_feature_strides = [v.stride for k, v in tf_ordered]
        self.transformer_num_feature_levels = len(self.transformer_in_features)

        chans = tf_channels[::-1] if self.transformer_num_feature_levels > 1 else [tf_channels[-1]]
        self.input_proj = nn.ModuleList([
            nn.Sequential(nn.Conv2d(c, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim)) for c in chans])
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)

        self.transformer = MSDeformAttnTransformerEncoderOnly(
            d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads,
            dim_feedforward=transformer_dim_feedforward, num_encoder_layers=transformer_enc_layers,
            num_feature_levels=self.transformer_num_feature_levels)
        self.pe_layer = PositionEmbeddingSine(conv_dim // 2, normalize=True)

        self.mask_dim = mask_dim
        self.mask_features = ConvNormAct(conv_dim, mask_dim, kernel_size=1, stride=1, padding=0)
        _c2_xavier_fill(self.mask_features)

        self.maskformer_num_feature_levels = 3
        self.common_stride = common_stride
        stride = min(self.transformer_feature_strides)
        self.num_fpn_levels = int(np.log2(stride) - np.log2(self.common_stride))
        lateral_convs, output_convs = [], []
        use_bias = norm == ""
        for idx, in_channels in enumerate(self.feature_channels[:self.num_fpn_levels]):
            lateral = ConvNormAct(in_channels, conv_dim, kernel_size=1, bias=use_bias,
                                  norm=_get_norm(norm, conv_dim))
            output = ConvNormAct(conv_dim, conv_dim, kernel_size=3, stride=1, padding=1, bias=use_bias,
                                 norm=_get_norm(norm, conv_dim), activation=F.relu)
            _c2_xavier_fill(lateral)
            _c2_xavier_fill(output)
            self.add_module("adapter_{}".format(idx + 1), lateral)
            self.add_module("layer_{}".format(idx + 1), output)
            lateral_convs.append(lateral)
            output_convs.append(output)
        self.lateral_convs = lateral_convs[::-1]
        self.output_convs = output_convs[::-1]

    @classmethod
    def from_config(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        """Same mapping as ref msdeformattn.py:294-312."""
        head = cfg.MODEL.SEM_SEG_HEAD
        return {
            "input_shape": {k: v for k, v in input_shape.items() if k in head.IN_FEATURES},
            "conv_dim": head.CONVS_DIM,
            "mask_dim": head.MASK_DIM,
            "norm": head.NORM,
            "transformer_dropout": cfg.MODEL.MASK_FORMER.DROPOUT,
            "transformer_nheads": cfg.MODEL.MASK_FORMER.NHEADS,
            "transformer_dim_feedforward": 1024,
            "transformer_enc_layers": head.TRANSFORMER_ENC_LAYERS,
            "transformer_in_features": head.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES,
            "common_stride": head.COMMON_STRIDE,
        }

    def forward_features(self, features):
        """features: dict name -> [B, C_f, H_f, W_f].  Runs in fp32 regardless of autocast (the
        reference disables autocast here, msdeformattn.py:314,320) and with TF32 off for the
        remaining library convolutions / GEMMs, so results keep fp32 parity with the reference."""
        with torch.autocast(device_type="cuda", enabled=False), fp32_math():
            return self._forward_features(features)

    def _forward_features(self, features):
        srcs, pos = [], []
        for idx, f in enumerate(self.transformer_in_features[::-1]):
            x = features[f].float()
            proj = self.input_proj[idx]
            srcs.append(ops.group_norm_cl(conv1x1_tokens(x, proj[0]), proj[1]))
            pos.append(self.pe_layer(x))
        y, spatial_shapes, _ = self.transformer(srcs, pos)
        bs = y.shape[0]
        shapes = spatial_shapes._mpf_host_shapes
        out, start = [], 0
        for (h, w) in shapes:
            z = y[:, start:start + h * w]
            start += h * w
            out.append(z.transpose(1, 2).reshape(bs, -1, h, w))   # logical NCHW, channels-last memory
        for idx, f in enumerate(self.in_features[:self.num_fpn_levels][::-1]):
            x = features[f].float()
            lateral = self.lateral_convs[idx]
            cur_fpn = lateral.norm_act(conv1x1_tokens(x, lateral))
            # cuDNN's fp32 3x3 convolution is ~7x slower on NHWC maps (46 ms vs 6.5 ms forward at [16,256,256,256],
            # profiles/README.md r1p): the sum is handed over in NCHW -- written that way by the fused
            # upsample+add kernel -- and the GroupNorm+ReLU behind the convolution returns to channels-last tokens.
            prev = out[-1].contiguous(memory_format=torch.channels_last)
            oconv = self.output_convs[idx]
            y = None
            if ops.conv3x3_cl_supported(cur_fpn, oconv):
                # everything channels-last: merge -> 3x3 conv on the tensor-core GEMM -> GroupNorm + ReLU
                merged = ops.upsample2x_add_cl(cur_fpn, prev)
                if merged is not None:
                    y = ops.conv3x3_cl(merged, oconv)
            if y is not None:
                out.append(oconv.norm_act(y))
                continue
            y = ops.upsample2x_add_to_nchw(cur_fpn, prev)
            if y is None:
                up = F.interpolate(prev, size=cur_fpn.shape[-2:], mode="bilinear", align_corners=False)
                y = (cur_fpn + up).contiguous()
            out.append(oconv(y))
        multi_scale_features = out[:self.maskformer_num_feature_levels]
        return conv1x1_tokens(out[-1], self.mask_features), out[0], multi_scale_features
