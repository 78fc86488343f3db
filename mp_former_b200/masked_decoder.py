"""Masked-attention transformer decoder (MP-Former / Mask2Former) with the reference's module API.

API mirrored (implementation is ours):
  * ``MultiScaleMaskedTransformerDecoder``        -- ref: transformer_decoder/
        mask2former_transformer_decoder.py:209-555
  * ``MultiScaleMaskedTransformerDecoderMaskDN``  -- ref: same file :558-1917, the class MP-Former
        trains with (run_50ep_no_noise_all_ly.sh:19), ``dn_mode='points'`` (prepare_for_dn_v5
        :968-1060, gen_mask_dn :1584-1622)
  ``forward(x, mask_features, mask=None, dn_args=None) -> {"pred_logits","pred_masks",
  "aux_outputs","dn_out"}``; constructor kwargs == the reference's ``from_config`` output.

State-dict keys are the reference's: the attention layers keep
``nn.MultiheadAttention`` modules purely as parameter containers (``in_proj_weight`` ...), the
arithmetic is done by this package's ops.

Differences by design (results unchanged):
  * the attention mask is ONE boolean map per (image, query) shared by all heads -- the reference
    materialises 8 identical copies (decoder :1875);
  * the "row fully masked -> attend everywhere" rule (decoder :1780) is applied inside the
    attention op from a per-row flag, without the reference's host-synchronising ``torch.where``;
  * DN (mask-piloted) preparation runs on the tensors' device without ``.cuda()`` uploads.
Other ``dn_mode`` values of the reference are experimental alternates outside the published recipe
 and raise ``NotImplementedError``.
"""
import logging

import torch
import torch.nn.functional as F
from torch import nn

from . import native, ops
from .position_encoding import PositionEmbeddingSine
from .registry import configurable, register_transformer_decoder


class _AttnParams(nn.Module):
    """post-norm attention block parameters: ``<name>.{in_proj_*, out_proj.*}`` + ``norm``."""

    def __init__(self, d_model, nhead, attn_name, dropout=0.0, normalize_before=False):
        super().__init__()
        setattr(self, attn_name, nn.MultiheadAttention(d_model, nhead, dropout=dropout))
        self.norm = nn.LayerNorm(d_model)
        self.nhead = nhead
        self.normalize_before = normalize_before
        if normalize_before:
            raise NotImplementedError("pre_norm=True is not used by any MP-Former config")
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class SelfAttentionLayer(_AttnParams):
    """ref decoder :19-74 (post-norm path)."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__(d_model, nhead, "self_attn", dropout, normalize_before)

    def forward(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        """tgt: [B, Q, C] (batch-first); tgt_mask: bool [Q, Q], True = not allowed."""
        assert tgt_key_padding_mask is None
        qk = tgt if query_pos is None else tgt + query_pos
        a = self.self_attn
        out = ops.self_attention(qk, tgt, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight,
                                 a.out_proj.bias, self.nhead, tgt_mask)
        return ops.add_layer_norm(tgt, out, self.norm)


class CrossAttentionLayer(_AttnParams):
    """ref decoder :77-137 (post-norm path)."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__(d_model, nhead, "multihead_attn", dropout, normalize_before)

    def forward(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None,
                query_pos=None):
        """tgt [B,Q,C]; memory / pos [B,HW,C] (pos may be [1,HW,C]); memory_mask: ops.PackedMask or
        bool [B,Q,HW] shared by all heads, True = not allowed; rows that are entirely True attend
        everywhere."""
        assert memory_key_padding_mask is None
        q_in = tgt if query_pos is None else tgt + query_pos
        a = self.multihead_attn
        out = ops.masked_cross_attention(q_in, memory, pos, a.in_proj_weight, a.in_proj_bias,
                                         a.out_proj.weight, a.out_proj.bias, self.nhead, memory_mask)
        return ops.add_layer_norm(tgt, out, self.norm)


class FFNLayer(nn.Module):
    """ref decoder :140-180 (post-norm path)."""

    def __init__(self, d_model, dim_feedforward=2048, dropout=0.0, activation="relu",
                 normalize_before=False):
        super().__init__()
        if normalize_before or activation != "relu":
            raise NotImplementedError("only post-norm relu FFN is used by MP-Former")
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, tgt):
        if self.dropout.p == 0.0 or not self.training:
            tgt2 = ops.ffn(tgt, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)
        else:
            hidden = ops.linear(tgt, self.linear1.weight, self.linear1.bias, relu=True)
            tgt2 = ops.linear(self.dropout(hidden), self.linear2.weight, self.linear2.bias)
        return ops.add_layer_norm(tgt, self.dropout(tgt2), self.norm)


class MLP(nn.Module):
    """ref decoder :194-206."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = ops.linear(x, layer.weight, layer.bias, relu=i < self.num_layers - 1)
        return x


class _HeadFeatures:
    """The per-head aliases of ``mask_features`` handed out in call order (ops.grad_fanout)."""

    def __init__(self, aliases, shared):
        self.aliases, self.shared, self.used = aliases, shared, 0

    def next(self):
        if self.used >= len(self.aliases):
            raise RuntimeError("more prediction heads than aliases of mask_features")
        a = self.aliases[self.used]
        self.used += 1
        return a, self.shared

    @property
    def shape(self):
        return self.aliases[0].shape

    @property
    def device(self):
        return self.aliases[0].device


class _MaskedDecoderBase(nn.Module):
    _version = 2

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        # checkpoint key migration of the reference (decoder :214-235): static_query -> query_feat
        version = local_metadata.get("version", None)
        if version is None or version < 2:
            renamed = False
            for k in list(state_dict.keys()):
                if k.startswith(prefix) and "static_query" in k:
                    state_dict[k.replace("static_query", "query_feat")] = state_dict.pop(k)
                    renamed = True
            if renamed:
                logging.getLogger(__name__).warning(
                    f"Weight format of {self.__class__.__name__} have changed! "
                    "Please upgrade your models. Applying automatic conversion now ...")
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)

    def _build(self, in_channels, mask_classification, num_classes, hidden_dim, num_queries, nheads,
               dim_feedforward, dec_layers, pre_norm, mask_dim, enforce_input_project):
        assert mask_classification, "Only support mask classification model"
        self.mask_classification = mask_classification
        self.pe_layer = PositionEmbeddingSine(hidden_dim // 2, normalize=True)
        self.num_heads = nheads
        self.num_classes = num_classes
        self.num_layers = dec_layers
        self.transformer_self_attention_layers = nn.ModuleList()
        self.transformer_cross_attention_layers = nn.ModuleList()
        self.transformer_ffn_layers = nn.ModuleList()
        for _ in range(dec_layers):
            self.transformer_self_attention_layers.append(
                SelfAttentionLayer(hidden_dim, nheads, dropout=0.0, normalize_before=pre_norm))
            self.transformer_cross_attention_layers.append(
                CrossAttentionLayer(hidden_dim, nheads, dropout=0.0, normalize_before=pre_norm))
            self.transformer_ffn_layers.append(
                FFNLayer(hidden_dim, dim_feedforward, dropout=0.0, normalize_before=pre_norm))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.num_queries = num_queries
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.num_feature_levels = 3
        self.level_embed = nn.Embedding(self.num_feature_levels, hidden_dim)
        self.input_proj = nn.ModuleList()
        for _ in range(self.num_feature_levels):
            if in_channels != hidden_dim or enforce_input_project:
                conv = nn.Conv2d(in_channels, hidden_dim, kernel_size=1)
                nn.init.kaiming_uniform_(conv.weight, a=1)
                nn.init.constant_(conv.bias, 0)
                self.input_proj.append(conv)
            else:
                self.input_proj.append(nn.Sequential())
        if self.mask_classification:
            self.class_embed = nn.Linear(hidden_dim, num_classes + 1)
        self.mask_embed = MLP(hidden_dim, hidden_dim, mask_dim, 3)

    @staticmethod
    def _base_config(cfg, in_channels, mask_classification):
        """ref decoder :697-727 / :339-367."""
        mf = cfg.MODEL.MASK_FORMER
        assert mf.DEC_LAYERS >= 1
        return {
            "in_channels": in_channels, "mask_classification": mask_classification,
            "num_classes": cfg.MODEL.SEM_SEG_HEAD.NUM_CLASSES, "hidden_dim": mf.HIDDEN_DIM,
            "num_queries": mf.NUM_OBJECT_QUERIES, "nheads": mf.NHEADS,
            "dim_feedforward": mf.DIM_FEEDFORWARD, "dec_layers": mf.DEC_LAYERS - 1,
            "pre_norm": mf.PRE_NORM, "enforce_input_project": mf.ENFORCE_INPUT_PROJ,
            "mask_dim": cfg.MODEL.SEM_SEG_HEAD.MASK_DIM,
        }

    # ---- shared forward pieces ---------------------------------------------------------------
    def _memory(self, x):
        """Per level: keys' source [B,HW,C] (= input_proj(x)+level_embed) and position [1,HW,C]."""
        src, pos, size_list = [], [], []
        for i in range(self.num_feature_levels):
            H, W = int(x[i].shape[-2]), int(x[i].shape[-1])
            size_list.append((H, W))
            pos.append(self.pe_layer.channels_last(H, W, x[i].device).flatten(1, 2))
            s = self.input_proj[i](x[i]).permute(0, 2, 3, 1).flatten(1, 2)
            src.append(s + self.level_embed.weight[i].view(1, 1, -1))
        return src, pos, size_list

    def _fanout_mask_features(self, mask_features):
        """One alias of ``mask_features`` per prediction head (num_layers + 1 of them), channels-last, with the
        heads' gradients accumulated in the GEMM epilogue (ops.grad_fanout)."""
        mask_features = mask_features.contiguous(memory_format=torch.channels_last)
        aliases, shared = ops.grad_fanout(mask_features, self.num_layers + 1)
        return _HeadFeatures(aliases, shared)

    def forward_prediction_heads(self, output, mask_features, attn_mask_target_size):
        """output [B,Q,C] -> (outputs_class [B,Q,K+1], outputs_mask [B,Q,H,W],
        attn_mask: ops.PackedMask, one bit per (image, query, key), shared by heads).
        ref decoder :1859-1877.  ``mask_features``: the map, or the ``_HeadFeatures`` of ``_fanout_mask_features``."""
        decoder_output = ops.add_layer_norm(output, None, self.decoder_norm)
        outputs_class = ops.linear(decoder_output, self.class_embed.weight, self.class_embed.bias)
        mask_embed = self.mask_embed(decoder_output)
        if isinstance(mask_features, _HeadFeatures):
            mf, shared = mask_features.next()
            outputs_mask = ops.mask_logits(mask_embed, mf, shared)
        else:
            outputs_mask = ops.mask_logits(mask_embed, mask_features)
        attn_mask = ops.attn_mask_from_logits(outputs_mask, attn_mask_target_size)
        return outputs_class, outputs_mask, attn_mask

    # Parity instrumentation (tests / bench.py's parity block; None in production).  A dict with
    #   "own":   list that receives, per cross-attention layer, the PackedMask this decoder derived itself
    #            (after the mask-piloted rows were written);
    #   "force": optional list of bool [B, Qt, hw] masks, one per layer, consumed INSTEAD of the decoder's own
    #            (teacher forcing with the oracle's masks, so that one flipped bit of a logit that sits on the
    #            threshold cannot compound over the following layers).
    mask_debug = None

    def _decode(self, output, src, pos, size_list, mask_features, tgt_mask, heads0, dn_hook=None):
        outputs_class, outputs_mask, attn_mask = heads0
        predictions_class, predictions_mask = [outputs_class], [outputs_mask]
        dbg = self.mask_debug
        for i in range(self.num_layers):
            li = i % self.num_feature_levels
            if dbg is not None:
                dbg["own"].append(attn_mask)
                if dbg.get("force") is not None:
                    attn_mask = ops.PackedMask.from_bool(dbg["force"][i].to(output.device))
            output = self.transformer_cross_attention_layers[i](
                output, src[li], memory_mask=attn_mask, pos=pos[li], query_pos=None)
            output = self.transformer_self_attention_layers[i](output, tgt_mask=tgt_mask)
            output = self.transformer_ffn_layers[i](output)
            level = (i + 1) % self.num_feature_levels
            outputs_class, outputs_mask, attn_mask = self.forward_prediction_heads(
                output, mask_features, size_list[level])
            if dn_hook is not None:
                attn_mask = dn_hook(i, level, attn_mask)
            predictions_class.append(outputs_class)
            predictions_mask.append(outputs_mask)
        assert len(predictions_class) == self.num_layers + 1
        return predictions_class, predictions_mask

    @torch.jit.unused
    def _set_aux_loss(self, outputs_class, outputs_seg_masks):
        if self.mask_classification:
            return [{"pred_logits": a, "pred_masks": b}
                    for a, b in zip(outputs_class[:-1], outputs_seg_masks[:-1])]
        return [{"pred_masks": b} for b in outputs_seg_masks[:-1]]


@register_transformer_decoder
class MultiScaleMaskedTransformerDecoder(_MaskedDecoderBase):
    @configurable
    def __init__(self, in_channels, mask_classification=True, *, num_classes: int, hidden_dim: int,
                 num_queries: int, nheads: int, dim_feedforward: int, dec_layers: int, pre_norm: bool,
                 mask_dim: int, enforce_input_project: bool):
        super().__init__()
        self._build(in_channels, mask_classification, num_classes, hidden_dim, num_queries, nheads,
                    dim_feedforward, dec_layers, pre_norm, mask_dim, enforce_input_project)

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        return cls._base_config(cfg, in_channels, mask_classification)

    def forward(self, x, mask_features, mask=None, dn_args=None):
        """ref decoder :427-523 (its ``dn_args`` branch calls the feature-DN ``prepare_for_dn``
        :369-419, which no MP-Former recipe uses)."""
        with torch.autocast(device_type="cuda", enabled=False):
            return self._forward([t.float() for t in x], mask_features.float(), mask, dn_args)

    def _forward(self, x, mask_features, mask=None, dn_args=None):
        assert len(x) == self.num_feature_levels
        del mask
        if dn_args:
            raise NotImplementedError("MultiScaleMaskedTransformerDecoder: feature-DN (dn_args) is not "
                                      "part of the MP-Former recipe; use ...DecoderMaskDN")
        src, pos, size_list = self._memory(x)
        bs = src[0].shape[0]
        mask_features = self._fanout_mask_features(mask_features)
        output = self.query_feat.weight.unsqueeze(0).repeat(bs, 1, 1)
        heads0 = self.forward_prediction_heads(output, mask_features, size_list[0])
        pc, pm = self._decode(output, src, pos, size_list, mask_features, None, heads0)
        _, pm = ops.collect_mask_heads(pm, 0, mask_features.shared)
        return {"pred_logits": pc[-1], "pred_masks": pm[-1],
                "aux_outputs": self._set_aux_loss(pc if self.mask_classification else None, pm),
                "dn_out": None}


@register_transformer_decoder
class MultiScaleMaskedTransformerDecoderMaskDN(_MaskedDecoderBase):
    @configurable
    def __init__(self, in_channels, mask_classification=True, *, num_classes: int, hidden_dim: int,
                 num_queries: int, nheads: int, dim_feedforward: int, dec_layers: int, pre_norm: bool,
                 mask_dim: int, enforce_input_project: bool, dn_mode="base", head_dn=False,
                 all_lys=False, dn_ratio=0.5, dn_label_noise_ratio=-1.0):
        super().__init__()
        self._build(in_channels, mask_classification, num_classes, hidden_dim, num_queries, nheads,
                    dim_feedforward, dec_layers, pre_norm, mask_dim, enforce_input_project)
        self.head_dn = head_dn
        self.dn_ratio = dn_ratio
        self.dn_label_noise_ratio = dn_label_noise_ratio
        self.dn_mode = dn_mode
        self.all_lys = all_lys
        self.matching_dict = dict()
        self.label_enc = nn.Embedding(num_classes, hidden_dim)
        if head_dn:
            raise NotImplementedError("head_dn=True (per-head GT noise) is not part of the MP-Former recipe")

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        ret = cls._base_config(cfg, in_channels, mask_classification)
        mf = cfg.MODEL.MASK_FORMER
        ret.update(dn_mode=mf.DN_MODE, head_dn=mf.HEAD_DN, all_lys=mf.ALL_LY_DN, dn_ratio=mf.DN_RATIO,
                   dn_label_noise_ratio=mf.LB_NOISE_RATIO)
        return ret

    # ---- mask-piloted (DN) preparation, on device --------------------------------------------
    @staticmethod
    def _gt_masked(targets, size, scalar, noise_scale, cache=None):
        """Area-downsampled GT masks -> True where the cell holds (almost) no GT pixel
        (ref decoder :986-987), optional point-flip noise (:994-998).  The reference recomputes the
        (deterministic) down-sampling for every layer; ``cache`` (dict, one forward) keeps it per size."""
        key = (int(size[0]), int(size[1]))
        if cache is not None and key in cache:
            masks = cache[key]
        else:
            masks = torch.cat([F.interpolate(t["masks"].float().unsqueeze(1), size=size, mode="area").flatten(1) <= 1e-8
                               for t in targets if len(t["masks"]) > 0]).repeat(scalar, 1)
            if cache is not None:
                cache[key] = masks
        if noise_scale == 0:
            return masks                      # the xor below would be with an all-False delta
        areas = (~masks).sum(1)
        ratio = areas * noise_scale / (size[0] * size[1])
        delta = torch.rand_like(masks, dtype=torch.float) < ratio[:, None]
        return torch.logical_xor(masks, delta)

    @staticmethod
    def _gt_packed(targets, size, scalar, known, bs, pad_size, cache):
        """Noise-free mask-piloted rows as packed bits [bs, pad_size, words], straight from the GT instance masks
        (one kernel per image instead of float conversion + area pooling + compare + scatter + pack); rows
        without a GT instance are fully masked, as in ref decoder :1036-1039."""
        key = ("bits", int(size[0]), int(size[1]))
        if key not in cache:
            rows = [native.gt_mask_area_bits(t["masks"], size) for t in targets if len(t["masks"]) > 0]
            gt = torch.cat(rows).repeat(scalar, 1)
            full = torch.full((bs, pad_size, gt.shape[-1]), -1, dtype=torch.int32, device=gt.device)
            full[known] = gt
            cache[key] = ops.PackedMask(full, int(size[0]) * int(size[1]))
        return cache[key]

    @staticmethod
    def _dn_indices(dn_args, num_boxes, scalar, single_pad, dev):
        """(batch index, slot index) of every mask-piloted query (ref decoder :1020-1029).  They depend only on the
        host-known instance counts; built once per ``dn_args`` and kept in it, so a second call with the same
        targets (e.g. under CUDA-graph capture) performs no host-to-device copy."""
        cache = dn_args.setdefault("_mpf_index_cache", {})
        key = (str(dev), int(scalar), int(single_pad), tuple(num_boxes))
        if key not in cache:
            bs = len(num_boxes)
            batch_idx = torch.repeat_interleave(torch.arange(bs), torch.as_tensor(num_boxes))
            idx = torch.cat([torch.arange(n) for n in num_boxes])
            map_idx = torch.cat([idx + single_pad * i for i in range(scalar)]).long()
            cache[key] = (batch_idx.repeat(scalar).to(dev), map_idx.to(dev))
        return cache[key]

    def prepare_for_dn_v5(self, mask_features, dn_args, size_list, cache=None):
        """ref decoder :968-1060.  Returns None when there is nothing to denoise."""
        targets, scalar, noise_scale = dn_args["tgt"], dn_args["scalar"], dn_args["noise_scale"]
        num_boxes = [len(t["boxes"]) for t in targets]
        single_pad = max_num = max(num_boxes)
        if scalar >= 100:
            scalar = scalar // max_num
        if max_num == 0 or scalar == 0:
            return None
        dev = mask_features.device
        pad_size = scalar * max_num
        dn_meta = {"max_num": max_num, "pad_size": pad_size}
        bs = len(num_boxes)
        hw0 = size_list[0][0] * size_list[0][1]
        cache = {} if cache is None else cache
        labels = torch.cat([t["labels"] for t in targets]).to(dev)
        known_labels = labels.repeat(scalar, 1).view(-1).clone()
        if self.dn_label_noise_ratio > 0:
            # same law as ref :1007-1015 (each label is replaced w.p. LB_NOISE_RATIO by a uniform class), written
            # without boolean-mask indexing so that it neither synchronises nor breaks CUDA-graph capture
            prob = torch.rand_like(known_labels.float())
            chosen = prob < self.dn_label_noise_ratio
            known_labels = torch.where(chosen, torch.randint_like(known_labels, 0, self.num_classes), known_labels)
        feats = self.label_enc(known_labels)
        known = self._dn_indices(dn_args, num_boxes, scalar, single_pad, dev)
        padding = torch.zeros(bs, pad_size, feats.shape[-1], device=dev, dtype=feats.dtype)
        padding = padding.index_put(known, feats)
        if noise_scale == 0:
            gt_rows = self._gt_packed(targets, size_list[0], scalar, known, bs, pad_size, cache)
        else:
            masks = self._gt_masked(targets, size_list[0], scalar, noise_scale, cache)
            padding_mask = torch.ones(bs, pad_size, hw0, dtype=torch.bool, device=dev)
            padding_mask[known] = masks
            gt_rows = ops.PackedMask.from_bool(padding_mask)
        output = torch.cat([padding, self.query_feat.weight.unsqueeze(0).repeat(bs, 1, 1)], 1)
        oc, om, attn_mask = self.forward_prediction_heads(output, mask_features, size_list[0])
        attn_mask = attn_mask.replace_rows(gt_rows, pad_size)
        tgt_size = pad_size + self.num_queries
        tgt_mask = torch.zeros(tgt_size, tgt_size, dtype=torch.bool, device=dev)
        tgt_mask[pad_size:, :pad_size] = True
        for i in range(scalar):
            tgt_mask[single_pad * i:single_pad * (i + 1), single_pad * (i + 1):pad_size] = True
            tgt_mask[single_pad * i:single_pad * (i + 1), :single_pad * i] = True
        return known, output, tgt_mask, dn_meta, scalar, (oc, om, attn_mask)

    def gen_mask_dn(self, dn_args, size, known, pad_size, scalar, cache=None):
        """ref decoder :1584-1622 -> bool [B, pad_size, h*w]."""
        bs = len(dn_args["tgt"])
        masks = self._gt_masked(dn_args["tgt"], size, scalar, dn_args["noise_scale"], cache)
        pm = torch.ones(bs, pad_size, size[0] * size[1], dtype=torch.bool, device=masks.device)
        pm[known] = masks
        return pm

    def forward(self, x, mask_features, mask=None, dn_args=None):
        """ref decoder :1706-1857.  The kernels compute in fp32 split precision whatever the trainer's autocast state
        is (the reference runs this module in fp16 under AMP; its pixel decoder already forces fp32,
        msdeformattn.py:314): autocast is switched off inside, so no library op of the path is re-cast behind the
        custom kernels' backs."""
        with torch.autocast(device_type="cuda", enabled=False):
            return self._forward([t.float() for t in x], mask_features.float(), mask, dn_args)

    def _forward(self, x, mask_features, mask=None, dn_args=None):
        assert len(x) == self.num_feature_levels
        del mask
        src, pos, size_list = self._memory(x)
        bs = src[0].shape[0]
        mask_features = self._fanout_mask_features(mask_features)
        res = None
        if dn_args is not None:
            if self.dn_mode != "points":
                raise NotImplementedError(
                    f"dn_mode={self.dn_mode!r}: only 'points' (the published MP-Former recipe) is implemented")
            gt_cache = {}
            res = self.prepare_for_dn_v5(mask_features, dn_args, size_list, gt_cache)
        dn_hook, tgt_mask, dn_meta = None, None, None
        if res is None:
            output = self.query_feat.weight.unsqueeze(0).repeat(bs, 1, 1)
            heads0 = self.forward_prediction_heads(output, mask_features, size_list[0])
        else:
            known, output, tgt_mask, dn_meta, scalar, heads0 = res
            pad_size = dn_meta["pad_size"]

            def dn_hook(i, level, attn_mask):
                if not (self.all_lys or i < 3):
                    return attn_mask
                if dn_args["noise_scale"] == 0:
                    packed = self._gt_packed(dn_args["tgt"], size_list[level], scalar, known, bs, pad_size, gt_cache)
                    return attn_mask.replace_rows(packed, pad_size)
                pm = self.gen_mask_dn(dn_args, size_list[level], known, pad_size, scalar, gt_cache)
                return attn_mask.replace_rows(ops.PackedMask.from_bool(pm), pad_size)

        pc, pm = self._decode(output, src, pos, size_list, mask_features, tgt_mask, heads0, dn_hook)
        if tgt_mask is not None:
            n_dn = pc[0].shape[1] - self.num_queries
            pc_split = [ops.split_queries(c, n_dn) for c in pc]
            dn_c, pc = [c[0] for c in pc_split], [c[1] for c in pc_split]
            # mask logits of all heads: split AND collected, so that the heads' backward GEMMs run batched
            dn_m, pm = ops.collect_mask_heads(pm, n_dn, mask_features.shared)
            dn_out = {"pred_logits": dn_c[-1], "pred_masks": dn_m[-1],
                      "aux_outputs": self._set_aux_loss(dn_c if self.mask_classification else None, dn_m),
                      "dn_args": dn_meta}
        else:
            dn_out = None
            _, pm = ops.collect_mask_heads(pm, 0, mask_features.shared)
            pc[-1] = pc[-1] + self.label_enc.weight[0, 0] * 0.0      # keeps label_enc in the DDP graph (ref :1846)
        return {"pred_logits": pc[-1], "pred_masks": pm[-1],
                "aux_outputs": self._set_aux_loss(pc if self.mask_classification else None, pm),
                "dn_out": dn_out}
