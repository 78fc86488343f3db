"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch, fp32/fp64) of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file; the product package ``mp_former_b200`` never does.

Every function cites the reference file:line it restates (paths relative to /root/reference/).
The restatement is pinned against outputs of the UNMODIFIED reference (imported with
``oracle/ref_loader.py`` in the authoring container) through the fixtures in ``tests/golden/``
-- see ``tests/golden/make_golden.py`` and ``tests/test_oracle_vs_golden.py``.  The reference itself
ships no golden vectors for this path; its only self-test pins the CUDA op to
``ms_deform_attn_core_pytorch``, which is what the goldens were generated with.

All functions are functional: parameters come from a state dict ``sd`` with the reference's key
names.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# a1/a4: MSDeformAttn core
# --------------------------------------------------------------------------------------------
def msda_core(value, spatial_shapes, sampling_locations, attention_weights):
    """out[b,q,m,:] = sum_{l,p} aw * bilinear_zero_pad(value_l, loc).

    Restates the CUDA kernel's arithmetic directly (ref: ops/src/cuda/ms_deform_im2col_cuda.cuh:38-89
    bilinear with per-corner zero padding, :277-303 pixel coords ``loc*size-0.5`` and the
    ``-1 < h_im < H`` sample gate), which the reference equates with
    ``ms_deform_attn_core_pytorch`` (ref: ops/functions/ms_deform_attn_func.py:52-72,
    ops/test.py:34-63).  ``spatial_shapes``: list of (H, W) python ints or an int tensor.
    value [N,S,M,D]; sampling_locations [N,Lq,M,L,P,2]; attention_weights [N,Lq,M,L,P] -> [N,Lq,M*D].
    Differentiable (autograd supplies the backward oracle, ref cuh:92-164)."""
    shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes)
                                             else spatial_shapes)]
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    out = value.new_zeros(N, Lq, M, D)
    start = 0
    b_idx = torch.arange(N, device=value.device).view(N, 1, 1, 1).expand(N, Lq, M, P)
    m_idx = torch.arange(M, device=value.device).view(1, 1, M, 1).expand(N, Lq, M, P)
    for l, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W].reshape(N, H, W, M, D)
        start += H * W
        loc = sampling_locations[:, :, :, l]                       # [N,Lq,M,P,2]
        w_im = loc[..., 0] * W - 0.5
        h_im = loc[..., 1] * H - 0.5
        inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h_low = torch.floor(h_im)
        w_low = torch.floor(w_im)
        lh, lw = h_im - h_low, w_im - w_low
        hh, hw = 1 - lh, 1 - lw
        h0, w0 = h_low.long(), w_low.long()
        acc = 0
        for dh, dw, wt in ((0, 0, hh * hw), (0, 1, hh * lw), (1, 0, lh * hw), (1, 1, lh * lw)):
            hi, wi = h0 + dh, w0 + dw
            ok = inside & (hi >= 0) & (hi <= H - 1) & (wi >= 0) & (wi <= W - 1)
            g = v[b_idx, hi.clamp(0, H - 1), wi.clamp(0, W - 1), m_idx]      # [N,Lq,M,P,D]
            acc = acc + (wt * ok.to(value.dtype)).unsqueeze(-1) * g
        out = out + (attention_weights[:, :, :, l].unsqueeze(-1) * acc).sum(3)
    return out.reshape(N, Lq, M * D)


# --------------------------------------------------------------------------------------------
# a8: sine position embedding
# --------------------------------------------------------------------------------------------
def msda_core_grid_sample(value, spatial_shapes, sampling_locations, attention_weights):
    """The reference's own CPU path, as it computes it (ref ops/functions/ms_deform_attn_func.py:52-72): per level one
    ``F.grid_sample`` (bilinear, zeros padding, align_corners=False) of the level's [N*M, D, H, W] view at the grid
    ``2 * loc - 1``, then the weighted sum over levels and points.  Same result as :func:`msda_core` (which restates
    the CUDA kernel) up to fp32 rounding; kept separately because it is the arithmetic whose SPEED on the host cores
    bench.py reports as the reference's CPU MSDeformAttn baseline."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes)
                                             else spatial_shapes)]
    grids = 2 * sampling_locations - 1
    sampled, start = [], 0
    for l, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W].flatten(2).transpose(1, 2).reshape(N * M, D, H, W)
        start += H * W
        g = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)                     # [N*M, Lq, P, 2]
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    w = attention_weights.transpose(1, 2).reshape(N * M, 1, Lq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * w).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()


def position_embedding_sine(B, H, W, num_pos_feats=128, temperature=10000, scale=2 * math.pi, device=None,
                            dtype=torch.float32):
    """ref: transformer_decoder/position_encoding.py:29-52 with mask=None, normalize=True."""
    y_embed = torch.arange(1, H + 1, dtype=torch.float32, device=device).view(1, H, 1).expand(B, H, W)
    x_embed = torch.arange(1, W + 1, dtype=torch.float32, device=device).view(1, 1, W).expand(B, H, W)
    eps = 1e-6
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2).to(dtype)


# --------------------------------------------------------------------------------------------
# a5: MSDeformAttn module
# --------------------------------------------------------------------------------------------
def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def msdeform_attn_module(sd, prefix, query, reference_points, input_flatten, shapes, n_heads, n_points,
                         padding_mask=None):
    """ref: ops/modules/ms_deform_attn.py:82-125 (2-d reference points branch :106-109)."""
    N, Lq, C = query.shape
    S = input_flatten.shape[1]
    L = len(shapes)
    value = _lin(sd, prefix + "value_proj", input_flatten)
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)
    value = value.view(N, S, n_heads, C // n_heads)
    off = _lin(sd, prefix + "sampling_offsets", query).view(N, Lq, n_heads, L, n_points, 2)
    aw = _lin(sd, prefix + "attention_weights", query).view(N, Lq, n_heads, L * n_points)
    aw = F.softmax(aw, -1).view(N, Lq, n_heads, L, n_points)
    normalizer = torch.tensor([[w, h] for h, w in shapes], dtype=query.dtype, device=query.device)
    loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    out = msda_core(value, shapes, loc, aw)
    return _lin(sd, prefix + "output_proj", out)


# --------------------------------------------------------------------------------------------
# a6/a7: pixel decoder
# --------------------------------------------------------------------------------------------
def encoder_reference_points(shapes, B, dtype=torch.float32, device=None):
    """ref: pixel_decoder/msdeformattn.py:141-153 with valid_ratios == 1."""
    pts = []
    for H, W in shapes:
        ry = torch.linspace(0.5, H - 0.5, H, dtype=torch.float32, device=device) / H
        rx = torch.linspace(0.5, W - 0.5, W, dtype=torch.float32, device=device) / W
        gy, gx = torch.meshgrid(ry, rx, indexing="ij")
        pts.append(torch.stack((gx.reshape(-1), gy.reshape(-1)), -1))
    ref = torch.cat(pts, 0)[None].expand(B, -1, -1)                    # [B,S,2]
    return ref[:, :, None, :].expand(-1, -1, len(shapes), -1).to(dtype)


def pixel_decoder_forward(sd, features, *, transformer_in_features=("res3", "res4", "res5"),
                          in_features=("res2", "res3", "res4", "res5"), n_heads=8, n_points=4,
                          enc_layers=6, prefix=""):
    """ref: pixel_decoder/msdeformattn.py:314-358 (forward_features), :61-89 (EncoderOnly.forward),
    :122-131 (encoder layer), :155-161 (encoder).  Returns (mask_features, out[0], multi_scale[3])."""
    p = prefix
    srcs, poss = [], []
    for idx, f in enumerate(transformer_in_features[::-1]):
        x = features[f].float()
        y = F.conv2d(x, sd[f"{p}input_proj.{idx}.0.weight"], sd[f"{p}input_proj.{idx}.0.bias"])
        y = F.group_norm(y, 32, sd[f"{p}input_proj.{idx}.1.weight"], sd[f"{p}input_proj.{idx}.1.bias"])
        srcs.append(y)
        poss.append(position_embedding_sine(x.shape[0], x.shape[2], x.shape[3], y.shape[1] // 2, device=x.device))
    B, C = srcs[0].shape[:2]
    shapes = [(s.shape[2], s.shape[3]) for s in srcs]
    src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    lvl = sd[p + "transformer.level_embed"]
    pos = torch.cat([q.flatten(2).transpose(1, 2) + lvl[i].view(1, 1, -1) for i, q in enumerate(poss)], 1)
    ref = encoder_reference_points(shapes, B, device=src.device)
    for i in range(enc_layers):
        lp = f"{p}transformer.encoder.layers.{i}."
        src2 = msdeform_attn_module(sd, lp + "self_attn.", src + pos, ref, src, shapes, n_heads, n_points,
                                    padding_mask=torch.zeros(B, src.shape[1], dtype=torch.bool, device=src.device))
        src = F.layer_norm(src + src2, (C,), sd[lp + "norm1.weight"], sd[lp + "norm1.bias"])
        ff = _lin(sd, lp + "linear2", F.relu(_lin(sd, lp + "linear1", src)))
        src = F.layer_norm(src + ff, (C,), sd[lp + "norm2.weight"], sd[lp + "norm2.bias"])
    out, start = [], 0
    for H, W in shapes:
        out.append(src[:, start:start + H * W].transpose(1, 2).reshape(B, C, H, W))
        start += H * W
    # one extra FPN level per stride octave between the finest transformer level and common stride
    n_fpn = 1
    for idx, f in enumerate(list(in_features[:n_fpn])[::-1]):
        k = n_fpn - idx
        x = features[f].float()
        cur = F.conv2d(x, sd[f"{p}adapter_{k}.weight"], sd.get(f"{p}adapter_{k}.bias"))
        cur = F.group_norm(cur, 32, sd[f"{p}adapter_{k}.norm.weight"], sd[f"{p}adapter_{k}.norm.bias"])
        y = cur + F.interpolate(out[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
        y = F.conv2d(y, sd[f"{p}layer_{k}.weight"], sd.get(f"{p}layer_{k}.bias"), padding=1)
        y = F.relu(F.group_norm(y, 32, sd[f"{p}layer_{k}.norm.weight"], sd[f"{p}layer_{k}.norm.bias"]))
        out.append(y)
    mask_features = F.conv2d(out[-1], sd[p + "mask_features.weight"], sd[p + "mask_features.bias"])
    return mask_features, out[0], out[:3]


# --------------------------------------------------------------------------------------------
# a9-a13: masked-attention transformer decoder
# --------------------------------------------------------------------------------------------
def mha(sd, prefix, query, key, value, n_heads, attn_mask=None):
    """torch.nn.MultiheadAttention forward as used by the reference (seq-first, dropout 0, bool
    attn_mask [B*h, Lq, Lk] with True = not allowed; ref decoder :105-108, :47-48).  Math of
    torch.nn.functional.multi_head_attention_forward: q,k,v in-projections, q scaled by
    1/sqrt(head_dim), masked scores -> softmax -> @v -> out_proj."""
    Lq, B, E = query.shape
    Lk = key.shape[0]
    hd = E // n_heads
    w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    q = q.reshape(Lq, B * n_heads, hd).transpose(0, 1) * math.sqrt(1.0 / hd)
    k = k.reshape(Lk, B * n_heads, hd).transpose(0, 1)
    v = v.reshape(Lk, B * n_heads, hd).transpose(0, 1)
    scores = torch.bmm(q, k.transpose(1, 2))
    if attn_mask is not None:
        scores = scores.masked_fill(attn_mask, float("-inf"))
    probs = F.softmax(scores, dim=-1)
    out = torch.bmm(probs, v).transpose(0, 1).reshape(Lq, B, E)
    return _lin(sd, prefix + "out_proj", out)


def prediction_heads(sd, p, output, mask_features, target_size, n_heads):
    """ref: transformer_decoder/mask2former_transformer_decoder.py:1859-1877 (== :525-542)."""
    C = output.shape[-1]
    dec = F.layer_norm(output, (C,), sd[p + "decoder_norm.weight"], sd[p + "decoder_norm.bias"])
    dec = dec.transpose(0, 1)
    outputs_class = _lin(sd, p + "class_embed", dec)
    e = F.relu(_lin(sd, p + "mask_embed.layers.0", dec))
    e = F.relu(_lin(sd, p + "mask_embed.layers.1", e))
    e = _lin(sd, p + "mask_embed.layers.2", e)
    outputs_mask = torch.einsum("bqc,bchw->bqhw", e, mask_features)
    attn_mask = attn_mask_from_logits(outputs_mask, target_size, n_heads)
    return outputs_class, outputs_mask, attn_mask


def attn_mask_from_logits(outputs_mask, target_size, n_heads):
    """The boolean stage of the heads (must be matched bit-exactly on identical logits):
    bilinear resize (align_corners=False) -> sigmoid -> ``< 0.5`` -> repeat over heads
    (ref decoder :1869-1875)."""
    a = F.interpolate(outputs_mask, size=target_size, mode="bilinear", align_corners=False)
    a = (a.sigmoid().flatten(2).unsqueeze(1).repeat(1, n_heads, 1, 1).flatten(0, 1) < 0.5).bool()
    return a.detach()


def _dn_gt_masks(targets, size, scalar):
    """ref decoder :986-987 / :1603-1605: area-downsampled GT masks, True where (almost) no GT."""
    return torch.cat([F.interpolate(t["masks"].float().unsqueeze(1), size=size, mode="area").flatten(1) <= 1e-8
                      for t in targets if len(t["masks"]) > 0]).repeat(scalar, 1)


def _dn_noise(masks, noise_scale, hw):
    """ref decoder :995-998: point-flip noise proportional to the mask area."""
    areas = (~masks).sum(1)
    ratio = areas * noise_scale / hw
    delta = torch.rand_like(masks, dtype=torch.float) < ratio[:, None]
    return torch.logical_xor(masks, delta)


def decoder_forward(sd, x, mask_features, *, num_queries, n_heads=8, dec_layers=9, num_classes=80,
                    dn_args=None, dn_label_noise_ratio=-1.0, all_lys=True, prefix="", trace=None):
    """``MultiScaleMaskedTransformerDecoderMaskDN.forward`` with ``dn_mode='points'``
    (ref decoder :1706-1857; prepare_for_normal :729-735; prepare_for_dn_v5 :968-1060;
    gen_mask_dn :1584-1622; postprocess_for_dn :1697-1703).  With ``dn_args=None`` this is also
    ``MultiScaleMaskedTransformerDecoder.forward`` (:427-523) up to the ``label_enc*0`` term.
    ``input_proj`` is the identity (in_channels == hidden_dim, no enforce_input_project).

    ``trace`` (test aid, optional dict): receives ``"masks"`` -- per cross-attention layer the boolean attention mask
    it consumed ([B, Qt, hw], head 0; the heads share it, ref :1875), after the mask-piloted rows were written and
    before the all-masked-row rule of :1780 -- and ``"resized"``, the bilinearly resized mask logits those bits were
    thresholded from ([B, Qt, hw], ref :1869), so that a test can teacher-force another implementation's layers with
    these masks and check that any bit it would have computed differently sits on the threshold."""
    p = prefix
    B, C = x[0].shape[:2]
    size_list, src, pos = [], [], []
    lvl = sd[p + "level_embed.weight"]
    for i in range(3):
        H, W = x[i].shape[-2:]
        size_list.append((H, W))
        pos.append(position_embedding_sine(B, H, W, C // 2, device=x[i].device).flatten(2).permute(2, 0, 1))
        src.append((x[i].flatten(2) + lvl[i][None, :, None]).permute(2, 0, 1))
    qf = sd[p + "query_feat.weight"]
    tgt_mask, dn_meta, known = None, None, None
    scalar = 0
    if dn_args is not None:
        targets, scalar, noise_scale = dn_args["tgt"], dn_args["scalar"], dn_args["noise_scale"]
        num_boxes = [len(t["boxes"]) for t in targets]
        max_num = max(num_boxes)
        if scalar >= 100:
            scalar = scalar // max_num
        if max_num == 0 or scalar == 0:
            dn_args = None
    if dn_args is None:
        output = qf.unsqueeze(1).repeat(1, B, 1)
        outputs_class, outputs_mask, attn_mask = prediction_heads(sd, p, output, mask_features, size_list[0], n_heads)
    else:
        single_pad = max_num
        pad_size = scalar * max_num
        dn_meta = {"max_num": max_num, "pad_size": pad_size}
        hw0 = size_list[0][0] * size_list[0][1]
        padding = torch.zeros(B, pad_size, C, device=x[0].device)
        padding_mask = torch.ones(B, pad_size, hw0, dtype=torch.bool, device=x[0].device)
        masks = _dn_noise(_dn_gt_masks(targets, size_list[0], scalar), noise_scale, hw0)
        labels = torch.cat([t["labels"] for t in targets])
        known_labels = labels.repeat(scalar, 1).view(-1).clone()
        if dn_label_noise_ratio > 0:
            prob = torch.rand_like(known_labels.float())
            chosen = prob < dn_label_noise_ratio
            known_labels[chosen] = torch.randint_like(known_labels[chosen], 0, num_classes)
        feats = sd[p + "label_enc.weight"][known_labels]
        batch_idx = torch.cat([torch.full_like(t["labels"].long(), i) for i, t in enumerate(targets)])
        known_bid = batch_idx.repeat(scalar, 1).view(-1)
        idx = torch.cat([torch.arange(n, device=x[0].device) for n in num_boxes])
        map_idx = torch.cat([idx + single_pad * i for i in range(scalar)]).long()
        known = (known_bid, map_idx)
        padding[known] = feats
        padding_mask[known] = masks
        padding_mask = padding_mask.unsqueeze(1).repeat(1, n_heads, 1, 1)
        output = torch.cat([padding.transpose(0, 1), qf.unsqueeze(1).repeat(1, B, 1)], 0)
        outputs_class, outputs_mask, attn_mask = prediction_heads(sd, p, output, mask_features, size_list[0], n_heads)
        attn_mask = attn_mask.view(B, n_heads, -1, attn_mask.shape[-1])
        attn_mask[:, :, :-num_queries] = padding_mask
        attn_mask = attn_mask.flatten(0, 1)
        tgt_size = pad_size + num_queries
        tgt_mask = torch.zeros(tgt_size, tgt_size, dtype=torch.bool, device=x[0].device)
        tgt_mask[pad_size:, :pad_size] = True
        for i in range(scalar):
            tgt_mask[single_pad * i:single_pad * (i + 1), single_pad * (i + 1):pad_size] = True
            tgt_mask[single_pad * i:single_pad * (i + 1), :single_pad * i] = True

    pred_class, pred_mask = [outputs_class], [outputs_mask]
    if trace is not None:
        trace["masks"], trace["resized"] = [], []
        resized = F.interpolate(outputs_mask, size=size_list[0], mode="bilinear", align_corners=False).flatten(2)
    for i in range(dec_layers):
        li = i % 3
        if trace is not None:
            trace["masks"].append(attn_mask.view(B, n_heads, -1, attn_mask.shape[-1])[:, 0].clone())
            trace["resized"].append(resized.detach())
        attn_mask = attn_mask.clone()
        attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False      # ref :1780
        cp = f"{p}transformer_cross_attention_layers.{i}."
        t2 = mha(sd, cp + "multihead_attn.", output, src[li] + pos[li], src[li], n_heads, attn_mask)
        output = F.layer_norm(output + t2, (C,), sd[cp + "norm.weight"], sd[cp + "norm.bias"])
        sp = f"{p}transformer_self_attention_layers.{i}."
        t2 = mha(sd, sp + "self_attn.", output, output, output, n_heads,
                 None if tgt_mask is None else tgt_mask[None].expand(B * n_heads, -1, -1))
        output = F.layer_norm(output + t2, (C,), sd[sp + "norm.weight"], sd[sp + "norm.bias"])
        fp = f"{p}transformer_ffn_layers.{i}."
        t2 = _lin(sd, fp + "linear2", F.relu(_lin(sd, fp + "linear1", output)))
        output = F.layer_norm(output + t2, (C,), sd[fp + "norm.weight"], sd[fp + "norm.bias"])
        level = (i + 1) % 3
        outputs_class, outputs_mask, attn_mask = prediction_heads(sd, p, output, mask_features, size_list[level], n_heads)
        if trace is not None:
            resized = F.interpolate(outputs_mask, size=size_list[level], mode="bilinear", align_corners=False).flatten(2)
        if dn_args is not None and (all_lys or i < 3):
            hw = size_list[level][0] * size_list[level][1]
            pm = torch.ones(B, dn_meta["pad_size"], hw, dtype=torch.bool, device=x[0].device)
            pm[known] = _dn_noise(_dn_gt_masks(dn_args["tgt"], size_list[level], scalar), dn_args["noise_scale"], hw)
            pm = pm.unsqueeze(1).repeat(1, n_heads, 1, 1)
            attn_mask = attn_mask.view(B, n_heads, -1, attn_mask.shape[-1])
            attn_mask[:, :, :-num_queries] = pm
            attn_mask = attn_mask.flatten(0, 1)
        pred_class.append(outputs_class)
        pred_mask.append(outputs_mask)

    def aux(cls, msk):
        return [{"pred_logits": a, "pred_masks": b} for a, b in zip(cls[:-1], msk[:-1])]

    if tgt_mask is not None:
        dn_cls = [c[:, :-num_queries] for c in pred_class]
        dn_msk = [m[:, :-num_queries] for m in pred_mask]
        pred_class = [c[:, -num_queries:] for c in pred_class]
        pred_mask = [m[:, -num_queries:] for m in pred_mask]
        dn_out = {"pred_logits": dn_cls[-1], "pred_masks": dn_msk[-1], "aux_outputs": aux(dn_cls, dn_msk),
                  "dn_args": dn_meta}
    else:
        dn_out = None
    return {"pred_logits": pred_class[-1], "pred_masks": pred_mask[-1],
            "aux_outputs": aux(pred_class, pred_mask), "dn_out": dn_out}


# --------------------------------------------------------------------------------------------
# deterministic parameters / inputs shared by golden generation and tests
# --------------------------------------------------------------------------------------------
def seeded_state_dict(template_sd, seed, scale=None):
    """Fills every floating-point entry of ``template_sd`` (name -> tensor, only shapes are used)
    with seeded values, iterating names in sorted order.  Norm weights are 1 + 0.1*randn, biases
    0.1*randn (0.5*randn for sampling_offsets so samples spread over several texels), matrices
    randn/sqrt(fan_in).  Used to give the reference module, the oracle and the product module the
    same parameters without storing them."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(template_sd.keys()):
        t = template_sd[name]
        if not torch.is_floating_point(t):
            out[name] = t.clone()
            continue
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        if t.dim() <= 1:
            if name.endswith("weight") and ("norm" in name or name.endswith(".1.weight")):
                v = 1.0 + 0.1 * r
            elif "sampling_offsets" in name:
                v = 0.5 * r
            else:
                v = 0.1 * r
        else:
            fan_in = t[0].numel() if t.dim() > 1 else t.numel()
            if "query_feat" in name or "level_embed" in name or "label_enc" in name:
                v = r
            elif "sampling_offsets" in name:
                v = r * (0.3 / math.sqrt(fan_in))
            else:
                v = r / math.sqrt(fan_in)
        out[name] = v.to(t.dtype)
    return out
