"""CPU checks of the host-side module logic (wiring, layouts, DN preparation, state-dict keys).

The product has no CPU path.  To exercise the Python modules without a GPU, THIS TEST swaps the
native MSDeformAttn op for the oracle and disables the CUDA-tensor guard -- a test-only monkeypatch;
the kernels themselves are validated by the ``-m gpu`` tests on the B200 box."""
import os

import pytest
import torch

import cases
import mp_former_b200 as M
from mp_former_b200 import MultiScaleDeformableAttention as MSDA
from mp_former_b200 import _lib
from oracle import torch_oracle as O
from test_oracle_vs_golden import (_cmp_out, close, decoder_template, load, pixel_decoder_template)


def cpu_pack_bits(mask):
    """bool [..., n] -> int32 [..., mask_words(n)] (bit i of word j = key 32j+i; padding = 1)."""
    from mp_former_b200 import native
    n = mask.shape[-1]
    w = native.mask_words(n)
    m = torch.ones(mask.shape[:-1] + (w * 32,), dtype=torch.bool)
    m[..., :n] = mask
    v = (m.view(*mask.shape[:-1], w, 32).to(torch.int64) << torch.arange(32)).sum(-1)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32)


def install_cpu_ops(setattr_fn):
    """Replaces every native kernel launcher by a CPU torch equivalent (test-only).
    ``setattr_fn(obj, name, value)`` performs the patch (pytest's monkeypatch.setattr or plain setattr)."""
    import torch.nn.functional as F
    from mp_former_b200 import native, ops

    def fwd(value, shapes, lsi, loc, aw, step, host_shapes=None):
        return O.msda_core(value, shapes, loc, aw)

    def linear(x, w, b=None, relu=False):
        y = F.linear(x, w, b)
        return F.relu(y) if relu else y

    def attn_mask_bits(logits, size):
        a = F.interpolate(logits, size=size, mode="bilinear", align_corners=False).flatten(2)
        return cpu_pack_bits(a <= ops.MASK_LOGIT_THRESHOLD)

    def xattn(q_in, memory, pos, w_in, b_in, w_out, b_out, nhead, mask):
        sd = {"in_proj_weight": w_in, "in_proj_bias": b_in, "out_proj.weight": w_out, "out_proj.bias": b_out}
        m = mask.to_bool() if isinstance(mask, ops.PackedMask) else mask
        am = None
        if m is not None:
            m = m & ~m.all(-1, keepdim=True)
            am = m[:, None].expand(-1, nhead, -1, -1).flatten(0, 1)
        y = O.mha(sd, "", q_in.transpose(0, 1), (memory + pos).transpose(0, 1), memory.transpose(0, 1),
                  nhead, am)
        return y.transpose(0, 1)

    def self_attn(qk_in, v_in, w_in, b_in, w_out, b_out, nhead, tgt_mask=None):
        sd = {"in_proj_weight": w_in, "in_proj_bias": b_in, "out_proj.weight": w_out, "out_proj.bias": b_out}
        B = qk_in.shape[0]
        am = None if tgt_mask is None else tgt_mask[None].expand(B * nhead, -1, -1)
        return O.mha(sd, "", qk_in.transpose(0, 1), qk_in.transpose(0, 1), v_in.transpose(0, 1), nhead,
                     am).transpose(0, 1)

    setattr_fn(ops, "self_attention", self_attn)
    setattr_fn(MSDA, "ms_deform_attn_forward", fwd)
    setattr_fn(MSDA, "enc_supported", lambda value, L, P: False)       # CPU: exercise the unfused module path
    setattr_fn(_lib, "require_cuda", lambda t, n: None)
    setattr_fn(ops, "linear", linear)
    setattr_fn(ops, "ffn", lambda x, w1, b1, w2, b2: F.linear(F.relu(F.linear(x, w1, b1)), w2, b2))
    setattr_fn(ops, "add_layer_norm", lambda x, r, norm: norm(x if r is None else x + r))
    setattr_fn(ops, "group_norm_cl", lambda x, gn, relu=False: F.relu(gn(x)) if relu else gn(x))
    setattr_fn(ops, "mask_logits", lambda e, f, shared=None: torch.einsum("bqc,bchw->bqhw", e, f))
    setattr_fn(native, "attn_mask_bits", attn_mask_bits)
    setattr_fn(native, "pack_bool_bits", cpu_pack_bits)
    setattr_fn(native, "gt_mask_area_bits", lambda masks, size: cpu_pack_bits(
        F.interpolate(masks.float().unsqueeze(1), size=size, mode="area").flatten(1) <= 1e-8))   # ref decoder :986
    setattr_fn(ops, "masked_cross_attention", xattn)
    setattr_fn(ops, "conv2d_fp32", lambda x, conv: conv._conv_forward(x, conv.weight, conv.bias))


@pytest.fixture()
def cpu_ops(monkeypatch):
    install_cpu_ops(monkeypatch.setattr)


def build_pixel_decoder():
    c = cases.PD_CFG
    shape = {k: M.ShapeSpec(channels=c["channels"][k], stride=c["strides"][k]) for k in c["channels"]}
    return M.MSDeformAttnPixelDecoder(
        shape, transformer_dropout=0.0, transformer_nheads=c["nheads"],
        transformer_dim_feedforward=c["dim_feedforward"], transformer_enc_layers=c["enc_layers"],
        conv_dim=c["conv_dim"], mask_dim=c["mask_dim"], norm="GN",
        transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval()


def build_decoder(cls=None, **kw):
    d = cases.DEC_CFG
    cls = cls or M.MultiScaleMaskedTransformerDecoderMaskDN
    extra = dict(dn_mode="points", all_lys=True, dn_label_noise_ratio=-1.0) \
        if cls is M.MultiScaleMaskedTransformerDecoderMaskDN else {}
    extra.update(kw)
    return cls(d["hidden_dim"], True, num_classes=d["num_classes"], hidden_dim=d["hidden_dim"],
               num_queries=d["num_queries"], nheads=d["nheads"], dim_feedforward=d["dim_feedforward"],
               dec_layers=d["dec_layers"], pre_norm=False, mask_dim=d["mask_dim"],
               enforce_input_project=False, **extra).eval()


def test_pixel_decoder_module_matches_reference_golden(golden_dir, cpu_ops):
    G = load(golden_dir, "pixel_decoder.pt")
    pd = build_pixel_decoder()
    assert sorted(pd.state_dict().keys()) == G["keys"]
    pd.load_state_dict(O.seeded_state_dict(pixel_decoder_template(), seed=41))
    with torch.no_grad():
        mf, enc0, ms = pd.forward_features(cases.pixel_decoder_features())
    close(mf, G["mask_features"], 5e-5)
    close(enc0, G["enc0"], 5e-5)
    assert len(ms) == 3
    for a, b in zip(ms, G["multi_scale"]):
        close(a, b, 5e-5)


@pytest.mark.parametrize("mode", ["plain", "dn", "dn2", "base"])
def test_decoder_module_matches_reference_golden(golden_dir, cpu_ops, mode):
    G = load(golden_dir, "decoder.pt")
    sd = O.seeded_state_dict(decoder_template(), seed=51)
    if mode == "base":
        dec = build_decoder(M.MultiScaleMaskedTransformerDecoder)
        sd = {k: v for k, v in sd.items() if not k.startswith("label_enc")}
    else:
        dec = build_decoder()
        assert sorted(dec.state_dict().keys()) == G["keys"]
    dec.load_state_dict(sd)
    x, mf = cases.decoder_inputs()
    dn_args = None
    if mode in ("dn", "dn2"):
        dn_args = {"tgt": cases.dn_targets(), "scalar": 1 if mode == "dn" else 2, "noise_scale": 0.0}
    with torch.no_grad():
        o = dec(x, mf, None, dn_args)
    _cmp_out(o, G[mode], 5e-5)
    if dn_args is not None:
        assert o["dn_out"]["dn_args"] == G[mode]["dn"]["dn_args"]
        _cmp_out(o["dn_out"], G[mode]["dn"], 5e-5)
    else:
        assert o["dn_out"] is None


def test_dn_label_noise_and_empty_targets(cpu_ops):
    dec = build_decoder(dn_label_noise_ratio=0.2)
    x, mf = cases.decoder_inputs()
    torch.manual_seed(0)
    o = dec(x, mf, None, {"tgt": cases.dn_targets(), "scalar": 1, "noise_scale": 0.4})
    assert o["dn_out"]["pred_masks"].shape[1] == 3
    assert torch.isfinite(o["pred_masks"]).all() and torch.isfinite(o["dn_out"]["pred_masks"]).all()
    empty = [{"labels": torch.zeros(0, dtype=torch.long), "masks": torch.zeros(0, 64, 64, dtype=torch.bool),
              "boxes": torch.zeros(0, 4)} for _ in range(2)]
    o = dec(x, mf, None, {"tgt": empty, "scalar": 1, "noise_scale": 0.0})
    assert o["dn_out"] is None


def test_split_queries_backward_equals_slicing():
    x = torch.randn(2, 7, 3, 5, requires_grad=True)
    a, b = M.ops.split_queries(x, 3)
    (a.sum() * 2 + (b ** 2).sum()).backward()
    y = x.detach().clone().requires_grad_(True)
    (y[:, :3].sum() * 2 + (y[:, 3:] ** 2).sum()).backward()
    assert torch.equal(x.grad, y.grad)
    x.grad = None
    M.ops.split_queries(x, 3)[1].sum().backward()            # only one half used: the other half's gradient is zero
    assert x.grad[:, :3].abs().sum() == 0 and x.grad[:, 3:].eq(1).all()


def test_static_query_checkpoint_migration():
    dec = build_decoder()
    sd = dec.state_dict()
    old = {k.replace("query_feat", "static_query"): v for k, v in sd.items()}
    dec2 = build_decoder()
    meta = getattr(old, "_metadata", None)
    dec2.load_state_dict(old)          # version-less checkpoint: static_query -> query_feat
    assert torch.equal(dec2.query_feat.weight, dec.query_feat.weight)


def test_cpu_tensors_are_rejected_loudly():
    value, shapes, loc, aw = cases.msda_inputs("testpy")
    st = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.tensor([0, 24])
    with pytest.raises(RuntimeError, match="CUDA"):
        MSDA.ms_deform_attn_forward(value, st, lsi, loc, aw, 2)
    with pytest.raises(RuntimeError, match="contiguous"):
        MSDA.ms_deform_attn_forward(value.transpose(2, 3), st, lsi, loc, aw, 2)


def test_registries_hold_reference_names():
    assert "MSDeformAttnPixelDecoder" in M.SEM_SEG_HEADS_REGISTRY
    assert "MultiScaleMaskedTransformerDecoder" in M.TRANSFORMER_DECODER_REGISTRY
    assert "MultiScaleMaskedTransformerDecoderMaskDN" in M.TRANSFORMER_DECODER_REGISTRY


def test_mask_threshold_equals_sigmoid_rule():
    """``sigmoid(x) < 0.5`` (fp32, the reference's boolean rule, decoder :1873) == ``x <= -0x1.7ffffep-23``
    for EVERY float32 in the only band where they could differ, plus ordinary values."""
    import numpy as np
    from mp_former_b200 import ops
    lo = int(np.float32(-1.1920929e-7).view(np.uint32))
    hi = int(np.float32(-2.3841858e-7).view(np.uint32))
    x = torch.from_numpy(np.arange(lo, hi + 1, dtype=np.uint32).view(np.float32).copy())
    assert torch.equal(x.sigmoid() < 0.5, x <= ops.MASK_LOGIT_THRESHOLD)
    g = torch.randn(1 << 20, generator=torch.Generator().manual_seed(0)) * torch.logspace(-9, 2, 1 << 20)
    g = torch.cat([g, torch.tensor([0.0, -0.0, float("inf"), -float("inf"), float("nan"), 1e-45, -1e-45])])
    assert torch.equal(g.sigmoid() < 0.5, g <= ops.MASK_LOGIT_THRESHOLD)


def test_grad_fanout_and_head_collector_fall_back_to_plain_autograd_sums():
    """ops.grad_fanout / ops.collect_mask_heads with consumers that do not speak the in-kernel accumulation protocol
    (here: einsum on CPU): gradients equal the plain autograd result, DN / matching split included, unused heads
    and missing DN gradients tolerated."""
    from mp_former_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, Qt, C, H, W, nH, n_dn = 2, 9, 8, 4, 6, 3, 4
    mf = torch.randn(B, C, H, W, generator=g, requires_grad=True)
    embeds = [torch.randn(B, Qt, C, generator=g, requires_grad=True) for _ in range(nH)]
    gouts = [torch.randn(B, Qt, H, W, generator=g) for _ in range(nH)]

    def loss_of(dn, mm):
        t = (mm[0] * gouts[0][:, n_dn:]).sum() + (dn[0] * gouts[0][:, :n_dn]).sum()
        return t + (mm[2] * gouts[2][:, n_dn:]).sum()                     # head 1 unused, head 2: no DN gradient

    aliases, shared = ops.grad_fanout(mf, nH)
    masks = [torch.einsum("bqc,bchw->bqhw", e, a) for e, a in zip(embeds, aliases)]
    dn, mm = ops.collect_mask_heads(masks, n_dn, shared)
    assert dn[0].shape == (B, n_dn, H, W) and mm[0].shape == (B, Qt - n_dn, H, W)
    loss_of(dn, mm).backward()
    got = [mf.grad.clone()] + [None if e.grad is None else e.grad.clone() for e in embeds]
    mf.grad = None
    for e in embeds:
        e.grad = None
    masks = [torch.einsum("bqc,bchw->bqhw", e, mf) for e in embeds]
    loss_of([m[:, :n_dn] for m in masks], [m[:, n_dn:] for m in masks]).backward()
    assert torch.allclose(got[0], mf.grad, atol=1e-5)
    for a, e in zip(got[1:], embeds):
        if e.grad is None:
            assert a is None or float(a.abs().max()) == 0.0
        else:
            assert torch.allclose(a, e.grad, atol=1e-5)
    # without autograd (inference) both helpers are pass-throughs
    with torch.no_grad():
        al, sh = ops.grad_fanout(mf, 2)
        assert sh is None and al[0] is mf
        d2, m2 = ops.collect_mask_heads([masks[0].detach()], 0, None)
        assert d2 is None and m2[0].shape == (B, Qt, H, W)


def test_head_collector_recognises_gradients_that_tile_one_buffer():
    """ops._HeadCollector._side_by_side: the per-head mask-logit gradients handed back by the criterion's joint
    evaluation (slices of ONE buffer along dim 1, in the criterion's head order, not the decoder's) are used in place;
    anything else (separate tensors, a hole, a wrong stride, a missing piece) falls back to the concatenation."""
    from mp_former_b200 import ops
    B, H, W = 2, 3, 4
    sizes = [2, 5, 2, 5, 2, 5]                                  # three heads: dn part, matching part
    buf = torch.arange(B * sum(sizes) * H * W, dtype=torch.float32).view(B, sum(sizes), H, W)
    pieces = list(buf.split(sizes, dim=1))
    # the criterion orders the heads (final, aux 0, aux 1) and each head (matching, dn); the collector (aux 0, aux 1,
    # final) x (dn, matching): a permutation of the same slices
    order = [3, 2, 5, 4, 1, 0]
    grads = [pieces[i] for i in order]
    got = ops._HeadCollector._side_by_side(grads, [sizes[i] for i in order], B, H, W)
    assert got is not None
    G2, rows = got
    assert G2.data_ptr() == buf.data_ptr() and G2.shape == (B, sum(sizes), H * W)
    assert torch.equal(G2, buf.view(B, sum(sizes), H * W))
    starts = [0, 2, 7, 9, 14, 16]
    assert rows == [starts[i] for i in order]
    for g, r, n in zip(grads, rows, [sizes[i] for i in order]):
        assert torch.equal(G2[:, r:r + n].reshape(B, n, H, W), g)
    # a view into the middle of a larger allocation works too (storage offset)
    big = torch.zeros(5 + buf.numel())
    inner = big[5:].view_as(buf)
    got = ops._HeadCollector._side_by_side(list(inner.split(sizes, dim=1)), sizes, B, H, W)
    assert got is not None and got[0].data_ptr() == inner.data_ptr()
    # fallbacks
    f = ops._HeadCollector._side_by_side
    assert f([p.clone() for p in pieces], sizes, B, H, W) is None                       # separate tensors
    assert f(pieces[:5] + [None], sizes, B, H, W) is None                               # a head without gradient
    assert f(pieces[:4] + [pieces[5], pieces[5]], sizes[:4] + [5, 5], B, H, W) is None  # a hole / an overlap
    wide = torch.zeros(B, sum(sizes) + 1, H, W)
    assert f(list(wide[:, :sum(sizes)].split(sizes, dim=1)), sizes, B, H, W) is None    # stride of a wider buffer
    assert f([p.double() for p in pieces], sizes, B, H, W) is None                      # dtype
