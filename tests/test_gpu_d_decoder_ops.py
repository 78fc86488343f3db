"""-m gpu: the decoder-side kernels (bit-packed mask stage, fused masked cross-attention, split-precision
linear / mask-logit GEMMs) against the oracle's torch restatement evaluated in fp64.
Tolerance TOL*max(1,|ref|) for fp32 results -- 1e-4 with the 3xTF32 GEMM, 5e-4 with the default bf16x3 GEMM
(16 significand bits per operand; the inputs here are unscaled N(0,1) so |ref| ~ 16 at K = 256) -- against
the contract's 1e-3; mask bits exact."""
import math

import pytest
import torch
import torch.nn.functional as F

import cases
from mp_former_b200 import native, ops
from oracle import torch_oracle as O
from test_oracle_vs_golden import load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4 if native.GEMM_MODE == "tf32x3" else 5e-4


def rel(a, b):
    return ((a.double() - b.double()).abs() / b.double().abs().clamp(min=1.0)).max().item()


def test_attn_mask_bits_bit_exact_vs_reference_golden(golden_dir):
    G = load(golden_dir, "attn_mask_bits.pt")
    logits = cases.threshold_logits().to(DEV)
    for key, ref in G["masks"].items():
        h, w = (int(v) for v in key.split("x"))
        pm = ops.attn_mask_from_logits(logits, (h, w))
        assert pm.bits.shape[-1] == native.mask_words(h * w)
        got = pm.to_bool().cpu()
        assert torch.equal(got, ref.view(1, 8, 4, h * w)[:, 0]), key
        # padding bits (keys >= h*w) are set
        full = native.unpack_bits(pm.bits, pm.bits.shape[-1] * 32).cpu()
        assert full[..., h * w:].all()


@pytest.mark.parametrize("H,W,h,w", [(256, 256, 32, 32), (256, 256, 64, 64), (256, 256, 128, 128),
                                     (64, 128, 8, 16), (20, 20, 7, 9), (16, 16, 16, 16)])
def test_attn_mask_bits_vs_torch_resize(H, W, h, w):
    g = torch.Generator(device=DEV).manual_seed(H + h)
    logits = torch.randn(2, 5, H, W, device=DEV, generator=g) * torch.logspace(-8, 1, 5, device=DEV).view(1, 5, 1, 1)
    ref = F.interpolate(logits.cpu(), size=(h, w), mode="bilinear", align_corners=False).flatten(2) <= ops.MASK_LOGIT_THRESHOLD
    got = ops.attn_mask_from_logits(logits, (h, w)).to_bool().cpu()
    mism = (got != ref).float().mean().item()
    if H % h == 0 and W % w == 0 and (H // h) in (1, 2, 4, 8):
        assert mism == 0.0                  # power-of-two factors: every product is exact -> bit exact
    else:
        assert mism < 1e-5                  # general factors: fma / rounding-order differences only


def test_pack_unpack_roundtrip():
    g = torch.Generator(device=DEV).manual_seed(0)
    for n in (4, 33, 64, 100, 1024):
        m = torch.rand(3, 7, n, device=DEV, generator=g) < 0.3
        bits = native.pack_bool_bits(m)
        assert bits.shape == (3, 7, native.mask_words(n))
        assert torch.equal(native.unpack_bits(bits, n), m)


def torch_xattn(q_in, memory, pos, w_in, b_in, w_out, b_out, nhead, mask, dtype=torch.float64):
    sd = {"in_proj_weight": w_in.to(dtype), "in_proj_bias": b_in.to(dtype), "out_proj.weight": w_out.to(dtype),
          "out_proj.bias": b_out.to(dtype)}
    m = mask & ~mask.all(-1, keepdim=True)
    am = m[:, None].expand(-1, nhead, -1, -1).flatten(0, 1)
    y = O.mha(sd, "", q_in.to(dtype).transpose(0, 1), (memory + pos).to(dtype).transpose(0, 1),
              memory.to(dtype).transpose(0, 1), nhead, am)
    return y.transpose(0, 1)


@pytest.mark.parametrize("B,Qt,HW", [(2, 100, 1024), (1, 130, 256), (3, 10, 4), (2, 37, 100), (1, 128, 64),
                                     (2, 120, 4096), (2, 37, 475), (1, 220, 950), (1, 9, 63), (2, 300, 2048)])
def test_masked_cross_attention_forward_backward(B, Qt, HW):
    """HW % 4 != 0 (19x25, 25x38 maps of images padded to multiples of 32): keys are padded inside the op."""
    E, nhead = 256, 8
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + Qt + HW)
    rn = lambda *s, sc=1.0: torch.randn(*s, device=DEV, generator=g) * sc
    q_in, memory, pos = rn(B, Qt, E), rn(B, HW, E), rn(1, HW, E)
    w_in, b_in = rn(3 * E, E, sc=1 / 16), rn(3 * E, sc=0.1)
    w_out, b_out = rn(E, E, sc=1 / 16), rn(E, sc=0.1)
    mask = torch.rand(B, Qt, HW, device=DEV, generator=g) < 0.7
    mask[0, 0] = True                      # fully masked row -> attends everywhere (ref decoder :1780)
    mask[0, 1] = True
    mask[0, 1, HW - 1] = False             # exactly one open key (the last one)
    mask[-1, Qt - 1] = False               # nothing masked
    leaves = [t.clone().requires_grad_(True) for t in (q_in, memory, w_in, b_in, w_out, b_out)]
    y = ops.masked_cross_attention(leaves[0], leaves[1], pos, leaves[2], leaves[3], leaves[4], leaves[5], nhead,
                                   ops.PackedMask.from_bool(mask))
    gy = rn(B, Qt, E)
    y.backward(gy)
    ref_leaves = [t.double().clone().requires_grad_(True) for t in (q_in, memory, w_in, b_in, w_out, b_out)]
    yr = torch_xattn(ref_leaves[0], ref_leaves[1], pos.double(), ref_leaves[2], ref_leaves[3], ref_leaves[4],
                     ref_leaves[5], nhead, mask)
    yr.backward(gy.double())
    assert torch.isfinite(y).all()
    assert rel(y, yr) < TOL, rel(y, yr)
    for name, a, b_ in zip(("q_in", "memory", "w_in", "b_in", "w_out", "b_out"), leaves, ref_leaves):
        scale = b_.grad.abs().max().item()
        err = (a.grad.double() - b_.grad).abs().max().item() / max(scale, 1e-6)
        assert err < 2 * TOL, (name, err)
    # a bool mask is accepted as well and gives the same result
    y2 = ops.masked_cross_attention(q_in, memory, pos, w_in, b_in, w_out, b_out, nhead, mask)
    assert torch.equal(y2, y.detach())


@pytest.mark.parametrize("B,Qt,HW,splits", [(2, 120, 4096, 1), (2, 120, 4096, 5), (2, 120, 4096, 64), (1, 130, 950, 3),
                                            (2, 9, 16384, 9)])
def test_masked_cross_attention_key_splits_agree(monkeypatch, B, Qt, HW, splits):
    """The key-split path (small batches: every (query tile, head, image) spread over several CTAs, log-sum-exp merge
    of the partial outputs, fixed-order sum of the partial dQ) against the unsplit kernels on identical inputs,
    including rows whose open keys all fall into ONE split and splits that are fully masked for a row."""
    E, nhead = 256, 8
    g = torch.Generator(device=DEV).manual_seed(B * 77 + Qt + HW + splits)
    rn = lambda *s, sc=1.0: torch.randn(*s, device=DEV, generator=g) * sc
    q_in, memory, pos = rn(B, Qt, E), rn(B, HW, E), rn(1, HW, E)
    w_in, b_in = rn(3 * E, E, sc=1 / 16), rn(3 * E, sc=0.1)
    w_out, b_out = rn(E, E, sc=1 / 16), rn(E, sc=0.1)
    mask = torch.rand(B, Qt, HW, device=DEV, generator=g) < 0.9
    mask[0, 0] = True                                   # fully masked row
    mask[0, 1] = True
    mask[0, 1, HW - 3] = False                          # a single open key, in the last split
    mask[0, 2] = True
    mask[0, 2, : min(64, HW)] = False                   # open keys only in the first key tile
    gy = rn(B, Qt, E)
    res = {}
    for tag, ks in (("unsplit", 1), ("split", splits)):
        monkeypatch.setattr(native, "XATTN_KEY_SPLITS", ks)
        leaves = [t.clone().requires_grad_(True) for t in (q_in, memory, w_in, b_in, w_out, b_out)]
        y = ops.masked_cross_attention(leaves[0], leaves[1], pos, leaves[2], leaves[3], leaves[4], leaves[5], nhead,
                                       ops.PackedMask.from_bool(mask))
        y.backward(gy)
        res[tag] = [y.detach()] + [t.grad for t in leaves]
    for a, b_ in zip(res["split"], res["unsplit"]):
        assert torch.isfinite(a).all()
        # both paths sit ~1e-5 .. 4e-5 (relative to the largest entry) from the fp64 result: 3xTF32 products summed in
        # a different order (benchmarks/debug_xattn_split.py prints both against fp64)
        assert (a - b_).abs().max().item() <= 2e-4 * max(1.0, b_.abs().max().item())


def test_linear_and_mask_logits_autograd():
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(2, 300, 256, device=DEV, generator=g, requires_grad=True)
    w = (torch.randn(288, 256, device=DEV, generator=g) / 16).requires_grad_(True)
    b = torch.randn(288, device=DEV, generator=g).requires_grad_(True)
    for relu in (False, True):
        for t in (x, w, b):
            t.grad = None
        y = ops.linear(x, w, b, relu=relu)
        gy = torch.randn(y.shape, device=DEV, generator=g)
        y.backward(gy)
        xr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, w, b))
        yr = F.linear(xr, wr, br)
        if relu:        # same activation pattern as the kernel: pre-activations within rounding of 0 may flip sign
            assert ((yr > 0) != (y > 0)).float().mean().item() < 1e-4
            yr = yr * (y.detach() > 0)
        yr.backward(gy.double())
        assert rel(y, yr) < TOL
        assert rel(x.grad, xr.grad) < TOL and rel(w.grad, wr.grad) < 1e-3 and rel(b.grad, br.grad) < 1e-4
    e = torch.randn(2, 100, 256, device=DEV, generator=g, requires_grad=True)
    f = torch.randn(2, 256, 32, 48, device=DEV, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = ops.mask_logits(e, f)
    go = torch.randn(out.shape, device=DEV, generator=g)
    out.backward(go)
    er, fr = e.detach().double().requires_grad_(True), f.detach().double().requires_grad_(True)
    outr = torch.einsum("bqc,bchw->bqhw", er, fr)
    outr.backward(go.double())
    assert rel(out, outr) < TOL and rel(e.grad, er.grad) < 1e-3 and rel(f.grad, fr.grad) < 1e-3
    # NCHW-contiguous features are accepted too (one layout copy)
    assert rel(ops.mask_logits(e.detach(), f.detach().contiguous()), outr) < TOL


def test_ffn_fused_relu_backward():
    g = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn(3, 500, 256, device=DEV, generator=g, requires_grad=True)
    w1 = (torch.randn(1024, 256, device=DEV, generator=g) / 16).requires_grad_(True)
    b1 = torch.randn(1024, device=DEV, generator=g).requires_grad_(True)
    w2 = (torch.randn(256, 1024, device=DEV, generator=g) / 32).requires_grad_(True)
    b2 = torch.randn(256, device=DEV, generator=g).requires_grad_(True)
    y = ops.ffn(x, w1, b1, w2, b2)
    gy = torch.randn(y.shape, device=DEV, generator=g)
    y.backward(gy)
    refs = [t.detach().double().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    # the reference uses the kernel's own activation pattern (the hidden layer of ops.ffn is bit-identical to
    # ops.linear(..., relu=True)): pre-activations within rounding of 0 may flip sign, which is not an error
    gate = ops.linear(x.detach(), w1.detach(), b1.detach(), relu=True) > 0
    hr = F.linear(refs[0], refs[1], refs[2])
    assert ((hr > 0) != gate).float().mean().item() < 1e-4
    yr = F.linear(hr * gate, refs[3], refs[4])
    yr.backward(gy.double())
    assert rel(y, yr) < TOL
    for name, a, r in zip(("x", "w1", "b1", "w2", "b2"), (x, w1, b1, w2, b2), refs):
        scale = r.grad.abs().max().item()
        assert (a.grad.double() - r.grad).abs().max().item() / scale < 2 * TOL, name


@pytest.mark.parametrize("H,W,h,w,n", [(1024, 1024, 32, 32, 5), (1024, 1024, 128, 128, 3), (512, 768, 64, 96, 4),
                                       (250, 333, 32, 32, 6), (100, 100, 7, 13, 2), (64, 64, 64, 64, 2)])
def test_gt_mask_area_bits_equal_area_interpolate_rule(H, W, h, w, n):
    """Mask-piloted rows from GT masks: bit = (F.interpolate(mask.float(), (h, w), mode='area') <= 1e-8)
    (ref decoder :986-987), including windows that do not divide the mask size and unaligned rows."""
    g = torch.Generator(device=DEV).manual_seed(H + W + h + w)
    masks = torch.zeros(n, H, W, dtype=torch.bool, device=DEV)
    for i in range(n):                                           # sparse blobs + isolated pixels + one empty mask
        if i == n - 1:
            break
        y0, x0 = int(torch.randint(0, H // 2, (1,), generator=g, device=DEV)), int(torch.randint(0, W // 2, (1,), generator=g, device=DEV))
        masks[i, y0:y0 + H // 3, x0:x0 + W // 4] = True
        pts = torch.randint(0, H * W, (5,), generator=g, device=DEV)
        masks[i].view(-1)[pts] = True
    bits = native.gt_mask_area_bits(masks, (h, w))
    ref = F.interpolate(masks.float().unsqueeze(1), size=(h, w), mode="area").flatten(1) <= 1e-8
    assert torch.equal(native.unpack_bits(bits, h * w), ref)
    assert torch.equal(bits, native.pack_bool_bits(ref))         # padding bits agree with the packer


@pytest.mark.parametrize("shape,with_r", [((16, 2150, 256), True), ((3, 120, 256), True), ((5, 7, 256), False),
                                          ((1000, 128), True), ((33, 512), True)])
def test_add_layer_norm_forward_backward(shape, with_r):
    """Fused residual + LayerNorm (ref msdeformattn.py:125-126, decoder :52,:112,:169) vs torch in fp64."""
    C = shape[-1]
    g = torch.Generator(device=DEV).manual_seed(sum(shape))
    x = (torch.randn(*shape, device=DEV, generator=g) * 2 + 0.3).requires_grad_(True)
    r = torch.randn(*shape, device=DEV, generator=g).requires_grad_(True) if with_r else None
    norm = torch.nn.LayerNorm(C).to(DEV)
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C, device=DEV, generator=g))
        norm.bias.copy_(torch.randn(C, device=DEV, generator=g))
    y = ops.add_layer_norm(x, r, norm)
    gy = torch.randn(*shape, device=DEV, generator=g)
    y.backward(gy)
    xr = x.detach().double().requires_grad_(True)
    rr = r.detach().double().requires_grad_(True) if with_r else None
    wr, br = norm.weight.detach().double().requires_grad_(True), norm.bias.detach().double().requires_grad_(True)
    yr = F.layer_norm(xr + rr if with_r else xr, (C,), wr, br, norm.eps)
    yr.backward(gy.double())
    assert rel(y, yr) < 1e-5
    assert rel(x.grad, xr.grad) < 1e-5
    if with_r:
        assert torch.equal(r.grad, x.grad)
    for a, b in ((norm.weight.grad, wr.grad), (norm.bias.grad, br.grad)):
        assert (a.double() - b).abs().max().item() / max(1.0, b.abs().max().item()) < 1e-5
    # the backward kernel also returns the column sums of dx (bias gradient of the Linear that produced r)
    with torch.no_grad():
        _, mean, rstd = native.add_layernorm_fwd(x.detach(), None if r is None else r.detach(), norm.weight, norm.bias,
                                                 norm.eps)
        dx, _, _, cs = native.add_layernorm_bwd(gy, x.detach(), None if r is None else r.detach(), norm.weight, mean,
                                                rstd, with_colsum=True)
    ref_cs = xr.grad.reshape(-1, C).sum(0)
    assert (cs.double() - ref_cs).abs().max().item() / max(1.0, ref_cs.abs().max().item()) < 1e-5


@pytest.mark.parametrize("rows,C,ld", [(344064, 256, 256), (5000, 1024, 1024), (4096, 288, 288), (3000, 100, 256),
                                       (100, 64, 64)])
def test_colsum(rows, C, ld):
    g = torch.Generator(device=DEV).manual_seed(rows + C)
    big = torch.randn(rows, ld, device=DEV, generator=g)
    x = big[:, :C]
    out = native.colsum(x)
    ref = x.double().sum(0)
    assert (out.double() - ref).abs().max().item() < 1e-4 * max(1.0, rows ** 0.5)


@pytest.mark.parametrize("B,C,H,W,G,relu", [(2, 256, 64, 64, 32, False), (2, 256, 48, 80, 32, True), (3, 128, 17, 9, 32, True),
                                            (1, 256, 256, 256, 32, True)])
def test_group_norm_channels_last(B, C, H, W, G, relu):
    """GroupNorm (+ReLU) on channels-last maps vs torch in fp64 (ref pixel_decoder/msdeformattn.py:216-219, :262-275)."""
    g = torch.Generator(device=DEV).manual_seed(B + C + H + W)
    x = (torch.randn(B, C, H, W, device=DEV, generator=g) * 1.5 + 0.7).contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    gn = torch.nn.GroupNorm(G, C).to(DEV)
    with torch.no_grad():
        gn.weight.copy_(torch.randn(C, device=DEV, generator=g))
        gn.bias.copy_(torch.randn(C, device=DEV, generator=g) * 0.3)
    y = ops.group_norm_cl(x, gn, relu=relu)
    assert y.shape == x.shape and y.permute(0, 2, 3, 1).is_contiguous()
    gy = torch.randn(B, C, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    xr = x.detach().double().requires_grad_(True)
    wr, br = gn.weight.detach().double().requires_grad_(True), gn.bias.detach().double().requires_grad_(True)
    yr = F.group_norm(xr, G, wr, br, gn.eps)
    if relu:
        assert ((yr > 0) != (y > 0)).float().mean().item() < 1e-4
        yr = yr * (y.detach() > 0)
    yr.backward(gy.double())
    assert rel(y, yr) < 1e-5
    assert rel(x.grad, xr.grad) < 2e-5
    for a, b in ((gn.weight.grad, wr.grad), (gn.bias.grad, br.grad)):
        assert (a.double() - b).abs().max().item() / max(1.0, b.abs().max().item()) < 2e-5


@pytest.mark.parametrize("B,C,H,W,G,relu", [(2, 256, 64, 64, 32, True), (2, 128, 48, 80, 32, False), (3, 64, 18, 12, 16, True),
                                            (1, 256, 256, 256, 32, True)])
def test_group_norm_nchw_to_channels_last(B, C, H, W, G, relu):
    """GroupNorm (+ReLU) reading the NCHW convolution output and writing channels-last tokens, vs torch in fp64
    (ref pixel_decoder/msdeformattn.py:262-275)."""
    g = torch.Generator(device=DEV).manual_seed(B + C + H + W)
    x = (torch.randn(B, C, H, W, device=DEV, generator=g) * 1.5 + 0.7).requires_grad_(True)
    gn = torch.nn.GroupNorm(G, C).to(DEV)
    with torch.no_grad():
        gn.weight.copy_(torch.randn(C, device=DEV, generator=g))
        gn.bias.copy_(torch.randn(C, device=DEV, generator=g) * 0.3)
    y = ops.group_norm_nchw_to_cl(x, gn, relu=relu)
    assert y is not None and y.shape == x.shape and y.permute(0, 2, 3, 1).is_contiguous()
    gy = torch.randn(B, C, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    assert x.grad.is_contiguous()
    xr = x.detach().double().requires_grad_(True)
    wr, br = gn.weight.detach().double().requires_grad_(True), gn.bias.detach().double().requires_grad_(True)
    yr = F.group_norm(xr, G, wr, br, gn.eps)
    if relu:
        assert ((yr > 0) != (y > 0)).float().mean().item() < 1e-4
        yr = yr * (y.detach() > 0)
    yr.backward(gy.double())
    assert rel(y, yr) < 1e-5
    assert rel(x.grad, xr.grad) < 2e-5
    for a, b in ((gn.weight.grad, wr.grad), (gn.bias.grad, br.grad)):
        assert (a.double() - b).abs().max().item() / max(1.0, b.abs().max().item()) < 2e-5
    # geometry outside the kernel's cover -> None (the caller falls back to the library op)
    assert ops.group_norm_nchw_to_cl(torch.randn(1, 96, 5, 5, device=DEV), torch.nn.GroupNorm(32, 96).to(DEV)) is None


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 8, 8), (1, 128, 34, 52), (2, 256, 64, 160), (1, 64, 2, 4), (1, 256, 256, 256)])
def test_upsample2x_add_to_nchw(B, C, H, W):
    """cur + bilinear x2 upsample(prev) (ref pixel_decoder/msdeformattn.py:349) as one CL,CL -> NCHW kernel, and its
    backward (transposition for cur, adjoint of the resize for prev) vs torch in fp64."""
    g = torch.Generator(device=DEV).manual_seed(B + C + H + W)
    cur = torch.randn(B, C, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    prev = torch.randn(B, C, H // 2, W // 2, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    prev.requires_grad_(True)
    y = ops.upsample2x_add_to_nchw(cur, prev)
    assert y is not None and y.is_contiguous() and y.shape == cur.shape
    ref32 = cur.detach() + F.interpolate(prev.detach(), size=(H, W), mode="bilinear", align_corners=False)
    assert (y.detach() - ref32).abs().max().item() < 2e-6          # same association as ATen: last-ulp agreement
    gy = torch.randn(B, C, H, W, device=DEV, generator=g)
    y.backward(gy)
    cr, pr = cur.detach().double().requires_grad_(True), prev.detach().double().requires_grad_(True)
    yr = cr + F.interpolate(pr, size=(H, W), mode="bilinear", align_corners=False)
    yr.backward(gy.double())
    assert rel(y, yr) < 1e-6
    assert rel(cur.grad, cr.grad) < 1e-6
    assert rel(prev.grad, pr.grad) < 1e-5
    assert ops.upsample2x_add_to_nchw(cur, torch.randn(B, C, H // 2 + 1, W // 2, device=DEV)) is None


def test_mask_feature_gradient_fanout_accumulates_in_the_epilogue():
    """ops.grad_fanout: the gradients the prediction heads send to mask_features are added inside the GEMM epilogue;
    result equals the plain autograd sum."""
    g = torch.Generator(device=DEV).manual_seed(5)
    B, Q, C, H, W = 2, 24, 64, 16, 24
    mf = torch.randn(B, C, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    embeds = [torch.randn(B, Q, C, device=DEV, generator=g).requires_grad_(True) for _ in range(4)]
    gouts = [torch.randn(B, Q, H, W, device=DEV, generator=g) for _ in range(4)]
    aliases, shared = ops.grad_fanout(mf, 4)
    loss = sum((ops.mask_logits(e, a, shared) * go).sum() for e, a, go in zip(embeds[:3], aliases, gouts))
    loss = loss + (aliases[3] * 0.5).sum()                         # a consumer outside the protocol
    loss.backward()
    got, got_e = mf.grad.clone(), [e.grad.clone() for e in embeds[:3]]
    mf.grad = None
    for e in embeds:
        e.grad = None
    loss = sum((ops.mask_logits(e, mf) * go).sum() for e, go in zip(embeds[:3], gouts)) + (mf * 0.5).sum()
    loss.backward()
    assert rel(got, mf.grad) < TOL
    for a, e in zip(got_e, embeds[:3]):
        assert rel(a, e.grad) < TOL
    assert shared.buf is None


@pytest.mark.parametrize("B,Cin,Cout,H,W,bias", [(2, 256, 256, 32, 32, False), (2, 96, 256, 18, 20, True),
                                                 (1, 2048, 256, 16, 8, True), (3, 512, 128, 12, 12, False)])
def test_conv1x1_nchw_to_channels_last(B, Cin, Cout, H, W, bias):
    """1x1 convolution of an NCHW backbone map into channels-last tokens without a layout copy (the map is the
    MN-major operand of the TN GEMM; ref pixel_decoder/msdeformattn.py:216-219, :262), forward and all gradients."""
    g = torch.Generator(device=DEV).manual_seed(B + Cin + H)
    x = torch.randn(B, Cin, H, W, device=DEV, generator=g).requires_grad_(True)
    w = (torch.randn(Cout, Cin, 1, 1, device=DEV, generator=g) / Cin ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, device=DEV, generator=g).requires_grad_(True) if bias else None
    y = ops.conv1x1_nchw_to_cl(x, w, b)
    assert y is not None and y.shape == (B, Cout, H, W) and y.permute(0, 2, 3, 1).is_contiguous()
    gy = torch.randn(B, Cout, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    br = b.detach().double().requires_grad_(True) if bias else None
    yr = F.conv2d(xr, wr, br)
    yr.backward(gy.double())
    assert rel(y, yr) < TOL
    assert rel(x.grad, xr.grad) < TOL and x.grad.is_contiguous()
    assert (w.grad.double() - wr.grad).abs().max().item() / (B * H * W) ** 0.5 < TOL
    if bias:
        assert (b.grad.double() - br.grad).abs().max().item() / (B * H * W) ** 0.5 < TOL
    # channels-last or tiny inputs are not this op's business
    assert ops.conv1x1_nchw_to_cl(torch.randn(1, 64, 4, 4, device=DEV), w[:, :64]) is None


@pytest.mark.parametrize("n_dn", [0, 7])
def test_collected_mask_heads_batched_backward(n_dn):
    """ops.collect_mask_heads: the heads' dE / dF GEMMs batched into two launches (mask_features read once, its
    gradient written once) give the gradients of the head-by-head path; one head receives no gradient for its DN
    part, one none at all."""
    g = torch.Generator(device=DEV).manual_seed(11 + n_dn)
    B, Qt, C, H, W, nH = 2, 27, 64, 16, 24, 4
    mf = torch.randn(B, C, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    embeds = [torch.randn(B, Qt, C, device=DEV, generator=g).requires_grad_(True) for _ in range(nH)]
    gouts = [torch.randn(B, Qt, H, W, device=DEV, generator=g) for _ in range(nH)]

    def loss_of(parts_dn, parts_m):
        total = 0.0
        for h in range(nH):
            if h == 2:
                continue                                            # head 2: unused by the loss
            total = total + (parts_m[h] * gouts[h][:, n_dn:]).sum()
            if n_dn and h != 1:                                     # head 1: no gradient for its DN rows
                total = total + (parts_dn[h] * gouts[h][:, :n_dn]).sum()
        return total

    aliases, shared = ops.grad_fanout(mf, nH)
    masks = [ops.mask_logits(e, a, shared) for e, a in zip(embeds, aliases)]
    dn_parts, m_parts = ops.collect_mask_heads(masks, n_dn, shared)
    assert (dn_parts is None) == (n_dn == 0) and m_parts[0].shape == (B, Qt - n_dn, H, W)
    loss_of(dn_parts, m_parts).backward()
    got_mf, got_e = mf.grad.clone(), [None if e.grad is None else e.grad.clone() for e in embeds]
    assert shared.buf is None and shared.dE is None
    mf.grad = None
    for e in embeds:
        e.grad = None
    masks = [ops.mask_logits(e, mf) for e in embeds]
    loss_of([m[:, :n_dn] for m in masks], [m[:, n_dn:] for m in masks]).backward()
    assert rel(got_mf, mf.grad) < TOL
    for h, (a, e) in enumerate(zip(got_e, embeds)):
        if e.grad is None:
            assert a is None or a.abs().max().item() == 0.0
        else:
            assert rel(a, e.grad) < TOL, h


class _Conv(torch.nn.Conv2d):
    pass


@pytest.mark.parametrize("B,Cin,Cout,H,W,bias", [(2, 64, 64, 8, 128, True), (1, 256, 256, 16, 256, False),
                                                 (2, 128, 64, 5, 128, False), (1, 64, 192, 33, 384, True),
                                                 (2, 64, 64, 7, 100, True), (1, 256, 256, 19, 200, False),
                                                 (1, 64, 128, 3, 9, False), (2, 128, 128, 40, 304, False)])
def test_conv3x3_channels_last_tensor_core(B, Cin, Cout, H, W, bias):
    """3x3 convolution as one GEMM with K = 9*Cin (taps = shifted TMA boxes, zero-filled outside the map; any map
    width: the last 128-pixel segment of a row is clipped by the 4-D output map), forward,
    input gradient (same kernel, flipped/transposed weights) and weight gradient (TN GEMM over all pixels) vs
    F.conv2d in fp64 (ref pixel_decoder/msdeformattn.py:268-275)."""
    g = torch.Generator(device=DEV).manual_seed(B + Cin + H + W)
    conv = _Conv(Cin, Cout, 3, padding=1, bias=bias).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / (9 * Cin) ** 0.5)
        if bias:
            conv.bias.copy_(torch.randn(Cout, device=DEV, generator=g))
    x = torch.randn(B, Cin, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert ops.conv3x3_cl_supported(x, conv)
    y = ops.conv3x3_cl(x, conv)
    assert y is not None and y.shape == (B, Cout, H, W) and y.permute(0, 2, 3, 1).is_contiguous()
    gy = torch.randn(B, Cout, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    xr = x.detach().double().requires_grad_(True)
    wr = conv.weight.detach().double().requires_grad_(True)
    br = conv.bias.detach().double().requires_grad_(True) if bias else None
    yr = F.conv2d(xr, wr, br, padding=1)
    yr.backward(gy.double())
    assert rel(y, yr) < TOL, rel(y, yr)
    assert rel(x.grad, xr.grad) < TOL, rel(x.grad, xr.grad)
    T = B * H * W
    assert (conv.weight.grad.double() - wr.grad).abs().max().item() / T ** 0.5 < TOL
    if bias:
        assert (conv.bias.grad.double() - br.grad).abs().max().item() / T ** 0.5 < TOL
    # layers outside the kernel's cover (channel counts that are not multiples of 64) -> None, the caller falls back
    assert ops.conv3x3_cl(torch.randn(1, 48, 8, 100, device=DEV).contiguous(memory_format=torch.channels_last),
                          _Conv(48, 64, 3, padding=1).to(DEV)) is None


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 8, 8), (1, 128, 34, 52), (2, 256, 64, 160), (1, 4, 2, 2)])
def test_upsample2x_add_channels_last(B, C, H, W):
    """The FPN merge with both sides channels-last (in front of the tensor-core 3x3 convolution) vs torch."""
    g = torch.Generator(device=DEV).manual_seed(B + C + H + W)
    cur = torch.randn(B, C, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    prev = torch.randn(B, C, H // 2, W // 2, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    prev.requires_grad_(True)
    y = ops.upsample2x_add_cl(cur, prev)
    assert y is not None and y.permute(0, 2, 3, 1).is_contiguous()
    ref32 = cur.detach() + F.interpolate(prev.detach(), size=(H, W), mode="bilinear", align_corners=False)
    assert (y.detach() - ref32).abs().max().item() < 2e-6
    gy = torch.randn(B, C, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    cr, pr = cur.detach().double().requires_grad_(True), prev.detach().double().requires_grad_(True)
    (cr + F.interpolate(pr, size=(H, W), mode="bilinear", align_corners=False)).backward(gy.double())
    assert rel(cur.grad, cr.grad) < 1e-6
    assert rel(prev.grad, pr.grad) < 1e-5


@pytest.mark.parametrize("B,Qt,with_mask", [(2, 120, True), (1, 100, False), (3, 7, True), (2, 320, True), (1, 33, False)])
def test_self_attention_core_forward_backward(B, Qt, with_mask):
    """csrc/self_attn.cu through ops.self_attention against nn.MultiheadAttention's arithmetic in fp64 (oracle mha),
    with the block mask of the mask-piloted groups (ref decoder :1051-1059)."""
    E, nhead = 256, 8
    g = torch.Generator(device=DEV).manual_seed(B * 31 + Qt)
    rn = lambda *s, sc=1.0: torch.randn(*s, device=DEV, generator=g) * sc
    x = rn(B, Qt, E)
    w_in, b_in, w_out, b_out = rn(3 * E, E, sc=1 / 16), rn(3 * E, sc=0.1), rn(E, E, sc=1 / 16), rn(E, sc=0.1)
    mask = None
    if with_mask:
        pad = max(1, Qt // 6)
        mask = torch.zeros(Qt, Qt, dtype=torch.bool, device=DEV)
        mask[pad:, :pad] = True                                    # matching queries do not see the dn group
        mask[: pad // 2, pad // 2:pad] = True
    leaves = [t.clone().requires_grad_(True) for t in (x, w_in, b_in, w_out, b_out)]
    before = dict(ops.ROUTES)
    y = ops.self_attention(leaves[0], leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], nhead, mask)
    assert ops.ROUTES["self_attention.core:native"] == before.get("self_attention.core:native", 0) + 1
    gy = rn(B, Qt, E)
    y.backward(gy)
    ref = [t.double().clone().requires_grad_(True) for t in (x, w_in, b_in, w_out, b_out)]
    sd = {"in_proj_weight": ref[1], "in_proj_bias": ref[2], "out_proj.weight": ref[3], "out_proj.bias": ref[4]}
    am = None if mask is None else mask[None].expand(B * nhead, -1, -1)
    xr = ref[0].transpose(0, 1)
    yr = O.mha(sd, "", xr, xr, xr, nhead, am).transpose(0, 1)
    yr.backward(gy.double())
    assert rel(y, yr) < TOL, rel(y, yr)
    for name, a, b_ in zip(("x", "w_in", "b_in", "w_out", "b_out"), leaves, ref):
        scale = b_.grad.abs().max().item()
        err = (a.grad.double() - b_.grad).abs().max().item() / max(scale, 1e-6)
        assert err < 2 * TOL, (name, err)
