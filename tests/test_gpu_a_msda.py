"""-m gpu: MSDeformAttn kernels (through the C ABI) vs the oracle / reference goldens.

Tolerances: fp32 forward 1e-5 * max(1,|ref|) (north_star allows 1e-3); fp64 1e-10; the reference's
own self-test uses rtol 1e-2 / atol 1e-3 for fp32 (ops/test.py:60)."""
import pytest
import torch

import cases
import mp_former_b200 as M
from mp_former_b200 import MultiScaleDeformableAttention as MSDA
from oracle import torch_oracle as O
from test_oracle_vs_golden import close, load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev_shapes(shapes, tiled=True):
    st = torch.as_tensor(shapes, dtype=torch.long, device=DEV)
    if tiled:
        st._mpf_host_shapes = tuple(shapes)
    lsi = torch.cat((st.new_zeros((1,)), st.prod(1).cumsum(0)[:-1]))
    return st, lsi


def run_gpu(value, shapes, loc, aw, gy=None, tiled=True):
    st, lsi = dev_shapes(shapes, tiled)
    v, l, a = (t.to(DEV).detach().requires_grad_(True) for t in (value, loc, aw))
    y = M.MSDeformAttnFunction.apply(v, st, lsi, l, a, 128)
    if gy is None:
        return y.detach().cpu()
    y.backward(gy.to(DEV))
    return y.detach().cpu(), v.grad.cpu(), l.grad.cpu(), a.grad.cpu()


@pytest.mark.parametrize("name", list(cases.MSDA_CASES))
@pytest.mark.parametrize("tiled", [True, False])
def test_golden_cases_forward_backward(golden_dir, name, tiled):
    G = load(golden_dir, "msda_core.pt")
    for dtype, tag, tol in ((torch.float32, "f32", 1e-5), (torch.float64, "f64", 1e-10)):
        key = f"{name}_{tag}"
        if key not in G:
            continue
        value, shapes, loc, aw = cases.msda_inputs(name, dtype)
        gy = torch.randn(G[key]["out"].shape, generator=torch.Generator().manual_seed(99)).to(dtype)
        y, gv, gl, ga = run_gpu(value, shapes, loc, aw, gy, tiled)
        close(y, G[key]["out"], tol)                      # vs the reference's own output
        close(gv, G[key]["grad_value"], tol * 10)
        close(ga, G[key]["grad_aw"], tol * 10)
        # oracle for d/d(loc) (follows the CUDA kernel's border rule, see test_oracle_vs_golden)
        v2, l2, a2 = (t.clone().requires_grad_(True) for t in (value, loc, aw))
        O.msda_core(v2, shapes, l2, a2).backward(gy)
        close(gl, l2.grad, 1e-4 if dtype == torch.float32 else 1e-9)


def test_reference_testpy_float_and_double_criteria():
    """ops/test.py:34-63: allclose default for fp64, rtol 1e-2/atol 1e-3 for fp32."""
    for dtype in (torch.float64, torch.float32):
        value, shapes, loc, aw = cases.msda_inputs("testpy", dtype)
        ref = O.msda_core(value, shapes, loc, aw)
        st, lsi = dev_shapes(shapes, False)
        got = M.MSDeformAttnFunction.apply(value.to(DEV), st, lsi, loc.to(DEV), aw.to(DEV), 2).cpu()
        if dtype == torch.float64:
            assert torch.allclose(got, ref)
        else:
            assert torch.allclose(got, ref, rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("channels", [30, 32, 64, 71, 1025, 2048, 3096])
def test_reference_gradcheck_channels(channels):
    """ops/test.py:66-89 gradcheck in fp64 over the reference's full list of channel counts (:88-89; they select
    its different backward kernels -- here fp64 always runs the one generic kernel)."""
    N, M_, Lq, L, P = 1, 2, 2, 2, 2
    shapes = [(6, 4), (3, 2)]
    S = sum(h * w for h, w in shapes)
    g = torch.Generator().manual_seed(3)
    value = (torch.rand(N, S, M_, channels, generator=g) * 0.01).double().to(DEV).requires_grad_(True)
    loc = torch.rand(N, Lq, M_, L, P, 2, generator=g).double().to(DEV).requires_grad_(True)
    aw = torch.rand(N, Lq, M_, L, P, generator=g) + 1e-5
    aw = (aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().to(DEV).requires_grad_(True)
    st, lsi = dev_shapes(shapes, False)
    assert torch.autograd.gradcheck(M.MSDeformAttnFunction.apply, (value, st, lsi, loc, aw, 2))


@pytest.mark.parametrize("geom", [
    dict(N=2, L=3, shapes=[(8, 8), (16, 16), (32, 32)], Lq=None),          # encoder-like, tiled
    dict(N=1, L=4, shapes=[(32, 32), (16, 16), (8, 8), (4, 4)], Lq=300),     # config-1 style, Lq != S
    dict(N=3, L=1, shapes=[(5, 37)], Lq=None),                             # single ragged level
    dict(N=1, L=2, shapes=[(1, 1), (2, 3)], Lq=1),                          # degenerate sizes
])
def test_vec_path_random_vs_oracle(geom):
    g = torch.Generator().manual_seed(7)
    M_, D, P = 8, 32, 4
    S = sum(h * w for h, w in geom["shapes"])
    Lq = geom["Lq"] or S
    value = torch.randn(geom["N"], S, M_, D, generator=g)
    loc = torch.rand(geom["N"], Lq, M_, geom["L"], P, 2, generator=g) * 1.4 - 0.2
    aw = torch.softmax(torch.randn(geom["N"], Lq, M_, geom["L"] * P, generator=g), -1).view(
        geom["N"], Lq, M_, geom["L"], P)
    gy = torch.randn(geom["N"], Lq, M_ * D, generator=g)
    v2, l2, a2 = (t.clone().requires_grad_(True) for t in (value, loc, aw))
    ref = O.msda_core(v2, geom["shapes"], l2, a2)
    ref.backward(gy)
    y, gv, gl, ga = run_gpu(value, geom["shapes"], loc, aw, gy)
    close(y, ref.detach(), 1e-5)
    close(gv, v2.grad, 1e-4)
    close(gl, l2.grad, 2e-4)
    close(ga, a2.grad, 1e-4)


def test_full_size_properties():
    """BASELINE config 2 geometry (B=2 here, S=21504, L=3): size-independent properties --
    linearity in value, tiled == linear ordering bit-for-bit, constant field reproduces the
    in-bounds weight mass, gradients of a constant field wrt value sum to the same mass."""
    shapes = [(32, 32), (64, 64), (128, 128)]
    S = sum(h * w for h, w in shapes)
    B, M_, D, L, P = 2, 8, 32, 3, 4
    g = torch.Generator(device=DEV).manual_seed(5)
    value = torch.randn(B, S, M_, D, device=DEV, generator=g)
    value2 = torch.randn(B, S, M_, D, device=DEV, generator=g)
    loc = torch.rand(B, S, M_, L, P, 2, device=DEV, generator=g) * 1.1 - 0.05
    aw = torch.softmax(torch.randn(B, S, M_, L * P, device=DEV, generator=g), -1).view(B, S, M_, L, P)
    st_t, lsi = dev_shapes(shapes, True)
    st_l, _ = dev_shapes(shapes, False)
    f = lambda v, st: MSDA.ms_deform_attn_forward(v, st, lsi, loc, aw, 128)
    y1, y2 = f(value, st_t), f(value2, st_t)
    assert torch.equal(y1, f(value, st_l))                                  # order independence
    y12 = f(value * 0.5 + value2 * 2.0, st_t)
    assert (y12 - (0.5 * y1 + 2.0 * y2)).abs().max().item() < 1e-4           # linearity
    ones = torch.ones_like(value)
    yc = f(ones, st_t).view(B, S, M_, D)
    assert (yc - yc[..., :1]).abs().max().item() == 0.0                     # channel independence
    assert yc.max().item() <= 1.0 + 1e-5 and yc.min().item() >= -1e-6       # weight mass in [0,1]
    gout = torch.ones(B, S, M_ * D, device=DEV)
    gv, gl, ga = MSDA.ms_deform_attn_backward(ones, st_t, lsi, loc, aw, gout, 128)
    assert abs(gv.sum().item() - yc.sum().item()) / yc.sum().item() < 1e-4  # adjoint identity
    gv2, gl2, ga2 = MSDA.ms_deform_attn_backward(ones, st_l, lsi, loc, aw, gout, 128)
    assert torch.equal(gl, gl2) and torch.equal(ga, ga2)
    close(gv.cpu(), gv2.cpu(), 1e-4)                                        # atomics: order differs


def test_error_behaviour_matches_reference():
    value, shapes, loc, aw = cases.msda_inputs("model_small")
    st, lsi = dev_shapes(shapes)
    v, l, a = value.to(DEV), loc.to(DEV), aw.to(DEV)
    with pytest.raises(RuntimeError, match="contiguous"):
        MSDA.ms_deform_attn_forward(torch.cat([v, v], -1)[..., ::2], st, lsi, l, a, 128)
    with pytest.raises(RuntimeError, match="must divide"):
        MSDA.ms_deform_attn_forward(torch.cat([v, v, v]), st, lsi, torch.cat([l, l, l]),
                                    torch.cat([a, a, a]), 2)
    with pytest.raises(RuntimeError):
        MSDA.ms_deform_attn_forward(v.half(), st, lsi, l.half(), a.half(), 128)


@pytest.mark.parametrize("geom", [
    dict(N=2, L=3, shapes=[(8, 8), (16, 16), (32, 32)], Lq=None, shared_ref=True),
    dict(N=2, L=4, shapes=[(16, 16), (8, 8), (4, 4), (2, 2)], Lq=77, shared_ref=False),
    dict(N=1, L=1, shapes=[(5, 36)], Lq=None, shared_ref=True),
])
def test_encoder_fused_softmax_and_locations(geom):
    """MSDeformAttnEncFunction (softmax + loc = ref + off/(W,H) inside the kernels) vs the unfused arithmetic
    of ref ops/modules/ms_deform_attn.py:102-118 evaluated with the oracle in fp64."""
    g = torch.Generator().manual_seed(17)
    M_, D, P, L = 8, 32, 4, geom["L"]
    shapes = geom["shapes"]
    S = sum(h * w for h, w in shapes)
    Lq = geom["Lq"] or S
    N = geom["N"]
    value = torch.randn(N, S, M_, D, generator=g)
    ow = torch.cat([torch.randn(N, Lq, M_ * L * P * 2, generator=g) * 2.0,
                    torch.randn(N, Lq, M_ * L * P, generator=g)], -1)
    ref = torch.rand(1 if geom["shared_ref"] else N, Lq, L, 2, generator=g) * 1.2 - 0.1
    gy = torch.randn(N, Lq, M_ * D, generator=g)

    def reference(value, ow, ref):
        n_off = M_ * L * P * 2
        off = ow[..., :n_off].reshape(N, Lq, M_, L, P, 2)
        aw = torch.softmax(ow[..., n_off:].reshape(N, Lq, M_, L * P), -1).view(N, Lq, M_, L, P)
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=ow.dtype)
        loc = ref.expand(N, -1, -1, -1)[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        return O.msda_core(value, shapes, loc, aw)

    v64, o64 = value.double().requires_grad_(True), ow.double().requires_grad_(True)
    yr = reference(v64, o64, ref.double())
    yr.backward(gy.double())
    st, lsi = dev_shapes(shapes, True)
    vd, od = value.to(DEV).requires_grad_(True), ow.to(DEV).requires_grad_(True)
    refd = ref.to(DEV).expand(N, -1, -1, -1) if geom["shared_ref"] else ref.to(DEV)
    y = M.msdeform_attn.MSDeformAttnEncFunction.apply(vd, st, lsi, od, refd, P)
    y.backward(gy.to(DEV))
    close(y.detach().cpu(), yr.detach().float(), 2e-5)
    close(vd.grad.cpu(), v64.grad.float(), 1e-4)
    scale = o64.grad.abs().max().item()
    assert (od.grad.cpu().double() - o64.grad).abs().max().item() / scale < 2e-4


@pytest.mark.parametrize("geom", [
    dict(N=2, shapes=[(32, 32), (64, 64), (128, 128)], off=3.0),      # bench geometry, offsets of a few texels
    dict(N=1, shapes=[(8, 8), (16, 16), (32, 32)], off=1.0),
    dict(N=2, shapes=[(16, 16), (32, 32), (64, 64)], off=40.0),       # huge offsets: most samples leave their region
    dict(N=1, shapes=[(20, 37), (10, 19)], off=2.0),                  # ragged tiles, two levels
    dict(N=1, shapes=[(12, 16)], off=0.3),                            # one level
])
def test_staged_tile_kernels_equal_the_l1_gather_kernels(geom):
    """csrc/msda_staged.cu (TMA-staged regions, LDS gathers, fixed-point shared-memory accumulation of grad_value) vs
    csrc/msda.cu on identical inputs: forward bit for bit (same association), grad_offsets/logits to fp32 rounding,
    grad_value within the fixed-point quantum (2^-21 of the tile's largest incoming gradient) + atomics order."""
    from mp_former_b200 import _lib
    g = torch.Generator().manual_seed(23)
    M_, D, P, L = 8, 32, 4, len(geom["shapes"])
    shapes = geom["shapes"]
    S = sum(h * w for h, w in shapes)
    N = geom["N"]
    value = torch.randn(N, S, M_, D, generator=g).to(DEV)
    # offsets in texels of each level, a directional bias per head like the module's initialisation
    ang = torch.arange(M_) * (2 * torch.pi / M_)
    bias = torch.stack([ang.cos(), ang.sin()], -1).view(1, 1, M_, 1, 1, 2) * torch.arange(1, P + 1).view(1, 1, 1, 1, P, 1)
    off = (torch.randn(N, S, M_, L, P, 2, generator=g) * 0.5 + bias) * geom["off"]
    ow = torch.cat([off.reshape(N, S, -1), torch.randn(N, S, M_ * L * P, generator=g)], -1).to(DEV)
    pts = []
    for h, w in shapes:
        ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
        pts.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    ref = torch.cat(pts)[None, :, None, :].expand(1, S, L, 2).contiguous().to(DEV)
    gy = torch.randn(N, S, M_ * D, generator=g).to(DEV)
    gy[:, : S // 3] *= 1e-3                                           # tiles with very different gradient scales
    st, lsi = dev_shapes(shapes, True)
    lib = _lib.load()
    outs = {}
    prev = lib.mpf_msda_set_staged(-1)
    try:
        for mode in (1, 0):
            lib.mpf_msda_set_staged(mode)
            y = MSDA.ms_deform_attn_enc_forward(value, st, lsi, ow, ref, P)
            gv, gow = MSDA.ms_deform_attn_enc_backward(value, st, lsi, ow, ref, gy, P)
            torch.cuda.synchronize()
            outs[mode] = (y, gv, gow)
    finally:
        lib.mpf_msda_set_staged(prev)
    (y1, gv1, go1), (y0, gv0, go0) = outs[1], outs[0]
    assert torch.equal(y1, y0)
    assert (go1 - go0).abs().max().item() <= 1e-5 * max(1.0, go0.abs().max().item())
    assert (gv1 - gv0).abs().max().item() <= 2e-5 * gy.abs().max().item()
    # and relative to each texel's own magnitude where the gradient is small (the 1e-3 part): quantum << value
    small = gv0[:, : shapes[0][0] * shapes[0][1]].abs().max().item()
    assert (gv1 - gv0)[:, : shapes[0][0] * shapes[0][1]].abs().max().item() <= 1e-3 * max(small, 1e-6) + 1e-5
