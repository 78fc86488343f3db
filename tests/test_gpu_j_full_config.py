"""-m gpu: parity at the FULL size of the BASELINE configurations (VERDICT r1 "What's weak" 1-3).

* config 2 head for one image -- 6 encoder layers, 9 decoder layers, 100 queries, mask-piloted (DN) group on, 1024^2 --
  against the CPU oracle: teacher-forced (arithmetic error of all ten heads within 1e-3, every bit the product would
  have set differently within eps of the threshold) and free-running (per-layer mask-bit flip rate);
* config 1 MSDeformAttn at full size (L=4, S=Lq=21760, distributions of ref ops/test.py:36-39, seed 3), forward and
  backward against the oracle;
* the forward bit-identical to the reference's own CUDA kernel (oracle/_ref/libmsda_stock.so: the UNMODIFIED
  ms_deform_im2col_cuda.cuh compiled where it lies), backward within fp32 atomics noise.
"""
import ctypes
import os

import pytest
import torch

import mp_former_b200 as M
from mp_former_b200 import MultiScaleDeformableAttention as MSDA
from mp_former_b200 import workload
from oracle import torch_oracle as O
from test_oracle_vs_golden import close
import parity_full

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def config2_report():
    torch.manual_seed(0)
    pd, dec = workload.build_head(num_queries=100, device=DEV, seed=0)
    pd.eval(), dec.eval()
    feats = workload.synthetic_features(1, 1024, 1024, seed=5)
    targets = workload.synthetic_targets(1, 1024, 1024, seed=5)
    return parity_full.compare(pd, dec, feats, targets, DEV, num_queries=100)


def test_config2_teacher_forced_all_heads_within_1e3(config2_report):
    r = config2_report
    assert r["layers"] == 9
    assert r["forced_max_rel_logits"] <= 1e-3, r["forced_per_head"]
    assert r["forced_max_rel_masks"] <= 1e-3, r["forced_per_head"]


def test_config2_flipped_bits_sit_on_the_threshold(config2_report):
    """Teacher-forced, the product derives its bits from logits within 1e-3 of the oracle's: a bit may differ only
    where the oracle's resized logit is within eps of 0 (the threshold, ops.MASK_LOGIT_THRESHOLD = -1.8e-7)."""
    r = config2_report
    assert r["forced_flip_max_dist"] <= r["eps"], r["forced_flip_rate"]
    assert max(r["forced_flip_rate"]) < 1e-3, r["forced_flip_rate"]


def test_config2_free_running_flip_rate(config2_report):
    """Free-running: layer 0 has identical inputs on both sides, so its flips obey the same bound; deeper layers
    inherit the consequences of earlier flips, so only the rate is bounded (and reported by bench.py)."""
    r = config2_report
    assert r["free_layer0_flip_max_dist"] <= r["eps"]
    assert r["free_layer0_rel_masks"] <= 1e-3
    assert r["free_flip_rate"][0] < 1e-3, r["free_flip_rate"]
    assert max(r["free_flip_rate"]) < 2e-2, r["free_flip_rate"]
    print("config 2 free-running mask flip rate per layer:", ["%.2e" % v for v in r["free_flip_rate"]],
          "final-head share beyond 1e-3: masks %.4f logits %.4f" % (r["free_frac_above_1e-3_masks"],
                                                                    r["free_frac_above_1e-3_logits"]))


# ------------------------------------------------------------------------------------------------
# config 1: MSDeformAttn at full size
# ------------------------------------------------------------------------------------------------
CONFIG1_SHAPES = [(128, 128), (64, 64), (32, 32), (16, 16)]          # S = 21760


def config1_inputs():
    """B=1, 256-d (M=8, D=32), L=4, P=4, Lq=S=21760; distributions of ref ops/test.py:36-39 under seed 3."""
    g = torch.Generator().manual_seed(3)
    S = sum(h * w for h, w in CONFIG1_SHAPES)
    value = torch.rand(1, S, 8, 32, generator=g) * 0.01
    loc = torch.rand(1, S, 8, 4, 4, 2, generator=g)
    aw = torch.rand(1, S, 8, 4, 4, generator=g) + 1e-5
    aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value, loc, aw


def _shapes_dev(tiled):
    st = torch.as_tensor(CONFIG1_SHAPES, dtype=torch.long, device=DEV)
    if tiled:
        st._mpf_host_shapes = tuple(CONFIG1_SHAPES)
    lsi = torch.cat((st.new_zeros((1,)), st.prod(1).cumsum(0)[:-1]))
    return st, lsi


@pytest.mark.parametrize("tiled", [True, False])
def test_config1_full_size_forward_backward_vs_oracle(tiled):
    value, loc, aw = config1_inputs()
    assert value.shape[1] == 21760
    gy = torch.randn(1, 21760, 256, generator=torch.Generator().manual_seed(4))
    v2, l2, a2 = (t.clone().requires_grad_(True) for t in (value, loc, aw))
    ref = O.msda_core(v2, CONFIG1_SHAPES, l2, a2)
    ref.backward(gy)
    st, lsi = _shapes_dev(tiled)
    v, l, a = (t.to(DEV).requires_grad_(True) for t in (value, loc, aw))
    y = M.MSDeformAttnFunction.apply(v, st, lsi, l, a, 128)
    y.backward(gy.to(DEV))
    close(y.detach().cpu(), ref.detach(), 1e-5)
    # gradients: relative to the largest reference entry (grad_value sums ~50 fp32 atomics per texel)
    for got, want, tol in ((v.grad, v2.grad, 1e-4), (l.grad, l2.grad, 2e-4), (a.grad, a2.grad, 1e-4)):
        scale = max(1.0, want.abs().max().item())
        assert (got.cpu() - want).abs().max().item() / scale <= tol


# ------------------------------------------------------------------------------------------------
# the reference's own CUDA kernel as the comparator
# ------------------------------------------------------------------------------------------------
def _stock():
    path = os.path.join(ROOT, "oracle", "_ref", "libmsda_stock.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libmsda_stock.so not built (needs /root/reference at build time)")
    lib = ctypes.CDLL(path)
    p, i = ctypes.c_void_p, ctypes.c_int
    lib.ref_msda_forward_f32.argtypes = [p] * 5 + [i] * 7 + [p, p]
    lib.ref_msda_backward_f32.argtypes = [p] * 6 + [i] * 7 + [p, p, p, p]
    return lib


@pytest.mark.parametrize("geom", ["config1", "config2_b2"])
def test_forward_bit_identical_to_stock_cuda_kernel(geom):
    lib = _stock()
    if geom == "config1":
        shapes = CONFIG1_SHAPES
        value, loc, aw = (t.to(DEV) for t in config1_inputs())
    else:
        shapes = [(32, 32), (64, 64), (128, 128)]
        g = torch.Generator(device=DEV).manual_seed(9)
        S = sum(h * w for h, w in shapes)
        value = torch.randn(2, S, 8, 32, device=DEV, generator=g)
        loc = torch.rand(2, S, 8, 3, 4, 2, device=DEV, generator=g) * 1.2 - 0.1
        aw = torch.softmax(torch.randn(2, S, 8, 12, device=DEV, generator=g), -1).view(2, S, 8, 3, 4)
    B, S, Mh, D = value.shape
    L, Lq, P = len(shapes), loc.shape[1], 4
    st = torch.as_tensor(shapes, dtype=torch.long, device=DEV)
    st._mpf_host_shapes = tuple(shapes)
    lsi = torch.cat((st.new_zeros((1,)), st.prod(1).cumsum(0)[:-1]))
    stream = torch.cuda.current_stream().cuda_stream
    ours = MSDA.ms_deform_attn_forward(value, st, lsi, loc, aw, 128)
    stock = torch.empty_like(ours)
    rc = lib.ref_msda_forward_f32(value.data_ptr(), st.data_ptr(), lsi.data_ptr(), loc.data_ptr(), aw.data_ptr(),
                                    B, S, Mh, D, L, Lq, P, stock.data_ptr(), stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(ours, stock)
    gy = torch.randn_like(ours)
    gv, gl, ga = MSDA.ms_deform_attn_backward(value, st, lsi, loc, aw, gy, 128)
    sv, sl, sa = torch.zeros_like(gv), torch.zeros_like(gl), torch.zeros_like(ga)
    rc = lib.ref_msda_backward_f32(gy.data_ptr(), value.data_ptr(), st.data_ptr(), lsi.data_ptr(), loc.data_ptr(),
                                     aw.data_ptr(), B, S, Mh, D, L, Lq, P, sv.data_ptr(), sl.data_ptr(),
                                     sa.data_ptr(), stream)
    assert rc == 0
    torch.cuda.synchronize()
    for got, want in ((gv, sv), (gl, sl), (ga, sa)):
        scale = max(1.0, want.abs().max().item())
        assert (got - want).abs().max().item() / scale <= 1e-4
