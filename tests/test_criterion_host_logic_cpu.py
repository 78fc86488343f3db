"""CPU: host logic of the device criterion (mp_former_b200/criterion.py) against the golden losses of the UNMODIFIED
reference SetCriterion.  The CUDA entry points it calls (native.point_sample_rows / point_sample_rows_bwd, which
address maps through device-pointer tables, and native.topk_gather_rows) are replaced -- in this test only -- by an emulation that reads / updates
the same addresses in host memory with torch's grid_sample, so that everything else (index bookkeeping, pointer
arithmetic, random-number consumption, dn assignment, loss keys and normalisation) is exercised without a GPU.  The
kernels themselves are covered by tests/test_gpu_h_criterion.py."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_criterion import CASES, CFG, inputs  # noqa: E402
from oracle import criterion_oracle as CO  # noqa: E402
from oracle import matcher_oracle as MO  # noqa: E402


def _host_map(ptr, is_f32, H, W):
    ct = ctypes.c_float if is_f32 else ctypes.c_uint8
    return np.ctypeslib.as_array((ct * (H * W)).from_address(int(ptr))).reshape(H, W)


def _emu_sample(map_ptrs, maps_are_f32, hw, coords, neg_abs=False, out=None):
    H, W = hw
    rows = [MO.point_sample(torch.from_numpy(_host_map(p, maps_are_f32, H, W).copy()).float()[None, None],
                            coords[r:r + 1])[0, 0] for r, p in enumerate(map_ptrs.tolist())]
    res = torch.stack(rows)
    res = -res.abs() if neg_abs else res
    if out is not None:
        out.copy_(res)
        return out
    return res


class _EmuMaskLossRows:
    """native.MaskLossRows (csrc/mask_loss.cu): per-row BCE mean and dice term, as torch ops (ref criterion.py:25-68)."""

    @staticmethod
    def apply(x, y):
        import torch.nn.functional as F
        p = x.sigmoid()
        dice = 1 - (2 * (p * y).sum(-1) + 1) / (p.sum(-1) + y.sum(-1) + 1)
        return F.binary_cross_entropy_with_logits(x, y, reduction="none").mean(1), dice


def _emu_sample_bwd(grad_map_ptrs, hw, coords, grad_out):
    H, W = hw
    for r, p in enumerate(grad_map_ptrs.tolist()):
        with torch.enable_grad():          # called from a once_differentiable backward (grad mode off)
            z = torch.zeros(1, 1, H, W, requires_grad=True)
            MO.point_sample(z, coords[r:r + 1]).backward(grad_out[r][None, None])
        _host_map(p, True, H, W)[...] += z.grad[0, 0].numpy()


def _emu_topk_gather(scores, payload, k):
    """native.topk_gather_rows: payload rows of the k largest scores per row, ascending index order."""
    idx = scores.topk(k, dim=1).indices.sort(dim=1).values
    return torch.gather(payload, 1, idx.unsqueeze(-1).expand(-1, -1, payload.shape[-1]))


class _OracleMatcher(torch.nn.Module):
    """Reference-contract matcher (list of CPU index pairs) for the host-logic test."""

    def __init__(self):
        super().__init__()

    def forward(self, outputs, targets):
        return MO.hungarian_match(outputs, targets, CFG["num_points"], CFG["cost_class"], CFG["cost_mask"],
                                  CFG["cost_dice"])[0]

    def __repr__(self, _repr_indent=4):
        return "Matcher oracle"


@pytest.fixture
def host_kernels(monkeypatch):
    from mp_former_b200 import _lib, native
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name: None)
    monkeypatch.setattr(native, "point_sample_rows", _emu_sample)
    monkeypatch.setattr(native, "point_sample_rows_bwd", _emu_sample_bwd)
    monkeypatch.setattr(native, "topk_gather_rows", _emu_topk_gather)
    monkeypatch.setattr(native, "MaskLossRows", _EmuMaskLossRows)


def _criterion(no_lb=False):
    from mp_former_b200.criterion import SetCriterion
    return SetCriterion(CFG["num_classes"], matcher=_OracleMatcher(), weight_dict={}, eos_coef=CFG["eos_coef"],
                        losses=["labels", "masks"], num_points=CFG["num_points"],
                        oversample_ratio=CFG["oversample_ratio"],
                        importance_sample_ratio=CFG["importance_sample_ratio"], dn_no_lb=no_lb)


@pytest.mark.parametrize("joint", [True, False])
def test_criterion_host_logic_matches_reference_golden(host_kernels, joint):
    """Both evaluation orders of the criterion -- all heads together (``_forward_heads``) and head by head -- against
    the losses of the unmodified reference under the same seed (same random-number consumption)."""
    G = torch.load(os.path.join(HERE, "golden", "criterion.pt"), weights_only=False)
    for name, with_dn, training, no_lb, seed in CASES:
        outputs, targets = inputs(with_dn=with_dn)
        crit = _criterion(no_lb).train(training)
        crit.joint_heads = joint
        torch.manual_seed(seed)
        got = crit(outputs, targets)
        assert crit.last_path == ("heads" if joint else "sequential"), name
        ref = G[name]
        assert sorted(got) == sorted(ref), (name, sorted(set(got) ^ set(ref)))
        for k in ref:
            assert torch.allclose(got[k].float(), ref[k].float(), rtol=1e-5, atol=1e-6), (name, k, got[k], ref[k])
    assert "num_points: 112" in repr(crit)


@pytest.mark.parametrize("joint", [True, False])
def test_criterion_host_logic_gradients_match_oracle(host_kernels, joint):
    """Gradients w.r.t. mask logits (a query slice of a larger tensor, like the decoder's [B, pad + Q, H, W]) and
    class logits, through the pointer-table backward, against autograd through the oracle."""
    outputs, targets = inputs(with_dn=True)
    full = torch.cat([torch.zeros(2, 3, 24, 32), outputs["pred_masks"]], 1)

    def run(fn, masks_leaf, logits_leaf):
        o = dict(outputs)
        o["pred_masks"], o["pred_logits"] = masks_leaf[:, 3:], logits_leaf
        torch.manual_seed(5)
        losses = fn(o)
        sum(v for k, v in sorted(losses.items()) if v.requires_grad).backward()
        return {k: v.detach() for k, v in losses.items()}, masks_leaf.grad, logits_leaf.grad

    crit = _criterion().train(True)
    crit.joint_heads = joint
    la, gm_a, gl_a = run(lambda o: crit(o, targets), full.clone().requires_grad_(True),
                         outputs["pred_logits"].clone().requires_grad_(True))
    lb, gm_b, gl_b = run(lambda o: CO.set_criterion(o, targets, losses=["labels", "masks"], training=True, **CFG),
                         full.clone().requires_grad_(True), outputs["pred_logits"].clone().requires_grad_(True))
    for k in lb:
        assert torch.allclose(la[k], lb[k], rtol=1e-5, atol=1e-6), k
    assert torch.allclose(gm_a, gm_b, rtol=1e-4, atol=1e-7) and float(gm_b.abs().sum()) > 0
    assert torch.allclose(gl_a, gl_b, rtol=1e-5, atol=1e-7)
    assert float(gm_a[:, :3].abs().sum()) == 0.0
    assert crit.last_path == ("heads" if joint else "sequential")
    # the trainer's weighted total from the stacked values equals the entry-by-entry sum
    torch.manual_seed(5)
    losses = crit(outputs, targets)
    wd = {k: 0.5 + 0.01 * i for i, k in enumerate(sorted(losses)) if "dice" not in k}
    want = sum(v * wd[k] for k, v in losses.items() if k in wd)
    assert torch.allclose(crit.weighted_total(losses, wd), want, rtol=1e-5)


def test_criterion_mixed_target_sizes_and_empty_images(host_kernels):
    """Unpadded targets of different sizes are zero-padded to the largest (top-left aligned) like the reference's
    nested tensor (utils/misc.py:48-73); an image without instances contributes nothing."""
    outputs, targets = inputs(with_dn=False)
    targets[0]["masks"] = targets[0]["masks"][:, :80, :100].contiguous()
    targets.append({"labels": torch.zeros(0, dtype=torch.int64), "masks": torch.zeros(0, 96, 128, dtype=torch.bool)})
    g = torch.Generator().manual_seed(1)
    o = {"pred_logits": torch.randn(3, 12, 6, generator=g), "pred_masks": torch.randn(3, 12, 24, 32, generator=g),
         "dn_out": None}
    crit = _criterion().train(False)
    torch.manual_seed(9)
    got = crit(o, targets)
    torch.manual_seed(9)
    ref = CO.set_criterion(o, targets, losses=["labels", "masks"], training=False, **CFG)
    assert sorted(got) == sorted(ref)
    for k in ref:
        assert torch.allclose(got[k], ref[k], rtol=1e-5, atol=1e-6), k
