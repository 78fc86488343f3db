"""GPU parity of the criterion's device path (mp_former_b200/criterion.py, csrc/point_sample.cu)
against the CPU/GPU-agnostic oracle (oracle/criterion_oracle.py, pinned to the unmodified reference by
tests/golden/criterion.pt) run on the same device with the same seed (identical random-number consumption), and of the
two sampling kernels against torch's grid_sample.  Tolerances: samples 1e-5 abs (fp32, same bilinear formula); losses
1e-4 relative (sums over the sampled points in a different order); gradients 1e-4 relative + 1e-6 abs (fp32 atomics)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_criterion import CASES, CFG, inputs  # noqa: E402
from oracle import criterion_oracle as CO  # noqa: E402
from oracle import matcher_oracle as MO  # noqa: E402

pytestmark = pytest.mark.gpu


def _to(x, dev="cuda"):
    if isinstance(x, torch.Tensor):
        return x.to(dev)
    if isinstance(x, dict):
        return {k: _to(v, dev) for k, v in x.items()}
    if isinstance(x, list):
        return [_to(v, dev) for v in x]
    return x


@pytest.mark.parametrize("as_float", [False, True])
def test_point_sample_rows_equals_grid_sample(as_float):
    from mp_former_b200 import native
    g = torch.Generator().manual_seed(3)
    maps = torch.rand(5, 40, 56, generator=g) > 0.5
    maps = (maps.float() * torch.randn(5, 40, 56, generator=g)) if as_float else maps
    rows = torch.tensor([4, 0, 0, 2, 3, 1, 4])
    coords = torch.rand(len(rows), 333, 2, generator=g)
    coords[0, :4] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.5 / 56, 0.5 / 40]])
    coords[1, 0] = torch.tensor([-0.3, 1.7])                       # outside the map: zeros
    ref = MO.point_sample(maps.float()[rows][:, None], coords)[:, 0]
    dm = maps.cuda() if as_float else maps.cuda().view(torch.uint8)
    ptrs = dm.data_ptr() + rows.cuda() * (40 * 56 * dm.element_size())
    got = native.point_sample_rows(ptrs, as_float, (40, 56), coords.cuda())
    assert torch.allclose(got.cpu(), ref, atol=1e-5)
    neg = native.point_sample_rows(ptrs, as_float, (40, 56), coords.cuda(), neg_abs=True)
    assert torch.equal(neg, -got.abs())


def test_point_sample_rows_autograd_on_a_strided_slice():
    from mp_former_b200 import native
    g = torch.Generator().manual_seed(4)
    full = torch.randn(3, 9, 20, 28, generator=g)
    rows = torch.tensor([0, 5, 5, 11, 17])                         # flat (b, q) over the [3, 6] slice, one repeated
    coords = torch.rand(len(rows), 200, 2, generator=g)
    w = torch.randn(len(rows), 200, generator=g)
    a = full.clone().requires_grad_(True)
    ref = MO.point_sample(a[:, 3:].reshape(18, 1, 20, 28)[rows], coords)[:, 0]
    (ref * w).sum().backward()
    b = full.cuda().requires_grad_(True)
    got = native.PointSampleRows.apply(b[:, 3:], rows.cuda(), coords.cuda())
    (got * w.cuda()).sum().backward()
    assert torch.allclose(got.detach().cpu(), ref.detach(), atol=1e-5)
    assert torch.allclose(b.grad.cpu(), a.grad, rtol=1e-4, atol=1e-6)
    with pytest.raises(RuntimeError):
        native.PointSampleRows.apply(full, rows, coords)           # CPU tensors: no fallback


def _device_criterion(no_lb=False, device_indices=True):
    from mp_former_b200.criterion import SetCriterion
    from mp_former_b200.matcher import HungarianMatcher
    m = HungarianMatcher(CFG["cost_class"], CFG["cost_mask"], CFG["cost_dice"], CFG["num_points"],
                         device_indices=device_indices)
    return SetCriterion(CFG["num_classes"], matcher=m, weight_dict={}, eos_coef=CFG["eos_coef"],
                        losses=["labels", "masks"], num_points=CFG["num_points"],
                        oversample_ratio=CFG["oversample_ratio"],
                        importance_sample_ratio=CFG["importance_sample_ratio"], dn_no_lb=no_lb).cuda()


@pytest.mark.parametrize("joint", [True, False])
def test_criterion_equals_oracle_on_device_all_cases(joint):
    """All heads evaluated together (the default, ``SetCriterion._forward_heads``) and head by head: same losses as the
    oracle under the same seed, i.e. the same random points."""
    for name, with_dn, training, no_lb, seed in CASES:
        outputs, targets = inputs(with_dn=with_dn)
        o, t = _to(outputs), _to(targets)
        crit = _device_criterion(no_lb).train(training)
        crit.joint_heads = joint
        torch.manual_seed(seed)
        got = crit(o, t)
        assert crit.last_path == ("heads" if joint else "sequential")
        torch.manual_seed(seed)
        ref = CO.set_criterion(o, t, losses=["labels", "masks"], training=training, dn_no_lb=no_lb, **CFG)
        assert sorted(got) == sorted(ref), (name, sorted(set(got) ^ set(ref)))
        for k in ref:
            assert torch.allclose(got[k], ref[k], rtol=1e-4, atol=1e-6), (name, k, got[k], ref[k])


@pytest.mark.parametrize("joint", [True, False])
def test_criterion_gradients_equal_oracle_on_device(joint):
    outputs, targets = inputs(with_dn=True)
    o, t = _to(outputs), _to(targets)
    full = torch.cat([torch.zeros(2, 3, 24, 32, device="cuda"), o["pred_masks"]], 1)

    def run(fn):
        masks_leaf, logits_leaf = full.clone().requires_grad_(True), o["pred_logits"].clone().requires_grad_(True)
        oo = dict(o)
        oo["pred_masks"], oo["pred_logits"] = masks_leaf[:, 3:], logits_leaf
        torch.manual_seed(5)
        losses = fn(oo)
        sum(v for k, v in sorted(losses.items()) if v.requires_grad).backward()
        return masks_leaf.grad, logits_leaf.grad

    crit = _device_criterion().train(True)
    crit.joint_heads = joint
    gm_a, gl_a = run(lambda oo: crit(oo, t))
    gm_b, gl_b = run(lambda oo: CO.set_criterion(oo, t, losses=["labels", "masks"], training=True, **CFG))
    assert torch.allclose(gm_a, gm_b, rtol=1e-4, atol=1e-6) and float(gm_b.abs().sum()) > 0
    assert torch.allclose(gl_a, gl_b, rtol=1e-4, atol=1e-6)
    assert crit.last_path == ("heads" if joint else "sequential")


def test_mask_loss_rows_kernel_equals_torch_losses():
    """csrc/mask_loss.cu against the reference's formulas (criterion.py:25-68) in torch, values and gradients; fp32,
    sums over 12544 points in a different order: 1e-5 relative."""
    import torch.nn.functional as F
    from mp_former_b200 import native
    g = torch.Generator(device="cuda").manual_seed(8)
    for R, P in ((37, 12544), (1, 1), (5, 255), (3, 1000)):
        x0 = torch.randn(R, P, device="cuda", generator=g) * 4
        x0[0, 0] = 60.0                                            # saturated logits stay finite
        x0[-1, -1] = -60.0
        y = (torch.rand(R, P, device="cuda", generator=g) > 0.6).float()
        y[R // 2] = 0.0                                            # a target that misses every point
        wb, wd = torch.randn(R, device="cuda", generator=g), torch.randn(R, device="cuda", generator=g)
        xa, xb = x0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
        bce, dice = native.MaskLossRows.apply(xa, y)
        p = xb.sigmoid()
        ref_bce = F.binary_cross_entropy_with_logits(xb, y, reduction="none").mean(1)
        ref_dice = 1 - (2 * (p * y).sum(-1) + 1) / (p.sum(-1) + y.sum(-1) + 1)
        assert torch.allclose(bce, ref_bce, rtol=1e-5, atol=1e-7) and torch.allclose(dice, ref_dice, rtol=1e-5, atol=1e-7)
        ((bce * wb).sum() + (dice * wd).sum()).backward()
        ((ref_bce * wb).sum() + (ref_dice * wd).sum()).backward()
        assert torch.allclose(xa.grad, xb.grad, rtol=1e-4, atol=1e-9), (xa.grad - xb.grad).abs().max()
    with pytest.raises(RuntimeError):
        native.MaskLossRows.apply(torch.zeros(2, 3), torch.zeros(2, 3))      # CPU tensors: no fallback


def test_criterion_bench_geometry_properties():
    """4 images at the bench geometry (100 queries + 2 dn groups, 256x256 logits, 1024x1024 bool GT masks, 12544
    points, 2 auxiliary layers): finite losses, every expected key, and mask-logit gradients only on matched / dn rows."""
    from mp_former_b200.criterion import SetCriterion
    from mp_former_b200.matcher import HungarianMatcher
    g = torch.Generator(device="cuda").manual_seed(6)
    B, Q, K, counts, groups = 4, 100, 80, [3, 11, 1, 20], 2
    max_num = max(counts)

    def head(q):
        return {"pred_logits": torch.randn(B, q, K + 1, device="cuda", generator=g),
                "pred_masks": (torch.randn(B, q, 256, 256, device="cuda", generator=g) * 3).requires_grad_(True)}

    targets = []
    for n in counts:
        m = torch.rand(n, 32, 32, device="cuda", generator=g) > 0.7
        targets.append({"labels": torch.randint(0, K, (n,), device="cuda", generator=g),
                        "masks": m.repeat_interleave(32, 1).repeat_interleave(32, 2)})
    out = head(Q)
    out["aux_outputs"] = [head(Q) for _ in range(2)]
    dn = head(groups * max_num)
    dn["aux_outputs"] = [head(groups * max_num) for _ in range(2)]
    dn["dn_args"] = {"pad_size": groups * max_num, "max_num": max_num}
    out["dn_out"] = dn
    crit = SetCriterion(K, matcher=HungarianMatcher(2.0, 5.0, 5.0, 12544, device_indices=True), weight_dict={},
                        eos_coef=0.1, losses=["labels", "masks"], num_points=12544, oversample_ratio=3.0,
                        importance_sample_ratio=0.75).cuda().train(True)
    losses = crit(out, targets)
    assert crit.last_path == "heads"
    assert len(losses) == 18 and all(torch.isfinite(v) for v in losses.values())
    sum(v for k, v in losses.items() if "mask" in k or "dice" in k).backward()
    rows = out["pred_masks"].grad.abs().flatten(2).sum(-1) > 0
    assert rows.sum(1).tolist() == counts
    dn_rows = dn["pred_masks"].grad.abs().flatten(2).sum(-1) > 0
    assert dn_rows.sum(1).tolist() == [groups * n for n in counts]
    for b, n in enumerate(counts):                                 # group g's query g*max_num + j <- target j
        expect = sorted(gq * max_num + j for gq in range(groups) for j in range(n))
        assert dn_rows[b].nonzero().flatten().tolist() == expect


@pytest.mark.parametrize("joint", [True, False])
@pytest.mark.parametrize("fault", ["nan_logits", "label_out_of_range"])
def test_criterion_failed_assignment_is_memory_safe_and_reported(fault, joint):
    """Diverged logits (NaN costs) or a label outside [0, num_classes): the reference gets scipy's ValueError (or a
    device assert).  Here the solve reports a status instead of a host round trip: the pairs stay in range (nothing
    is read or written outside the prediction / gradient buffers -- run under compute-sanitizer memcheck by
    scripts/gpu_sanitizer.sh), every loss of the step is NaN, and ``check_status`` raises the ValueError."""
    outputs, targets = inputs(with_dn=True)
    o, t = _to(outputs), _to(targets)
    guard = torch.zeros(2, 3, 24, 32, device="cuda")
    full = torch.cat([guard, o["pred_masks"]], 1)                  # rows in FRONT of the maps the criterion may touch
    if fault == "nan_logits":
        full[0, 3 + 2] = float("nan")
    else:
        t[0]["labels"] = t[0]["labels"].clone()
        t[0]["labels"][0] = CFG["num_classes"] + 5
    leaf = full.clone().requires_grad_(True)
    o["pred_masks"] = leaf[:, 3:]
    crit = _device_criterion().train(True)
    crit.joint_heads = joint
    losses = crit(o, t)
    assert all(bool(torch.isnan(v)) for v in losses.values())
    sum(v for v in losses.values() if v.requires_grad).backward()
    torch.cuda.synchronize()
    assert float(leaf.grad[:, :3].abs().sum()) == 0.0              # nothing landed in front of the prediction maps
    with pytest.raises(ValueError, match="invalid numeric entries"):
        crit.check_status()
    crit.check_status()                                            # reported once; a healthy step follows
    o2, t2 = _to(inputs(with_dn=True)[0]), _to(inputs(with_dn=True)[1])
    good = crit(o2, t2)
    assert all(bool(torch.isfinite(v)) for v in good.values())
    crit.check_status()


@pytest.mark.parametrize("R,n,k,w", [(7, 37632, 9408, 2), (3, 1000, 1, 2), (2, 4096, 4096, 1), (5, 777, 300, 2),
                                     (1, 49152, 12345, 1)])
def test_topk_gather_rows_selects_the_same_set_as_torch_topk(R, n, k, w):
    from mp_former_b200 import native
    g = torch.Generator(device="cuda").manual_seed(R * 1000 + k)
    scores = -torch.randn(R, n, device="cuda", generator=g).abs()          # the criterion's -|logit|
    scores[0, : n // 7] = scores[0, 0]                                     # a block of exact ties (crosses the threshold
    scores[-1, 5] = float("inf")                                           #  for some parameter sets), +inf and NaN
    scores[-1, 9 % n] = float("nan")
    payload = torch.randn(R, n, w, device="cuda", generator=g)
    payload[..., 0] = torch.arange(n, device="cuda")                      # channel 0 carries the index (exact in fp32)
    got = native.topk_gather_rows(scores, payload, k)
    assert got.shape == (R, k, w)
    ref_idx = scores.topk(k, dim=1).indices
    for r in range(R):
        inf = float("inf")
        sel_vals = torch.sort(torch.nan_to_num(scores[r][ref_idx[r]], nan=inf, posinf=inf)).values
        idx = got[r, :, 0].long()
        assert bool(((idx >= 0) & (idx < n)).all()), (r, idx[:8])
        assert torch.equal(got[r], payload[r][idx])                        # whole payload rows were gathered
        assert bool((idx[1:] > idx[:-1]).all()), (r, idx[:8])              # ascending index order, no duplicates
        mine = torch.sort(torch.nan_to_num(scores[r][idx], nan=inf, posinf=inf)).values
        assert torch.equal(mine, sel_vals), (r, mine[:4], sel_vals[:4])    # same multiset of scores as torch.topk
    again = native.topk_gather_rows(scores, payload, k)
    assert torch.equal(again, got)                                         # deterministic
