"""GPU parity of the device Hungarian matcher (mp_former_b200/matcher.py, csrc/matcher.cu)
against the CPU oracle (oracle/matcher_oracle.py, oracle/lsap_oracle.py), the golden assignments of the unmodified
reference matcher (tests/golden/matcher.pt) and scipy (the reference's solver).

Tolerances: cost matrices are fp32 sums over up to 12544 points in a different association than the reference's
einsum -> 2e-4 relative + 2e-4 absolute; index pairs are compared exactly (the LSAP kernel is exact on equal costs;
on the seeded inputs the optimum is separated from the runner-up by far more than the cost tolerance)."""
import os
import sys

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_matcher import inputs  # noqa: E402
from oracle import matcher_oracle as MO  # noqa: E402

pytestmark = pytest.mark.gpu

COST_RTOL, COST_ATOL = 2e-4, 2e-4


def _dev(outputs, targets, dev="cuda"):
    o = {k: v.to(dev) for k, v in outputs.items()}
    t = [{k: v.to(dev) for k, v in tt.items()} for tt in targets]
    return o, t


def _reference_points(seed, bs, P):
    """The points the reference drew for the golden run: torch.rand(1, P, 2) per image from the seeded CPU generator."""
    torch.manual_seed(seed)
    return torch.cat([torch.rand(1, P, 2) for _ in range(bs)])


def _oracle_costs(outputs, targets, coords, w):
    Q = outputs["pred_logits"].shape[1]
    return [MO.matching_cost(outputs["pred_logits"][b], outputs["pred_masks"][b], targets[b]["labels"],
                             targets[b]["masks"], coords[b:b + 1], *w) if len(targets[b]["labels"]) else
            torch.zeros(Q, 0) for b in range(len(targets))]


def _split_costs(flat, Q, counts):
    out, o = [], 0
    for n in counts:
        out.append(flat[o:o + Q * n].view(Q, n))
        o += Q * n
    return out


@pytest.mark.parametrize("mode", ["gather", "sorted_gather", "sorted_streamed"])
def test_matcher_equals_reference_golden(mode):
    """The three ways the prediction logits are sampled -- gathered at the points as drawn, gathered at the row-major
    ordered points, streamed through shared memory (native.sample_shared_points) -- against the costs and assignments
    of the unmodified reference."""
    from mp_former_b200.matcher import HungarianMatcher
    G = torch.load(os.path.join(HERE, "golden", "matcher.pt"), weights_only=False)
    outputs, targets = inputs()
    o, t = _dev(outputs, targets)
    Q = outputs["pred_logits"].shape[1]
    for case in G["cases"]:
        wc, wm, wd = case["weights"]
        coords = _reference_points(case["seed"], len(targets), case["num_points"])
        m = HungarianMatcher(cost_class=wc, cost_mask=wm, cost_dice=wd, num_points=case["num_points"],
                             sort_points=mode != "gather")
        m.stream_samples = mode == "sorted_streamed"
        qi, ti, counts, cost, status = m.match_device(o, t, point_coords=coords.cuda())
        assert int(status.item()) == 0
        ref_costs = _oracle_costs(outputs, targets, coords, (wc, wm, wd))
        for got, ref in zip(_split_costs(cost.cpu(), Q, counts), ref_costs):
            assert torch.allclose(got, ref, rtol=COST_RTOL, atol=COST_ATOL), (got - ref).abs().max()
        sizes = [min(Q, n) for n in counts]
        for (i, j), (gi, gj) in zip(zip(torch.split(qi.cpu(), sizes), torch.split(ti.cpu(), sizes)), case["indices"]):
            assert torch.equal(i, gi) and torch.equal(j, gj)


def test_matcher_forward_contract_and_rng_consumption():
    """forward() returns CPU int64 pairs like the reference and consumes one rand(1, P, 2) per image from the device
    generator; the device_indices variant returns the same pairs without the host copy."""
    from mp_former_b200.matcher import HungarianMatcher
    outputs, targets = inputs()
    o, t = _dev(outputs, targets)
    m = HungarianMatcher(2.0, 5.0, 5.0, num_points=500)
    torch.manual_seed(123)
    res = m(o, t)
    torch.manual_seed(123)
    coords = torch.cat([torch.rand(1, 500, 2, device="cuda") for _ in range(len(t))])
    after = torch.rand(1, device="cuda")
    torch.manual_seed(123)
    m(o, t)
    assert torch.equal(after, torch.rand(1, device="cuda"))          # same generator position afterwards
    ref = [linear_sum_assignment(c.numpy()) for c in _oracle_costs(outputs, targets, coords.cpu(), (2.0, 5.0, 5.0))]
    assert len(res) == len(t)
    for (i, j), (ri, rj), tt in zip(res, ref, targets):
        assert i.device.type == "cpu" and i.dtype == torch.int64 and j.dtype == torch.int64
        assert len(i) == len(j) == min(outputs["pred_logits"].shape[1], len(tt["labels"]))
        assert np.array_equal(i.numpy(), ri) and np.array_equal(j.numpy(), rj)
    md = HungarianMatcher(2.0, 5.0, 5.0, num_points=500, device_indices=True)
    torch.manual_seed(123)
    resd = md(o, t)
    for (i, j), (di, dj) in zip(res, resd):
        assert di.is_cuda and torch.equal(di.cpu(), i) and torch.equal(dj.cpu(), j)
    assert "cost_class: 2.0" in repr(m)


@pytest.mark.parametrize("H,W,Q,P,sort", [(256, 256, 5, 12544, True), (100, 36, 3, 777, True), (100, 32, 4, 300, False),
                                          (300, 128, 2, 5000, True)])
def test_sample_shared_points_equals_gathered_samples(H, W, Q, P, sort):
    """Streaming sampler against the gather kernel (same bilinear formula: 1e-6), for points in row-major order with
    the band table (the fast path), for maps whose height is not a multiple of the band, and for UNSORTED points with
    a band table that does not describe them (every footprint outside its band falls back to global memory)."""
    from mp_former_b200 import native
    from mp_former_b200.matcher import HungarianMatcher
    g = torch.Generator(device="cuda").manual_seed(H * 7 + P)
    B = 3
    full = torch.randn(B, Q + 2, H, W, device="cuda", generator=g)
    maps = full[:, 2:]                                             # a query slice of a larger tensor
    pts = torch.rand(B, P, 2, device="cuda", generator=g)
    pts[0, :8] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [0.0, 1.0], [1.0, 0.0], [0.5, 0.0], [0.5, 1.0], [0.0, 0.5],
                               [1.0, 0.5]], device="cuda")       # footprints hanging over every edge
    rows, n_bands = native.shared_point_bands(H, W)
    rows = min(rows, 37) if H > 37 else rows                       # several bands also on small maps
    n_bands = -(-H // rows)
    if sort:
        pts, band_lo = HungarianMatcher.row_major_order(pts, H, W, band_rows=rows)
        assert band_lo.shape == (B, n_bands + 1) and int(band_lo[0, 0]) == 0 and int(band_lo[0, -1]) == P
    else:
        band_lo = torch.tensor([[0] + [P] * n_bands] * B, dtype=torch.int32, device="cuda")   # "everything in band 0"
    got = native.sample_shared_points(maps, pts, band_lo, rows)
    ptrs = (maps.data_ptr() + 4 * (torch.arange(B, device="cuda").view(B, 1) * maps.stride(0) +
                                   torch.arange(Q, device="cuda").view(1, Q) * maps.stride(1))).reshape(-1)
    ref = native.point_sample_rows(ptrs, True, (H, W), pts.repeat_interleave(Q, 0)).view(B, Q, P)
    assert torch.allclose(got, ref, rtol=0, atol=1e-6), (got - ref).abs().max()


def test_match_device_heads_equals_head_by_head():
    """Several prediction heads solved by one launch of the assignment kernel: the same pairs as one call per head
    with the same points."""
    from mp_former_b200.matcher import HungarianMatcher
    outputs, targets = inputs()
    o, t = _dev(outputs, targets)
    g = torch.Generator(device="cuda").manual_seed(17)
    heads = [o] + [{"pred_logits": torch.randn(o["pred_logits"].shape, device="cuda", generator=g),
                    "pred_masks": torch.randn(o["pred_masks"].shape, device="cuda", generator=g) * 3} for _ in range(3)]
    m = HungarianMatcher(2.0, 5.0, 5.0, num_points=300)
    pts = [m.draw_points(len(t), "cuda") for _ in heads]
    q, tt, status = m.match_device_heads(heads, t, pts)
    assert int(status.item()) == 0
    per_q, per_t = [], []
    for h, p in zip(heads, pts):
        qi, ti, _, _, st = m.match_device(h, t, point_coords=p)
        assert int(st.item()) == 0
        per_q.append(qi), per_t.append(ti)
    assert torch.equal(q, torch.cat(per_q)) and torch.equal(tt, torch.cat(per_t))
    heads[2]["pred_logits"][1, 0, 0] = float("nan")              # a failed problem is reported with its (head, image)
    _, _, status = m.match_device_heads(heads, t, pts)
    assert int(status.item()) == 1 + 2 * len(t) + 1


def _lsap_cases():
    rng = np.random.default_rng(11)
    out = []
    for (r, c) in ((1, 1), (1, 7), (7, 1), (5, 5), (20, 6), (6, 20), (100, 13), (13, 100), (40, 40), (100, 37),
                   (200, 150), (130, 300)):
        out.append(rng.standard_normal((r, c)).astype(np.float32))
        out.append(rng.integers(0, 3, (r, c)).astype(np.float32))
        out.append(rng.integers(0, 2, (r, c)).astype(np.float32))
    out.append(np.ones((6, 6), np.float32))
    out.append(np.zeros((3, 9), np.float32))
    out.append(np.zeros((9, 3), np.float32))
    return out


def test_lsap_kernel_equals_scipy_including_ties():
    from mp_former_b200 import native
    for C in _lsap_cases():
        Q, n = C.shape
        cost = torch.from_numpy(C).cuda().reshape(-1)
        offs = torch.tensor([0, n], dtype=torch.int32, device="cuda")
        qi, ti, status = native.lsap(cost, offs, [n], Q)
        ri, ci = linear_sum_assignment(C)
        assert int(status.item()) == 0
        assert np.array_equal(qi.cpu().numpy(), ri) and np.array_equal(ti.cpu().numpy(), ci), C.shape


def test_lsap_kernel_batched_ragged_and_invalid():
    from mp_former_b200 import native
    rng = np.random.default_rng(5)
    Q, counts = 50, [7, 0, 50, 64, 1]
    mats = [rng.standard_normal((Q, n)).astype(np.float32) for n in counts]
    flat = torch.from_numpy(np.concatenate([m.reshape(-1) for m in mats])).cuda()
    offs = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32, device="cuda")
    qi, ti, status = native.lsap(flat, offs, counts, Q)
    assert int(status.item()) == 0
    sizes = [min(Q, n) for n in counts]
    for i, j, M in zip(torch.split(qi.cpu(), sizes), torch.split(ti.cpu(), sizes), mats):
        ri, ci = linear_sum_assignment(M)
        assert np.array_equal(i.numpy(), ri) and np.array_equal(j.numpy(), ci)
    bad = mats[0].copy()
    bad[3, 2] = np.nan
    bad[:, 4] = np.nan          # a whole column invalid: no feasible complete assignment of the 7 targets
    qi, ti, status = native.lsap(torch.from_numpy(bad.reshape(-1)).cuda(),
                                 torch.tensor([0, 7], dtype=torch.int32, device="cuda"), [7], Q)
    # failure is the status word; the pairs are in-range placeholders so that a consumer cannot leave its buffers
    assert int(status.item()) == 1
    assert torch.equal(qi.cpu(), torch.arange(7)) and torch.equal(ti.cpu(), torch.arange(7))
    # scipy rejects ANY NaN / -inf entry, also one that a complete assignment could avoid
    for val in (np.nan, -np.inf):
        bad = mats[0].copy()
        bad[3, 2] = val
        with pytest.raises(ValueError):
            linear_sum_assignment(bad)
        _, _, status = native.lsap(torch.from_numpy(bad.reshape(-1)).cuda(),
                                   torch.tensor([0, 7], dtype=torch.int32, device="cuda"), [7], Q)
        assert int(status.item()) == 1
    ok = mats[0].copy()
    ok[3, 2] = np.inf           # +inf is a legal "forbidden pair" for scipy
    qi, ti, status = native.lsap(torch.from_numpy(ok.reshape(-1)).cuda(),
                                 torch.tensor([0, 7], dtype=torch.int32, device="cuda"), [7], Q)
    ri, ci = linear_sum_assignment(ok)
    assert int(status.item()) == 0 and np.array_equal(qi.cpu().numpy(), ri) and np.array_equal(ti.cpu().numpy(), ci)


@pytest.mark.parametrize("tgt_float", [False, True])
def test_match_cost_two_query_tiles_two_target_tiles_strided(tgt_float):
    """Q = 100 (tiles of 64 + 36), 40 and 33 targets (two target tiles), 12544 points (98 chunks), prediction maps
    that are a query slice of a larger tensor (the decoder's [B, pad + Q, H, W] with the DN part in front)."""
    from mp_former_b200.matcher import HungarianMatcher
    g = torch.Generator().manual_seed(4)
    B, Q, K, H, W, Hg, Wg, P = 2, 100, 80, 64, 64, 256, 256, 12544
    full = torch.randn(B, Q + 9, H, W, generator=g) * 4
    logits = torch.randn(B, Q, K + 1, generator=g)
    targets = []
    for n in (40, 33):
        m = torch.rand(n, Hg // 8, Wg // 8, generator=g) > 0.6
        m = m.repeat_interleave(8, 1).repeat_interleave(8, 2)
        targets.append({"labels": torch.randint(0, K, (n,), generator=g), "masks": m.float() if tgt_float else m})
    coords = torch.rand(B, P, 2, generator=g)
    coords[0, 0] = torch.tensor([0.0, 0.0])                       # map corners: zero-padding rule
    coords[0, 1] = torch.tensor([1.0, 1.0])
    outputs = {"pred_logits": logits, "pred_masks": full[:, 9:]}
    o = {"pred_logits": logits.cuda(), "pred_masks": full.cuda()[:, 9:]}
    t = [{k: v.cuda() for k, v in tt.items()} for tt in targets]
    m = HungarianMatcher(2.0, 5.0, 5.0, num_points=P)
    qi, ti, counts, cost, status = m.match_device(o, t, point_coords=coords.cuda())
    cost2 = m.match_device(o, t, point_coords=coords.cuda())[3]
    assert torch.equal(cost, cost2)                                # no atomics: bit-reproducible
    ref = _oracle_costs(outputs, targets, coords, (2.0, 5.0, 5.0))
    sizes = [min(Q, n) for n in counts]
    for got, r, i, j in zip(_split_costs(cost.cpu(), Q, counts), ref, torch.split(qi.cpu(), sizes),
                            torch.split(ti.cpu(), sizes)):
        assert torch.allclose(got, r, rtol=COST_RTOL, atol=COST_ATOL), (got - r).abs().max()
        ri, ci = linear_sum_assignment(got.numpy())               # exact on the kernel's own costs
        assert np.array_equal(i.numpy(), ri) and np.array_equal(j.numpy(), ci)
        oi, oj = linear_sum_assignment(r.numpy())                 # and equal to the oracle's assignment
        assert np.array_equal(i.numpy(), oi) and np.array_equal(j.numpy(), oj)


def test_matcher_without_targets_and_mixed_mask_sizes():
    from mp_former_b200.matcher import HungarianMatcher
    g = torch.Generator().manual_seed(8)
    B, Q, K = 3, 10, 5
    o = {"pred_logits": torch.randn(B, Q, K + 1, generator=g).cuda(),
         "pred_masks": torch.randn(B, Q, 16, 16, generator=g).cuda()}
    empty = [{"labels": torch.zeros(0, dtype=torch.int64).cuda(), "masks": torch.zeros(0, 64, 64, dtype=torch.bool).cuda()}
             for _ in range(B)]
    res = HungarianMatcher(1, 1, 1, num_points=100)(o, empty)
    assert len(res) == B and all(len(i) == 0 and len(j) == 0 for i, j in res)
    # per-image mask sizes (unpadded targets): matched image by image, still on the device
    sizes = [(64, 64), (32, 48), (64, 64)]
    ns = [3, 0, 12]
    tg = [{"labels": torch.randint(0, K, (n,), generator=g), "masks": torch.rand(n, h, w, generator=g) > 0.5}
          for n, (h, w) in zip(ns, sizes)]
    coords = torch.rand(B, 100, 2, generator=g)
    m = HungarianMatcher(1, 1, 1, num_points=100)
    qi, ti, counts, cost, status = m.match_device(o, [{k: v.cuda() for k, v in t.items()} for t in tg],
                                                  point_coords=coords.cuda())
    oc = {k: v.cpu() for k, v in o.items()}
    ref = _oracle_costs(oc, tg, coords, (1, 1, 1))
    sz = [min(Q, n) for n in counts]
    for got, r, i, j in zip(_split_costs(cost.cpu(), Q, counts), ref, torch.split(qi.cpu(), sz), torch.split(ti.cpu(), sz)):
        assert torch.allclose(got, r.reshape(got.shape), rtol=COST_RTOL, atol=COST_ATOL)
        ri, ci = linear_sum_assignment(r.reshape(got.shape).numpy())
        assert np.array_equal(i.numpy(), ri) and np.array_equal(j.numpy(), ci)


def test_matcher_full_size_properties():
    """BASELINE geometry (16 images, 100 queries, 256x256 logits, 1024x1024 GT masks, 12544 points): every image's
    pairs form a one-to-one assignment whose total cost equals scipy's optimum on the same matrix."""
    from mp_former_b200.matcher import HungarianMatcher
    g = torch.Generator(device="cuda").manual_seed(2)
    B, Q, K = 16, 100, 80
    o = {"pred_logits": torch.randn(B, Q, K + 1, device="cuda", generator=g),
         "pred_masks": torch.randn(B, Q, 256, 256, device="cuda", generator=g) * 3}
    t = []
    for b in range(B):
        n = 1 + (7 * b) % 20
        m = torch.rand(n, 32, 32, device="cuda", generator=g) > 0.7
        t.append({"labels": torch.randint(0, K, (n,), device="cuda", generator=g),
                  "masks": m.repeat_interleave(32, 1).repeat_interleave(32, 2)})
    m = HungarianMatcher(2.0, 5.0, 5.0, num_points=12544)
    qi, ti, counts, cost, status = m.match_device(o, t)
    assert int(status.item()) == 0
    sizes = [min(Q, n) for n in counts]
    for C, i, j in zip(_split_costs(cost.cpu(), Q, counts), torch.split(qi.cpu(), sizes), torch.split(ti.cpu(), sizes)):
        assert len(set(i.tolist())) == len(i) and sorted(j.tolist()) == list(range(C.shape[1]))
        ri, ci = linear_sum_assignment(C.numpy())
        assert np.array_equal(i.numpy(), ri) and np.array_equal(j.numpy(), ci)


def test_matcher_rejects_cpu_tensors():
    from mp_former_b200.matcher import HungarianMatcher
    outputs, targets = inputs()
    with pytest.raises(RuntimeError):
        HungarianMatcher(1, 1, 1, num_points=10)(outputs, targets)
