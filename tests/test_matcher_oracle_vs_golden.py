"""CPU: the matcher restatement (oracle/matcher_oracle.py, groundwork for SURVEY.md §8f rank 1) against the golden
assignments / cost terms produced by the UNMODIFIED reference matcher (tests/golden/make_golden_matcher.py)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_matcher import inputs  # noqa: E402
from oracle import matcher_oracle as MO  # noqa: E402


def test_matcher_oracle_matches_reference_golden():
    G = torch.load(os.path.join(HERE, "golden", "matcher.pt"), weights_only=False)
    outputs, targets = inputs()
    for case in G["cases"]:
        wc, wm, wd = case["weights"]
        torch.manual_seed(case["seed"])
        idx, costs = MO.hungarian_match(outputs, targets, case["num_points"], wc, wm, wd)
        assert len(idx) == len(case["indices"])
        for (i, j), (gi, gj), C, t in zip(idx, case["indices"], costs, targets):
            assert torch.equal(i, gi) and torch.equal(j, gj)
            assert C.shape == (outputs["pred_logits"].shape[1], len(t["labels"]))
            assert len(i) == min(C.shape)
    g = torch.Generator().manual_seed(9)
    a, t = torch.randn(6, 50, generator=g) * 2, (torch.rand(4, 50, generator=g) > 0.5).float()
    assert torch.allclose(MO.batch_dice_cost(a, t), G["dice"], atol=1e-6)
    assert torch.allclose(MO.batch_sigmoid_ce_cost(a, t), G["ce"], atol=1e-6)


def test_point_sample_is_bilinear_at_unit_square_coordinates():
    """The detectron2 helper restated in the oracle: pixel centres map to (i + 0.5) / size."""
    m = torch.arange(12.0).view(1, 1, 3, 4)
    xs = (torch.arange(4) + 0.5) / 4
    ys = (torch.arange(3) + 0.5) / 3
    pts = torch.stack(torch.meshgrid(xs, ys, indexing="xy"), -1).reshape(1, -1, 2)
    got = MO.point_sample(m, pts).view(3, 4)
    assert torch.allclose(got, m[0, 0], atol=1e-6)


def test_row_major_point_order_is_a_permutation_and_leaves_costs_unchanged():
    """HungarianMatcher(sort_points=True) only reorders an image's points; the cost sums (hence the assignment) do not
    depend on the order beyond fp32 rounding."""
    from mp_former_b200.matcher import HungarianMatcher
    outputs, targets = inputs()
    g = torch.Generator().manual_seed(2)
    coords = torch.rand(3, 500, 2, generator=g)
    H, W = outputs["pred_masks"].shape[-2:]
    srt = HungarianMatcher.row_major_order(coords, H, W)
    key = (srt[..., 1] * H - 0.5).floor().clamp(0, H - 1) * W + (srt[..., 0] * W - 0.5).floor().clamp(0, W - 1)
    assert bool((key[:, 1:] >= key[:, :-1]).all())
    for b in range(3):
        assert torch.equal(srt[b][srt[b][:, 0].argsort(stable=True)], coords[b][coords[b][:, 0].argsort(stable=True)])
        a = MO.matching_cost(outputs["pred_logits"][b], outputs["pred_masks"][b], targets[b]["labels"],
                             targets[b]["masks"], coords[b:b + 1], 2.0, 5.0, 5.0)
        c = MO.matching_cost(outputs["pred_logits"][b], outputs["pred_masks"][b], targets[b]["labels"],
                             targets[b]["masks"], srt[b:b + 1], 2.0, 5.0, 5.0)
        assert torch.allclose(a, c, rtol=1e-5, atol=1e-5)


def test_row_major_order_band_table_for_the_streamed_sampler():
    """``row_major_order(..., band_rows=r)``: the same ordering plus, per image, the index of the first point of every
    band of r map rows (what native.sample_shared_points walks): bands are contiguous, cover all points, and every
    point of band k has its footprint's top row in [k r, (k + 1) r) -- points above the map join band 0."""
    from mp_former_b200 import native
    from mp_former_b200.matcher import HungarianMatcher
    g = torch.Generator().manual_seed(9)
    for H, W, rows in ((256, 256, 95), (100, 36, 37), (24, 32, 24), (7, 8, 2)):
        coords = torch.rand(3, 777, 2, generator=g)
        coords[0, :4] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [0.5, 0.0], [0.5, 1.0]])
        srt, band_lo = HungarianMatcher.row_major_order(coords, H, W, band_rows=rows)
        assert torch.equal(srt, HungarianMatcher.row_major_order(coords, H, W))
        n_bands = -(-H // rows)
        assert band_lo.dtype == torch.int32 and band_lo.shape == (3, n_bands + 1)
        assert bool((band_lo[:, 0] == 0).all()) and bool((band_lo[:, -1] == 777).all())
        assert bool((band_lo[:, 1:] >= band_lo[:, :-1]).all())
        y0 = (srt[..., 1] * H - 0.5).floor().clamp(0, H - 1)
        for b in range(3):
            for k in range(n_bands):
                lo, hi = int(band_lo[b, k]), int(band_lo[b, k + 1])
                if hi > lo:
                    assert float(y0[b, lo:hi].min()) >= k * rows and float(y0[b, lo:hi].max()) < (k + 1) * rows
    r, n = native.shared_point_bands(256, 256)
    assert (r + 1) * 256 * 4 <= native.SHARED_POINT_BAND_BYTES and r * n >= 256 and r * (n - 1) < 256
    assert native.shared_point_bands(8, 8) == (8, 1)
