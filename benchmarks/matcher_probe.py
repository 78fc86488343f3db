"""Device Hungarian matcher at the bench geometry (16 images, 100 queries, 256x256 mask logits, 1024x1024 GT masks,
12544 points, 1..20 instances per image): CUDA-event time of one head's matching (cost kernels + LSAP kernel, no host
round trip) next to the reference's formulation run with stock PyTorch ops on the same GPU (per image: two
grid_samples, the softplus/sigmoid maps, three einsums, the cost matrix copied to the host and scipy's solve --
mask2former/modeling/matcher.py:96-157), wall-clock timed because it synchronises per image.  One JSON line."""
import json
import os
import statistics
import sys
import time

import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mp_former_b200.matcher import HungarianMatcher  # noqa: E402

DEV = "cuda:0"
B, Q, K, P = int(os.environ.get("MPF_B", "16")), 100, 80, 12544
REPS = int(os.environ.get("MPF_REPS", "10"))


def stock_match(outputs, targets, coords, wc, wm, wd):
    """The reference's per-image formulation with library ops (timing comparator only)."""
    res = []
    for b in range(len(targets)):
        prob = outputs["pred_logits"][b].softmax(-1)
        c_class = -prob[:, targets[b]["labels"]]
        om = outputs["pred_masks"][b][:, None]
        tm = targets[b]["masks"].to(om)[:, None]
        grid = 2.0 * coords[b:b + 1].unsqueeze(2) - 1.0
        t = F.grid_sample(tm, grid.repeat(tm.shape[0], 1, 1, 1), align_corners=False).squeeze(3).squeeze(1)
        x = F.grid_sample(om, grid.repeat(om.shape[0], 1, 1, 1), align_corners=False).squeeze(3).squeeze(1)
        pos = F.binary_cross_entropy_with_logits(x, torch.ones_like(x), reduction="none")
        neg = F.binary_cross_entropy_with_logits(x, torch.zeros_like(x), reduction="none")
        c_mask = (torch.einsum("nc,mc->nm", pos, t) + torch.einsum("nc,mc->nm", neg, 1 - t)) / x.shape[1]
        s = x.sigmoid()
        c_dice = 1 - (2 * torch.einsum("nc,mc->nm", s, t) + 1) / (s.sum(-1)[:, None] + t.sum(-1)[None, :] + 1)
        C = (wm * c_mask + wc * c_class + wd * c_dice).cpu()
        res.append(linear_sum_assignment(C))
    return res


def main():
    g = torch.Generator(device=DEV).manual_seed(0)
    outputs = {"pred_logits": torch.randn(B, Q, K + 1, device=DEV, generator=g),
               "pred_masks": torch.randn(B, Q, 256, 256, device=DEV, generator=g) * 3}
    targets = []
    for b in range(B):
        n = 1 + (7 * b) % 20
        m = torch.rand(n, 32, 32, device=DEV, generator=g) > 0.7
        targets.append({"labels": torch.randint(0, K, (n,), device=DEV, generator=g),
                        "masks": m.repeat_interleave(32, 1).repeat_interleave(32, 2)})
    coords = torch.rand(B, P, 2, device=DEV, generator=g)
    sort_points = os.environ.get("MPF_SORT_POINTS", "0") == "1"
    m = HungarianMatcher(2.0, 5.0, 5.0, num_points=P, device_indices=True, sort_points=sort_points)
    m.stream_samples = os.environ.get("MPF_STREAM", "1") == "1"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(3):
        m.match_device(outputs, targets, point_coords=coords)
    ts = []
    for _ in range(REPS):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        qi, ti, counts, cost, status = m.match_device(outputs, targets, point_coords=coords)
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    ours_ms = statistics.median(ts)
    if os.environ.get("MPF_PROFILE"):
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m.match_device(outputs, targets, point_coords=coords)
            torch.cuda.synchronize()
        for e in sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total):
            if e.self_device_time_total > 0:
                print(f"# {e.self_device_time_total:9.1f} us  x{e.count:3d}  {e.key[:110]}")
    for _ in range(2):
        ref = stock_match(outputs, targets, coords, 2.0, 5.0, 5.0)
    ws = []
    for _ in range(max(3, REPS // 2)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref = stock_match(outputs, targets, coords, 2.0, 5.0, 5.0)
        torch.cuda.synchronize()
        ws.append((time.perf_counter() - t0) * 1e3)
    stock_ms = statistics.median(ws)
    sizes = [min(Q, n) for n in counts]
    same = all(list(r[0]) == i.tolist() and list(r[1]) == j.tolist()
               for r, i, j in zip(ref, torch.split(qi.cpu(), sizes), torch.split(ti.cpu(), sizes)))
    ntot = sum(counts)
    # gathers: 2 sectors of 32 B per (query, point) and per (target, point)
    sector_bytes = 64.0 * P * (B * Q + ntot)
    print(json.dumps({"probe": "hungarian_matcher_one_head", "B": B, "Q": Q, "points": P, "targets_total": ntot,
                      "device_ms": ours_ms, "stock_torch_scipy_ms_wall": stock_ms, "speedup": stock_ms / ours_ms,
                      "same_assignment_as_stock": bool(same), "gather_sector_GBps": sector_bytes / ours_ms / 1e6,
                      "host_syncs": {"device": 0, "stock": B}, "sort_points": sort_points, "stream_samples": bool(sort_points and m.stream_samples)}))


if __name__ == "__main__":
    main()
