"""Device criterion at the bench geometry (16 images, 100 queries + one dn group, 256x256 mask logits of 10 heads,
1024x1024 bool GT masks, 12544 points): CUDA-event time of one SetCriterion forward and of forward + backward to the
mask / class logits, launches issued by libmpformer_b200, and a breakdown matcher / classification / mask losses
(each timed alone on the last head).  One JSON line.  (The reference's formulation as a comparator for the matching
part: benchmarks/matcher_probe.py.)"""
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mp_former_b200 import _lib, workload  # noqa: E402

DEV = "cuda:0"
B, Q, K = int(os.environ.get("MPF_B", "16")), 100, 80
REPS = int(os.environ.get("MPF_REPS", "5"))


def timed(fn, reps=REPS, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    return statistics.median(ts)


def stock_loss_masks(pred_masks, targets, b, q, t, num_masks, num_points=12544, k=3.0, beta=0.75):
    """The reference's formulation of SetCriterion.loss_masks with library ops (timing comparator only; ref
    criterion.py:141-191): gather of the matched prediction maps, float zero-padded copy of every GT mask of the batch,
    grid_sample at the candidate points, top-k, two more grid_samples, BCE + dice."""
    src = pred_masks[b, q][:, None]
    nmax = max(len(x["labels"]) for x in targets)
    hg, wg = targets[0]["masks"].shape[-2:]
    padded = torch.zeros(len(targets), nmax, hg, wg, dtype=torch.bool, device=src.device)
    for i, x in enumerate(targets):
        padded[i, :len(x["labels"])] = x["masks"]
    tgt = padded.to(src)[b, t][:, None]

    def sample(m, c):
        return F.grid_sample(m, 2.0 * c.unsqueeze(2) - 1.0, align_corners=False).squeeze(3)

    with torch.no_grad():
        R, n_over, n_unc = src.shape[0], int(num_points * k), int(beta * num_points)
        cand = torch.rand(R, n_over, 2, device=src.device)
        unc = -sample(src, cand).abs()[:, 0]
        idx = unc.topk(n_unc, dim=1)[1] + n_over * torch.arange(R, device=src.device)[:, None]
        coords = torch.cat([cand.view(-1, 2)[idx.view(-1)].view(R, n_unc, 2),
                            torch.rand(R, num_points - n_unc, 2, device=src.device)], 1)
        labels = sample(tgt, coords).squeeze(1)
    x = sample(src, coords).squeeze(1)
    ce = F.binary_cross_entropy_with_logits(x, labels, reduction="none").mean(1).sum() / num_masks
    p = x.sigmoid()
    dice = (1 - (2 * (p * labels).sum(-1) + 1) / (p.sum(-1) + labels.sum(-1) + 1)).sum() / num_masks
    return ce + dice


def main():
    g = torch.Generator(device=DEV).manual_seed(0)
    targets = workload.synthetic_targets(B, 1024, 1024, num_classes=K, seed=0, device=DEV)
    max_num = max(len(t["labels"]) for t in targets)

    def head(q):
        return {"pred_logits": torch.randn(B, q, K + 1, device=DEV, generator=g).requires_grad_(True),
                "pred_masks": (torch.randn(B, q, 256, 256, device=DEV, generator=g) * 3).requires_grad_(True)}

    out = head(Q)
    out["aux_outputs"] = [head(Q) for _ in range(9)]
    dn = head(max_num)
    dn["aux_outputs"] = [head(max_num) for _ in range(9)]
    dn["dn_args"] = {"pad_size": max_num, "max_num": max_num}
    out["dn_out"] = dn
    crit, weighted_sum = workload.build_criterion(num_classes=K, device=DEV)
    crit.train(True)
    leaves = [out["pred_masks"], out["pred_logits"]] + [a[k] for a in out["aux_outputs"] for k in a] + \
             [dn["pred_masks"], dn["pred_logits"]] + [a[k] for a in dn["aux_outputs"] for k in a]

    def fwd():
        return weighted_sum(crit(out, targets))

    def fwd_bwd():
        for t in leaves:
            t.grad = None
        fwd().backward()

    l0 = _lib.launch_count()
    loss = fwd()
    launches_fwd = _lib.launch_count() - l0
    fwd_ms, step_ms = timed(fwd), timed(fwd_bwd)
    main_head = {k: v.detach() for k, v in out.items() if k in ("pred_logits", "pred_masks")}
    step = crit._step_state(targets, torch.device(DEV))
    idx = crit._match(main_head, targets, step)
    match_ms = timed(lambda: crit.matcher.match_device(main_head, targets))
    labels_ms = timed(lambda: crit.loss_labels(main_head, step, idx, 1.0))
    masks_ms = timed(lambda: crit.loss_masks(main_head, step, idx, 1.0))
    # one head's mask losses, forward + backward into the mask logits: ours vs the reference's formulation
    pm = out["pred_masks"]

    def ours_fb():
        pm.grad = None
        d = crit.loss_masks({"pred_masks": pm}, step, idx, 1.0)
        (d["loss_mask"] + d["loss_dice"]).backward()

    def stock_fb():
        pm.grad = None
        stock_loss_masks(pm, targets, idx[0], idx[1], idx[2], 1.0).backward()

    ours_fb_ms, stock_fb_ms = timed(ours_fb), timed(stock_fb)
    print(json.dumps({"probe": "set_criterion_bench_geometry", "B": B, "Q": Q, "heads": 10, "dn_queries": max_num,
                      "targets_total": sum(len(t["labels"]) for t in targets), "loss": float(loss),
                      "forward_ms": fwd_ms, "forward_backward_ms": step_ms, "native_launches_forward": launches_fwd,
                      "one_head": {"matcher_ms": match_ms, "loss_labels_ms": labels_ms, "loss_masks_fwd_ms": masks_ms,
                                   "loss_masks_fwd_bwd_ms": ours_fb_ms, "stock_loss_masks_fwd_bwd_ms": stock_fb_ms,
                                   "loss_masks_speedup": stock_fb_ms / ours_fb_ms},
                      "host_syncs_forward": 0}))


if __name__ == "__main__":
    main()
