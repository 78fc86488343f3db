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
    print(json.dumps({"probe": "set_criterion_bench_geometry", "B": B, "Q": Q, "heads": 10, "dn_queries": max_num,
                      "targets_total": sum(len(t["labels"]) for t in targets), "loss": float(loss),
                      "forward_ms": fwd_ms, "forward_backward_ms": step_ms, "native_launches_forward": launches_fwd,
                      "one_head": {"matcher_ms": match_ms, "loss_labels_ms": labels_ms, "loss_masks_fwd_ms": masks_ms},
                      "host_syncs_forward": 0}))


if __name__ == "__main__":
    main()
