"""Generates tests/golden/criterion.pt from the UNMODIFIED reference SetCriterion + HungarianMatcher (authoring
container only):

    python tests/golden/make_golden_criterion.py

Inputs are seeded; the reference draws its random points from the global torch generator, so the seed set right before
each call is part of the fixture.  The mask-piloted ("dn") branch of the reference hard-codes ``.cuda()``
(criterion.py:249-255); it runs on the CPU here under oracle.ref_loader.cuda_is_identity."""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

CFG = dict(num_classes=5, eos_coef=0.1, num_points=112, oversample_ratio=3.0, importance_sample_ratio=0.75,
           cost_class=2.0, cost_mask=5.0, cost_dice=5.0)


def inputs(seed=31, B=2, Q=12, K=5, H=24, W=32, Hg=96, Wg=128, counts=(3, 5), groups=2, aux=2, with_dn=True):
    g = torch.Generator().manual_seed(seed)

    def head(q):
        return {"pred_logits": torch.randn(B, q, K + 1, generator=g), "pred_masks": torch.randn(B, q, H, W, generator=g) * 3}

    targets = []
    ys, xs = torch.arange(Hg).view(-1, 1).float(), torch.arange(Wg).view(1, -1).float()
    for b in range(B):
        n = counts[b % len(counts)]
        cy, cx = torch.rand(n, generator=g) * Hg, torch.rand(n, generator=g) * Wg
        ry, rx = (torch.rand(n, generator=g) * 0.3 + 0.05) * Hg, (torch.rand(n, generator=g) * 0.3 + 0.05) * Wg
        masks = (((ys[None] - cy.view(-1, 1, 1)) / ry.view(-1, 1, 1)) ** 2 +
                 ((xs[None] - cx.view(-1, 1, 1)) / rx.view(-1, 1, 1)) ** 2) <= 1.0
        targets.append({"labels": torch.randint(0, K, (n,), generator=g), "masks": masks})
    out = head(Q)
    out["aux_outputs"] = [head(Q) for _ in range(aux)]
    if with_dn:
        max_num = max(counts[:B])
        dn = head(groups * max_num)
        dn["aux_outputs"] = [head(groups * max_num) for _ in range(aux)]
        dn["dn_args"] = {"pad_size": groups * max_num, "max_num": max_num}
        out["dn_out"] = dn
    else:
        out["dn_out"] = None
    return out, targets


CASES = (  # name, with_dn, training, dn_no_lb, seed
    ("train_dn_aux", True, True, False, 17),
    ("train_dn_no_lb", True, True, True, 18),
    ("eval_no_dn", False, False, False, 19),
    ("train_without_dn_out", False, True, False, 20),
)


def main():
    warnings.simplefilter("ignore")
    crit_mod = ref_loader.load_criterion()
    mat_mod = ref_loader.load_matcher()
    res = {}
    for name, with_dn, training, no_lb, seed in CASES:
        outputs, targets = inputs(with_dn=with_dn)
        matcher = mat_mod.HungarianMatcher(cost_class=CFG["cost_class"], cost_mask=CFG["cost_mask"],
                                           cost_dice=CFG["cost_dice"], num_points=CFG["num_points"])
        crit = crit_mod.SetCriterion(CFG["num_classes"], matcher=matcher, weight_dict={}, eos_coef=CFG["eos_coef"],
                                     losses=["labels", "masks"], num_points=CFG["num_points"],
                                     oversample_ratio=CFG["oversample_ratio"],
                                     importance_sample_ratio=CFG["importance_sample_ratio"], dn_no_lb=no_lb)
        crit.train(training)
        torch.manual_seed(seed)
        with ref_loader.cuda_is_identity():
            losses = crit(outputs, targets)
        res[name] = {k: v.detach().clone() for k, v in losses.items()}
        print(name, len(losses), {k: round(float(v), 5) for k, v in list(losses.items())[:4]})
    torch.save(res, os.path.join(HERE, "criterion.pt"))


if __name__ == "__main__":
    main()
