"""Generates tests/golden/matcher.pt from the UNMODIFIED reference matcher (run in the authoring container only):

    python tests/golden/make_golden_matcher.py

Inputs are seeded; the reference draws its random points from the global torch generator, so the seed set right
before the call is part of the fixture."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref_loader  # noqa: E402


def inputs(seed=71, B=3, Q=20, K=7, H=32, W=48, Hg=128, Wg=192):
    g = torch.Generator().manual_seed(seed)
    outputs = {"pred_logits": torch.randn(B, Q, K + 1, generator=g), "pred_masks": torch.randn(B, Q, H, W, generator=g) * 3}
    targets = []
    for b in range(B):
        n = [4, 1, 9][b % 3]
        ys = torch.arange(Hg).view(-1, 1).float()
        xs = torch.arange(Wg).view(1, -1).float()
        cy, cx = torch.rand(n, generator=g) * Hg, torch.rand(n, generator=g) * Wg
        ry, rx = (torch.rand(n, generator=g) * 0.3 + 0.05) * Hg, (torch.rand(n, generator=g) * 0.3 + 0.05) * Wg
        masks = (((ys[None] - cy.view(-1, 1, 1)) / ry.view(-1, 1, 1)) ** 2 +
                 ((xs[None] - cx.view(-1, 1, 1)) / rx.view(-1, 1, 1)) ** 2) <= 1.0
        targets.append({"labels": torch.randint(0, K, (n,), generator=g), "masks": masks})
    return outputs, targets


def main():
    m = ref_loader.load_matcher()
    outputs, targets = inputs()
    out = {"cases": []}
    for (wc, wm, wd, P, seed) in ((2.0, 5.0, 5.0, 12544, 5), (1.0, 1.0, 1.0, 300, 6)):
        matcher = m.HungarianMatcher(cost_class=wc, cost_mask=wm, cost_dice=wd, num_points=P)
        torch.manual_seed(seed)
        idx = matcher(outputs, targets)
        out["cases"].append({"weights": (wc, wm, wd), "num_points": P, "seed": seed,
                             "indices": [(i.clone(), j.clone()) for i, j in idx]})
    # the two cost terms on fixed points (no RNG involved)
    g = torch.Generator().manual_seed(9)
    a, t = torch.randn(6, 50, generator=g) * 2, (torch.rand(4, 50, generator=g) > 0.5).float()
    out["dice"] = m.batch_dice_loss(a, t)
    out["ce"] = m.batch_sigmoid_ce_loss(a, t)
    torch.save(out, os.path.join(HERE, "matcher.pt"))
    print("wrote matcher.pt:", [c["indices"][0] for c in out["cases"]])


if __name__ == "__main__":
    main()
