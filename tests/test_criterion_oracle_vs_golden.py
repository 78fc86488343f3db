"""CPU: the criterion restatement (oracle/criterion_oracle.py) against the losses of the UNMODIFIED reference
SetCriterion + HungarianMatcher on seeded inputs (tests/golden/make_golden_criterion.py): same keys, values within 1e-5
(same random-number consumption, call for call)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_criterion import CASES, CFG, inputs  # noqa: E402
from oracle import criterion_oracle as CO  # noqa: E402


def test_criterion_oracle_matches_reference_golden():
    G = torch.load(os.path.join(HERE, "golden", "criterion.pt"), weights_only=False)
    for name, with_dn, training, no_lb, seed in CASES:
        outputs, targets = inputs(with_dn=with_dn)
        torch.manual_seed(seed)
        got = CO.set_criterion(outputs, targets, losses=["labels", "masks"], training=training, dn_no_lb=no_lb, **CFG)
        ref = G[name]
        assert sorted(got) == sorted(ref), (name, sorted(set(got) ^ set(ref)))
        for k in ref:
            assert torch.allclose(got[k].float().cpu(), ref[k].float(), rtol=1e-5, atol=1e-6), (name, k, got[k], ref[k])


def test_criterion_oracle_gradients_flow_to_matched_rows_only():
    outputs, targets = inputs(with_dn=False)
    outputs["pred_masks"].requires_grad_(True)
    torch.manual_seed(3)
    got = CO.set_criterion({k: v for k, v in outputs.items() if k != "aux_outputs"}, targets,
                           losses=["masks"], training=False, **CFG)
    (got["loss_mask"] + got["loss_dice"]).backward()
    rows = outputs["pred_masks"].grad.abs().flatten(2).sum(-1) > 0
    assert rows.sum(1).tolist() == [len(t["labels"]) for t in targets]
