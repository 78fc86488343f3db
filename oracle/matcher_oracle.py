"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Hungarian matching cost (groundwork for the next
row of the coverage contract, SURVEY.md §8f rank 1: "criterion point sampling + Hungarian cost").  Not used by the
product path or by bench.py.

Restates mask2former/modeling/matcher.py of the reference:
  * ``batch_dice_loss``        matcher.py:15-30
  * ``batch_sigmoid_ce_loss``  matcher.py:38-62
  * ``HungarianMatcher.memory_efficient_forward``  matcher.py:96-157 (one shared set of random points per image,
    class cost = -softmax probability of the target class, LSAP by scipy)
and the one third-party function on that path, detectron2's ``point_sample`` (unpinned "git master" dependency,
INSTALL.md:36-38; not vendored): bilinear ``F.grid_sample`` at ``2 * coords - 1``, ``align_corners=False``.
Pinned by ``tests/golden/matcher.pt`` (generated from the unmodified reference by ``tests/golden/make_golden_matcher.py``).
"""
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment


def point_sample(inp, point_coords, align_corners=False):
    """inp [N, C, H, W], point_coords [N, P, 2] in [0, 1] (x, y) -> [N, C, P]."""
    return F.grid_sample(inp, 2.0 * point_coords.unsqueeze(2) - 1.0, align_corners=align_corners).squeeze(3)


def batch_dice_cost(inputs, targets):
    """inputs [Q, P] logits, targets [n, P] in {0, 1} -> [Q, n]   (matcher.py:15-30)."""
    inputs = inputs.sigmoid().flatten(1)
    numerator = 2 * torch.einsum("nc,mc->nm", inputs, targets)
    denominator = inputs.sum(-1)[:, None] + targets.sum(-1)[None, :]
    return 1 - (numerator + 1) / (denominator + 1)


def batch_sigmoid_ce_cost(inputs, targets):
    """[Q, P] logits, [n, P] targets -> [Q, n] mean binary cross-entropy over the points (matcher.py:38-62)."""
    hw = inputs.shape[1]
    pos = F.binary_cross_entropy_with_logits(inputs, torch.ones_like(inputs), reduction="none")
    neg = F.binary_cross_entropy_with_logits(inputs, torch.zeros_like(inputs), reduction="none")
    return (torch.einsum("nc,mc->nm", pos, targets) + torch.einsum("nc,mc->nm", neg, 1 - targets)) / hw


def matching_cost(pred_logits, pred_masks, labels, masks, point_coords, cost_class=1.0, cost_mask=1.0, cost_dice=1.0):
    """One image.  pred_logits [Q, K+1], pred_masks [Q, H, W], labels [n], masks [n, Hg, Wg] (bool / float),
    point_coords [1, P, 2] -> cost matrix [Q, n]   (matcher.py:105-149)."""
    out_prob = pred_logits.softmax(-1)
    c_class = -out_prob[:, labels]
    out_mask = pred_masks[:, None]
    tgt_mask = masks.to(out_mask)[:, None]
    tgt_pts = point_sample(tgt_mask, point_coords.repeat(tgt_mask.shape[0], 1, 1)).squeeze(1)
    out_pts = point_sample(out_mask, point_coords.repeat(out_mask.shape[0], 1, 1)).squeeze(1)
    out_pts, tgt_pts = out_pts.float(), tgt_pts.float()
    c_mask = batch_sigmoid_ce_cost(out_pts, tgt_pts)
    c_dice = batch_dice_cost(out_pts, tgt_pts)
    return cost_mask * c_mask + cost_class * c_class + cost_dice * c_dice


@torch.no_grad()
def hungarian_match(outputs, targets, num_points, cost_class=1.0, cost_mask=1.0, cost_dice=1.0):
    """``HungarianMatcher.forward``: consumes ``torch.rand(1, num_points, 2)`` once per image, in image order, like
    the reference (seed the global generator to reproduce it).  Returns [(index_i, index_j)] int64 pairs and, for
    tests, the cost matrices."""
    bs, num_queries = outputs["pred_logits"].shape[:2]
    indices, costs = [], []
    for b in range(bs):
        coords = torch.rand(1, num_points, 2, device=outputs["pred_masks"].device)
        C = matching_cost(outputs["pred_logits"][b], outputs["pred_masks"][b], targets[b]["labels"],
                          targets[b]["masks"], coords, cost_class, cost_mask, cost_dice)
        C = C.reshape(num_queries, -1).cpu()
        costs.append(C)
        i, j = linear_sum_assignment(C)
        indices.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
    return indices, costs
