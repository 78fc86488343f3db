"""TEST INFRASTRUCTURE ONLY -- restatement (plain PyTorch, device-agnostic) of the reference's SetCriterion, the caller
that consumes the decoder's outputs.  Not used by the product path or by bench.py's timed legs.

Follows mask2former/modeling/criterion.py of the reference:
  * ``dice_loss`` :21-41, ``sigmoid_ce_loss`` :49-67, ``calculate_uncertainty`` :73-87
  * ``SetCriterion.loss_labels`` :123-141, ``loss_masks`` :141-191, permutation indices :193-203
  * ``SetCriterion.forward`` :213-304 (main matching, num_masks normalisation, the mask-piloted "dn" losses with the
    fixed GT assignment :243-262, the per-layer auxiliary losses :276-299, ``dn_no_lb`` :300-301)
and detectron2's ``get_uncertain_point_coords_with_randomness`` (PointRend importance sampling; third-party, unpinned
"git master", INSTALL.md:36-38 -- restated from the published algorithm, see oracle/ref_loader.py).
Random-number consumption (global generator of the masks' device) is the reference's, call for call, so that seeded
runs are comparable: matcher (rand(1, P, 2) per image) -> per ``loss_masks`` call rand(R, k*N, 2) then rand(R, N-bN, 2).
Pinned by ``tests/golden/criterion.pt`` (generated from the unmodified reference by tests/golden/make_golden_criterion.py).
"""
import torch
import torch.nn.functional as F

from .matcher_oracle import hungarian_match, point_sample


def uncertain_points(logits, num_points, oversample_ratio, importance_sample_ratio):
    """logits [R, 1, H, W] -> point coordinates [R, num_points, 2]."""
    R = logits.shape[0]
    n_over = int(num_points * oversample_ratio)
    cand = torch.rand(R, n_over, 2, device=logits.device, dtype=logits.dtype)
    unc = -point_sample(logits, cand).abs()[:, 0]                       # criterion.py:73-87
    n_unc = int(importance_sample_ratio * num_points)
    top = unc.topk(n_unc, dim=1).indices
    coords = torch.gather(cand, 1, top[..., None].expand(-1, -1, 2))
    if num_points - n_unc > 0:
        coords = torch.cat([coords, torch.rand(R, num_points - n_unc, 2, device=logits.device)], 1)
    return coords


def flat_indices(indices):
    """[(src_b, tgt_b)] -> (batch index, src index, tgt index), images back to back (criterion.py:193-203)."""
    b = torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(indices)])
    return b, torch.cat([s for s, _ in indices]), torch.cat([t for _, t in indices])


def loss_labels(outputs, targets, indices, num_classes, empty_weight):
    logits = outputs["pred_logits"].float()
    b, s, t = flat_indices(indices)
    tc = torch.full(logits.shape[:2], num_classes, dtype=torch.int64, device=logits.device)
    tc[b, s] = torch.cat([tt["labels"][j] for tt, (_, j) in zip(targets, indices)])
    return {"loss_ce": F.cross_entropy(logits.transpose(1, 2), tc, empty_weight.to(logits.device))}


def loss_masks(outputs, targets, indices, num_masks, num_points, oversample_ratio, importance_sample_ratio):
    b, s, t = flat_indices(indices)
    src = outputs["pred_masks"][b, s][:, None]
    hg = max(tt["masks"].shape[-2] for tt in targets)
    wg = max(tt["masks"].shape[-1] for tt in targets)
    # zero-padded to the largest map, top-left aligned (utils/misc.py:48-73)
    tgt = torch.cat([F.pad(tt["masks"][j], (0, wg - tt["masks"].shape[-1], 0, hg - tt["masks"].shape[-2]))
                     for tt, (_, j) in zip(targets, indices)]).to(src)[:, None]
    with torch.no_grad():
        coords = uncertain_points(src, num_points, oversample_ratio, importance_sample_ratio)
        labels = point_sample(tgt, coords).squeeze(1)
    x = point_sample(src, coords).squeeze(1)
    ce = F.binary_cross_entropy_with_logits(x, labels, reduction="none").mean(1).sum() / num_masks
    p = x.sigmoid()
    dice = (1 - (2 * (p * labels).sum(-1) + 1) / (p.sum(-1) + labels.sum(-1) + 1)).sum() / num_masks
    return {"loss_mask": ce, "loss_dice": dice}


def dn_indices(targets, dn_args, device):
    """Fixed assignment of the mask-piloted queries: group g's query g*max_num + j belongs to target j (:246-257)."""
    scalar = dn_args["pad_size"] // dn_args["max_num"]
    out = []
    for tt in targets:
        n = len(tt["labels"])
        t = torch.arange(n, device=device).repeat(scalar)
        o = (torch.arange(scalar, device=device) * dn_args["max_num"]).repeat_interleave(n) + t
        out.append((o, t))
    return out, scalar


def set_criterion(outputs, targets, *, num_classes, eos_coef, losses, num_points, oversample_ratio,
                  importance_sample_ratio, cost_class, cost_mask, cost_dice, training=True, dn_no_lb=False,
                  world_size=1, global_num_masks=None, matcher=None):
    """``SetCriterion.forward`` (criterion.py:213-304).  ``matcher(outputs, targets) -> [(i, j)]`` defaults to the
    matcher oracle with the given cost weights.  ``global_num_masks`` / ``world_size``: the all-reduced target count of a
    data-parallel run (this restatement does not communicate)."""
    dev = outputs["pred_masks"].device
    if matcher is None:
        def matcher(o, t):
            return [(i.to(dev), j.to(dev)) for i, j in
                    hungarian_match(o, t, num_points, cost_class, cost_mask, cost_dice)[0]]
    empty_weight = torch.ones(num_classes + 1)
    empty_weight[-1] = eos_coef

    def all_losses(o, idx, nm):
        d = {}
        for name in losses:
            if name == "labels":
                d.update(loss_labels(o, targets, idx, num_classes, empty_weight))
            elif name == "masks":
                d.update(loss_masks(o, targets, idx, nm, num_points, oversample_ratio, importance_sample_ratio))
            else:
                raise AssertionError(f"do you really want to compute {name} loss?")
        return d

    # criterion.py:231-237: the number of targets is summed over the ranks (all_reduce) and divided by the world size
    total = float(sum(len(t["labels"]) for t in targets)) if global_num_masks is None else float(global_num_masks)
    num_masks = max(total / world_size, 1.0)
    dn_out = outputs.get("dn_out")
    main = {k: v for k, v in outputs.items() if k not in ("aux_outputs", "dn_out")}
    out = all_losses(main, matcher(main, targets), num_masks)
    use_dn = bool(training and dn_out)
    zero = torch.zeros((), device=dev)

    def dn_losses(o, suffix):
        if use_dn:
            return {k + "_dn" + suffix: v for k, v in all_losses(o, dn_idx, num_masks * scalar).items()}
        return {k + suffix: zero for k in ("loss_mask_dn", "loss_dice_dn", "loss_ce_dn")}

    if use_dn:
        dn_idx, scalar = dn_indices(targets, dn_out["dn_args"], dev)
        out.update(dn_losses({k: v for k, v in dn_out.items() if k != "aux_outputs"}, ""))
    else:
        out.update(dn_losses(None, ""))
    for i, aux in enumerate(outputs.get("aux_outputs", [])):
        out.update({f"{k}_{i}": v for k, v in all_losses(aux, matcher(aux, targets), num_masks).items()})
        out.update(dn_losses(dn_out["aux_outputs"][i] if use_dn else None, f"_{i}"))
    if dn_no_lb:
        out = {k: v for k, v in out.items() if not k.startswith("loss_ce_dn")}
    return out
