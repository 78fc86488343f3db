"""Mirror of the reference's ``SetCriterion`` (mask2former/modeling/criterion.py:90-322): same constructor, same
``forward(outputs, targets)`` contract, same loss keys (``loss_ce / loss_mask / loss_dice`` + ``_dn`` + ``_{i}``
suffixes), same per-call consumption of the random generator (the candidate points of ``loss_masks`` and the matcher's
points are drawn like the reference's; END-TO-END seed parity with the reference does not hold, because the DN
preparation upstream draws differently -- see masked_decoder.prepare_for_dn_v5) -- arranged for the device:

  * the matcher's index pairs stay on the device when the matcher is ``mp_former_b200.matcher.HungarianMatcher``
    (``match_device``); any matcher with the reference's ``forward`` contract (list of CPU index pairs) works too;
  * predictions are point-sampled where the prediction heads wrote them and GT masks where the loader put them
    (``native.point_sample_rows`` through device pointer tables): no ``pred_masks[idx]`` gather, no float / zero-padded
    copy of every GT mask of the batch per loss call (criterion.py:152-155; 1.3 GB at the bench geometry), and the
    backward adds straight into a dense mask-logit gradient (``native.PointSampleRows``);
  * the fixed assignment of the mask-piloted ("dn") queries (criterion.py:246-257) and the batch index vectors are
    built once per step on the host from the target counts instead of per image with ``.cuda()`` uploads;
  * ``num_masks`` is a host number (single process) or stays a device tensor after the all-reduce (no ``.item()``).

The loss arithmetic itself (cross-entropy, BCE, dice, top-k of the uncertainty scores) is library PyTorch.
There is no CPU path: CUDA tensors only."""
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from . import native
from .matcher import PackedTargets, targets_key


def _world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class SetCriterion(nn.Module):
    """Loss of the mask-classification model (ref criterion.py:90-322): Hungarian assignment between ground truth and
    predictions, then classification and point-sampled mask losses for every matched pair, for the mask-piloted
    queries with their fixed assignment, and for every auxiliary decoder layer."""

    def __init__(self, num_classes, matcher, weight_dict, eos_coef, losses, num_points, oversample_ratio,
                 importance_sample_ratio, dn_no_lb=False):
        super().__init__()
        self.num_classes = num_classes
        self.matcher = matcher
        self.weight_dict = weight_dict
        self.eos_coef = eos_coef
        self.losses = losses
        self.dn_no_lb = dn_no_lb
        empty_weight = torch.ones(self.num_classes + 1)
        empty_weight[-1] = self.eos_coef
        self.register_buffer("empty_weight", empty_weight)
        self.num_points = num_points
        self.oversample_ratio = oversample_ratio
        self.importance_sample_ratio = importance_sample_ratio
        assert oversample_ratio >= 1 and 0 <= importance_sample_ratio <= 1
        self._step_key = None
        self._step = None
        self._status = None          # int32 [1] on the device: != 0 once a matching of this step has failed

    def check_status(self):
        """Raises the reference's error for a failed assignment (scipy's ``ValueError: matrix contains invalid numeric
        entries`` / ``cost matrix is infeasible`` at matcher.py:153 -- NaN/-inf costs from diverged logits, or a label
        outside [0, num_classes)).  The device matcher reports failure through a status word instead of a host round
        trip per head, so the forward itself cannot raise: it poisons every loss of the step with NaN (nothing is
        silently trained on a made-up assignment) and this method -- one 4-byte device read; call it wherever the
        trainer already synchronises, e.g. next to the loss logging -- raises."""
        st, self._status = self._status, None
        if st is not None and int(st.item()) != 0:
            raise ValueError(f"HungarianMatcher: cost matrix of image {int(st.item()) - 1} contains invalid numeric "
                             "entries (NaN / -inf) or is infeasible")

    # -- per-step state shared by the 20 loss evaluations of one forward --------------------------------------------
    class _Step:
        def __init__(self, targets, device, packed=None):
            self.packed = packed if packed is not None else PackedTargets(targets, device)
            self.device = device
            self.counts = self.packed.counts
            offs = [0]
            for n in self.counts:
                offs.append(offs[-1] + n)
            self.offsets = torch.tensor(offs[:-1], dtype=torch.int64, device=device)
            self.total = offs[-1]
            # device copy of the local target count, made once per step-state: the all-reduce below then needs no
            # host->device upload per forward (and the forward stays CUDA-graph capturable)
            self.total_dev = torch.tensor([float(self.total)], dtype=torch.float, device=device)
            if self.packed.uniform():
                self.masks = self.packed.masks
            else:       # zero-padded to the largest map, top-left aligned (utils/misc.py:48-73); rare: the model pads
                hg = max(s[0] for s in self.packed.sizes)
                wg = max(s[1] for s in self.packed.sizes)
                self.masks = [F.pad(m, (0, wg - m.shape[-1], 0, hg - m.shape[-2])) for m in self.packed.masks]
            self.hw = tuple(self.masks[0].shape[-2:]) if self.masks else (1, 1)
            self.elsize = self.masks[0].element_size() if self.masks else 1
            self.mask_ptrs = torch.tensor([m.data_ptr() if m.shape[0] else 0 for m in self.masks], dtype=torch.int64,
                                          device=device)
            self._batch_index = {}
            self._dn = {}
            self.num_masks = None

        def batch_index(self, sizes):
            key = tuple(sizes)
            if key not in self._batch_index:
                idx = [b for b, s in enumerate(sizes) for _ in range(s)]
                self._batch_index[key] = torch.tensor(idx, dtype=torch.int64, device=self.device)
            return self._batch_index[key]

        def dn_indices(self, dn_args):
            """(batch, query, target) of the mask-piloted queries: group g's query g*max_num + j <- target j."""
            max_num = int(dn_args["max_num"])
            scalar = int(dn_args["pad_size"]) // max_num
            key = (max_num, scalar)
            if key not in self._dn:
                b, q, t = [], [], []
                for i, n in enumerate(self.counts):
                    for g in range(scalar):
                        b += [i] * n
                        q += [g * max_num + j for j in range(n)]
                        t += list(range(n))
                mk = (lambda x: torch.tensor(x, dtype=torch.int64, device=self.device))
                self._dn[key] = (mk(b), mk(q), mk(t), scalar)
            return self._dn[key]

    def _step_state(self, targets, device):
        key = targets_key(targets)
        if key != self._step_key:
            packed = self.matcher.pack_targets(targets, device) if hasattr(self.matcher, "pack_targets") else None
            self._step = SetCriterion._Step(targets, device, packed)
            self._step_key = key
        return self._step

    # -- assignment -------------------------------------------------------------------------------------------------
    def _match(self, outputs, targets, step):
        """-> (batch index, query index, target index), int64 device vectors over all matched pairs."""
        if hasattr(self.matcher, "match_device"):
            q, t, counts, _, status = self.matcher.match_device(outputs, targets)
            # a failed solve returns in-range placeholder pairs (memory-safe) and a non-zero status
            self._status = status if self._status is None else torch.maximum(self._status, status)
            nq = outputs["pred_logits"].shape[1]
            return step.batch_index([min(nq, n) for n in counts]), q, t
        indices = self.matcher(outputs, targets)
        dev = step.device
        sizes = [len(i) for i, _ in indices]
        cat = (lambda xs: torch.cat([torch.as_tensor(x, dtype=torch.int64) for x in xs]).to(dev) if xs else
               torch.zeros(0, dtype=torch.int64, device=dev))
        return step.batch_index(sizes), cat([i for i, _ in indices]), cat([j for _, j in indices])

    # -- the two losses ---------------------------------------------------------------------------------------------
    def loss_labels(self, outputs, step, idx, num_masks):
        """Classification loss (ref criterion.py:123-139): matched queries take their target's class, all others the
        no-object class, which is down-weighted by ``eos_coef``."""
        b, q, t = idx
        logits = outputs["pred_logits"].float()
        classes = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=logits.device)
        # (labels outside [0, num_classes) make the matcher fail -- status, NaN losses --; clamped so that the
        # class-index gather of the cross-entropy kernel stays in range until check_status() raises)
        classes[b, q] = step.packed.labels[step.offsets[b] + t].clamp(0, self.num_classes)
        return {"loss_ce": F.cross_entropy(logits.transpose(1, 2), classes, self.empty_weight)}

    def loss_masks(self, outputs, step, idx, num_masks):
        """Point-sampled binary cross-entropy and dice losses of the matched masks (ref criterion.py:141-191) with
        PointRend's importance sampling of the points (detectron2 get_uncertain_point_coords_with_randomness)."""
        b, q, t = idx
        pm = outputs["pred_masks"].float()
        B, Q, H, W = pm.shape
        R = int(b.numel())
        if R == 0:
            zero = pm.sum() * 0.0
            return {"loss_mask": zero, "loss_dice": zero}
        if pm.stride(-1) != 1 or pm.stride(-2) != W:
            pm = pm.contiguous()
        n_over = int(self.num_points * self.oversample_ratio)
        n_unc = int(self.importance_sample_ratio * self.num_points)
        n_rand = self.num_points - n_unc
        with torch.no_grad():
            src_ptrs = pm.data_ptr() + 4 * (b * pm.stride(0) + q * pm.stride(1))
            cand = torch.rand(R, n_over, 2, device=pm.device, dtype=pm.dtype)
            unc = native.point_sample_rows(src_ptrs, True, (H, W), cand, neg_abs=True)     # -|logit|, :73-87
            if n_over <= native.TOPK_GATHER_MAX_N:
                # the SET of the n_unc most uncertain candidates (the losses are sums over the points): radix select
                # in shared memory + ordered gather, one kernel instead of a segmented sort, top-k and a gather
                coords = native.topk_gather_rows(unc, cand, n_unc)
            else:
                top = unc.topk(n_unc, dim=1).indices
                coords = torch.gather(cand, 1, top.unsqueeze(-1).expand(-1, -1, 2))
            if n_rand > 0:
                coords = torch.cat([coords, torch.rand(R, n_rand, 2, device=pm.device)], dim=1)
            coords = coords.contiguous()
            tgt_ptrs = step.mask_ptrs[b] + t * (step.hw[0] * step.hw[1] * step.elsize)
            labels = native.point_sample_rows(tgt_ptrs, step.packed.is_f32, step.hw, coords)
        x = native.PointSampleRows.apply(pm, b * Q + q, coords)
        ce = F.binary_cross_entropy_with_logits(x, labels, reduction="none").mean(1).sum() / num_masks
        p = x.sigmoid()
        dice = (1 - (2 * (p * labels).sum(-1) + 1) / (p.sum(-1) + labels.sum(-1) + 1)).sum() / num_masks
        return {"loss_mask": ce, "loss_dice": dice}

    def get_loss(self, loss, outputs, step, idx, num_masks):
        loss_map = {"labels": self.loss_labels, "masks": self.loss_masks}
        assert loss in loss_map, f"do you really want to compute {loss} loss?"
        return loss_map[loss](outputs, step, idx, num_masks)

    def _all_losses(self, outputs, step, idx, num_masks):
        out = {}
        for loss in self.losses:
            out.update(self.get_loss(loss, outputs, step, idx, num_masks))
        return out

    # -- forward ----------------------------------------------------------------------------------------------------
    def forward(self, outputs, targets):
        """outputs: {"pred_logits", "pred_masks", optional "aux_outputs": [...], "dn_out": None | {"pred_logits",
        "pred_masks", "aux_outputs", "dn_args": {"pad_size", "max_num"}}}; targets: list of {"labels", "masks"}
        (ref criterion.py:213-304).  Returns the dict of unweighted losses."""
        main = {k: v for k, v in outputs.items() if k != "aux_outputs" and k != "dn_out"}
        dn_out = outputs.get("dn_out")
        dev = main["pred_masks"].device
        step = self._step_state(targets, dev)
        self._status = None
        if step.num_masks is None:
            # average number of target masks across ranks (ref criterion.py:262-269).  It depends on the targets
            # alone, so it is part of the per-step state: one all-reduce when a new batch of targets arrives, kept on
            # the device (no .item() sync), and nothing to communicate when a step with the same targets is replayed
            # from a CUDA graph.
            ws = _world_size()
            if ws > 1:
                nm = step.total_dev.clone()
                dist.all_reduce(nm)
                step.num_masks = torch.clamp(nm / ws, min=1)[0]
            else:
                step.num_masks = max(float(step.total), 1.0)
        num_masks = step.num_masks

        losses = self._all_losses(main, step, self._match(main, targets, step), num_masks)

        use_dn = bool(self.training and dn_out)
        zero = torch.zeros((), device=dev)           # (a fill kernel: capturable, unlike a host scalar upload)
        if use_dn:
            db, dq, dt, scalar = step.dn_indices(dn_out["dn_args"])
            dn_idx = (db, dq, dt)

        def dn_losses(o, suffix):
            if use_dn:
                d = self._all_losses(o, step, dn_idx, num_masks * scalar)
                return {k + "_dn" + suffix: v for k, v in d.items()}
            return {k + suffix: zero for k in ("loss_mask_dn", "loss_dice_dn", "loss_ce_dn")}

        losses.update(dn_losses({k: v for k, v in dn_out.items() if k != "aux_outputs"} if use_dn else None, ""))
        if "aux_outputs" in outputs:
            for i, aux in enumerate(outputs["aux_outputs"]):
                d = self._all_losses(aux, step, self._match(aux, targets, step), num_masks)
                losses.update({k + f"_{i}": v for k, v in d.items()})
                losses.update(dn_losses(dn_out["aux_outputs"][i] if use_dn else None, f"_{i}"))
        if self.dn_no_lb:
            losses = {k: losses[k] for k in losses if not k.startswith("loss_ce_dn")}
        if self._status is not None:        # failed assignment -> every loss NaN (see check_status)
            poison = torch.where(self._status[0] == 0, 0.0, float("nan"))
            losses = {k: v + poison for k, v in losses.items()}
        return losses

    def __repr__(self):
        head = "Criterion " + self.__class__.__name__
        body = [
            "matcher: {}".format(self.matcher.__repr__(_repr_indent=8)),
            "losses: {}".format(self.losses),
            "weight_dict: {}".format(self.weight_dict),
            "num_classes: {}".format(self.num_classes),
            "eos_coef: {}".format(self.eos_coef),
            "num_points: {}".format(self.num_points),
            "oversample_ratio: {}".format(self.oversample_ratio),
            "importance_sample_ratio: {}".format(self.importance_sample_ratio),
        ]
        return "\n".join([head] + [" " * 4 + line for line in body])
