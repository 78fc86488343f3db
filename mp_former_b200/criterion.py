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

  * ALL prediction heads of a step (final + auxiliary layers, matching and mask-piloted queries: 20 loss evaluations
    in the published recipe) are evaluated together (``_forward_heads``): the random numbers are drawn head by head
    in the reference's order, everything else runs once over the rows of all heads -- one assignment launch, one
    sampling launch per direction, one fused BCE + dice launch per direction (``native.MaskLossRows``,
    csrc/mask_loss.cu), one cross-entropy over the concatenated class logits -- and the mask-logit gradients of all
    heads land in ONE buffer (``native.PointSampleViews``) that the decoder's batched head backward consumes as is.
    The head-by-head evaluation (``_forward_sequential``, the reference's loop) remains for inputs the joint path
    does not cover (mixed target sizes, no targets at all, other ``losses`` lists).

The cross-entropy is library PyTorch; the rest of the loss arithmetic are this package's kernels.
There is no CPU path: CUDA tensors only."""
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from . import native
from .matcher import PackedTargets, targets_key


class LossDict(dict):
    """The criterion's result: loss name -> scalar, like the reference's dict.  ``vectors`` additionally holds the
    same values as a few stacked tensors ``[(names, values [n])]`` when the heads were evaluated together, so that the
    trainer's weighted total (ref maskformer_model.py:225-231) is a handful of launches instead of one multiply and
    one add per entry: see ``SetCriterion.weighted_total``."""
    vectors = None


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
        self.joint_heads = True      # evaluate all prediction heads together when the inputs allow (_forward_heads)
        self.chunk_bytes = 96 << 20  # candidate points drawn / ranked per pass of the joint path (stays in the L2)
        self._weights = {}
        self.last_path = None        # "heads" | "sequential": which evaluation the last forward took

    def check_status(self):
        """Raises the reference's error for a failed assignment (scipy's ``ValueError: matrix contains invalid numeric
        entries`` / ``cost matrix is infeasible`` at matcher.py:153 -- NaN/-inf costs from diverged logits, or a label
        outside [0, num_classes)).  The device matcher reports failure through a status word instead of a host round
        trip per head, so the forward itself cannot raise: it poisons every loss of the step with NaN (nothing is
        silently trained on a made-up assignment) and this method -- one 4-byte device read; call it wherever the
        trainer already synchronises, e.g. next to the loss logging -- raises."""
        st, self._status = self._status, None
        if st is not None and int(st.item()) != 0:
            raise ValueError(f"HungarianMatcher: cost matrix of assignment problem {int(st.item()) - 1} (image, or "
                             "head * batch + image when the heads are solved together) contains invalid numeric "
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
            self.joint = {}
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
        use_dn = bool(self.training and dn_out)
        heads = [main] + list(outputs.get("aux_outputs", []))
        dn_heads = None
        if use_dn:
            dn_heads = [{k: v for k, v in dn_out.items() if k != "aux_outputs"}]
            dn_heads += list(dn_out["aux_outputs"])[:len(heads) - 1]
        self.last_path = "heads" if self.joint_heads and self._joint_ok(heads, dn_heads, step) else "sequential"
        if self.last_path == "heads":
            losses = self._forward_heads(heads, dn_heads, dn_out, targets, step, num_masks)
        else:
            losses = self._forward_sequential(heads, dn_heads, dn_out, targets, step, num_masks)
        if self.dn_no_lb:
            for k in [k for k in losses if k.startswith("loss_ce_dn")]:
                del losses[k]
        return losses

    def _forward_sequential(self, heads, dn_heads, dn_out, targets, step, num_masks):
        """Head by head, as the reference loops (criterion.py:271-304)."""
        dev = step.device
        use_dn = dn_heads is not None
        zero = torch.zeros((), device=dev)           # (a fill kernel: capturable, unlike a host scalar upload)
        if use_dn:
            db, dq, dt, scalar = step.dn_indices(dn_out["dn_args"])
            dn_idx = (db, dq, dt)
        losses = LossDict()
        for e, head in enumerate(heads):
            suffix = "" if e == 0 else f"_{e - 1}"
            d = self._all_losses(head, step, self._match(head, targets, step), num_masks)
            losses.update({k + suffix: v for k, v in d.items()})
            if use_dn:
                d = self._all_losses(dn_heads[e], step, dn_idx, num_masks * scalar)
                losses.update({k + "_dn" + suffix: v for k, v in d.items()})
            else:
                losses.update({k + suffix: zero for k in ("loss_mask_dn", "loss_dice_dn", "loss_ce_dn")})
        if self._status is not None:        # failed assignment -> every loss NaN (see check_status)
            poison = torch.where(self._status[0] == 0, 0.0, float("nan"))
            for k in losses:
                losses[k] = losses[k] + poison
        return losses

    # -- all heads together -----------------------------------------------------------------------------------------
    def _joint_ok(self, heads, dn_heads, step):
        """Whether ``_forward_heads`` covers these inputs; everything else takes the head-by-head path."""
        if sorted(self.losses) != ["labels", "masks"] or not step.packed.uniform() or step.total == 0:
            return False
        if int(self.num_points * self.oversample_ratio) > native.TOPK_GATHER_MAX_N:
            return False
        if int(self.importance_sample_ratio * self.num_points) <= 0:
            return False
        m0, l0 = heads[0]["pred_masks"], heads[0]["pred_logits"]
        if m0.dim() != 4 or m0.shape[1] == 0:
            return False
        H, W = m0.shape[-2:]
        for group in ([heads, dn_heads] if dn_heads else [heads]):
            g0 = group[0]
            for o in group:
                m, lg = o["pred_masks"], o["pred_logits"]
                if m.dtype != torch.float32 or lg.dtype != l0.dtype or m.shape != g0["pred_masks"].shape or \
                        lg.shape != g0["pred_logits"].shape or m.stride(-1) != 1 or m.stride(-2) != W or \
                        m.stride(1) != H * W or tuple(m.shape[-2:]) != (H, W) or m.shape[0] != m0.shape[0]:
                    return False
        return dn_heads is None or len(dn_heads) == len(heads)

    def _forward_heads(self, heads, dn_heads, dn_out, targets, step, num_masks):
        """All loss evaluations of the step in one pass (see the module docstring).  Row layout: for head e = 0 (final),
        1 .. (auxiliary 0 ..): its R_m matched pairs, then its R_d mask-piloted pairs -- the order in which the
        reference draws their random points (criterion.py:271-304), so a seed gives the same points.  Column layout
        of the concatenated class logits / mask-logit gradients: per head its Q matching queries, then its ``pad``
        mask-piloted queries."""
        dev = step.device
        nE = len(heads)
        use_dn = dn_heads is not None
        m0 = heads[0]["pred_masks"]
        B, Q, H, W = m0.shape
        sizes_m = [min(Q, n) for n in step.counts]
        R_m = sum(sizes_m)
        scalar, R_d, pad = 1, 0, 0
        if use_dn:
            db, dq, dt, scalar = step.dn_indices(dn_out["dn_args"])
            R_d, pad = int(db.numel()), dn_heads[0]["pred_masks"].shape[1]
        per, width = R_m + R_d, Q + pad
        total = nE * width
        n_over = int(self.num_points * self.oversample_ratio)
        n_unc = int(self.importance_sample_ratio * self.num_points)
        n_rand = self.num_points - n_unc
        P = self.num_points
        device_matcher = hasattr(self.matcher, "match_device_heads")

        # per-step tables that depend on the target counts and the head geometry only
        key = ("heads", nE, Q, pad, scalar, R_d)
        tab = step.joint.get(key)
        if tab is None:
            bm = step.batch_index(sizes_m)
            b_per = torch.cat([bm, db]) if use_dn else bm
            e_rows = torch.arange(nE, device=dev).repeat_interleave(per)
            first = e_rows * width
            if use_dn:       # the mask-piloted rows of every head start Q columns further
                first = first + torch.cat([torch.zeros(R_m, dtype=torch.int64, device=dev),
                                           torch.full((R_d,), Q, dtype=torch.int64, device=dev)]).repeat(nE)
            b_all = b_per.repeat(nE)
            tab = step.joint[key] = {"b_all": b_all, "slot0": b_all * total + first}
        b_all = tab["b_all"]

        maps, logit_list = [], []
        for e in range(nE):
            maps.append(heads[e]["pred_masks"])
            logit_list.append(heads[e]["pred_logits"])
            if use_dn:
                maps.append(dn_heads[e]["pred_masks"])
                logit_list.append(dn_heads[e]["pred_logits"])
        s0_set = {m.stride(0) for m in maps}
        s1 = H * W

        R = nE * per
        q_all = torch.empty(R, dtype=torch.int64, device=dev)
        t_all = torch.empty(R, dtype=torch.int64, device=dev)
        base = torch.empty(R, dtype=torch.int64, device=dev)
        s0_rows = torch.empty(R, dtype=torch.int64, device=dev) if len(s0_set) > 1 else None
        coords = torch.empty((R, P, 2), dtype=torch.float32, device=dev)
        labels = torch.empty((R, P), dtype=torch.float32, device=dev)
        q2, t2, base2 = q_all.view(nE, per), t_all.view(nE, per), base.view(nE, per)
        if use_dn:
            q2[:, R_m:] = dq
            t2[:, R_m:] = dt
        for e in range(nE):             # device addresses of the heads' maps: fills (kernel arguments; capturable)
            base2[e, :R_m].fill_(maps[(2 if use_dn else 1) * e].data_ptr())
            if use_dn:
                base2[e, R_m:].fill_(maps[2 * e + 1].data_ptr())
            if s0_rows is not None:
                s0_rows.view(nE, per)[e, :R_m].fill_(maps[(2 if use_dn else 1) * e].stride(0))
                if use_dn:
                    s0_rows.view(nE, per)[e, R_m:].fill_(maps[2 * e + 1].stride(0))
        s0 = s0_rows if s0_rows is not None else next(iter(s0_set))

        heads_per_pass = max(1, min(nE, self.chunk_bytes // max(1, per * n_over * 12)))
        with torch.no_grad():
            for e0 in range(0, nE, heads_per_pass):
                e1 = min(nE, e0 + heads_per_pass)
                n = e1 - e0
                r0, r1 = e0 * per, e1 * per
                cand = torch.empty((n * per, n_over, 2), dtype=torch.float32, device=dev)
                rnd = torch.empty((n * per, n_rand, 2), dtype=torch.float32, device=dev) if n_rand > 0 else None
                pts, generic = [], []
                # 1. the random numbers of these heads, in the reference's order (matcher.py:120; criterion.py:162,
                #    get_uncertain_point_coords_with_randomness: candidates, then the purely random points)
                for e in range(e0, e1):
                    if device_matcher:
                        pts.append(self.matcher.draw_points(B, dev))
                    else:
                        generic.append(self._match(heads[e], targets, step))
                    lo = (e - e0) * per
                    for a, b in ((lo, lo + R_m), (lo + R_m, lo + per)):
                        if b > a:
                            torch.rand(b - a, n_over, 2, out=cand[a:b])
                            if rnd is not None:
                                torch.rand(b - a, n_rand, 2, out=rnd[a:b])
                # 2. assignment of these heads' matching queries
                if device_matcher:
                    qm, tm, status = self.matcher.match_device_heads(heads[e0:e1], targets, pts)
                    self._status = status if self._status is None else torch.maximum(self._status, status)
                    q2[e0:e1, :R_m] = qm.view(n, R_m)
                    t2[e0:e1, :R_m] = tm.view(n, R_m)
                else:
                    for e, (_, qe, te) in zip(range(e0, e1), generic):
                        q2[e, :R_m] = qe
                        t2[e, :R_m] = te
                # 3. importance sampling of the points: the n_unc most uncertain of the candidates + random ones
                qs, ts, bs_ = q_all[r0:r1], t_all[r0:r1], b_all[r0:r1]
                src = base[r0:r1] + 4 * (bs_ * (s0 if s0_rows is None else s0[r0:r1]) + qs * s1)
                unc = native.point_sample_rows(src, True, (H, W), cand, neg_abs=True)
                top = native.topk_gather_rows(unc, cand, n_unc)
                if rnd is not None:
                    torch.cat([top, rnd], dim=1, out=coords[r0:r1])
                else:
                    coords[r0:r1] = top
                # 4. the targets at those points
                tgt_ptrs = step.mask_ptrs[bs_] + ts * (step.hw[0] * step.hw[1] * step.elsize)
                native.point_sample_rows(tgt_ptrs, step.packed.is_f32, step.hw, coords[r0:r1], out=labels[r0:r1])
            src_all = base + 4 * (b_all * s0 + q_all * s1)
            slot = tab["slot0"] + q_all
            tgt_classes = step.packed.labels[step.offsets[b_all] + t_all].clamp(0, self.num_classes)

        # mask losses of all rows
        x = native.PointSampleViews.apply(coords, src_all, slot, *maps)
        bce, dice = native.MaskLossRows.apply(x, labels)
        bce, dice = bce.view(nE, per), dice.view(nE, per)
        # classification loss of all heads (ref criterion.py:123-139): weighted mean per head = sum(w nll) / sum(w)
        logits = torch.cat(logit_list, dim=1).float()                     # [B, total, K + 1]
        classes = torch.full((B * total,), self.num_classes, dtype=torch.int64, device=dev)
        classes[slot] = tgt_classes
        nll = F.cross_entropy(logits.view(B * total, -1), classes, self.empty_weight, reduction="none")
        nll = nll.view(B, nE, width)
        wts = self.empty_weight[classes].view(B, nE, width)

        vec = {"loss_ce": nll[:, :, :Q].sum((0, 2)) / wts[:, :, :Q].sum((0, 2)),
               "loss_mask": bce[:, :R_m].sum(1) / num_masks, "loss_dice": dice[:, :R_m].sum(1) / num_masks}
        if use_dn:
            vec["loss_ce_dn"] = nll[:, :, Q:].sum((0, 2)) / wts[:, :, Q:].sum((0, 2))
            vec["loss_mask_dn"] = bce[:, R_m:].sum(1) / (num_masks * scalar)
            vec["loss_dice_dn"] = dice[:, R_m:].sum(1) / (num_masks * scalar)
        else:
            zeros = torch.zeros(nE, device=dev)
            vec.update({k: zeros for k in ("loss_ce_dn", "loss_mask_dn", "loss_dice_dn")})
        if self._status is not None:        # failed assignment -> every loss NaN (see check_status)
            poison = torch.where(self._status[0] == 0, 0.0, float("nan"))
            vec = {k: v + poison for k, v in vec.items()}
        losses = LossDict()
        losses.vectors = []
        suffixes = [""] + [f"_{i}" for i in range(nE - 1)]
        for k, v in vec.items():
            if self.dn_no_lb and k == "loss_ce_dn":
                continue
            names = [k + sfx for sfx in suffixes]
            losses.vectors.append((names, v))
            losses.update(zip(names, v.unbind(0)))
        return losses

    def weighted_total(self, losses, weight_dict=None):
        """sum_k weight_dict[k] * losses[k] over the losses that have a weight: the scalar the trainer back-propagates
        (ref maskformer_model.py:225-231 followed by the trainer's ``sum(loss_dict.values())``).  With the stacked
        values of the joint path this is one multiply-and-sum per loss kind instead of one multiply and one add per
        entry (60 in the published recipe)."""
        wd = self.weight_dict if weight_dict is None else weight_dict
        vec = getattr(losses, "vectors", None)
        if not vec:
            return sum(v * wd[k] for k, v in losses.items() if k in wd)
        total = None
        for names, values in vec:
            wkey = tuple(float(wd.get(k, 0.0)) for k in names)
            w = self._weights.get((wkey, values.device))
            if w is None:       # built on first use (eagerly, before any graph capture) and kept
                w = self._weights[(wkey, values.device)] = torch.tensor(wkey, dtype=values.dtype, device=values.device)
            term = torch.dot(values, w)
            total = term if total is None else total + term
        return total

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
