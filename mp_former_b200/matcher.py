"""Device-resident mirror of the reference's ``HungarianMatcher`` (mask2former/modeling/matcher.py:70-189): same
constructor, same ``forward(outputs, targets)`` contract and result format, same consumption of the global random
generator PER CALL (one ``torch.rand(1, num_points, 2)`` per image, in image order, matcher.py:120; end-to-end seed
parity with a reference training run does not hold, see criterion.py) -- but the whole batch is
matched in three kernel launches (native.match_cost + native.lsap) with ONE device->host copy of the finished index
pairs, instead of, per image, two grid_samples, three einsums, a cost-matrix copy to the host (a stream sync) and a
scipy solve.  With ``device_indices=True`` the pairs stay on the device and nothing synchronises at all (the number of
pairs per image, min(Q, n_b), is known on the host from the target shapes).

There is no CPU path: CUDA tensors only (the extension must be present; see mp_former_b200._lib)."""
import torch
from torch import nn

from . import _lib, native


def targets_key(targets):
    """Identity of a step's targets for the per-step caches: storage addresses, shapes, dtypes and the tensors'
    in-place modification counters (the caches hold the tensors, so an address cannot be recycled under them)."""
    return tuple((t["masks"].data_ptr(), tuple(t["masks"].shape), t["masks"].dtype, t["masks"]._version,
                  t["labels"].data_ptr(), t["labels"]._version) for t in targets)


class PackedTargets:
    """The per-step, head-independent half of the matcher's inputs: label vector, per-image offsets and the device
    table of mask pointers.  Built once per step and reused by the ten prediction heads' matchings."""

    def __init__(self, targets, device):
        masks = []
        for t in targets:
            m = t["masks"]
            _lib.require_cuda(m, "targets[i]['masks']")
            if m.dim() != 3:
                raise RuntimeError("targets[i]['masks'] must be [n, H, W]")
            if m.dtype == torch.bool:
                m = m.contiguous().view(torch.uint8)
            elif m.dtype == torch.uint8 or m.dtype == torch.float32:
                m = m.contiguous()
            else:                                   # the reference converts with `.to(out_mask)` (matcher.py:115)
                m = m.to(torch.float32).contiguous()
            masks.append(m)
        kinds = {m.dtype for m in masks}
        if len(kinds) > 1:
            masks = [m.to(torch.float32) for m in masks]
        self.masks = masks                          # keeps the storage behind the pointer table alive
        self._mask_refs = [t["masks"] for t in targets]
        self.is_f32 = bool(masks) and masks[0].dtype == torch.float32
        self.counts = [int(m.shape[0]) for m in masks]
        self.sizes = [tuple(m.shape[-2:]) for m in masks]
        self._label_refs = [t["labels"] for t in targets]      # pinned for the cache key's sake
        labels = [t["labels"].to(device=device, dtype=torch.int64) for t in targets]
        for lab, n in zip(labels, self.counts):
            if lab.numel() != n:
                raise RuntimeError("targets[i]['labels'] and targets[i]['masks'] disagree on the number of instances")
        self.labels = torch.cat(labels) if labels else torch.zeros(0, dtype=torch.int64, device=device)
        self.device = device

    def uniform(self):
        return len(set(self.sizes)) <= 1

    def tables(self, lo, hi):
        """(ptrs int64 [hi-lo], offsets int32 [hi-lo+1], labels, counts) of images lo..hi-1."""
        counts = self.counts[lo:hi]
        offs = [0]
        for n in counts:
            offs.append(offs[-1] + n)
        ptrs = torch.tensor([m.data_ptr() if m.shape[0] else 0 for m in self.masks[lo:hi]], dtype=torch.int64,
                            device=self.device)
        offsets = torch.tensor(offs, dtype=torch.int32, device=self.device)
        first = sum(self.counts[:lo])
        return ptrs, offsets, self.labels[first:first + offs[-1]], counts


class HungarianMatcher(nn.Module):
    """Assignment between the targets and the predictions of the network (ref matcher.py:70-189).

    ``forward`` returns ``[(index_i, index_j)] * batch`` -- int64 tensors with ``len == min(num_queries, n_b)``,
    ``index_i`` the selected predictions in ascending order, ``index_j`` the matched targets -- on the CPU like the
    reference (matcher.py:153-156), or on the device with ``device_indices=True``."""

    def __init__(self, cost_class: float = 1, cost_mask: float = 1, cost_dice: float = 1, num_points: int = 0,
                 device_indices: bool = False, sort_points=None):
        super().__init__()
        self.cost_class = cost_class
        self.cost_mask = cost_mask
        self.cost_dice = cost_dice
        assert cost_class != 0 or cost_mask != 0 or cost_dice != 0, "all costs cant be 0"
        self.num_points = num_points
        self.device_indices = device_indices
        # Visit the points of an image in row-major pixel order: the random gathers of the cost kernel then walk every
        # map once, row by row, instead of pulling a 32-byte DRAM sector per corner (the 419 MB of mask logits of a
        # head do not fit the L2).  Measured on the B200 with the round-2 cost kernel (profiles/r2f_matcher_probe*.json,
        # 16 images x 100 queries x 12544 points): 0.77 ms per head with the sort (argsort included) vs 1.10 ms
        # without; same assignments.  The cost sums do not depend on the order beyond fp32 rounding.  None (default):
        # sort when the maps of a head exceed the L2 (>= 8 images' worth of 100 x 256 x 256 logits); for a couple of
        # images the argsort costs more than the gathers it tidies.
        self.sort_points = sort_points
        # With sorted points: sample the predictions by streaming every map once through shared memory
        # (native.sample_shared_points) instead of gathering.  Measured on the B200 at 16 images x 100 queries x 12544
        # points, cold L2 (profiles/r2q_matcher_probe_*.txt, r2r_*): the streamed sampler takes 0.18 ms and the cost
        # kernel on its samples 0.30 ms, against 0.41 ms for the cost kernel gathering at the sorted points itself
        # (1.03 ms unsorted) -- what remains of that kernel is the sampling of the GT masks and its reduction over the
        # points, not the prediction gathers.  Off by default for that reason; kept as a tested option.
        self.stream_samples = False
        self._packed_key = None
        self._packed = None
        self._tables = {}
        self.last_status = None

    # -- inputs shared by the heads of one step -------------------------------------------------------------------
    def pack_targets(self, targets, device):
        key = targets_key(targets)
        if key != self._packed_key:
            self._packed = PackedTargets(targets, device)
            self._packed_key = key
            self._tables = {}
        return self._packed

    def _tables_for(self, packed, lo, hi):
        if (lo, hi) not in self._tables:
            self._tables[(lo, hi)] = packed.tables(lo, hi)
        return self._tables[(lo, hi)]

    @staticmethod
    def row_major_order(point_coords, H, W, band_rows=None):
        """point_coords [B, P, 2] (x, y) in [0, 1] -> the same points of every image, ordered row-major by the
        top-left pixel of their bilinear footprint on the H x W map (``floor(c * size - 0.5)``): the 32 points of a
        warp then touch 6.0 cache lines per gather instruction on average instead of 31.8 (12544 uniform points on a
        256 x 256 fp32 map; ordering by the pixel containing the point gives 11.1, tiled orders 6.1-6.5).
        With ``band_rows``: also returns int32 [B, n_bands + 1], the index of the first point of every band of
        ``band_rows`` map rows (for ``native.sample_shared_points``)."""
        key = ((point_coords[..., 1] * H - 0.5).floor().clamp(0, H - 1) * W +
               (point_coords[..., 0] * W - 0.5).floor().clamp(0, W - 1))
        if band_rows is None:
            order = key.argsort(dim=1, stable=True)
            return torch.gather(point_coords, 1, order.unsqueeze(-1).expand(-1, -1, 2))
        skey, order = key.sort(dim=1, stable=True)
        n_bands = -(-H // band_rows)
        edges = (torch.arange(n_bands + 1, device=key.device, dtype=key.dtype) * float(band_rows * W))
        band_lo = torch.searchsorted(skey, edges.expand(key.shape[0], -1).contiguous()).to(torch.int32)
        return torch.gather(point_coords, 1, order.unsqueeze(-1).expand(-1, -1, 2)), band_lo

    def _head_cost(self, logits, masks, pts, ptrs, packed, hw, labels, offsets, counts):
        """Cost matrices of one prediction head over a group of images (flat, see native.match_cost)."""
        logits = logits.float()                     # autocast heads: the costs are computed in fp32 (matcher.py:134-136)
        masks = masks.float()
        sort = self.sort_points
        if sort is None:
            sort = masks.numel() * masks.element_size() >= (200 << 20)
        sampled = None
        if sort:
            H, W = masks.shape[-2:]
            if self.stream_samples and W % 4 == 0 and masks.stride(-1) == 1 and masks.stride(-2) == W and \
                    masks.stride(0) % 4 == 0 and masks.stride(1) % 4 == 0 and masks.data_ptr() % 16 == 0:
                # every map is streamed once through shared memory instead of being gathered from (see __init__)
                rows, _ = native.shared_point_bands(H, W)
                pts, band_lo = self.row_major_order(pts, H, W, band_rows=rows)
                sampled = native.sample_shared_points(masks, pts, band_lo, rows)
            else:
                pts = self.row_major_order(pts, H, W)
        return native.match_cost(logits, masks, ptrs, packed.is_f32, hw, labels, offsets, counts, pts,
                                 self.cost_class, self.cost_mask, self.cost_dice, sampled=sampled)

    def draw_points(self, bs, device):
        """The matcher's random points of one call: all masks of an image share one set, drawn per image like the
        reference (``torch.rand(1, num_points, 2)`` inside the loop over the batch, matcher.py:120)."""
        return torch.cat([torch.rand(1, self.num_points, 2, device=device) for _ in range(bs)])

    # -- the matching ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def match_device_heads(self, heads, targets, point_coords):
        """``match_device`` for SEVERAL prediction heads of one step (``heads``: list of outputs dicts over the same
        batch and targets; ``point_coords``: their [B, P, 2] point sets, e.g. from ``draw_points``): the cost matrices
        of all heads are solved by ONE launch of the assignment kernel (one CTA per (head, image) problem instead of
        one launch of B CTAs per head).  Target maps of one size only (``PackedTargets.uniform()``).
        Returns (query_idx, target_idx, status): the pairs of head 0's images, then head 1's, ...; status: 0, or
        1 + (head * B + image) of a failed problem."""
        bs, num_queries = heads[0]["pred_logits"].shape[:2]
        if len(targets) != bs:
            raise RuntimeError(f"{len(targets)} targets for a batch of {bs}")
        dev = heads[0]["pred_masks"].device
        packed = self.pack_targets(targets, dev)
        if not packed.uniform():
            raise RuntimeError("match_device_heads: target maps of one size expected")
        ptrs, offsets, labels, counts = self._tables_for(packed, 0, bs)
        ntot = sum(counts)
        nh = len(heads)
        if ntot == 0:
            z = torch.zeros(0, dtype=torch.int64, device=dev)
            return z, z, torch.zeros(1, dtype=torch.int32, device=dev)
        costs = []
        for o, pts in zip(heads, point_coords):
            logits, masks = o["pred_logits"], o["pred_masks"]
            _lib.require_cuda(masks, "outputs['pred_masks']")
            _lib.require_cuda(logits, "outputs['pred_logits']")
            costs.append(self._head_cost(logits, masks, pts, ptrs, packed, packed.sizes[0], labels, offsets, counts))
        key = ("heads", nh)
        if key not in self._tables:     # offsets of the nh * B problems in the concatenated cost buffer
            offs, acc = [0], 0
            for _ in range(nh):
                for n in counts:
                    acc += n
                    offs.append(acc)
            self._tables[key] = torch.tensor(offs, dtype=torch.int32, device=dev)
        q, t, status = native.lsap(costs[0] if nh == 1 else torch.cat(costs), self._tables[key], list(counts) * nh,
                                   num_queries)
        self.last_status = status
        return q, t, status

    @torch.no_grad()
    def match_device(self, outputs, targets, point_coords=None):
        """Returns (query_idx, target_idx, counts, cost, status): flat int64 device tensors holding the pairs of all
        images back to back (min(Q, n_b) each), the host list n_b, the flat cost matrices and the solver status
        (int32 [1]: 0, or 1 + the index of an image whose cost matrix holds NaN / -inf or is infeasible -- where scipy
        raises ValueError in the reference; that image's pairs are then the in-range placeholders (k, k)).
        ``point_coords`` [B, P, 2] overrides the random points (tests)."""
        logits, masks = outputs["pred_logits"], outputs["pred_masks"]
        _lib.require_cuda(masks, "outputs['pred_masks']")
        _lib.require_cuda(logits, "outputs['pred_logits']")
        bs, num_queries = logits.shape[:2]
        if len(targets) != bs:
            raise RuntimeError(f"{len(targets)} targets for a batch of {bs}")
        dev = masks.device
        packed = self.pack_targets(targets, dev)
        if point_coords is None:
            # all masks of an image share one set of points; drawn per image like the reference (matcher.py:120)
            point_coords = self.draw_points(bs, dev)
        groups = [(0, bs)] if packed.uniform() else [(b, b + 1) for b in range(bs)]
        qi, ti, costs, stats = [], [], [], []
        for lo, hi in groups:
            ptrs, offsets, labels, counts = self._tables_for(packed, lo, hi)
            if sum(counts) == 0:
                continue
            hw = packed.sizes[lo]
            cost = self._head_cost(logits[lo:hi], masks[lo:hi], point_coords[lo:hi], ptrs, packed, hw, labels, offsets,
                                   counts)
            q, t, status = native.lsap(cost, offsets, counts, num_queries)
            qi.append(q), ti.append(t), costs.append(cost), stats.append(status)
        if not qi:
            z = torch.zeros(0, dtype=torch.int64, device=dev)
            return z, z, packed.counts, torch.zeros(0, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
        cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs))
        status = stats[0] if len(stats) == 1 else torch.stack(stats).max(0).values
        self.last_status = status
        return cat(qi), cat(ti), packed.counts, cat(costs), status

    @torch.no_grad()
    def memory_efficient_forward(self, outputs, targets):
        num_queries = outputs["pred_logits"].shape[1]
        q, t, counts, _, status = self.match_device(outputs, targets)
        sizes = [min(num_queries, n) for n in counts]
        if self.device_indices:
            return list(zip(torch.split(q, sizes), torch.split(t, sizes)))
        # ONE device->host copy for the batch (the reference: one per image, plus the solve on the host)
        host = torch.cat([q, t, status.to(torch.int64)]).cpu()
        m = q.numel()
        if int(host[-1]) != 0:
            raise ValueError(f"matrix of image {int(host[-1]) - 1} contains invalid numeric entries or is infeasible")
        return list(zip(torch.split(host[:m], sizes), torch.split(host[m:2 * m], sizes)))

    @torch.no_grad()
    def forward(self, outputs, targets):
        """outputs: {"pred_logits": [B, Q, K+1], "pred_masks": [B, Q, H, W]}; targets: list of B dicts with "labels"
        [n_b] and "masks" [n_b, H_gt, W_gt] (ref matcher.py:159-179)."""
        return self.memory_efficient_forward(outputs, targets)

    def __repr__(self, _repr_indent=4):
        head = "Matcher " + self.__class__.__name__
        body = [
            "cost_class: {}".format(self.cost_class),
            "cost_mask: {}".format(self.cost_mask),
            "cost_dice: {}".format(self.cost_dice),
        ]
        lines = [head] + [" " * _repr_indent + line for line in body]
        return "\n".join(lines)
