"""2-D sine position embedding behind the reference's interface
(ref: transformer_decoder/position_encoding.py:13-64).

The embedding is separable: channel c of the first half depends on the row only, of the second half on the column
only.  For an unpadded map (``mask=None``, the only case on this path) the two 1-D tables ``[H, F]`` and ``[W, F]``
are evaluated with the reference's sequence of fp32 operations (count -> / (last + eps) * scale -> / dim_t -> sin |
cos, so the values are the reference's bit for bit) and broadcast into the ``[1, H, W, 2F]`` channels-last map the
kernels consume.  Maps are kept in a small LRU per module: fixed-crop training hits one key per level, variable-size
evaluation cannot grow the cache beyond ``CACHE_ENTRIES`` maps (the reference recomputes per call).  (A captured CUDA
graph reads the cached map it was captured with: keep the number of distinct captured geometries per module below
``CACHE_ENTRIES``, or raise it.)"""
import math
from collections import OrderedDict

import torch
from torch import nn


class PositionEmbeddingSine(nn.Module):
    CACHE_ENTRIES = 8

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.scale = 2 * math.pi if scale is None else scale
        self._cache = OrderedDict()

    def _waves(self, coord):
        """coord [...] (already normalised) -> [..., F]: sin on even channels, cos on odd ones, channel pairs sharing
        a wavelength (ref :43-52)."""
        F_ = self.num_pos_feats
        c = torch.arange(F_, dtype=torch.float32, device=coord.device)
        wavelength = self.temperature ** (2 * torch.div(c, 2, rounding_mode="floor") / F_)
        angle = coord[..., None] / wavelength
        even = (torch.arange(F_, device=coord.device) % 2 == 0)
        return torch.where(even, angle.sin(), angle.cos())

    def _normalised(self, count, last):
        return count / (last + 1e-6) * self.scale if self.normalize else count

    def _table(self, n, device):
        count = torch.arange(1, n + 1, dtype=torch.float32, device=device)       # cumsum of an all-valid line
        return self._waves(self._normalised(count, count[-1:]))                # [n, F]

    def channels_last(self, H, W, device):
        """[1, H, W, 2F] embedding of an all-valid H x W map (LRU-cached)."""
        key = (int(H), int(W), str(device))
        hit = self._cache.get(key)
        if hit is not None:
            self._cache.move_to_end(key)
            return hit
        rows, cols = self._table(H, device), self._table(W, device)
        F_ = self.num_pos_feats
        pos = torch.empty((1, H, W, 2 * F_), dtype=torch.float32, device=device)
        pos[0, :, :, :F_] = rows[:, None, :]
        pos[0, :, :, F_:] = cols[None, :, :]
        self._cache[key] = pos
        while len(self._cache) > self.CACHE_ENTRIES:
            self._cache.popitem(last=False)
        return pos

    def forward(self, x, mask=None):
        """Returns [B, 2F, H, W] like the reference."""
        if mask is None:
            pos = self.channels_last(x.size(2), x.size(3), x.device)
            return pos.permute(0, 3, 1, 2).expand(x.size(0), -1, -1, -1)
        valid = ~mask                                                          # padded batch: counts per image
        ys = valid.cumsum(1, dtype=torch.float32)
        xs = valid.cumsum(2, dtype=torch.float32)
        py = self._waves(self._normalised(ys, ys[:, -1:, :]))
        px = self._waves(self._normalised(xs, xs[:, :, -1:]))
        return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)

    def extra_repr(self):
        return (f"num_pos_feats={self.num_pos_feats}, temperature={self.temperature}, normalize={self.normalize}, "
                f"scale={self.scale}")
