"""2-D sine position embedding with the reference's interface
(ref: transformer_decoder/position_encoding.py:13-64).

With ``mask=None`` the embedding depends only on (H, W): it is built once per shape/device and
cached instead of being recomputed by six elementwise chains per call."""
import math

import torch
from torch import nn


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.scale = 2 * math.pi if scale is None else scale
        self._cache = {}

    def _build(self, not_mask, device):
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            eps = 1e-6
            y_embed = y_embed / (y_embed[:, -1:, :] + eps) * self.scale
            x_embed = x_embed / (x_embed[:, :, -1:] + eps) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).flatten(3)
        pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).flatten(3)
        return torch.cat((pos_y, pos_x), dim=3)          # [B, H, W, 2*num_pos_feats]  (channels last)

    def channels_last(self, H, W, device):
        """[1, H, W, C] embedding for an all-valid H x W map (cached)."""
        key = (H, W, str(device))
        if key not in self._cache:
            ones = torch.ones((1, H, W), dtype=torch.bool, device=device)
            self._cache[key] = self._build(ones, device)
        return self._cache[key]

    def forward(self, x, mask=None):
        """Returns [B, 2*num_pos_feats, H, W] like the reference."""
        if mask is None:
            pos = self.channels_last(x.size(2), x.size(3), x.device)
            return pos.permute(0, 3, 1, 2).expand(x.size(0), -1, -1, -1)
        return self._build(~mask, x.device).permute(0, 3, 1, 2)

    def __repr__(self, _repr_indent=4):
        head = "Positional encoding " + self.__class__.__name__
        body = [f"num_pos_feats: {self.num_pos_feats}", f"temperature: {self.temperature}",
                f"normalize: {self.normalize}", f"scale: {self.scale}"]
        return "\n".join([head] + [" " * _repr_indent + line for line in body])
