"""Debug aid: masked cross-attention with / without the key split against the fp64 restatement, per tensor."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from mp_former_b200 import native, ops  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402

DEV = "cuda:0"


def ref_xattn(q_in, memory, pos, w_in, b_in, w_out, b_out, nhead, mask):
    sd = {"in_proj_weight": w_in, "in_proj_bias": b_in, "out_proj.weight": w_out, "out_proj.bias": b_out}
    m = mask & ~mask.all(-1, keepdim=True)
    am = m[:, None].expand(-1, nhead, -1, -1).flatten(0, 1)
    y = O.mha(sd, "", q_in.transpose(0, 1), (memory + pos).transpose(0, 1), memory.transpose(0, 1), nhead, am)
    return y.transpose(0, 1)


def main():
    E, nhead = 256, 8
    for B, Qt, HW, splits in ((2, 9, 16384, 9), (2, 9, 4096, 9), (2, 120, 16384, 9), (1, 9, 16384, 3)):
        g = torch.Generator(device=DEV).manual_seed(B * 77 + Qt + HW + splits)
        rn = lambda *s, sc=1.0: torch.randn(*s, device=DEV, generator=g) * sc  # noqa: E731
        q_in, memory, pos = rn(B, Qt, E), rn(B, HW, E), rn(1, HW, E)
        w_in, b_in, w_out, b_out = rn(3 * E, E, sc=1 / 16), rn(3 * E, sc=0.1), rn(E, E, sc=1 / 16), rn(E, sc=0.1)
        mask = torch.rand(B, Qt, HW, device=DEV, generator=g) < 0.9
        mask[0, 0] = True
        mask[0, 1] = True
        mask[0, 1, HW - 3] = False
        mask[0, 2] = True
        mask[0, 2, :64] = False
        gy = rn(B, Qt, E)
        names = ("y", "q_in", "memory", "w_in", "b_in", "w_out", "b_out")
        ref = [t.double().clone().requires_grad_(True) for t in (q_in, memory, w_in, b_in, w_out, b_out)]
        yr = ref_xattn(ref[0], ref[1], pos.double(), ref[2], ref[3], ref[4], ref[5], nhead, mask)
        yr.backward(gy.double())
        refs = [yr.detach()] + [t.grad for t in ref]
        for ks in (1, splits):
            native.XATTN_KEY_SPLITS = ks
            leaves = [t.clone().requires_grad_(True) for t in (q_in, memory, w_in, b_in, w_out, b_out)]
            y = ops.masked_cross_attention(leaves[0], leaves[1], pos, leaves[2], leaves[3], leaves[4], leaves[5], nhead,
                                           ops.PackedMask.from_bool(mask))
            y.backward(gy)
            got = [y.detach()] + [t.grad for t in leaves]
            errs = {n: float((a.double() - b).abs().max() / b.abs().max().clamp(min=1e-9)) for n, a, b in zip(names, got, refs)}
            print(f"B={B} Qt={Qt} HW={HW} key_splits={ks}:", {k: f"{v:.2e}" for k, v in errs.items()}, flush=True)
            if errs["y"] > 1e-3:
                d = (y.detach().double() - yr.detach()).abs().amax(-1)
                print("   rows with large forward error (b, q):", (d > 1e-3).nonzero().tolist()[:20], flush=True)


if __name__ == "__main__":
    main()
