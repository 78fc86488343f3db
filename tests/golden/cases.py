"""Seeded input builders shared by ``make_golden.py`` (which runs the UNMODIFIED reference on them,
in the authoring container only) and by the tests (which run the oracle / the CUDA path on the
same inputs anywhere).  Pure CPU torch; deterministic given the seeds."""
import torch

# ---- MSDeformAttn core cases -------------------------------------------------------------------
MSDA_CASES = {
    # reference ops/test.py:24-39 geometry and distributions, torch.manual_seed(3)
    "testpy": dict(N=1, M=2, D=2, Lq=2, L=2, P=2, shapes=[(6, 4), (3, 2)], seed=3, lo=0.0, hi=1.0),
    # model-like geometry (M=8, D=32, L=3, P=4), queries = pixels, locations spill over the border
    "model_small": dict(N=1, M=8, D=32, Lq=None, L=3, P=4, shapes=[(4, 4), (8, 8), (16, 16)], seed=11,
                        lo=-0.15, hi=1.15),
    # BASELINE config-1 style: L=4, Lq != S
    "l4": dict(N=1, M=8, D=32, Lq=50, L=4, P=4, shapes=[(16, 16), (8, 8), (4, 4), (2, 2)], seed=12,
               lo=0.0, hi=1.0),
    # odd channel count / points (generic kernel path), non-square levels
    "odd": dict(N=2, M=3, D=5, Lq=7, L=2, P=3, shapes=[(5, 3), (2, 7)], seed=13, lo=-0.3, hi=1.3),
}


def msda_inputs(name, dtype=torch.float32):
    c = MSDA_CASES[name]
    g = torch.Generator().manual_seed(c["seed"])
    S = sum(h * w for h, w in c["shapes"])
    Lq = c["Lq"] or S
    value = torch.rand(c["N"], S, c["M"], c["D"], generator=g) * 0.01
    loc = torch.rand(c["N"], Lq, c["M"], c["L"], c["P"], 2, generator=g) * (c["hi"] - c["lo"]) + c["lo"]
    aw = torch.rand(c["N"], Lq, c["M"], c["L"], c["P"], generator=g) + 1e-5
    aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    if name == "model_small":
        # a few samples exactly on texel centres / exactly on the -1 and H borders
        loc[0, 0, 0, 0, 0] = torch.tensor([0.5 / 4, 0.5 / 4])
        loc[0, 0, 0, 0, 1] = torch.tensor([-0.5 / 4, 0.5])        # w_im == -1 -> excluded
        loc[0, 0, 0, 0, 2] = torch.tensor([1.0 + 0.5 / 4, 0.5])   # w_im == W  -> excluded
        loc[0, 0, 0, 0, 3] = torch.tensor([1.0, 1.0])
    return value.to(dtype), c["shapes"], loc.to(dtype), aw.to(dtype)


# ---- module-level cases ------------------------------------------------------------------------
PD_CFG = dict(conv_dim=256, mask_dim=256, nheads=8, dim_feedforward=128, enc_layers=2,
              channels={"res2": 32, "res3": 48, "res4": 64, "res5": 96},
              strides={"res2": 4, "res3": 8, "res4": 16, "res5": 32})
DEC_CFG = dict(hidden_dim=256, num_queries=10, nheads=8, dim_feedforward=128, dec_layers=3,
               num_classes=7, mask_dim=256)


def pixel_decoder_features(B=2, base=64, seed=21):
    g = torch.Generator().manual_seed(seed)
    feats = {}
    for k in ("res2", "res3", "res4", "res5"):
        s = PD_CFG["strides"][k]
        feats[k] = torch.randn(B, PD_CFG["channels"][k], base // s, base // s, generator=g)
    return feats


def decoder_inputs(B=2, seed=22):
    g = torch.Generator().manual_seed(seed)
    C = DEC_CFG["hidden_dim"]
    x = [torch.randn(B, C, s, s, generator=g) for s in (2, 4, 8)]
    mask_features = torch.randn(B, DEC_CFG["mask_dim"], 16, 16, generator=g) * 0.5
    return x, mask_features


def dn_targets(B=2, seed=23, size=64):
    """Synthetic GT: rectangles as masks, labels, dummy boxes (only len() is used by the reference,
    ref mask2former_transformer_decoder.py:970-971)."""
    g = torch.Generator().manual_seed(seed)
    counts = [3, 1] if B == 2 else [int(torch.randint(1, 4, (1,), generator=g)) for _ in range(B)]
    tg = []
    for n in counts:
        masks = torch.zeros(n, size, size, dtype=torch.bool)
        for i in range(n):
            y0, x0 = [int(v) for v in torch.randint(0, size // 2, (2,), generator=g)]
            h, w = [int(v) for v in torch.randint(4, size // 2, (2,), generator=g)]
            masks[i, y0:y0 + h, x0:x0 + w] = True
        labels = torch.randint(0, DEC_CFG["num_classes"], (n,), generator=g)
        tg.append({"labels": labels, "masks": masks, "boxes": torch.zeros(n, 4)})
    return tg


def threshold_logits(seed=24):
    """Mask logits [1, 4, 16, 16] with values engineered around the sigmoid(x) < 0.5 boundary."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 4, 16, 16, generator=g)
    x[0, 0] *= 1e-7
    x[0, 1] *= 1e-4
    specials = torch.tensor([0.0, -0.0, 1e-8, -1e-8, 5.9e-8, -5.9e-8, 6.1e-8, -6.1e-8, 1.2e-7, -1.2e-7,
                             2.4e-7, -2.4e-7, 1e-38, -1e-38, 3e-7, -3e-7])
    x[0, 2, 0, :] = specials
    x[0, 2, 1, :] = specials
    return x
