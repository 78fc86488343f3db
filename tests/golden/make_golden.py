"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py) on the seeded inputs of cases.py.

Run in the authoring container only:   python tests/golden/make_golden.py
The fixtures (*.pt, fp32/fp64 tensors, a few MB in total) are committed; tests never need the
reference tree.  Parameters are NOT stored: they are regenerated from seeds with
oracle.torch_oracle.seeded_state_dict on both sides.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle.ref_loader import cuda_is_identity, load_reference  # noqa: E402
from oracle.torch_oracle import seeded_state_dict  # noqa: E402


def checksum(*tensors):
    return float(sum(t.double().abs().sum() for t in tensors))


def main():
    torch.set_num_threads(8)
    R = load_reference()
    meta = {"torch": torch.__version__}

    # ---- MSDeformAttn core: reference ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:52-72)
    out = {}
    for name in cases.MSDA_CASES:
        for dtype in (torch.float32, torch.float64):
            if dtype == torch.float64 and name not in ("testpy", "odd"):
                continue
            value, shapes, loc, aw = cases.msda_inputs(name, dtype)
            value.requires_grad_(True); loc.requires_grad_(True); aw.requires_grad_(True)
            y = R.ms_deform_attn_core_pytorch(value, shapes, loc, aw)
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(99), dtype=torch.float32).to(dtype)
            gv, gl, ga = torch.autograd.grad(y, (value, loc, aw), gy)
            key = f"{name}_{'f32' if dtype == torch.float32 else 'f64'}"
            out[key] = {"out": y.detach(), "grad_value": gv, "grad_loc": gl, "grad_aw": ga,
                        "in_checksum": checksum(value.detach(), loc.detach(), aw.detach())}
    torch.save(out, os.path.join(HERE, "msda_core.pt"))

    # ---- MSDeformAttn module (ops/modules/ms_deform_attn.py:82-125), CPU path of the reference
    m = R.MSDeformAttn(d_model=256, n_levels=3, n_heads=8, n_points=4).eval()
    sd = seeded_state_dict(m.state_dict(), seed=31)
    m.load_state_dict(sd)
    value, shapes, _, _ = cases.msda_inputs("model_small")
    g = torch.Generator().manual_seed(32)
    S = value.shape[1]
    query = torch.randn(1, S, 256, generator=g)
    src = torch.randn(1, S, 256, generator=g)
    ref_pts = torch.rand(1, S, 3, 2, generator=g)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    with torch.no_grad():
        y = m(query, ref_pts, src, shapes_t, lsi, None)
    torch.save({"out": y, "in_checksum": checksum(query, src, ref_pts)}, os.path.join(HERE, "msda_module.pt"))

    # ---- pixel decoder (pixel_decoder/msdeformattn.py:314-358)
    cfg = cases.PD_CFG
    shape = {k: R.ShapeSpec(channels=cfg["channels"][k], stride=cfg["strides"][k]) for k in cfg["channels"]}
    pd = R.MSDeformAttnPixelDecoder(
        shape, transformer_dropout=0.0, transformer_nheads=cfg["nheads"],
        transformer_dim_feedforward=cfg["dim_feedforward"], transformer_enc_layers=cfg["enc_layers"],
        conv_dim=cfg["conv_dim"], mask_dim=cfg["mask_dim"], norm="GN",
        transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval()
    sd = seeded_state_dict(pd.state_dict(), seed=41)
    pd.load_state_dict(sd)
    feats = cases.pixel_decoder_features()
    with torch.no_grad():
        mf, enc0, ms = pd.forward_features(feats)
    torch.save({"mask_features": mf, "enc0": enc0, "multi_scale": ms, "keys": sorted(sd.keys()),
                "in_checksum": checksum(*feats.values())}, os.path.join(HERE, "pixel_decoder.pt"))

    # ---- transformer decoder (transformer_decoder/mask2former_transformer_decoder.py:1706-1857)
    dc = cases.DEC_CFG
    common = dict(num_classes=dc["num_classes"], hidden_dim=dc["hidden_dim"], num_queries=dc["num_queries"],
                  nheads=dc["nheads"], dim_feedforward=dc["dim_feedforward"], dec_layers=dc["dec_layers"],
                  pre_norm=False, mask_dim=dc["mask_dim"], enforce_input_project=False)
    dec = R.MultiScaleMaskedTransformerDecoderMaskDN(dc["hidden_dim"], True, **common, dn_mode="points",
                                                     all_lys=True, dn_label_noise_ratio=-1.0).eval()
    sd = seeded_state_dict(dec.state_dict(), seed=51)
    dec.load_state_dict(sd)
    x, mask_features = cases.decoder_inputs()

    def pack(o):
        r = {"pred_logits": o["pred_logits"], "pred_masks": o["pred_masks"],
             "aux_logits": [a["pred_logits"] for a in o["aux_outputs"]],
             "aux_masks": [a["pred_masks"] for a in o["aux_outputs"]]}
        if o.get("dn_out") is not None:
            d = o["dn_out"]
            r["dn"] = {"pred_logits": d["pred_logits"], "pred_masks": d["pred_masks"],
                       "aux_logits": [a["pred_logits"] for a in d["aux_outputs"]],
                       "aux_masks": [a["pred_masks"] for a in d["aux_outputs"]],
                       "dn_args": d["dn_args"]}
        return r

    with torch.no_grad():
        o_plain = dec(x, mask_features, None, None)
        with cuda_is_identity():
            o_dn = dec(x, mask_features, None,
                       {"tgt": cases.dn_targets(), "scalar": 1, "noise_scale": 0.0})
            o_dn2 = dec(x, mask_features, None,
                        {"tgt": cases.dn_targets(), "scalar": 2, "noise_scale": 0.0})
    base = R.MultiScaleMaskedTransformerDecoder(dc["hidden_dim"], True, **common).eval()
    base.load_state_dict({k: v for k, v in sd.items() if not k.startswith("label_enc")})
    with torch.no_grad():
        o_base = base(x, mask_features, None)
    torch.save({"plain": pack(o_plain), "dn": pack(o_dn), "dn2": pack(o_dn2), "base": pack(o_base),
                "keys": sorted(sd.keys()), "in_checksum": checksum(*x, mask_features)},
               os.path.join(HERE, "decoder.pt"))

    # ---- boolean stage of the heads on engineered logits -- bit-exact target.  These are the same three
    # torch calls the reference makes at decoder :1869-1875 (interpolate, sigmoid, < 0.5, head repeat);
    # they are restated here because the method offers no way to inject logits.
    logits = cases.threshold_logits()
    hm = {}
    for size in ((2, 2), (4, 4), (8, 8), (16, 16)):
        a = torch.nn.functional.interpolate(logits, size=size, mode="bilinear", align_corners=False)
        a = (a.sigmoid().flatten(2).unsqueeze(1).repeat(1, 8, 1, 1).flatten(0, 1) < 0.5).bool()
        hm[f"{size[0]}x{size[1]}"] = a
    torch.save({"masks": hm, "in_checksum": checksum(logits)}, os.path.join(HERE, "attn_mask_bits.pt"))

    # ---- position embedding (position_encoding.py:29-52)
    pe = R.PositionEmbeddingSine(128, normalize=True)
    torch.save({"pos_3x5": pe(torch.zeros(1, 256, 3, 5)), "pos_8x8": pe(torch.zeros(2, 256, 8, 8))},
               os.path.join(HERE, "position_embedding.pt"))
    torch.save(meta, os.path.join(HERE, "meta.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
