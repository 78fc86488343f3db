"""Generator of tests/golden/state_dict_shapes.json: key -> shape of the state dicts of the UNMODIFIED reference
modules at the published recipe (R50 and Swin-L input channels; 6 encoder layers, 9 decoder layers; 100 / 200 queries;
80 / 133 classes), dumped through oracle/ref_loader.py.  Runs in the authoring container only (/root/reference);
tests/test_registry_cpu.py strict-loads state dicts of exactly these keys and shapes into the product modules.

    python tests/golden/make_golden_state_dict_shapes.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

CHANNELS = {"r50": {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048},
            "swin_l": {"res2": 192, "res3": 384, "res4": 768, "res5": 1536}}
STRIDES = {"res2": 4, "res3": 8, "res4": 16, "res5": 32}


def main():
    R = ref_loader.load_reference()
    out = {}
    for backbone, queries, classes in (("r50", 100, 80), ("swin_l", 200, 133)):
        shape = {k: R.ShapeSpec(channels=c, stride=STRIDES[k]) for k, c in CHANNELS[backbone].items()}
        pd = R.MSDeformAttnPixelDecoder(shape, transformer_dropout=0.0, transformer_nheads=8,
                                        transformer_dim_feedforward=1024, transformer_enc_layers=6, conv_dim=256,
                                        mask_dim=256, norm="GN", transformer_in_features=["res3", "res4", "res5"],
                                        common_stride=4)
        dec = R.MultiScaleMaskedTransformerDecoderMaskDN(
            256, True, num_classes=classes, hidden_dim=256, num_queries=queries, nheads=8, dim_feedforward=2048,
            dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, dn_mode="points", all_lys=True,
            dn_label_noise_ratio=0.2)
        out[backbone] = {"pixel_decoder": {k: list(v.shape) for k, v in pd.state_dict().items()},
                         "decoder": {k: list(v.shape) for k, v in dec.state_dict().items()},
                         "queries": queries, "classes": classes}
    with open(os.path.join(HERE, "state_dict_shapes.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print({k: (len(v["pixel_decoder"]), len(v["decoder"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
