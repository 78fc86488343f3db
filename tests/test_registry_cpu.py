"""Boundary b2: the module / registry API through which the reference selects and drives the path.

* ``from_config`` + ``build_pixel_decoder`` / ``build_transformer_decoder`` from a yacs-like config carrying the keys of
  the published recipe (ref run_50ep_no_noise_all_ly.sh:9-22 on top of configs/coco/instance-segmentation/
  maskformer2_R50_bs16_50ep.yaml), as ``MaskFormerHead.from_config`` calls them (ref mask_former_head.py:87-115);
* strict ``load_state_dict`` of state dicts whose keys and shapes were dumped from the UNMODIFIED reference modules
  (tests/golden/state_dict_shapes.json, generator committed next to it), R50 and Swin-L;
* ``install_into_detectron2`` against stub registries with Detectron2's ``Registry`` surface;
* with /root/reference present (authoring container): the reference's own ``MaskFormerHead`` built around the product
  modules and driven through ``head(features, mask, dn_args=...)`` (ref mask_former_head.py:117-132), kernels swapped
  for CPU torch ops (tests/test_host_logic_cpu.install_cpu_ops), against the oracle.
"""
import json
import os
import sys
import types

import pytest
import torch

import mp_former_b200 as M
from mp_former_b200 import registry
from oracle import ref_loader, synthetic
from oracle import torch_oracle as O
from test_host_logic_cpu import install_cpu_ops

HERE = os.path.dirname(os.path.abspath(__file__))


def ns(**kw):
    return types.SimpleNamespace(**kw)


def recipe_cfg(queries=100, classes=80):
    """The config keys the two ``from_config`` read, at the values of the published MP-Former recipe."""
    return ns(MODEL=ns(
        SEM_SEG_HEAD=ns(PIXEL_DECODER_NAME="MSDeformAttnPixelDecoder", IN_FEATURES=["res2", "res3", "res4", "res5"],
                        CONVS_DIM=256, MASK_DIM=256, NORM="GN", TRANSFORMER_ENC_LAYERS=6,
                        DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES=["res3", "res4", "res5"], COMMON_STRIDE=4,
                        NUM_CLASSES=classes, IGNORE_VALUE=255, LOSS_WEIGHT=1.0),
        MASK_FORMER=ns(TRANSFORMER_DECODER_NAME="MultiScaleMaskedTransformerDecoderMaskDN", DROPOUT=0.0, NHEADS=8,
                       HIDDEN_DIM=256, NUM_OBJECT_QUERIES=queries, DIM_FEEDFORWARD=2048, DEC_LAYERS=10,
                       PRE_NORM=False, ENFORCE_INPUT_PROJ=False, TRANSFORMER_IN_FEATURE="multi_scale_pixel_decoder",
                       DN_MODE="points", HEAD_DN=False, ALL_LY_DN=True, DN_RATIO=0.5, LB_NOISE_RATIO=0.2)))


def input_shape(backbone):
    return {k: M.ShapeSpec(channels=c, stride=synthetic.STRIDES[k])
            for k, c in synthetic.BACKBONE_CHANNELS[backbone].items()}


@pytest.mark.parametrize("backbone", ["r50", "swin_l"])
def test_builders_from_recipe_config_and_reference_state_dict_shapes(backbone):
    golden = json.load(open(os.path.join(HERE, "golden", "state_dict_shapes.json")))[backbone]
    cfg = recipe_cfg(golden["queries"], golden["classes"])
    pd = registry.build_pixel_decoder(cfg, input_shape(backbone))
    dec = registry.build_transformer_decoder(cfg, cfg.MODEL.SEM_SEG_HEAD.CONVS_DIM, mask_classification=True)
    assert isinstance(pd, M.MSDeformAttnPixelDecoder)
    assert isinstance(dec, M.MultiScaleMaskedTransformerDecoderMaskDN)
    assert dec.num_layers == 9 and dec.num_queries == golden["queries"] and dec.dn_mode == "points"
    assert dec.all_lys is True and dec.dn_label_noise_ratio == 0.2
    assert len(pd.transformer.encoder.layers) == 6
    for mod, want, tmpl in ((pd, golden["pixel_decoder"], synthetic.pixel_decoder_template(backbone)),
                            (dec, golden["decoder"], synthetic.decoder_template(golden["queries"],
                                                                                golden["classes"]))):
        # a checkpoint written by the reference module loads strictly: same keys, same shapes
        sd = {k: torch.full(shape, 0.25) for k, shape in want.items()}
        result = mod.load_state_dict(sd, strict=True)
        assert not result.missing_keys and not result.unexpected_keys
        assert all(torch.equal(v, sd[k]) for k, v in mod.state_dict().items())
        # and the import-free templates of the bench's reference arm describe the same state dicts
        assert {k: list(v.shape) for k, v in tmpl.items()} == want


def test_unknown_names_raise_like_detectron2_registry():
    cfg = recipe_cfg()
    cfg.MODEL.SEM_SEG_HEAD.PIXEL_DECODER_NAME = "BasePixelDecoder"
    with pytest.raises(KeyError, match="SEM_SEG_HEADS"):
        registry.build_pixel_decoder(cfg, input_shape("r50"))
    cfg.MODEL.MASK_FORMER.TRANSFORMER_DECODER_NAME = "StandardTransformerDecoder"
    with pytest.raises(KeyError, match="TRANSFORMER_MODULE"):
        registry.build_transformer_decoder(cfg, 256)


class _D2Registry:
    """detectron2.utils.registry.Registry's surface (fvcore): ``_obj_map``, ``register`` (asserts on duplicates)."""

    def __init__(self, name):
        self._name, self._obj_map = name, {}

    def register(self, obj=None):
        assert obj.__name__ not in self._obj_map, f"An object named '{obj.__name__}' was already registered"
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        return self._obj_map[name]


def test_install_into_detectron2_overrides_the_reference_classes(monkeypatch):
    heads, decoders = _D2Registry("SEM_SEG_HEADS"), _D2Registry("TRANSFORMER_MODULE")
    stock = type("MSDeformAttnPixelDecoder", (), {})
    heads.register(stock)                                            # the reference's own class is already there
    decoders.register(type("MultiScaleMaskedTransformerDecoderMaskDN", (), {}))
    mod = types.ModuleType("mask2former.modeling.transformer_decoder.maskformer_transformer_decoder")
    mod.TRANSFORMER_DECODER_REGISTRY = decoders
    monkeypatch.setitem(sys.modules, mod.__name__, mod)
    monkeypatch.setattr(registry, "HAVE_DETECTRON2", True)
    monkeypatch.setattr(registry, "_D2_HEADS", heads)
    registry.install_into_detectron2()
    assert heads.get("MSDeformAttnPixelDecoder") is M.MSDeformAttnPixelDecoder
    assert decoders.get("MultiScaleMaskedTransformerDecoderMaskDN") is M.MultiScaleMaskedTransformerDecoderMaskDN
    assert decoders.get("MultiScaleMaskedTransformerDecoder") is M.MultiScaleMaskedTransformerDecoder
    with pytest.raises(AssertionError):                              # without override a duplicate is refused
        registry.install_into_detectron2(override=False)
    monkeypatch.setattr(registry, "HAVE_DETECTRON2", False)
    with pytest.raises(RuntimeError, match="detectron2"):
        registry.install_into_detectron2()


def test_synthetic_workload_duplicates_agree_with_the_product_workload():
    """bench.py's reference arm builds its inputs without importing the product package: same seeds, same tensors."""
    from mp_former_b200 import workload
    for backbone, h, w in (("r50", 64, 96), ("swin_l", 64, 64)):
        a, b = workload.synthetic_features(2, h, w, backbone=backbone, seed=3), synthetic.features(2, h, w, backbone, 3)
        assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    ta, tb = workload.synthetic_targets(3, 64, 96, num_classes=8, seed=4), synthetic.targets(3, 64, 96, 8, seed=4)
    assert all(torch.equal(x[k], y[k]) for x, y in zip(ta, tb) for k in ("labels", "masks", "boxes"))
    crit_w = synthetic.recipe_weight_dict()
    assert len(crit_w) == 60 and crit_w["loss_mask_dn_8"] == 5.0 and crit_w["loss_ce"] == 2.0


@pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference (authoring container)")
def test_reference_maskformer_head_drives_the_product_modules(monkeypatch):
    """The reference's own caller: ``MaskFormerHead(...)`` around the product modules, ``head(features, mask,
    dn_args=...)`` -> ``pixel_decoder.forward_features`` -> ``predictor(multi_scale, mask_features, mask, dn_args=)``
    (ref mask_former_head.py:117-121).  Small geometry, kernels replaced by CPU torch ops; result vs the oracle."""
    import cases
    from test_host_logic_cpu import build_decoder, build_pixel_decoder
    from test_oracle_vs_golden import decoder_template, pixel_decoder_template
    R = ref_loader.load_head()
    install_cpu_ops(monkeypatch.setattr)
    pd, dec = build_pixel_decoder(), build_decoder()
    psd = O.seeded_state_dict(pixel_decoder_template(), seed=41)
    dsd = O.seeded_state_dict(decoder_template(), seed=51)
    c = cases.PD_CFG
    shape = {k: M.ShapeSpec(channels=c["channels"][k], stride=c["strides"][k]) for k in c["channels"]}
    head = R.MaskFormerHead(shape, num_classes=cases.DEC_CFG["num_classes"], pixel_decoder=pd, loss_weight=1.0,
                            ignore_value=-1, transformer_predictor=dec,
                            transformer_in_feature="multi_scale_pixel_decoder")
    # checkpoint keys as the reference's head stores them: sem_seg_head.{pixel_decoder,predictor}.*
    sd = {"pixel_decoder." + k: v for k, v in psd.items()}
    sd.update({"predictor." + k: v for k, v in dsd.items()})
    result = head.load_state_dict(sd, strict=True)
    assert not result.missing_keys and not result.unexpected_keys
    feats = cases.pixel_decoder_features()
    dn = {"tgt": cases.dn_targets(), "scalar": 1, "noise_scale": 0.0}
    with torch.no_grad():
        out = head(feats, None, dn_args=dn)
        d = cases.DEC_CFG
        omf, _, oms = O.pixel_decoder_forward(psd, feats, n_heads=c["nheads"], enc_layers=c["enc_layers"])
        ref = O.decoder_forward(dsd, oms, omf, num_queries=d["num_queries"], n_heads=d["nheads"],
                                dec_layers=d["dec_layers"], num_classes=d["num_classes"], dn_args=dn)
    assert set(out) == {"pred_logits", "pred_masks", "aux_outputs", "dn_out"}
    for key in ("pred_logits", "pred_masks"):
        assert torch.allclose(out[key], ref[key], rtol=1e-3, atol=1e-3), key
        assert torch.allclose(out["dn_out"][key], ref["dn_out"][key], rtol=1e-3, atol=1e-3), key
    assert len(out["aux_outputs"]) == d["dec_layers"]
