"""TEST / BENCH INFRASTRUCTURE ONLY -- synthetic workloads of the BASELINE.json configurations for the CPU legs of
bench.py (``--impl reference`` and ``cpu_baseline``) WITHOUT importing the product package: state-dict templates with
the reference's key names and shapes (pinned against the reference modules' own key sets by
tests/test_oracle_vs_golden.py and, for these full-size templates, against the product modules by
tests/test_registry_cpu.py), seeded parameters, R50 / Swin-L shaped backbone features and COCO-like targets.

``features`` / ``targets`` are deliberate duplicates of ``mp_former_b200.workload.synthetic_features / _targets``
(same seeds -> same tensors; tests/test_host_logic_cpu.py checks it), so that the two arms of the bench see the same
data without one importing the other.
"""
import torch

from . import torch_oracle as O

BACKBONE_CHANNELS = {
    "r50": {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048},
    "swin_l": {"res2": 192, "res3": 384, "res4": 768, "res5": 1536},
}
STRIDES = {"res2": 4, "res3": 8, "res4": 16, "res5": 32}

# BASELINE.json `configs` (index + 1) -> geometry.  `global_batch` is the recipe's SOLVER.IMS_PER_BATCH (16; the
# Cityscapes stress case of configs[4] names 8).  ref: configs/coco/instance-segmentation/maskformer2_R50_bs16_50ep.yaml,
# configs/coco/panoptic-segmentation/swin/maskformer2_swin_large_IN21k_384_bs16_100ep.yaml:5-18,
# configs/cityscapes/instance-segmentation/maskformer2_R50_bs16_90k.yaml.
PRESETS = {
    2: dict(name="config2: R50, COCO-instance head, 100 queries, 1024x1024", backbone="r50", queries=100,
            classes=80, height=1024, width=1024, global_batch=16),
    3: dict(name="config3: config2 data-parallel, global batch 16, SetCriterion + HungarianMatcher", backbone="r50",
            queries=100, classes=80, height=1024, width=1024, global_batch=16),
    4: dict(name="config4: Swin-L, COCO-panoptic head, 200 queries, 1024x1024", backbone="swin_l", queries=200,
            classes=133, height=1024, width=1024, global_batch=16),
    5: dict(name="config5: R50, Cityscapes-instance head, 100 queries, 1024x2048 (32768 keys on the finest level)",
            backbone="r50", queries=100, classes=8, height=1024, width=2048, global_batch=8),
}


def pixel_decoder_template(backbone="r50", conv_dim=256, mask_dim=256, d_ffn=1024, enc_layers=6, heads=8, levels=3,
                           points=4):
    """name -> shape of ``MSDeformAttnPixelDecoder.state_dict()`` (ref pixel_decoder/msdeformattn.py:166-311)."""
    ch = BACKBONE_CHANNELS[backbone]
    C = conv_dim
    t = {}
    for i, f in enumerate(("res5", "res4", "res3")):
        t[f"input_proj.{i}.0.weight"] = (C, ch[f], 1, 1)
        t[f"input_proj.{i}.0.bias"] = (C,)
        t[f"input_proj.{i}.1.weight"] = (C,)
        t[f"input_proj.{i}.1.bias"] = (C,)
    t["transformer.level_embed"] = (levels, C)
    n_off, n_att = heads * levels * points * 2, heads * levels * points
    for i in range(enc_layers):
        p = f"transformer.encoder.layers.{i}."
        for name, shape in (("self_attn.sampling_offsets", (n_off, C)), ("self_attn.attention_weights", (n_att, C)),
                            ("self_attn.value_proj", (C, C)), ("self_attn.output_proj", (C, C)),
                            ("linear1", (d_ffn, C)), ("linear2", (C, d_ffn))):
            t[p + name + ".weight"] = shape
            t[p + name + ".bias"] = (shape[0],)
        for name in ("norm1", "norm2"):
            t[p + name + ".weight"] = (C,)
            t[p + name + ".bias"] = (C,)
    t["mask_features.weight"] = (mask_dim, C, 1, 1)
    t["mask_features.bias"] = (mask_dim,)
    t["adapter_1.weight"] = (C, ch["res2"], 1, 1)
    t["layer_1.weight"] = (C, C, 3, 3)
    for name in ("adapter_1", "layer_1"):
        t[name + ".norm.weight"] = (C,)
        t[name + ".norm.bias"] = (C,)
    return {k: torch.empty(s) for k, s in t.items()}


def decoder_template(queries=100, classes=80, hidden=256, d_ffn=2048, dec_layers=9, mask_dim=256):
    """name -> shape of ``MultiScaleMaskedTransformerDecoderMaskDN.state_dict()`` (ref decoder :601-695)."""
    C = hidden
    t = {}
    for i in range(dec_layers):
        for kind, attn in (("cross", "multihead_attn"), ("self", "self_attn")):
            p = f"transformer_{kind}_attention_layers.{i}."
            t[p + attn + ".in_proj_weight"] = (3 * C, C)
            t[p + attn + ".in_proj_bias"] = (3 * C,)
            t[p + attn + ".out_proj.weight"] = (C, C)
            t[p + attn + ".out_proj.bias"] = (C,)
            t[p + "norm.weight"] = (C,)
            t[p + "norm.bias"] = (C,)
        p = f"transformer_ffn_layers.{i}."
        t[p + "linear1.weight"] = (d_ffn, C)
        t[p + "linear1.bias"] = (d_ffn,)
        t[p + "linear2.weight"] = (C, d_ffn)
        t[p + "linear2.bias"] = (C,)
        t[p + "norm.weight"] = (C,)
        t[p + "norm.bias"] = (C,)
    t["decoder_norm.weight"] = (C,)
    t["decoder_norm.bias"] = (C,)
    t["query_feat.weight"] = (queries, C)
    t["level_embed.weight"] = (3, C)
    t["class_embed.weight"] = (classes + 1, C)
    t["class_embed.bias"] = (classes + 1,)
    for j in range(3):
        t[f"mask_embed.layers.{j}.weight"] = (mask_dim if j == 2 else C, C)
        t[f"mask_embed.layers.{j}.bias"] = (mask_dim if j == 2 else C,)
    t["label_enc.weight"] = (classes, C)
    return {k: torch.empty(s) for k, s in t.items()}


def head_state_dicts(backbone="r50", queries=100, classes=80, seed=0):
    """Seeded parameters of the whole head (pixel decoder, decoder) as two state dicts."""
    return (O.seeded_state_dict(pixel_decoder_template(backbone), seed=seed),
            O.seeded_state_dict(decoder_template(queries, classes), seed=seed + 1))


def features(batch, height=1024, width=1024, backbone="r50", seed=0):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(batch, c, height // STRIDES[k], width // STRIDES[k], generator=g)
            for k, c in BACKBONE_CHANNELS[backbone].items()}


def targets(batch, height=1024, width=1024, num_classes=80, seed=0, max_inst=20):
    """Per image n in [1, max_inst] instances: axis-aligned ellipses as masks, labels, boxes."""
    g = torch.Generator().manual_seed(seed + 1)
    ys = torch.arange(height).view(-1, 1).float()
    xs = torch.arange(width).view(1, -1).float()
    out = []
    for _ in range(batch):
        n = int(torch.randint(1, max_inst + 1, (1,), generator=g))
        cy = torch.rand(n, generator=g) * height
        cx = torch.rand(n, generator=g) * width
        ry = (torch.rand(n, generator=g) * 0.25 + 0.03) * height
        rx = (torch.rand(n, generator=g) * 0.25 + 0.03) * width
        masks = (((ys[None] - cy.view(-1, 1, 1)) / ry.view(-1, 1, 1)) ** 2 +
                 ((xs[None] - cx.view(-1, 1, 1)) / rx.view(-1, 1, 1)) ** 2) <= 1.0
        labels = torch.randint(0, num_classes, (n,), generator=g)
        boxes = torch.stack([cx / width, cy / height, 2 * rx / width, 2 * ry / height], -1)
        out.append({"labels": labels, "masks": masks, "boxes": boxes})
    return out


def recipe_weight_dict(dec_layers=10, class_weight=2.0, mask_weight=5.0, dice_weight=5.0):
    """Loss weights as ``MaskFormer.from_config`` assembles them (ref maskformer_model.py:105-147)."""
    w = {"loss_ce": class_weight, "loss_mask": mask_weight, "loss_dice": dice_weight}
    w.update({k + "_dn": v for k, v in list(w.items())})
    aux = {}
    for i in range(dec_layers - 1):
        aux.update({k + f"_{i}": v for k, v in w.items()})
    w.update(aux)
    return w
