"""Synthetic COCO-shaped workloads for bench.py / tests (BASELINE.json configs).

No datasets or checkpoints exist offline: backbone features are seeded ``randn`` tensors of the
shapes a ResNet-50 / Swin-L backbone produces, weights are module-default initialisations perturbed
so that sampling is non-degenerate (the stock init has zero attention logits and fixed offsets,
ref ops/modules/ms_deform_attn.py:61-76).
"""
import torch

from .masked_decoder import MultiScaleMaskedTransformerDecoderMaskDN
from .pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec

BACKBONE_CHANNELS = {
    "r50": {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048},
    "swin_l": {"res2": 192, "res3": 384, "res4": 768, "res5": 1536},
}
STRIDES = {"res2": 4, "res3": 8, "res4": 16, "res5": 32}


def build_head(backbone="r50", num_queries=100, num_classes=80, device="cuda", seed=0,
               pixel_decoder_cls=MSDeformAttnPixelDecoder, decoder_cls=MultiScaleMaskedTransformerDecoderMaskDN):
    """COCO-instance head of MP-Former (configs/coco/instance-segmentation/maskformer2_R50_bs16_50ep.yaml
    + run_50ep_no_noise_all_ly.sh:9-22): 6 encoder layers, 9 decoder layers, DN_MODE points,
    ALL_LY_DN, LB_NOISE_RATIO 0.2."""
    torch.manual_seed(seed)
    ch = BACKBONE_CHANNELS[backbone]
    shape = {k: ShapeSpec(channels=ch[k], stride=STRIDES[k]) for k in ch}
    pd = pixel_decoder_cls(shape, transformer_dropout=0.0, transformer_nheads=8,
                           transformer_dim_feedforward=1024, transformer_enc_layers=6, conv_dim=256,
                           mask_dim=256, norm="GN", transformer_in_features=["res3", "res4", "res5"],
                           common_stride=4)
    dec = decoder_cls(256, True, num_classes=num_classes, hidden_dim=256, num_queries=num_queries,
                      nheads=8, dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256,
                      enforce_input_project=False, dn_mode="points", all_lys=True,
                      dn_label_noise_ratio=0.2)
    with torch.no_grad():
        for m in pd.modules():
            if m.__class__.__name__ == "MSDeformAttn":
                m.attention_weights.weight.normal_(std=0.02)
                m.sampling_offsets.weight.normal_(std=0.02)
    return pd.to(device), dec.to(device)


def synthetic_features(batch, height=1024, width=1024, backbone="r50", seed=0, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    ch = BACKBONE_CHANNELS[backbone]
    feats = {}
    for k, c in ch.items():
        t = torch.randn(batch, c, height // STRIDES[k], width // STRIDES[k], generator=g)
        if pin:
            t = t.pin_memory()
        feats[k] = t if device == "cpu" else t.to(device)
    return feats


def synthetic_targets(batch, height=1024, width=1024, num_classes=80, seed=0, device="cpu", max_inst=20):
    """Per image n in [1, max_inst] instances: axis-aligned ellipses as masks, labels, boxes
    (only ``len(boxes)`` is read by the decoder, ref decoder :970-971)."""
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
        out.append({"labels": labels.to(device), "masks": masks.to(device), "boxes": boxes.to(device)})
    return out


def build_criterion(num_classes=80, dec_layers=10, class_weight=2.0, mask_weight=5.0, dice_weight=5.0,
                    no_object_weight=0.1, num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75,
                    device="cuda", device_indices=True):
    """The criterion of the COCO-instance recipe as ``MaskFormer.from_config`` assembles it (ref
    maskformer_model.py:105-147: weights 2 / 5 / 5, eos 0.1, 12544 points, oversampling 3, importance ratio 0.75,
    deep supervision over ``dec_layers - 1`` auxiliary layers, dn losses weighted like the matching ones), on the
    device matcher.  Returns (criterion, weighted_sum) with ``weighted_sum(losses)`` the scalar the trainer
    back-propagates (maskformer_model.py:225-231)."""
    from .criterion import SetCriterion
    from .matcher import HungarianMatcher
    matcher = HungarianMatcher(cost_class=class_weight, cost_mask=mask_weight, cost_dice=dice_weight,
                               num_points=num_points, device_indices=device_indices)
    weight_dict = {"loss_ce": class_weight, "loss_mask": mask_weight, "loss_dice": dice_weight}
    weight_dict.update({k + "_dn": v for k, v in list(weight_dict.items())})
    aux = {}
    for i in range(dec_layers - 1):
        aux.update({k + f"_{i}": v for k, v in weight_dict.items()})
    weight_dict.update(aux)
    crit = SetCriterion(num_classes, matcher=matcher, weight_dict=weight_dict, eos_coef=no_object_weight,
                        losses=["labels", "masks"], num_points=num_points, oversample_ratio=oversample_ratio,
                        importance_sample_ratio=importance_sample_ratio).to(device)

    def weighted_sum(losses):
        return crit.weighted_total(losses, weight_dict)

    return crit, weighted_sum
