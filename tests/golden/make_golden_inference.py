"""Generates tests/golden/inference.pt by running the UNMODIFIED reference MaskFormer.forward in eval mode
(mask2former/maskformer_model.py:232-279 and the *_inference methods :301-401) around a stand-in backbone / head that
returns seeded ``pred_logits`` / ``pred_masks`` (authoring container only):

    python tests/golden/make_golden_inference.py

Two images of different sizes (padding to the size divisibility, crop, resize to the requested output resolution)."""
import os
import sys
import types
import warnings

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

CFG = dict(num_queries=14, num_classes=6, topk=9, size_divisibility=32, thing_ids=(0, 2, 3),
           object_mask_threshold=0.3, overlap_threshold=0.6)
IMAGES = [  # (input h, w), (requested output h, w)
    ((64, 96), (64, 96)),          # no padding, output = input resolution (resize is the identity)
    ((50, 70), (75, 105)),         # padded to 64 x 96, cropped, resized x1.5
]


def inputs(seed=41, stride=4):
    g = torch.Generator().manual_seed(seed)
    B, Q, K = len(IMAGES), CFG["num_queries"], CFG["num_classes"]
    d = CFG["size_divisibility"]
    Hp = max((s[0] + d - 1) // d * d for s, _ in IMAGES)
    Wp = max((s[1] + d - 1) // d * d for s, _ in IMAGES)
    out = {"pred_logits": torch.randn(B, Q, K + 1, generator=g) * 2,
           "pred_masks": torch.randn(B, Q, Hp // stride, Wp // stride, generator=g) * 3}
    batched = [{"image": torch.rand(3, h, w, generator=g) * 255, "height": oh, "width": ow}
               for (h, w), (oh, ow) in IMAGES]
    return out, batched


def panoptic_inputs(seed=43, stride=4):
    """Structured predictions for the panoptic branch: elliptic blobs as mask logits, several queries of the same
    "stuff" class (merged into one segment, :345-350), overlapping blobs (overlap-threshold rejections, :341-342),
    no-object and low-confidence queries (:311)."""
    g = torch.Generator().manual_seed(seed)
    out, batched = inputs()
    B, Q, K = out["pred_logits"].shape[0], CFG["num_queries"], CFG["num_classes"]
    h, w = out["pred_masks"].shape[-2:]
    classes = torch.tensor([[0, 1, 1, 2, 5, 5, 3, 6, 4, 1, 0, 6, 5, 2], [1, 1, 4, 4, 0, 3, 3, 6, 6, 5, 2, 2, 1, 0]])
    logits = torch.randn(B, Q, K + 1, generator=g) * 0.7
    logits.scatter_add_(2, classes[..., None], torch.full((B, Q, 1), 4.0))
    logits[:, 9] *= 0.1                                            # a low-confidence query
    ys, xs = torch.arange(h).view(-1, 1).float(), torch.arange(w).view(1, -1).float()
    cy, cx = torch.rand(B, Q, generator=g) * h, torch.rand(B, Q, generator=g) * w
    ry, rx = (torch.rand(B, Q, generator=g) * 0.3 + 0.12) * h, (torch.rand(B, Q, generator=g) * 0.3 + 0.12) * w
    inside = (((ys - cy[..., None, None]) / ry[..., None, None]) ** 2 +
              ((xs - cx[..., None, None]) / rx[..., None, None]) ** 2) <= 1.0
    amp = torch.rand(B, Q, 1, 1, generator=g) * 3 + 2
    masks = torch.where(inside, amp, -amp) + torch.randn(B, Q, h, w, generator=g) * 0.5
    return {"pred_logits": logits, "pred_masks": masks}, batched


class _Head(nn.Module):
    def __init__(self, outputs, num_classes):
        super().__init__()
        self.outputs, self.num_classes = outputs, num_classes

    def forward(self, features, **kw):
        return dict(self.outputs)


def main():
    warnings.simplefilter("ignore")
    mod = ref_loader.load_meta_arch()
    outputs, batched = inputs()
    res = {}
    for name, semantic_on, instance_on, before, panoptic_on in (
            ("instance", False, True, True, False), ("semantic", True, False, True, False),
            ("semantic_after", True, False, False, False), ("panoptic", False, True, True, True),
            ("panoptic_structured", False, True, True, True)):
        if name == "panoptic_structured":
            outputs, batched = panoptic_inputs()
        backbone = nn.Identity()
        backbone.size_divisibility = CFG["size_divisibility"]
        model = mod.MaskFormer(backbone=backbone, sem_seg_head=_Head(outputs, CFG["num_classes"]), criterion=None,
                               num_queries=CFG["num_queries"], object_mask_threshold=CFG["object_mask_threshold"],
                               overlap_threshold=CFG["overlap_threshold"],
                               metadata=types.SimpleNamespace(
                                   thing_dataset_id_to_contiguous_id={10 * i: i for i in CFG["thing_ids"]}),
                               size_divisibility=CFG["size_divisibility"], sem_seg_postprocess_before_inference=before,
                               pixel_mean=[123.675, 116.28, 103.53], pixel_std=[58.395, 57.12, 57.375],
                               semantic_on=semantic_on, panoptic_on=panoptic_on, instance_on=instance_on,
                               test_topk_per_image=CFG["topk"], scalar=1, noise_scale=0.0).eval()
        with torch.no_grad():
            out = model(batched)
        packed = []
        for r in out:
            d = {}
            if "instances" in r:
                i = r["instances"]
                d.update({"image_size": i.image_size, "pred_masks": i.pred_masks.bool(), "scores": i.scores,
                          "pred_classes": i.pred_classes})
            if "sem_seg" in r:
                d["sem_seg"] = r["sem_seg"]
            if "panoptic_seg" in r:
                d["panoptic_seg"], d["segments_info"] = r["panoptic_seg"]
            packed.append(d)
        res[name] = packed
        print(name, [tuple(v.shape) if hasattr(v, "shape") else v for v in packed[1].values()][:6])
        if "segments_info" in packed[0]:
            print("   segments:", [[(d["id"], d["isthing"], d["category_id"]) for d in p["segments_info"]] for p in packed])
    torch.save(res, os.path.join(HERE, "inference.pt"))


if __name__ == "__main__":
    main()
