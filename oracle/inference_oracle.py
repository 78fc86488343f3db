"""TEST INFRASTRUCTURE ONLY -- restatement (plain PyTorch) of the reference's inference epilogue, the consumer of the
decoder's outputs at test time.  Not used by the product path or by bench.py.

Follows mask2former/maskformer_model.py of the reference:
  * :236-243  pred_masks resized to the padded input size (bilinear, align_corners=False)
  * :247-260  per image: detectron2 ``sem_seg_postprocess`` (crop the padding away, resize to the requested output
              resolution; third-party, restated from its published source in oracle/ref_loader.py)
  * :300-304  ``semantic_inference``; :365-401 ``instance_inference``
Pinned by ``tests/golden/inference.pt``: the UNMODIFIED ``MaskFormer.forward`` in eval mode around a stand-in backbone /
head (tests/golden/make_golden_inference.py)."""
import torch
import torch.nn.functional as F


def full_resolution_masks(mask_pred, padded_size, image_size, out_size):
    """mask_pred [Q, h, w] -> [Q, out_h, out_w]  (:236-243, :256-259)."""
    up = F.interpolate(mask_pred[None], size=tuple(padded_size), mode="bilinear", align_corners=False)[0]
    up = up[:, :image_size[0], :image_size[1]]
    return F.interpolate(up[None], size=tuple(out_size), mode="bilinear", align_corners=False)[0]


def semantic_inference(mask_cls, mask_pred_full):
    return torch.einsum("qc,qhw->chw", F.softmax(mask_cls, dim=-1)[..., :-1], mask_pred_full.sigmoid())


def instance_inference(mask_cls, mask_pred_full, num_classes, topk, thing_ids=None):
    """-> dict(pred_masks bool [k, H, W], scores [k], pred_classes [k], query [k])   (:365-401)."""
    scores = F.softmax(mask_cls, dim=-1)[:, :-1]
    s, idx = scores.flatten(0, 1).topk(topk, sorted=False)
    labels, query = idx % num_classes, idx // num_classes
    if thing_ids is not None:                      # panoptic models keep the "thing" classes only (:381-388)
        keep = torch.isin(labels, torch.as_tensor(sorted(thing_ids), device=labels.device))
        s, labels, query = s[keep], labels[keep], query[keep]
    m = mask_pred_full[query]
    fg = (m > 0).float()
    mask_scores = (m.sigmoid().flatten(1) * fg.flatten(1)).sum(1) / (fg.flatten(1).sum(1) + 1e-6)
    return {"pred_masks": fg.bool(), "scores": s * mask_scores, "pred_classes": labels, "query": query}
