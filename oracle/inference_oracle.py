"""TEST INFRASTRUCTURE ONLY -- restatement (plain PyTorch) of the reference's inference epilogue, the consumer of the
decoder's outputs at test time.  Not used by the product path or by bench.py.

Follows mask2former/maskformer_model.py of the reference:
  * :239-244  pred_masks resized to the padded input size (bilinear, align_corners=False)
  * :247-260  per image: detectron2 ``sem_seg_postprocess`` (crop the padding away, resize to the requested output
              resolution; third-party, restated from its published source in oracle/ref_loader.py)
  * :301-305  ``semantic_inference``; :365-401 ``instance_inference``
Pinned by ``tests/golden/inference.pt``: the UNMODIFIED ``MaskFormer.forward`` in eval mode around a stand-in backbone /
head (tests/golden/make_golden_inference.py)."""
import torch
import torch.nn.functional as F


def full_resolution_masks(mask_pred, padded_size, image_size, out_size):
    """mask_pred [Q, h, w] -> [Q, out_h, out_w]  (:239-244, :257-259)."""
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


def panoptic_inference(mask_cls, mask_pred_full, num_classes, thing_ids, object_mask_threshold, overlap_threshold):
    """-> (panoptic_seg int32 [H, W], segments_info)   (:307-363, segment by segment like the reference)."""
    scores, labels = F.softmax(mask_cls, dim=-1).max(-1)
    prob = mask_pred_full.sigmoid()
    keep = labels.ne(num_classes) & (scores > object_mask_threshold)
    cur_scores, cur_classes, cur_masks = scores[keep], labels[keep], prob[keep]
    seg = torch.zeros(prob.shape[-2:], dtype=torch.int32, device=prob.device)
    info = []
    if cur_masks.shape[0] == 0:
        return seg, info
    ids = (cur_scores.view(-1, 1, 1) * cur_masks).argmax(0)
    stuff, current = {}, 0
    for k in range(cur_classes.shape[0]):
        c = int(cur_classes[k])
        isthing = c in thing_ids
        area = int((ids == k).sum())
        original = int((cur_masks[k] >= 0.5).sum())
        m = (ids == k) & (cur_masks[k] >= 0.5)
        if area > 0 and original > 0 and int(m.sum()) > 0:
            if area / original < overlap_threshold:
                continue
            if not isthing:
                if c in stuff:
                    seg[m] = stuff[c]
                    continue
                stuff[c] = current + 1
            current += 1
            seg[m] = current
            info.append({"id": current, "isthing": bool(isthing), "category_id": c})
    return seg, info
