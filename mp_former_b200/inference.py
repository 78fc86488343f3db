"""Inference epilogue of the mask-classification model on the device: what the reference's
``MaskFormer.forward`` does after the head in eval mode (mask2former/maskformer_model.py:232-279, :301-305, :365-401).

``instance_inference`` takes the decoder's LOW-RESOLUTION mask logits: the resize to the padded input size, the crop
and resize of detectron2's ``sem_seg_postprocess``, the gather of the top-k queries, the ``> 0`` threshold and the
foreground-probability score are one kernel (``native.instance_masks``) that writes each binary mask once -- the
reference materialises Q full-resolution fp32 maps per image (420 MB at 100 queries, 1024 x 1024) and re-reads them
four times.  ``semantic_inference`` is the reference's arithmetic on library ops (resize, softmax, einsum).
``panoptic_inference`` (:307-363) computes the per-query pixel counts the reference reads back one ``.item()`` at a
time with two bincounts and a row sum, and synchronises once.

A maintainer's patch in the reference: in ``MaskFormer.forward`` drop the ``F.interpolate`` at :239-244 and call
``instance_inference(mask_cls_result, low_res_mask_result, images.tensor.shape[-2:], image_size, (height, width),
num_classes, test_topk_per_image)`` in place of :274-276."""
import torch
import torch.nn.functional as F

from . import native


class InstanceResult(dict):
    """What the reference stores in a detectron2 ``Instances`` (:390-400): ``pred_masks`` [k, H, W], ``scores`` [k],
    ``pred_classes`` [k], ``pred_boxes`` (zeros [k, 4], as in the reference, :393) and ``image_size``; attribute access
    like ``Instances``."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None


@torch.no_grad()
def instance_inference(mask_cls, mask_pred, padded_size, image_size, out_size, num_classes, topk, thing_ids=None,
                       mask_dtype=torch.float32):
    """mask_cls [Q, K+1], mask_pred [Q, h, w] low-resolution logits of one image -> InstanceResult.
    ``mask_dtype=torch.float32`` reproduces the reference's 0/1 float masks (:392); ``torch.uint8`` writes a quarter of
    the bytes."""
    scores = F.softmax(mask_cls.float(), dim=-1)[:, :-1]
    k = min(int(topk), scores.numel())
    s, idx = scores.flatten(0, 1).topk(k, sorted=False)
    labels = idx % num_classes                      # == arange(K).repeat(Q)[idx]   (:371-374)
    query = idx // num_classes
    if thing_ids is not None:                       # panoptic models keep the "thing" classes only (:381-388)
        keep = torch.isin(labels, torch.as_tensor(sorted(thing_ids), device=labels.device))
        s, labels, query = s[keep], labels[keep], query[keep]
    masks, sums = native.instance_masks(mask_pred.float(), query, padded_size, image_size, out_size, mask_dtype)
    mask_scores = sums[:, 0] / (sums[:, 1] + 1e-6)  # average foreground probability (:398)
    return InstanceResult(image_size=tuple(int(v) for v in out_size), pred_masks=masks, scores=s * mask_scores,
                          pred_classes=labels, pred_boxes=torch.zeros(masks.shape[0], 4, device=masks.device))


@torch.no_grad()
def semantic_inference(mask_cls, mask_pred, padded_size, image_size, out_size, postprocess_before_inference=True):
    """mask_cls [Q, K+1], mask_pred [Q, h, w] -> sem_seg [K, out_h, out_w]   (:239-244, :257-267, :301-305)."""
    up = F.interpolate(mask_pred.float()[None], size=tuple(padded_size), mode="bilinear", align_corners=False)[0]

    def post(x):
        x = x[:, :image_size[0], :image_size[1]]
        return F.interpolate(x[None], size=tuple(out_size), mode="bilinear", align_corners=False)[0]

    cls = F.softmax(mask_cls.float(), dim=-1)[..., :-1]
    if postprocess_before_inference:
        return torch.einsum("qc,qhw->chw", cls, post(up).sigmoid())
    return post(torch.einsum("qc,qhw->chw", cls, up.sigmoid()))


@torch.no_grad()
def panoptic_inference(mask_cls, mask_pred, num_classes, thing_ids, object_mask_threshold, overlap_threshold):
    """mask_cls [Q, K+1], mask_pred [Q, H, W] full-resolution logits -> (panoptic_seg int32 [H, W], segments_info)
    (ref maskformer_model.py:307-363).

    The reference walks the kept queries one by one and reads three pixel counts per query back to the host
    (``.item()`` at :335-340: up to 3 Q stream synchronisations per image).  Here the three counts of every kept query
    come from two ``bincount``s and one row sum, ONE device->host copy brings (class, counts) of all kept queries, the
    sequential part that is inherently host-side -- segment numbering with the merge of same-class "stuff" segments,
    and ``segments_info``, a list of Python dicts -- runs on those few numbers, and one lookup-table gather paints the
    segment ids."""
    scores, labels = F.softmax(mask_cls.float(), dim=-1).max(-1)
    keep = labels.ne(num_classes) & (scores > object_mask_threshold)
    H, W = mask_pred.shape[-2:]
    seg = torch.zeros((H, W), dtype=torch.int32, device=mask_pred.device)
    kept = keep.nonzero().flatten()                 # (data-dependent size: the reference's boolean indexing, :311-315)
    n = int(kept.numel())
    if n == 0:
        return seg, []
    prob = mask_pred.float()[kept].sigmoid()
    ids = (scores[kept].view(-1, 1, 1) * prob).argmax(0)                       # :331
    fg = prob >= 0.5
    sel = fg.gather(0, ids[None])[0]                                           # pixel claimed by its arg-max mask
    flat = ids.flatten()
    area = torch.bincount(flat, minlength=n)                                   # (ids == k).sum()            :336
    inter = torch.bincount(flat[sel.flatten()], minlength=n)                   # ((ids == k) & fg_k).sum()   :338-340
    original = fg.flatten(1).sum(1)                                            # (mask_k >= 0.5).sum()       :337
    host = torch.stack([labels[kept], area, original, inter]).cpu().tolist()   # the one synchronisation
    thing_ids = set(int(t) for t in thing_ids)
    lut, info, stuff, current = [0] * n, [], {}, 0
    for k, (c, a, o, i) in enumerate(zip(*host)):
        if not (a > 0 and o > 0 and i > 0) or a / o < overlap_threshold:
            continue
        isthing = c in thing_ids
        if not isthing:
            if c in stuff:                          # merge stuff regions of one class (:345-348)
                lut[k] = stuff[c]
                continue
            stuff[c] = current + 1
        current += 1
        lut[k] = current
        info.append({"id": current, "isthing": bool(isthing), "category_id": int(c)})
    lut_t = torch.tensor(lut, dtype=torch.int32, device=seg.device)
    seg = torch.where(sel, lut_t[ids], seg)
    return seg, info
