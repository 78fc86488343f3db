"""Shared by tests/test_gpu_j_full_config.py and bench.py's ``parity`` block: runs ONE image of a BASELINE
configuration through the product head (pixel decoder + masked decoder, CUDA) and through the CPU oracle
(oracle/torch_oracle.py, pinned to the unmodified reference) with identical weights and inputs, and separates the two
ways the outputs can differ:

  * arithmetic error -- measured with TEACHER FORCING: every cross-attention layer of the product decoder consumes the
    oracle's attention-mask bits, so both sides see the same discrete decisions and all ten heads must agree within
    ``1e-3 * max(1, |ref|)`` (north_star's tolerance);
  * mask-bit flips -- the boolean stage thresholds a resized logit at 0 (ref decoder :1869-1875); a logit within float
    noise of the threshold may land on the other side.  Under teacher forcing both sides derive the bits from (nearly)
    the same logits, so every flipped bit must belong to a logit within ``eps`` of the threshold; free-running, the
    flip rate per layer is reported (a flip in layer i changes the inputs of layer i+1, so later layers legitimately
    diverge on the affected rows).

Test infrastructure: imports ``oracle``; never imported by the product package.
"""
import torch

from oracle import torch_oracle as O


def to_dev(o, dev):
    if torch.is_tensor(o):
        return o.to(dev)
    if isinstance(o, dict):
        return {k: to_dev(v, dev) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(to_dev(v, dev) for v in o)
    return o


def rel_err(a, b):
    """max |a-b| / max(1, |b|)"""
    return float(((a - b).abs() / b.abs().clamp(min=1.0)).max())


def heads_of(out):
    """[(name, pred_logits, pred_masks)] over the ten heads (+ the dn heads) in layer order."""
    rows = []
    for tag, o in (("match", out), ("dn", out.get("dn_out"))):
        if o is None:
            continue
        for i, a in enumerate(o["aux_outputs"]):
            rows.append((f"{tag}{i}", a["pred_logits"], a["pred_masks"]))
        rows.append((f"{tag}{len(o['aux_outputs'])}", o["pred_logits"], o["pred_masks"]))
    return rows


def oracle_run(pd, dec, feats, targets, num_queries, enc_layers, dec_layers):
    """CPU oracle forward of one image; returns (out, trace)."""
    psd = {k: v.detach().cpu() for k, v in pd.state_dict().items()}
    dsd = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
    dn = None if targets is None else {"tgt": targets, "scalar": 1, "noise_scale": 0.0}
    trace = {}
    with torch.no_grad():
        mf, _, ms = O.pixel_decoder_forward(psd, feats, enc_layers=enc_layers)
        out = O.decoder_forward(dsd, ms, mf, num_queries=num_queries, dec_layers=dec_layers,
                                num_classes=dec.num_classes, dn_args=dn, dn_label_noise_ratio=-1.0, trace=trace)
    return out, trace


def product_run(pd, dec, feats, targets, dev, force=None):
    """Product forward on ``dev``; returns (out on CPU, own masks as bool [B,Qt,hw] per layer on CPU)."""
    dn = None if targets is None else {"tgt": to_dev(targets, dev), "scalar": 1, "noise_scale": 0.0}
    dbg = {"own": [], "force": force}
    dec.mask_debug = dbg
    try:
        with torch.no_grad():
            mf, _, ms = pd.forward_features(to_dev(feats, dev))
            out = dec(ms, mf, None, dn)
    finally:
        dec.mask_debug = None
    own = [m.to_bool().cpu() for m in dbg["own"]]

    def cpu(o):
        if torch.is_tensor(o):
            return o.detach().cpu()
        if isinstance(o, dict):
            return {k: cpu(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return type(o)(cpu(v) for v in o)
        return o
    return cpu(out), own


def compare(pd, dec, feats, targets, dev, num_queries, enc_layers=6, dec_layers=9, eps=1e-3):
    """Runs the oracle once and the product twice (teacher-forced, free-running).  Returns a dict of plain numbers:

    forced_max_rel_logits / forced_max_rel_masks   worst head under teacher forcing (arithmetic error)
    forced_flip_rate[i], forced_flip_max_dist      bits the product would have set differently, and how far the
                                                   oracle's resized logit of the worst such bit is from 0
    free_flip_rate[i]                              per-layer flip rate of the free-running product
    free_frac_above_1e-3_masks / _logits           share of final-head entries beyond the tolerance, free-running
    """
    saved = dec.dn_label_noise_ratio
    dec.dn_label_noise_ratio = -1.0            # the label noise draws from the device generator: off on both sides
    try:
        ref, trace = oracle_run(pd, dec, feats, targets, num_queries, enc_layers, dec_layers)
        forced, own_f = product_run(pd, dec, feats, targets, dev, force=trace["masks"])
        free, own = product_run(pd, dec, feats, targets, dev, force=None)
    finally:
        dec.dn_label_noise_ratio = saved
    res = {"layers": len(trace["masks"]), "eps": eps}
    worst_l = worst_m = 0.0
    per_head = {}
    for (name, rl, rm), (_, pl, pm) in zip(heads_of(ref), heads_of(forced)):
        el, em = rel_err(pl, rl), rel_err(pm, rm)
        per_head[name] = (el, em)
        worst_l, worst_m = max(worst_l, el), max(worst_m, em)
    res["forced_per_head"] = per_head
    res["forced_max_rel_logits"], res["forced_max_rel_masks"] = worst_l, worst_m
    n_dn = ref["dn_out"]["pred_masks"].shape[1] if ref.get("dn_out") is not None else 0
    rates, dist = [], 0.0
    for i, (ob, rb) in enumerate(zip(own_f, trace["masks"])):
        assert ob.shape == rb.shape, (ob.shape, rb.shape)
        flip = ob != rb
        assert not bool(flip[:, :n_dn].any()), f"layer {i}: mask-piloted (GT) rows differ"
        rates.append(float(flip.float().mean()))
        if bool(flip.any()):
            dist = max(dist, float(trace["resized"][i][flip].abs().max()))
    res["forced_flip_rate"], res["forced_flip_max_dist"] = rates, dist
    res["free_flip_rate"] = [float((ob != rb).float().mean()) for ob, rb in zip(own, trace["masks"])]
    flip0 = own[0] != trace["masks"][0]
    res["free_layer0_flip_max_dist"] = float(trace["resized"][0][flip0].abs().max()) if bool(flip0.any()) else 0.0
    fm = ((free["pred_masks"] - ref["pred_masks"]).abs() / ref["pred_masks"].abs().clamp(min=1.0))
    fl = ((free["pred_logits"] - ref["pred_logits"]).abs() / ref["pred_logits"].abs().clamp(min=1.0))
    res["free_frac_above_1e-3_masks"] = float((fm > 1e-3).float().mean())
    res["free_frac_above_1e-3_logits"] = float((fl > 1e-3).float().mean())
    res["free_layer0_rel_masks"] = rel_err(heads_of(free)[0][2], heads_of(ref)[0][2])
    return res
