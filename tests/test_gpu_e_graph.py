"""-m gpu: one CUDA graph for forward + loss + backward of the path (mp_former_b200/graphs.py) reproduces the
eagerly issued step: same loss, same parameter gradients (up to the order of the backward's fp32 atomics), also
after new inputs are loaded into the static buffers."""
import pytest
import torch

import cases
from mp_former_b200 import graphs
from oracle import torch_oracle as O
from test_host_logic_cpu import build_decoder, build_pixel_decoder
from test_oracle_vs_golden import decoder_template, pixel_decoder_template

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_graphed_step_matches_eager_step():
    pd = build_pixel_decoder().to(DEV)
    pd.load_state_dict(O.seeded_state_dict(pixel_decoder_template(), seed=41))
    dec = build_decoder(dn_label_noise_ratio=-1.0).to(DEV)
    dec.load_state_dict(O.seeded_state_dict(decoder_template(), seed=51))
    feats = {k: v.to(DEV) for k, v in cases.pixel_decoder_features().items()}
    dn = {"tgt": [{k: v.to(DEV) for k, v in t.items()} for t in cases.dn_targets()], "scalar": 1, "noise_scale": 0.0}
    params = list(pd.parameters()) + list(dec.parameters())

    def step_fn(inp):
        mf, _, ms = pd.forward_features(inp)
        out = dec(ms, mf, None, dn)
        return out["pred_masks"].square().mean() + out["pred_logits"].square().mean() + \
            out["dn_out"]["pred_masks"].square().mean()

    def eager(inp):
        for p in params:
            p.grad = None
        loss = step_fn(inp)
        loss.backward()
        return loss.item(), [None if p.grad is None else p.grad.clone() for p in params]

    feats2 = {k: v * 0.5 + 0.1 for k, v in feats.items()}
    l1, g1 = eager(feats)
    l2, g2 = eager(feats2)
    gs = graphs.GraphedStep(step_fn, feats, params, warmup=2)
    assert gs.launches_per_replay > 50                       # the native kernels are inside the graph
    for inp, lref, gref in ((feats, l1, g1), (feats2, l2, g2), (feats, l1, g1)):
        loss = gs(inp)
        torch.cuda.synchronize()
        assert abs(loss.item() - lref) <= 1e-5 * max(1.0, abs(lref))
        for p, r in zip(params, gref):
            if r is None:
                continue
            scale = max(1e-6, r.abs().max().item())
            assert (p.grad - r).abs().max().item() / scale < 5e-4      # fp32 atomics: summation order differs run to run


def test_weight_operand_cache_equals_per_use_splits_and_follows_weight_updates():
    """native.WeightOperandCache (one launch refreshing every split weight operand of the step): same step as with a
    split per use; entries are bypassed after an in-place weight update until ``begin_step`` refreshes them; and a
    captured graph that contains ``begin_step`` follows weight updates on replay."""
    from mp_former_b200 import native
    pd = build_pixel_decoder().to(DEV)
    pd.load_state_dict(O.seeded_state_dict(pixel_decoder_template(), seed=41))
    dec = build_decoder(dn_label_noise_ratio=-1.0).to(DEV)
    dec.load_state_dict(O.seeded_state_dict(decoder_template(), seed=51))
    feats = {k: v.to(DEV) for k, v in cases.pixel_decoder_features().items()}
    dn = {"tgt": [{k: v.to(DEV) for k, v in t.items()} for t in cases.dn_targets()], "scalar": 1, "noise_scale": 0.0}
    params = list(pd.parameters()) + list(dec.parameters())

    def step_fn(inp):
        mf, _, ms = pd.forward_features(inp)
        out = dec(ms, mf, None, dn)
        return out["pred_masks"].square().mean() + out["pred_logits"].square().mean() + \
            out["dn_out"]["pred_masks"].square().mean()

    def run(begin=None):
        for p in params:
            p.grad = None
        if begin is not None:
            begin()
        loss = step_fn(feats)
        loss.backward()
        return loss.item(), [None if p.grad is None else p.grad.clone() for p in params]

    def same(a, b):
        assert abs(a[0] - b[0]) <= 1e-6 * max(1.0, abs(b[0]))
        for x, y in zip(a[1], b[1]):
            if y is not None:
                assert (x - y).abs().max().item() <= 5e-4 * max(1e-6, y.abs().max().item())

    ref = run()
    cache = native.set_weight_cache(native.WeightOperandCache(params))
    try:
        same(run(cache.begin_step), ref)                         # first step: everything recorded, served per use
        assert cache.hits == 0 and cache.pending
        same(run(cache.begin_step), ref)                         # second step: arena built, operands come from it
        assert cache.hits > 50 and not cache.pending and len(cache.entries) > 50
        with torch.no_grad():
            for p in params:
                p.mul_(1.01)
        native.set_weight_cache(None)
        ref2 = run()
        native.set_weight_cache(cache)
        hits = cache.hits
        same(run(), ref2)                                        # stale entries are bypassed without begin_step
        assert cache.hits == hits
        same(run(cache.begin_step), ref2)                        # ... and used again after it
        assert cache.hits > hits
        # inside a CUDA graph: begin_step is captured, so the replay refreshes the arena from the current weights
        gs = graphs.GraphedStep(lambda inp: (cache.begin_step(), step_fn(inp))[1], feats, params, warmup=2)
        gs(feats)
        torch.cuda.synchronize()
        same((gs.static_loss.item(), [p.grad for p in params]), ref2)
        with torch.no_grad():
            for p in params:
                p.div_(1.01)
        gs(feats)
        torch.cuda.synchronize()
        got = (gs.static_loss.item(), [None if p.grad is None else p.grad.clone() for p in params])
        native.set_weight_cache(None)
        for p in params:
            p.grad = None
        ref3 = run()
        assert abs(got[0] - ref3[0]) <= 1e-5 * max(1.0, abs(ref3[0]))
        for x, y in zip(got[1], ref3[1]):
            if y is not None:
                assert (x - y).abs().max().item() <= 5e-4 * max(1e-6, y.abs().max().item())
    finally:
        native.set_weight_cache(None)


def test_head_gradients_read_in_place_equal_the_concatenated_route():
    """Decoder -> criterion -> backward with the criterion's two evaluation orders: all heads in one pass hands the
    heads' mask-logit gradients back as slices of ONE buffer, which the decoder's batched head backward reads in place
    (in the criterion's head order: the small operands are permuted instead); head by head they arrive as separate
    tensors and are concatenated.  Same loss, same gradients of every decoder parameter and of mask_features."""
    from mp_former_b200 import ops, workload
    dec = build_decoder(dn_label_noise_ratio=-1.0).to(DEV).train()
    dec.load_state_dict(O.seeded_state_dict(decoder_template(), seed=51))
    x, mf0 = cases.decoder_inputs()
    x = [t.to(DEV) for t in x]
    targets = [{k: v.to(DEV) for k, v in t.items()} for t in cases.dn_targets()]
    K, L = cases.DEC_CFG["num_classes"], cases.DEC_CFG["dec_layers"]
    crit, weighted_sum = workload.build_criterion(num_classes=K, dec_layers=L + 1, num_points=112, device=DEV)
    crit.train(True)
    params = [p for p in dec.parameters() if p.requires_grad]

    def run(joint):
        for p in params:
            p.grad = None
        mf = mf0.to(DEV).requires_grad_(True)
        crit.joint_heads = joint
        torch.manual_seed(3)
        out = dec(x, mf, None, {"tgt": targets, "scalar": 2, "noise_scale": 0.0})
        loss = weighted_sum(crit(out, targets))
        loss.backward()
        crit.check_status()
        return loss.item(), [mf.grad.clone()] + [None if p.grad is None else p.grad.clone() for p in params]

    key_in, key_cat = "mask_heads.backward:gradients_in_place", "mask_heads.backward:gradients_concatenated"
    n_in, n_cat = ops.ROUTES[key_in], ops.ROUTES[key_cat]
    la, ga = run(True)
    if ops.ROUTES[key_in] != n_in + 1:
        pytest.skip("the batched head backward did not take the in-place route at this geometry")
    lb, gb = run(False)
    assert crit.last_path == "sequential" and ops.ROUTES[key_cat] == n_cat + 1
    assert abs(la - lb) <= 1e-5 * max(1.0, abs(lb))
    checked = 0
    for a, b in zip(ga, gb):
        assert (a is None) == (b is None)
        if b is None:
            continue
        assert (a - b).abs().max().item() <= 1e-3 * max(1e-6, b.abs().max().item())
        checked += 1
    assert checked > 20 and float(gb[0].abs().sum()) > 0
