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
