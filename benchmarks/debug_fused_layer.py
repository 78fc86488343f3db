"""Debug aid: fused encoder layer vs module path, run-to-run determinism (prints worst parameter-gradient gaps)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
from mp_former_b200 import ops  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402
from test_host_logic_cpu import build_pixel_decoder  # noqa: E402
from test_oracle_vs_golden import pixel_decoder_template  # noqa: E402

DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = bool(int(os.environ.get("TF32", "1")))
pd = build_pixel_decoder().to(DEV)
pd.load_state_dict(O.seeded_state_dict(pixel_decoder_template(), seed=43))
feats = {k: v.to(DEV).requires_grad_(True) for k, v in cases.pixel_decoder_features().items()}


def run(only_mf=False):
    for p in pd.parameters():
        p.grad = None
    for v in feats.values():
        v.grad = None
    mf, enc, ms = pd.forward_features(feats)
    outs = [mf] if only_mf else [mf, *ms]
    w = [torch.randn(t.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(i)) for i, t in enumerate(outs)]
    sum((t * wi).sum() for t, wi in zip(outs, w)).backward()
    g = {n: p.grad.clone() for n, p in pd.named_parameters()}
    g.update({"d/d" + k: v.grad.clone() for k, v in feats.items()})
    return [t.detach().clone() for t in [mf, *ms]], g


def gap(a, b, tag):
    errs = {n: ((a[n] - b[n]).abs().max().item() / max(1e-3, b[n].abs().max().item())) for n in b}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print(tag, [(n, f"{e:.2e}") for n, e in worst], flush=True)


for only_mf in (False, True):
    print("== loss on", "mask_features only" if only_mf else "all outputs")
    ops.NO_FUSED_ENCODER_LAYER = False
    of1, f1 = run(only_mf)
    of2, f2 = run(only_mf)
    ops.NO_FUSED_ENCODER_LAYER = True
    om1, m1 = run(only_mf)
    om2, m2 = run(only_mf)
    gap(f1, f2, "fused vs fused   ")
    gap(m1, m2, "module vs module ")
    gap(f1, m1, "fused vs module  ")
    print("forward max abs diff", [f"{(a - b).abs().max().item():.2e}" for a, b in zip(of1, om1)])
