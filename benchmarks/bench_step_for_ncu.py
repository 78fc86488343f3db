"""One eager training step of the bench workload for `ncu` launch lists (`--metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`): two warm-up steps, then ONE step bracketed by
cudaProfilerStart/Stop (run ncu with `--profile-from-start off`).  Same head, inputs and criterion as `bench.py`'s
default line (MPF_B images, default 16)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mp_former_b200 import native, workload  # noqa: E402

DEV = "cuda:0"
B = int(os.environ.get("MPF_B", "16"))
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
pd, dec = workload.build_head(device=DEV)
feats = workload.synthetic_features(B, device=DEV)
targets = workload.synthetic_targets(B, device=DEV)
dn = {"tgt": targets, "scalar": 1, "noise_scale": 0.0}
criterion, weighted_sum = workload.build_criterion(device=DEV)
criterion.train(True)
params = list(pd.parameters()) + list(dec.parameters())
wcache = native.set_weight_cache(native.WeightOperandCache(params))


def step():
    for p in params:
        p.grad = None
    wcache.begin_step()
    mf, _, ms = pd.forward_features(feats)
    loss = weighted_sum(criterion(dec(ms, mf, None, dn), targets))
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
