"""torch.profiler breakdown of one bench step (which aten ops / custom Functions own the device time)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile, record_function

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import pseudo_loss  # noqa: E402
from mp_former_b200 import workload  # noqa: E402

DEV = "cuda:0"
B = int(os.environ.get("MPF_B", "16"))
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
pd, dec = workload.build_head(device=DEV)
feats = workload.synthetic_features(B, device=DEV)
dn = {"tgt": workload.synthetic_targets(B, device=DEV), "scalar": 1, "noise_scale": 0.0}
params = list(pd.parameters()) + list(dec.parameters())
loss_of = pseudo_loss
if os.environ.get("MPF_LOSS", "criterion") == "criterion":
    criterion, weighted_sum = workload.build_criterion(device=DEV)
    criterion.train(True)
    loss_of = lambda out: weighted_sum(criterion(out, dn["tgt"]))  # noqa: E731


def step():
    for p in params:
        p.grad = None
    with record_function("FWD_pixel_decoder"):
        mf, _, ms = pd.forward_features(feats)
    with record_function("FWD_decoder"):
        out = dec(ms, mf, None, dn)
    with record_function("LOSS"):
        loss = loss_of(out)
    with record_function("BWD"):
        loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=bool(os.environ.get("MPF_SHAPES"))) as prof:
    step()
    torch.cuda.synchronize()
if os.environ.get("MPF_KERNELS"):
    # every device kernel of the step, one line each: total us, calls, name (complete, unlike the top-N tables)
    rows = [(e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0]
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows if not r[2].startswith(("FWD_", "BWD", "LOSS")))
    print(f"# total self device time (kernels + memcpy/memset): {tot / 1e3:.3f} ms")
    for us, n, k in rows:
        print(f"{us:10.1f}\t{n:5d}\t{k[:150]}")
elif os.environ.get("MPF_SHAPES"):
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=60,
                                                             max_name_column_width=40, max_shapes_column_width=90))
else:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
