"""Timing decomposition of the 3xTF32 GEMM (MPF_GEMM_DEBUG bit mask, see gemm_tf32x3.cu GemmArgs::debug):
runs the encoder GEMM shapes with parts of the kernel switched off to see which stage bounds a tile."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mp_former_b200 import native  # noqa: E402

DEV = "cuda:0"
B, S = 16, 21504
KERNELS = os.environ.get("MPF_PROBE_KERNELS", "bf16x3,tf32x3").split(",")
BNS = os.environ.get("MPF_PROBE_BNS", "").split(",")


def timeit(fn, reps=5, warm=2):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    g = torch.Generator(device=DEV).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
    flags = [0, 128, 1, 64, 2, 4, 8, 16, 32, 48, 2 | 4, 2 | 4 | 8, 2 | 4 | 48, 2 | 4 | 8 | 48]
    shapes = ((B * S, 256, 256, "value_proj"), (B * S, 288, 256, "offsets+logits"), (B * S, 1024, 256, "ffn.linear1"),
              (B * S, 256, 1024, "ffn.linear2"))
    for (m, n, k, tag) in shapes:
        a, w, bias = rn(m, k), rn(n, k) / 16, rn(n)
        for kernel in KERNELS:
            wh, wl = native.split_bf16(w) if kernel == "bf16x3" else native.split_tf32(w)
            for bn in BNS:
                if bn:
                    os.environ["MPF_GEMM_BN"] = bn
                row = {}
                for f in flags:
                    os.environ["MPF_GEMM_DEBUG"] = str(f)
                    row[f] = round(timeit(lambda: native.gemm(a, wh, wl, bias)), 4)
                print(json.dumps({"kernel": kernel, "shape": f"{tag} M={m} N={n} K={k}", "BN": bn or "auto",
                                  "ms_by_debug_flags": row}), flush=True)
        del a
    os.environ.pop("MPF_GEMM_DEBUG", None)
    os.environ.pop("MPF_GEMM_BN", None)


if __name__ == "__main__":
    main()
