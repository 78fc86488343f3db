"""Encoder MSDeformAttn (mpf_msda_enc_*: softmax + sampling locations fused) at the bench geometry: TMA-staged tile
kernels (csrc/msda_staged.cu) vs the L1-gather kernels (csrc/msda.cu), forward and backward, for two offset
distributions -- "init": the module's initialisation (one direction per head, 1..4 texels, + small noise; what the
bench's random-init head produces), "wide": isotropic N(0, 3 texel) offsets (a trained model's spread, larger regions).
One JSON line per case; times are CUDA-event medians with an L2 flush between launches."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mp_former_b200 import MultiScaleDeformableAttention as MSDA  # noqa: E402
from mp_former_b200 import _lib  # noqa: E402

DEV = "cuda:0"


def timeit(fn, reps=12, warm=3):
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
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    lib = _lib.load()
    M, D, P = 8, 32, 4
    for B, shapes in ((16, [(32, 32), (64, 64), (128, 128)]), (2, [(32, 32), (64, 64), (128, 128)]),
                      (8, [(32, 64), (64, 128), (128, 256)])):
        L = len(shapes)
        S = sum(h * w for h, w in shapes)
        g = torch.Generator(device=DEV).manual_seed(0)
        value = torch.randn(B, S, M, D, device=DEV, generator=g)
        st = torch.as_tensor(shapes, dtype=torch.long, device=DEV)
        st._mpf_host_shapes = tuple(shapes)
        lsi = torch.cat((st.new_zeros((1,)), st.prod(1).cumsum(0)[:-1]))
        ref = torch.cat([torch.stack(torch.meshgrid((torch.arange(w, device=DEV) + 0.5) / w,
                                                    (torch.arange(h, device=DEV) + 0.5) / h, indexing="xy"),
                                     -1).reshape(-1, 2) for h, w in shapes])[None, :, None, :].expand(1, S, L, 2).contiguous()
        gy = torch.randn(B, S, M * D, device=DEV, generator=g)
        ang = torch.arange(M, device=DEV) * (2 * torch.pi / M)
        dirs = torch.stack([ang.cos(), ang.sin()], -1)
        dirs = dirs / dirs.abs().max(-1, keepdim=True).values
        bias = dirs.view(1, 1, M, 1, 1, 2) * torch.arange(1, P + 1, device=DEV).view(1, 1, 1, 1, P, 1)
        for dist in ("init", "wide"):
            noise = torch.randn(B, S, M, L, P, 2, device=DEV, generator=g)
            off = bias + 0.3 * noise if dist == "init" else 3.0 * noise
            ow = torch.cat([off.reshape(B, S, -1), torch.randn(B, S, M * L * P, device=DEV, generator=g)], -1).contiguous()
            alg_f = 4 * (S * M * D + 2 * S * M * L * P + S * M * L * P + S * M * D) * B
            alg_b = alg_f + 4 * (S * M * D + 3 * S * M * L * P) * B
            res = {"probe": "msda_enc", "B": B, "shapes": shapes, "offsets": dist, "alg_GB_fwd": alg_f / 1e9,
                   "alg_GB_bwd": alg_b / 1e9}
            outs = {}
            for mode, tag in ((1, "staged"), (0, "gather")):
                lib.mpf_msda_set_staged(mode)
                f = timeit(lambda: MSDA.ms_deform_attn_enc_forward(value, st, lsi, ow, ref, P))
                bw = timeit(lambda: MSDA.ms_deform_attn_enc_backward(value, st, lsi, ow, ref, gy, P))
                res[tag + "_fwd_ms"], res[tag + "_bwd_ms"] = f, bw
                res[tag + "_fwd_frac_hbm"] = alg_f / f / 1e6 / peaks["hbm_gbs"]
                res[tag + "_bwd_frac_hbm"] = alg_b / bw / 1e6 / peaks["hbm_gbs"]
                outs[mode] = (MSDA.ms_deform_attn_enc_forward(value, st, lsi, ow, ref, P),
                              *MSDA.ms_deform_attn_enc_backward(value, st, lsi, ow, ref, gy, P))
            lib.mpf_msda_set_staged(1)
            res["fwd_bit_identical"] = bool(torch.equal(outs[1][0], outs[0][0]))
            res["grad_value_max_abs_diff"] = (outs[1][1] - outs[0][1]).abs().max().item()
            res["grad_value_scale"] = outs[0][1].abs().max().item()
            res["grad_ow_max_abs_diff"] = (outs[1][2] - outs[0][2]).abs().max().item()
            res["speedup_fwd"], res["speedup_bwd"] = res["gather_fwd_ms"] / res["staged_fwd_ms"], \
                res["gather_bwd_ms"] / res["staged_bwd_ms"]
            print(json.dumps(res), flush=True)
        del value, gy
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
