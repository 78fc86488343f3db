"""MSDeformAttn micro-benchmark on one GPU: our sm_100a kernels vs the UNMODIFIED reference CUDA
kernels (oracle/_ref/libmsda_stock.so, built from /root/reference at build time), forward and
backward, BASELINE geometries, two sampling-location distributions:
  uniform : loc ~ U[0,1)           (reference ops/test.py distribution; worst case for locality)
  grid    : pixel centre + N(0, 2 texel) offsets  (what the encoder produces)
Prints one JSON object per line; also checks ours == stock numerically."""
import ctypes
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mp_former_b200 import MultiScaleDeformableAttention as MSDA  # noqa: E402

DEV = "cuda:0"
GEOMS = {
    "model_L3_1024": [(32, 32), (64, 64), (128, 128)],
    "config1_L4": [(128, 128), (64, 64), (32, 32), (16, 16)],
    "cityscapes_L3": [(32, 64), (64, 128), (128, 256)],
}


def stock():
    p = os.path.join(ROOT, "oracle", "_ref", "libmsda_stock.so")
    if not os.path.exists(p):
        return None
    lib = ctypes.CDLL(p)
    vp, i = ctypes.c_void_p, ctypes.c_int
    lib.ref_msda_forward_f32.argtypes = [vp] * 5 + [i] * 7 + [vp, vp]
    lib.ref_msda_backward_f32.argtypes = [vp] * 6 + [i] * 7 + [vp] * 3 + [vp]
    return lib


def timeit(fn, reps=20, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                      # evict L2 between timed launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)


def main():
    B = int(os.environ.get("MPF_B", "16"))
    lib = stock()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    M, D, P = 8, 32, 4
    for gname, shapes in GEOMS.items():
        L = len(shapes)
        S = sum(h * w for h, w in shapes)
        g = torch.Generator(device=DEV).manual_seed(0)
        value = torch.randn(B, S, M, D, device=DEV, generator=g)
        st = torch.as_tensor(shapes, dtype=torch.long, device=DEV)
        lsi = torch.cat((st.new_zeros((1,)), st.prod(1).cumsum(0)[:-1]))
        st_tiled = st.clone(); st_tiled._mpf_host_shapes = tuple(shapes)
        aw = torch.softmax(torch.randn(B, S, M, L * P, device=DEV, generator=g), -1).view(B, S, M, L, P)
        gout = torch.randn(B, S, M * D, device=DEV, generator=g)
        for dist in ("grid", "uniform"):
            if dist == "uniform":
                loc = torch.rand(B, S, M, L, P, 2, device=DEV, generator=g)
            else:
                ref = torch.cat([torch.stack(torch.meshgrid(
                    (torch.arange(w, device=DEV) + 0.5) / w, (torch.arange(h, device=DEV) + 0.5) / h,
                    indexing="xy"), -1).reshape(-1, 2) for h, w in shapes])          # [S,2] (x,y)
                norm = torch.tensor([[w, h] for h, w in shapes], device=DEV, dtype=torch.float32)
                off = torch.randn(B, S, M, L, P, 2, device=DEV, generator=g) * 2.0
                loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
            loc = loc.contiguous()
            alg_f = 4 * (S * M * D + 2 * S * M * L * P + S * M * L * P + S * M * D) * B
            alg_b = alg_f + 4 * (S * M * D + 3 * S * M * L * P) * B
            res = {"geom": gname, "dist": dist, "B": B, "S": S, "L": L, "alg_bytes_fwd": alg_f, "alg_bytes_bwd": alg_b}
            for tag, stt in (("ours_tiled", st_tiled), ("ours_linear", st)):
                med, best = timeit(lambda: MSDA.ms_deform_attn_forward(value, stt, lsi, loc, aw, 128))
                res[tag + "_fwd_ms"] = med
                res[tag + "_fwd_GBs"] = alg_f / med / 1e6
                res[tag + "_fwd_frac_of_measured_hbm"] = alg_f / med / 1e6 / peaks["hbm_gbs"]
                med, best = timeit(lambda: MSDA.ms_deform_attn_backward(value, stt, lsi, loc, aw, gout, 128))
                res[tag + "_bwd_ms"] = med
                res[tag + "_bwd_GBs"] = alg_b / med / 1e6
            if lib is not None:
                out = torch.zeros(B, S, M * D, device=DEV)
                gv, gl, ga = torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(aw)
                s = torch.cuda.current_stream().cuda_stream

                def f_stock():
                    out.zero_()                                    # reference host zero-fills (cuda.cu:59)
                    lib.ref_msda_forward_f32(value.data_ptr(), st.data_ptr(), lsi.data_ptr(), loc.data_ptr(),
                                             aw.data_ptr(), B, S, M, D, L, S, P, out.data_ptr(), s)

                def b_stock():
                    gv.zero_(); gl.zero_(); ga.zero_()             # cuda.cu:126-128
                    lib.ref_msda_backward_f32(gout.data_ptr(), value.data_ptr(), st.data_ptr(), lsi.data_ptr(),
                                              loc.data_ptr(), aw.data_ptr(), B, S, M, D, L, S, P,
                                              gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), s)
                res["stock_fwd_ms"], _ = timeit(f_stock)
                res["stock_bwd_ms"], _ = timeit(b_stock)
                res["speedup_fwd"] = res["stock_fwd_ms"] / res["ours_tiled_fwd_ms"]
                res["speedup_bwd"] = res["stock_bwd_ms"] / res["ours_tiled_bwd_ms"]
                f_stock(); b_stock()
                mine = MSDA.ms_deform_attn_forward(value, st_tiled, lsi, loc, aw, 128)
                mgv, mgl, mga = MSDA.ms_deform_attn_backward(value, st_tiled, lsi, loc, aw, gout, 128)
                res["max_abs_diff_vs_stock"] = {
                    "out": (mine - out).abs().max().item(), "grad_value": (mgv - gv).abs().max().item(),
                    "grad_loc": (mgl - gl).abs().max().item(), "grad_aw": (mga - ga).abs().max().item(),
                    "out_scale": out.abs().max().item(), "grad_value_scale": gv.abs().max().item()}
            print(json.dumps(res), flush=True)
        del value, aw, gout
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
