"""Runs each native kernel at the bench geometry (B images of 1024x1024, R50 head) a few times and prints
CUDA-event timings with the derived roofline numbers, one JSON object per kernel.  Also the target of the
`ncu --set full -k regex:...` captures committed under profiles/."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mp_former_b200 import MultiScaleDeformableAttention as MSDA  # noqa: E402
from mp_former_b200 import native  # noqa: E402

DEV = "cuda:0"
B = int(os.environ.get("MPF_B", "16"))
REPS = int(os.environ.get("MPF_REPS", "10"))
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
TF32_PEAK = PEAKS["bf16_tflops"] / 2.0       # dense TF32 = half the measured bf16 cuBLAS rate


def timeit(fn, reps=REPS, warm=2):
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


def emit(name, ms, bytes_=None, mma_flops=None, **kw):
    r = {"kernel": name, "ms": ms, "B": B}
    if bytes_ is not None:
        r.update(algorithmic_GB=bytes_ / 1e9, GBs=bytes_ / ms / 1e6, frac_of_measured_hbm=bytes_ / ms / 1e6 / PEAKS["hbm_gbs"])
    if mma_flops is not None:
        r.update(tensor_TFLOPs_issued=mma_flops / ms / 1e9,
                 frac_of_tf32_peak=mma_flops / ms / 1e9 / TF32_PEAK, tf32_peak_TFLOPs=TF32_PEAK)
    r.update(kw)
    print(json.dumps(r), flush=True)


def main():
    g = torch.Generator(device=DEV).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
    which = os.environ.get("MPF_PROBE", "msda,gemm,masklogits,layernorm,conv,fpn,maskbits,xattn").split(",")
    S, M, D, L, P = 21504, 8, 32, 3, 4
    shapes = [(32, 32), (64, 64), (128, 128)]
    if "msda" in which:
        value = rn(B, S, M, D)
        st = torch.as_tensor(shapes, dtype=torch.long, device=DEV)
        st._mpf_host_shapes = tuple(shapes)
        lsi = torch.cat((st.new_zeros((1,)), st.prod(1).cumsum(0)[:-1]))
        ref = torch.cat([torch.stack(torch.meshgrid((torch.arange(w, device=DEV) + 0.5) / w,
                                                    (torch.arange(h, device=DEV) + 0.5) / h, indexing="xy"), -1).reshape(-1, 2)
                         for h, w in shapes])
        norm = torch.tensor([[w, h] for h, w in shapes], device=DEV, dtype=torch.float32)
        loc = (ref[None, :, None, None, None, :] + rn(B, S, M, L, P, 2) * 2.0 / norm[None, None, None, :, None, :]).contiguous()
        aw = torch.softmax(rn(B, S, M, L * P), -1).view(B, S, M, L, P)
        gout = rn(B, S, M * D)
        alg_f = 4 * (S * M * D + 2 * S * M * L * P + S * M * L * P + S * M * D) * B
        alg_b = alg_f + 4 * (S * M * D + 3 * S * M * L * P) * B
        emit("msda_fwd_vec_kernel<8>", timeit(lambda: MSDA.ms_deform_attn_forward(value, st, lsi, loc, aw, 128)), alg_f)
        emit("msda_bwd_vec_kernel<8>", timeit(lambda: MSDA.ms_deform_attn_backward(value, st, lsi, loc, aw, gout, 128)), alg_b)
        ow = torch.cat([rn(B, S, M * L * P * 2) * 2.0, rn(B, S, M * L * P)], -1).contiguous()
        refp = ref[None, :, None, :].expand(1, S, L, 2).contiguous()
        alg_ef = 4 * (S * M * D + S * M * L * P * 3 + S * M * D) * B
        emit("msda_enc_fwd_kernel<8> (softmax+loc fused)",
             timeit(lambda: MSDA.ms_deform_attn_enc_forward(value, st, lsi, ow, refp, P)), alg_ef)
        emit("msda_enc_bwd_kernel<8> (softmax+loc fused)",
             timeit(lambda: MSDA.ms_deform_attn_enc_backward(value, st, lsi, ow, refp, gout, P)),
             alg_ef + 4 * (S * M * D + S * M * L * P * 3) * B)
        del value, loc, aw, gout, ow
    if "gemm" in which:
        for (m, n, k, tag) in ((B * S, 1024, 256, "ffn.linear1"), (B * S, 256, 1024, "ffn.linear2"),
                               (B * S, 256, 256, "value_proj"), (B * S, 288, 256, "offsets+weights")):
            a, w, bias = rn(m, k), rn(n, k) / 16, rn(n)
            wh, wl = native.split_b(w)
            ms = timeit(lambda: native.gemm(a, wh, wl, bias))
            emit(f"gemm_{native.GEMM_MODE} {tag} M={m} N={n} K={k}", ms, 4 * (m * k + m * n) + 4 * n * k, 3 * 2.0 * m * n * k,
                 fp32_equiv_TFLOPs=2.0 * m * n * k / ms / 1e9)
            del a
        # epilogue variants at the FFN-backward / residual shapes
        m = B * S
        a, w, hid = rn(m, 256), rn(1024, 256) / 16, rn(m, 1024)
        wh, wl = native.split_b(w)
        ms = timeit(lambda: native.gemm_general(a, wh, b_lo=wl, gate=hid))
        emit(f"gemm_{native.GEMM_MODE} ffn.d_hidden (gated) M={m} N=1024 K=256", ms, 4 * (m * 256 + 2 * m * 1024 + 2 * 1024 * 256),
             3 * 2.0 * m * 1024 * 256)
        w2 = rn(256, 1024) / 32
        w2h, w2l = native.split_b(w2)
        ms = timeit(lambda: native.gemm(hid, w2h, w2l))
        emit(f"gemm_{native.GEMM_MODE} ffn.d_x M={m} N=256 K=1024", ms, 4 * (m * 1024 + m * 256 + 2 * 1024 * 256), 3 * 2.0 * m * 1024 * 256)
        ms = timeit(lambda: native.matmul_tn(hid, a))
        emit(f"gemm_{native.GEMM_MODE}_tn ffn.d_w1 (token reduction, split-K) [1024 x 256] over M={m}", ms, 4 * (m * 1024 + m * 256),
             3 * 2.0 * m * 1024 * 256)
        del a, hid
    if "gemm_small" in which:
        # decoder-side products (few M tiles): B * Qt query rows, and the K / V projections of one memory level
        Qt = 220
        for (m, n, k, tag) in ((B * Qt, 256, 256, "decoder linear (queries)"), (B * Qt, 2048, 256, "decoder ffn.linear1"),
                               (B * Qt, 256, 2048, "decoder ffn.linear2"), (B * 1024, 256, 256, "K/V projection HW=1024"),
                               (B * 4096, 256, 256, "K/V projection HW=4096"), (B * 16384, 256, 256, "K/V projection HW=16384")):
            a, w, bias = rn(m, k), rn(n, k) / 16, rn(n)
            wh, wl = native.split_b(w)
            ms = timeit(lambda: native.gemm(a, wh, wl, bias))
            emit(f"gemm_{native.GEMM_MODE} {tag} M={m} N={n} K={k}", ms, 4 * (m * k + m * n) + 4 * n * k, 3 * 2.0 * m * n * k)
            gy = rn(m, n)
            ms = timeit(lambda: native.matmul_tn(gy, a))
            emit(f"gemm_{native.GEMM_MODE}_tn weight grad of {tag} [{n} x {k}] over T={m}", ms, 4 * (m * n + m * k), 3 * 2.0 * m * n * k)
            ms = timeit(lambda: native.split_b(w))
            emit(f"split_bf16 weight [{n} x {k}]", ms, 8 * n * k)
            del a, gy
    if "masklogits" in which:
        Q, C, HW = 120, 256, 65536
        e, f = rn(B, Q, C), rn(B, HW, C)
        eh, el = native.split_b(e)
        ms = timeit(lambda: native.gemm(f, eh, el, transpose_c=True))
        emit(f"gemm_{native.GEMM_MODE} mask_logits B={B} Q={Q} HW={HW}", ms, 4 * B * (HW * C + Q * HW + 2 * Q * C), 3 * 2.0 * B * Q * C * HW,
             fp32_equiv_TFLOPs=2.0 * B * Q * C * HW / ms / 1e9)
        del f
    if "layernorm" in which:
        x, r, gy = rn(B * S, 256), rn(B * S, 256), rn(B * S, 256)
        gm, bt = rn(256), rn(256)
        ms = timeit(lambda: native.add_layernorm_fwd(x, r, gm, bt, 1e-5))
        emit("add_layernorm_fwd [344064 x 256]", ms, 4 * 3 * B * S * 256)
        _, mean, rstd = native.add_layernorm_fwd(x, r, gm, bt, 1e-5)
        ms = timeit(lambda: native.add_layernorm_bwd(gy, x, r, gm, mean, rstd))
        emit("add_layernorm_bwd [344064 x 256]", ms, 4 * 4 * B * S * 256)
        ms = timeit(lambda: native.colsum(gy))
        emit("colsum [344064 x 256]", ms, 4 * B * S * 256)
        del x, r, gy
    if "conv" in which:
        # 3x3 convolution of the FPN stage on the tensor-core GEMMs ([16, 256, 256, 256] channels-last, 256 -> 256)
        C, H, W = 256, 256, 256
        x, gy = rn(B, H, W, C), rn(B, H, W, C)
        w = rn(C, C, 3, 3) / (9 * C) ** 0.5
        w_hi, w_lo = native.split_bf16(w.permute(0, 2, 3, 1).reshape(C, 9 * C))
        fl = 2.0 * B * H * W * C * 9 * C
        ms = timeit(lambda: native.conv3x3_cl(x, w_hi, w_lo))
        emit("conv3x3_cl fwd / dgrad (gemm_bf16x3_kernel, K = 2304)", ms, 4 * (2 * B * H * W * C + 9 * C * C), 3 * fl,
             fp32_equiv_TFLOPs=fl / ms / 1e9)
        ms = timeit(lambda: native.conv3x3_cl_wgrad(gy, x))
        emit("conv3x3_cl wgrad (gemm_bf16x3_tn_kernel, N = 2304)", ms, 4 * (2 * B * H * W * C + 9 * C * C), 3 * fl,
             fp32_equiv_TFLOPs=fl / ms / 1e9)
        del x, gy
    if "fpn" in which:
        # FPN-stage layout-crossing kernels at the bench geometry ([16, 256, 256, 256] maps)
        C, H, W = 256, 256, 256
        cur, prev = rn(B, H, W, C), rn(B, H // 2, W // 2, C)
        ms = timeit(lambda: native.upsample2x_add_nchw_fwd(cur, prev))
        emit("upsample2x_add_nchw_fwd [16,256,256,256]", ms, 4 * B * C * H * W * (2 + 0.25))
        gout = rn(B, C, H, W)
        ms = timeit(lambda: native.upsample2x_add_nchw_bwd(gout))
        emit("upsample2x_add_nchw_bwd [16,256,256,256]", ms, 4 * B * C * H * W * (2 + 0.25))
        del cur, prev
        x3, gm, bt = gout.view(B, C, H * W), rn(C), rn(C)
        cur, prev = rn(B, H, W, C), rn(B, H // 2, W // 2, C)
        ms = timeit(lambda: native.upsample2x_add_cl_fwd(cur, prev))
        emit("upsample2x_add_cl_fwd [16,256,256,256]", ms, 4 * B * C * H * W * (2 + 0.25))
        ms = timeit(lambda: native.upsample2x_cl_bwd(cur))
        emit("upsample2x_cl_bwd [16,256,256,256]", ms, 4 * B * C * H * W * (1 + 0.25))
        del cur, prev
        ms = timeit(lambda: native.groupnorm_nchw2cl_fwd(x3, gm, bt, 1e-5, 32, True))
        emit("groupnorm_nchw2cl_fwd (stats + apply) [16,256,65536]", ms, 4 * B * C * H * W * 3)
        y, mean, rstd = native.groupnorm_nchw2cl_fwd(x3, gm, bt, 1e-5, 32, True)
        ms = timeit(lambda: native.groupnorm_nchw2cl_bwd(y, x3, gm, bt, mean, rstd, 32, True))
        emit("groupnorm_nchw2cl_bwd (stats + apply) [16,256,65536]", ms, 4 * B * C * H * W * 5)
        ms = timeit(lambda: native.groupnorm_cl_fwd(y, gm, bt, 1e-5, 32, False))
        emit("groupnorm_cl_fwd (stats + apply) [16,65536,256]", ms, 4 * B * C * H * W * 3)
        del x3, y, gout
    if "maskbits" in which:
        logits = rn(B, 120, 256, 256)
        for hw in (32, 64, 128):
            ms = timeit(lambda: native.attn_mask_bits(logits, (hw, hw)))
            emit(f"attn_mask_bits 256->{hw}", ms, None, None, outputs_per_us=B * 120 * hw * hw / ms / 1e3)
        del logits
    if "xattn" in which:
        E, heads, Qt = 256, 8, 120
        for HW in ([int(os.environ["MPF_HW"])] if os.environ.get("MPF_HW") else [1024, 4096, 16384]):
            q, k, vt = rn(B, Qt, E), rn(B, HW, E), rn(B, E, HW)
            qh, ql = native.split_tf32(q); kh, kl = native.split_tf32(k); vh, vl = native.split_tf32(vt)
            bits = native.pack_bool_bits(torch.rand(B, Qt, HW, device=DEV, generator=g) < 0.7)
            ms = timeit(lambda: native.masked_xattn_fwd(qh, ql, kh, kl, vh, vl, bits, None, heads))
            mma = 3 * 2.0 * B * heads * 128 * HW * 32 * 2          # 128-row tiles, S and PV, 3 passes
            emit(f"masked_xattn_fwd B={B} Qt={Qt} HW={HW}", ms, 4 * B * (4 * HW * E + 2 * Qt * E) + B * Qt * HW // 8, mma,
                 useful_fp32_TFLOPs=4.0 * B * Qt * HW * E / ms / 1e9)

    if "xattn_bwd" in which:
        E, heads, Qt = 256, 8, 120
        for HW in (4096, 16384):
            q, k, v = rn(B, Qt, E), rn(B, HW, E), rn(B, HW, E)
            qh, ql = native.split_tf32(q); kh, kl = native.split_tf32(k); vh, vl = native.split_tf32(v)
            kth, ktl = native.split_tf32(k.transpose(1, 2).contiguous())
            bits = native.pack_bool_bits(torch.rand(B, Qt, HW, device=DEV, generator=g) < 0.7)
            d_o = rn(B, Qt, E)
            lse2, delta = rn(B, heads, Qt).abs() + 6.0, rn(B, heads, Qt) * 0.01
            ms = timeit(lambda: native.masked_xattn_bwd(qh, ql, kh, kl, kth, ktl, vh, vl, d_o, bits, None, lse2, delta, heads))
            mma = 3 * 2.0 * B * heads * 128 * HW * 32 * 5          # S, dP (dq kernel) ; S^T, dP^T, dV, dK (dkv kernel) ~ 5 products
            emit(f"masked_xattn_bwd (dq + dkv kernels, host-side transposes included) B={B} Qt={Qt} HW={HW}", ms,
                 4 * B * (8 * HW * E + 2 * HW * E) + B * Qt * HW // 8, mma)


if __name__ == "__main__":
    main()
