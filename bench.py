#!/usr/bin/env python
"""bench.py -- img/s (forward+backward) of the MP-Former hot path on N B200s, one JSON line.

A "step" is one training pass of the hot path -- MSDeformAttn pixel decoder + masked-attention transformer decoder with
the mask-piloted (DN) query group, the recipe's SetCriterion + HungarianMatcher on top, backward to every parameter of
the head, gradient all-reduce -- over one batch of synthetic backbone features (the backbone is upstream of the path).

  python bench.py [--gpus N] [--steps K] [--warmup W]     our arm; default preset = BASELINE configs[1]/[2]:
                                                          R50, 1024x1024, GLOBAL batch 16 split over the N ranks
  python bench.py --config {1,2,3,4,5}                    another BASELINE.json configuration (1 = MSDeformAttn alone)
  python bench.py --weak                                  16 images per GPU instead of 16 in total (round-1 mode)
  python bench.py --loss pseudo                           linear pseudo-loss instead of the criterion (round-1 mode)
  python bench.py --impl reference [...]                  the reference's CPU path (oracle port) on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...       one rank per GPU

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
The reference arm imports nothing of the product package (oracle/ only).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec fwd+bwd, MP-Former head (MSDeformAttn pixel decoder + masked decoder + criterion), synthetic"
UNIT = "img/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4, 5],
                    help="BASELINE.json configs[i-1]; 0 (default) = 2 on one GPU / 3 on several (same head, global "
                         "batch 16)")
    ap.add_argument("--weak", action="store_true", help="weak scaling: the global batch of the preset PER GPU")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (overrides the preset's split)")
    ap.add_argument("--loss", default="criterion", choices=["criterion", "pseudo"])
    ap.add_argument("--no-dn", action="store_true", help="drop the mask-piloted (DN) query group")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weight-cache", action="store_true", help="split every weight operand where it is used")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-stock", action="store_true")
    return ap.parse_args()


def world_info():
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def resolve(a, world):
    """Preset -> workload description shared by both arms (identical `config` objects)."""
    from oracle import synthetic as SY
    idx = a.config if a.config else (2 if world == 1 else 3)
    p = dict(SY.PRESETS[idx])
    gb = p["global_batch"]
    if a.batch:
        per_gpu, scaling = a.batch, "weak"
    elif a.weak:
        per_gpu, scaling = gb, "weak"
    else:
        if gb % world:
            raise SystemExit(f"global batch {gb} does not split over {world} ranks")
        per_gpu, scaling = gb // world, "strong"
    p.update(per_gpu=per_gpu, scaling=scaling, preset=idx)
    loss = ("SetCriterion + HungarianMatcher (recipe weights 2/5/5, eos 0.1, 12544 points, 10 heads + dn)"
            if a.loss == "criterion" else "linear pseudo-loss")
    p["config"] = {
        "workload": (f"{p['name']}: backbone maps -> MSDeformAttnPixelDecoder(6 layers) + "
                     f"MultiScaleMaskedTransformerDecoderMaskDN(9 layers, {p['queries']} queries, "
                     f"{'DN points' if not a.no_dn else 'no DN'}), fwd+bwd, {loss}"),
        "preset": f"BASELINE.json configs[{idx - 1}]", "images_per_gpu": per_gpu, "global_batch": per_gpu * world,
        "parallelism": f"dp{world}" if world > 1 else "single", "loss": a.loss,
        "l2": "inputs + activations of a step (> 250 MB per image) exceed the 126 MB L2; no explicit flush",
    }
    return p


def pseudo_loss(out):
    terms = [out["pred_masks"].float().mean(), out["pred_logits"].float().mean()]
    for a in out["aux_outputs"]:
        terms += [a["pred_masks"].float().mean(), a["pred_logits"].float().mean()]
    dn = out.get("dn_out")
    if dn is not None:
        terms += [dn["pred_masks"].float().mean(), dn["pred_logits"].float().mean()]
        for a in dn["aux_outputs"]:
            terms += [a["pred_masks"].float().mean(), a["pred_logits"].float().mean()]
    return sum(terms)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path on the host cores.
# Imports oracle/ only.
# ------------------------------------------------------------------------------------------------
def cpu_port_step_fn(a, p):
    """Returns (fn, sample_description): fn() runs ONE image fwd+bwd through the CPU oracle."""
    import torch
    from oracle import synthetic as SY
    from oracle import torch_oracle as O
    torch.set_num_threads(min(32, os.cpu_count() or 1))   # >32 threads slow the gather-heavy port down
    psd, dsd = SY.head_state_dicts(p["backbone"], p["queries"], p["classes"], seed=0)
    psd = {k: v.requires_grad_(v.is_floating_point()) for k, v in psd.items()}
    dsd = {k: v.requires_grad_(v.is_floating_point()) for k, v in dsd.items()}
    feats = SY.features(1, p["height"], p["width"], p["backbone"])
    tg = SY.targets(1, p["height"], p["width"], p["classes"])
    dn = None if a.no_dn else {"tgt": tg, "scalar": 1, "noise_scale": 0.0}
    loss_of = pseudo_loss
    if a.loss == "criterion":
        from oracle import criterion_oracle as CO
        wd = SY.recipe_weight_dict()

        def loss_of(out):
            losses = CO.set_criterion(out, tg, num_classes=p["classes"], eos_coef=0.1, losses=["labels", "masks"],
                                      num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75,
                                      cost_class=2.0, cost_mask=5.0, cost_dice=5.0, training=True)
            return sum(v * wd[k] for k, v in losses.items() if k in wd)

    def fn():
        mf, _, ms = O.pixel_decoder_forward(psd, feats)
        out = O.decoder_forward(dsd, ms, mf, num_queries=p["queries"], num_classes=p["classes"], dn_args=dn,
                                dn_label_noise_ratio=0.2)
        loss = loss_of(out)
        loss.backward()
        for sd in (psd, dsd):
            for v in sd.values():
                v.grad = None
        return float(loss.detach())

    return fn, "1 image of the same configuration fwd+bwd per step through oracle/torch_oracle.py (+ criterion_oracle.py)"


def config1_inputs(device="cpu"):
    import torch
    shapes = [(128, 128), (64, 64), (32, 32), (16, 16)]              # S = Lq = 21760 (BASELINE configs[0])
    g = torch.Generator().manual_seed(3)
    S = sum(h * w for h, w in shapes)
    value = torch.rand(1, S, 8, 32, generator=g) * 0.01                # distributions of ref ops/test.py:36-39
    loc = torch.rand(1, S, 8, 4, 4, 2, generator=g)
    aw = torch.rand(1, S, 8, 4, 4, generator=g) + 1e-5
    aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return shapes, value.to(device), loc.to(device), aw.to(device)


def msda_alg_bytes(B, S, Lq, M, D, L, P, backward=False):
    fwd = 4 * (S * M * D + 2 * Lq * M * L * P + Lq * M * L * P + Lq * M * D) * B
    if not backward:
        return fwd
    return fwd + 4 * (S * M * D + 3 * Lq * M * L * P) * B


def cpu_msda_baseline(shapes, reps=5, backward=False):
    """The reference's CPU MSDeformAttn path (``ms_deform_attn_core_pytorch``, ref ops/functions/ms_deform_attn_func.py:
    52-72, restated in oracle/torch_oracle.py::msda_core_grid_sample) on the host cores: ONE image, one call over
    S = Lq = all pixels of ``shapes``, M=8, D=32, P=4."""
    import torch
    from oracle import torch_oracle as O
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    S = sum(h * w for h, w in shapes)
    M_, D, L, P = 8, 32, len(shapes), 4
    g = torch.Generator().manual_seed(3)
    value = (torch.rand(1, S, M_, D, generator=g) * 0.01).requires_grad_(backward)
    loc = torch.rand(1, S, M_, L, P, 2, generator=g).requires_grad_(backward)
    aw = torch.rand(1, S, M_, L, P, generator=g) + 1e-5
    aw = (aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)).requires_grad_(backward)

    def call():
        y = O.msda_core_grid_sample(value, shapes, loc, aw)
        if backward:
            y.sum().backward()
            value.grad = loc.grad = aw.grad = None
    with torch.set_grad_enabled(backward):
        for _ in range(2):
            call()
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        ms = (time.perf_counter() - t0) * 1e3 / reps
    alg = msda_alg_bytes(1, S, S, M_, D, L, P, backward)
    return {"kernel": "ms_deform_attn_core_pytorch (reference CPU path, per-level F.grid_sample), "
                      + ("forward+backward" if backward else "forward"),
            "ms_per_image_per_layer": ms, "achieved": alg / (ms / 1e3) / 1e9, "unit": "GB/s (algorithmic bytes)",
            "algorithmic_bytes": alg, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 image, S=Lq={S}, L={L}, M=8, D=32, P=4, {reps} calls"}


def run_reference_arm(a):
    world, rank, _ = world_info()
    if rank != 0:
        return
    import torch
    if a.config == 1:
        shapes = config1_inputs()[0]
        t0 = time.perf_counter()
        r = cpu_msda_baseline(shapes, reps=max(1, a.steps), backward=True)
        print(json.dumps({
            "impl": "reference", "metric": "MSDeformAttn fwd+bwd, BASELINE configs[0] (B=1, L=4, S=Lq=21760), "
            "algorithmic GB/s", "value": r["achieved"], "unit": "GB/s", "n_gpus": a.gpus, "steps": max(1, a.steps),
            "warmup": 2, "ms_per_step": r["ms_per_image_per_layer"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": "config1 MSDeformAttn"},
            "cpu_baseline": {"value": r["achieved"], "unit": "GB/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["achieved"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}), flush=True)
        return
    p = resolve(a, world)
    fn, sample = cpu_port_step_fn(a, p)
    for _ in range(a.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    dt = time.perf_counter() - t0
    v = a.steps / dt
    cores = torch.get_num_threads()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": p["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": p["config"],
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, path):
        self.path, self.proc, self.idx = path, None, gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            q = [x.strip() for x in line.split(",")]
            if len(q) < 9:
                continue
            try:
                sm.append(float(q[1])); mx.append(float(q[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), q[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(pk)) if os.path.exists(pk) else {}
    src = "MEASURED_PEAKS.json" if d else "fallback of B200_PROFILING.md (6650 GB/s, 1400 TFLOP/s sustained)"
    return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), src


def run_config1(a):
    """BASELINE configs[0] on the GPU: MSDeformAttn forward+backward at L=4, S=Lq=21760 through the reference's
    extension-module API (mp_former_b200.MultiScaleDeformableAttention), checked against the CPU oracle in place."""
    import torch
    from mp_former_b200 import MultiScaleDeformableAttention as MSDA
    from mp_former_b200 import _lib
    from oracle import torch_oracle as O
    dev = torch.device("cuda", 0)
    shapes, value, loc, aw = config1_inputs(dev)
    st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    st._mpf_host_shapes = tuple(shapes)
    lsi = torch.cat((st.new_zeros((1,)), st.prod(1).cumsum(0)[:-1]))
    gy = torch.randn(1, value.shape[1], 256, device=dev)
    ref = O.msda_core(value.cpu(), shapes, loc.cpu(), aw.cpu())
    err = float((MSDA.ms_deform_attn_forward(value, st, lsi, loc, aw, 128).cpu() - ref).abs().max())
    assert err < 1e-6, err

    def step():
        MSDA.ms_deform_attn_forward(value, st, lsi, loc, aw, 128)
        MSDA.ms_deform_attn_backward(value, st, lsi, loc, aw, gy, 128)
    for _ in range(max(a.warmup, 3)):
        step()
    sampler = ClockSampler(0, os.path.join(ROOT, "gpurun_out", "clocks_rank0.csv"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    sampler.start()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2: written between timed steps
    l0 = _lib.launch_count()
    ms = []
    for _ in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    launches = _lib.launch_count() - l0
    clocks = sampler.stop()
    hbm, _, src = peaks()
    S = value.shape[1]
    alg = msda_alg_bytes(1, S, S, 8, 32, 4, 4, True) + msda_alg_bytes(1, S, S, 8, 32, 4, 4, False)
    t = statistics.mean(ms)
    cpu = None if a.no_cpu_baseline else cpu_msda_baseline(shapes, reps=5, backward=True)
    # e2e: host tensors in, host gradients out, through the same extension-module API
    hv, hl, ha, hg = (x.cpu().pin_memory() for x in (value, loc, aw, gy))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        v, l, w, g = (x.to(dev, non_blocking=True) for x in (hv, hl, ha, hg))
        y = MSDA.ms_deform_attn_forward(v, st, lsi, l, w, 128)
        gv, gl, ga = MSDA.ms_deform_attn_backward(v, st, lsi, l, w, g, 128)
        outs = [x.cpu() for x in (y, gv, gl, ga)]
    e1.record()
    torch.cuda.synchronize()
    te = e0.elapsed_time(e1) / a.steps
    h2d = sum(x.numel() * 4 for x in (hv, hl, ha, hg))
    d2h = sum(x.numel() * 4 for x in outs)
    print(json.dumps({
        "metric": "MSDeformAttn fwd+bwd, BASELINE configs[0] (B=1, L=4, S=Lq=21760), algorithmic GB/s",
        "value": alg / (t / 1e3) / 1e9, "unit": "GB/s", "n_gpus": 1, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": "config1 MSDeformAttn", "l2": "256 MB written between timed steps"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": alg / (te / 1e3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "roofline": {"bound": "hbm", "achieved": alg / (t / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": alg / (t / 1e3) / 1e9 / hbm, "traffic": None, "peak_source": src,
                     "kernel": "msda_fwd_vec_kernel<8> + msda_bwd_vec_kernel<8> (one image: 170 CTAs x 8 heads)",
                     "parity_max_abs_err_vs_oracle": err},
        "cpu_baseline": None if cpu is None else {"value": cpu["achieved"], "unit": "GB/s", "cores": cpu["cores"],
                                                  "kind": "port", "sample": cpu["sample"]}}), flush=True)


def stock_forward_block(pd, dec, p, feats, B):
    """Forward only, eval, no DN: this package vs the STOCK CUDA path -- the reference's eager module code (the
    oracle's restatement run on CUDA tensors: same library calls) with the UNMODIFIED reference MSDeformAttn CUDA
    kernel (oracle/_ref/libmsda_stock.so, built from /root/reference's .cuh at build time) plugged in.  A second stock
    variant runs the decoder under fp16 autocast, as the reference recipe trains it (SOLVER.AMP.ENABLED, pixel decoder
    forced to fp32 at msdeformattn.py:314)."""
    import torch
    from oracle import torch_oracle as O
    so = os.path.join(ROOT, "oracle", "_ref", "libmsda_stock.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/libmsda_stock.so not built"}
    lib = ctypes.CDLL(so)
    vp, i = ctypes.c_void_p, ctypes.c_int
    lib.ref_msda_forward_f32.argtypes = [vp] * 5 + [i] * 7 + [vp, vp]

    def stock_core(value, shapes, loc, aw):
        N, S, Mh, D = value.shape
        Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
        st = torch.as_tensor(shapes, dtype=torch.long, device=value.device)
        lsi = torch.cat((st.new_zeros((1,)), st.prod(1).cumsum(0)[:-1]))
        out = torch.zeros(N, Lq, Mh * D, device=value.device)         # reference zero-fills (cuda.cu:59)
        value, loc, aw = value.float().contiguous(), loc.float().contiguous(), aw.float().contiguous()
        rc = lib.ref_msda_forward_f32(value.data_ptr(), st.data_ptr(), lsi.data_ptr(), loc.data_ptr(),
                                      aw.data_ptr(), N, S, Mh, D, L, Lq, P, out.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        return out

    def timeit(fn, reps):
        fn()
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    psd = {k: v.detach() for k, v in pd.state_dict().items()}
    dsd = {k: v.detach() for k, v in dec.state_dict().items()}
    was_training = pd.training, dec.training
    pd.eval(), dec.eval()
    orig = O.msda_core
    O.msda_core = stock_core
    try:
        with torch.no_grad():
            def ours():
                mf, _, ms = pd.forward_features(feats)
                return dec(ms, mf)

            def stock(amp=False):
                mf, _, ms = O.pixel_decoder_forward(psd, feats)
                with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
                    return O.decoder_forward(dsd, ms, mf, num_queries=p["queries"], num_classes=p["classes"])
            a_, b_ = ours(), stock()
            r0 = ((a_["aux_outputs"][0]["pred_masks"] - b_["aux_outputs"][0]["pred_masks"]).abs()
                  / b_["aux_outputs"][0]["pred_masks"].abs().clamp(min=1.0)).max().item()
            del a_, b_
            t_ours, t_stock, t_amp = timeit(ours, 5), timeit(stock, 3), timeit(lambda: stock(True), 3)
    finally:
        O.msda_core = orig
        pd.train(was_training[0]), dec.train(was_training[1])
    torch.cuda.empty_cache()
    return {"what": "forward only (eval, no DN), TF32 off; stock = reference eager modules on CUDA + the unmodified "
                    "reference MSDeformAttn CUDA kernel", "images": B, "ours_ms": t_ours, "stock_fp32_ms": t_stock,
            "stock_decoder_fp16_amp_ms": t_amp, "speedup_vs_stock_fp32": t_stock / t_ours,
            "speedup_vs_stock_decoder_fp16_amp": t_amp / t_ours, "head0_pred_masks_max_rel_diff": r0}


def run_ours(a):
    import torch
    import torch.distributed as dist
    import mp_former_b200 as M
    from mp_former_b200 import MultiScaleDeformableAttention as MSDA
    from mp_former_b200 import _lib, graphs, native, workload

    world, rank, local = world_info()
    assert torch.cuda.is_available(), "bench.py (our arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False          # fp32 path, like the reference's pixel decoder
    torch.backends.cudnn.allow_tf32 = False

    p = resolve(a, world)
    B, H, W = p["per_gpu"], p["height"], p["width"]
    pd, dec = workload.build_head(backbone=p["backbone"], num_queries=p["queries"], num_classes=p["classes"],
                                  device=dev, seed=0)

    class Head(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.pixel_decoder, self.predictor = pd, dec

        def forward(self, feats, dn_args):
            mf, _, ms = self.pixel_decoder.forward_features(feats)
            return self.predictor(ms, mf, None, dn_args)

    head = Head()
    params = [q for q in head.parameters()]
    feats = workload.synthetic_features(B, H, W, backbone=p["backbone"], seed=rank, device=dev)
    targets = workload.synthetic_targets(B, H, W, num_classes=p["classes"], seed=rank, device=dev)
    dn_args = None if a.no_dn else {"tgt": targets, "scalar": 1, "noise_scale": 0.0}

    loss_of, criterion = pseudo_loss, None
    if a.loss == "criterion":
        criterion, weighted_sum = workload.build_criterion(num_classes=p["classes"], device=dev)
        criterion.train(True)

        def loss_of(out):
            return weighted_sum(criterion(out, targets))

    # Gradient all-reduce (the path's only collective): one flat NCCL all-reduce of the head's
    # gradients (~20 M parameters) after the backward, averaged over ranks like DistributedDataParallel.
    def allreduce_grads():
        graphs.allreduce_gradients(params, world)

    # Split weight operands of the step refreshed by one launch at its start (native.WeightOperandCache) instead of a
    # split + transposing copy per use (--no-weight-cache: the per-use path).
    wcache = None if a.no_weight_cache else native.set_weight_cache(native.WeightOperandCache(params))

    def forward_loss(f):
        if wcache is not None:
            wcache.begin_step()
        return loss_of(head(f, dn_args))

    def eager_step(f):
        for q in params:
            q.grad = None
        loss = forward_loss(f)
        loss.backward()
        allreduce_grads()
        return loss

    for _ in range(max(a.warmup, 3)):
        eager_step(feats)
    torch.cuda.synchronize()
    if criterion is not None:
        criterion.check_status()

    # One CUDA graph for forward + loss + backward (mp_former_b200/graphs.py).  Falls back to eager stepping (and says
    # so) if capture is refused.
    gs, graph_note = None, "disabled (--no-graph)"
    if not a.no_graph:
        try:
            gs = graphs.GraphedStep(forward_loss, feats, params, warmup=2)
            graph_note = "forward+loss+backward captured once, replayed per step"
        except Exception as e:  # noqa: BLE001
            gs, graph_note = None, f"capture failed, eager stepping: {type(e).__name__}: {str(e)[:160]}"
            torch.cuda.synchronize()

    def step(f=None):
        """One step on device-resident inputs (f None: the graph's static input buffers)."""
        if gs is None:
            return eager_step(feats if f is None else f)
        if f is not None:
            gs.load_inputs(f)
        loss = gs.replay()
        graphs.allreduce_gradients(params, world, flat=gs.flat_grad)
        return loss

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    sync()

    # ---- timed region: device-resident inputs ------------------------------------------------
    sampler = ClockSampler(local, os.path.join(ROOT, "gpurun_out", f"clocks_rank{rank}.csv")) if rank == 0 else None
    if sampler:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        sampler.start()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    sync()
    ms_total = ev0.elapsed_time(ev1)
    launches = (_lib.launch_count() - l0) if gs is None else gs.launches_per_replay * a.steps
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = B * world * a.steps / (ms_total / 1e3)

    # ---- e2e: host (pinned) inputs, H2D inside the timed region, loss read back ---------------
    e2e = None
    if not a.no_e2e:
        host = workload.synthetic_features(B, H, W, backbone=p["backbone"], seed=rank, device="cpu", pin=True)
        h2d = sum(t_.numel() * t_.element_size() for t_ in host.values())
        # Every step's inputs come from pinned host memory; the copy of step i+1 runs on a side stream while
        # step i computes (double buffering, as a data loader would), and each step's loss is read back.
        copy_stream = torch.cuda.Stream(device=dev)

        def upload():
            with torch.cuda.stream(copy_stream):
                f = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return f, ev

        def e2e_run(n):
            nxt = upload()
            last = 0.0
            for i in range(n):
                f, ev = nxt
                torch.cuda.current_stream().wait_event(ev)
                if i + 1 < n:
                    nxt = upload()
                loss = step(f)
                for t_ in f.values():
                    t_.record_stream(torch.cuda.current_stream())
                last = float(loss.item())            # D2H read of the step's result
            return last
        e2e_run(2)
        sync()
        n_e2e = max(3, min(a.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(n_e2e)
        e1.record()
        sync()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * n_e2e / (float(te.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": n_e2e}

    # ---- per-kernel device times (CUDA events around each launch) over two eager steps --------------------------
    # taken outside the timed region: events cannot be read inside a captured graph, and they serialise nothing
    n_prof = 2
    MSDA.profile_begin()
    native.profile_begin()
    for _ in range(n_prof):
        eager_step(feats)
    prof = MSDA.profile_end()
    gprof = native.profile_end()

    if rank == 0:
        hbm, tens, src = peaks()
        step_ms = ms_total / a.steps

        def ncu_traffic(kernel):
            """DRAM bytes per launch of `kernel` (dram__bytes_read.sum + dram__bytes_write.sum, mean over its launches
            in one training step) from the committed ncu launch list of THIS workload (config 2, 16 images, criterion:
            profiles/r2E_launch_summary_step_b16.txt, made by scripts/gpu_r2_final1.sh); None for other workloads --
            ncu cannot run inside the timed bench."""
            path = os.path.join(ROOT, "profiles", "r2E_launch_summary_step_b16.txt")
            cfg_idx = a.config if a.config else (2 if world == 1 else 3)
            if cfg_idx != 2 or B != 16 or a.loss != "criterion" or not os.path.exists(path):
                return None
            for ln in open(path):
                f = ln.split()
                if len(f) >= 7 and not ln.startswith("#") and kernel in " ".join(f[6:]):
                    return (float(f[4]) + float(f[5])) * 1e6
            return None

        def gemm_roof(kind, what, mma_per_product=3.0):
            r = gprof.get(kind)
            if not r or r["ms"] <= 0:
                return None
            n = r["launches"]
            tf = r["flops"] / (r["ms"] / 1e3) / 1e12
            gb = r["bytes"] / (r["ms"] / 1e3) / 1e9
            # which ceiling binds this mix of shapes: time at HBM peak for the algorithmic bytes vs time at the
            # tensor peak for the MMAs the split arithmetic issues; the larger one is the roofline
            t_hbm = r["bytes"] / (hbm * 1e9)
            t_tensor = mma_per_product * r["flops"] / (tens * 1e12)
            common = {"kernel": f"{kind} ({what})", "traffic": ncu_traffic(kind),
                      "tensor_TFLOPs_algorithmic": tf, "tensor_frac_algorithmic": tf / tens,
                      "tensor_issue_frac": mma_per_product * tf / tens, "achieved_hbm_GBs": gb, "hbm_frac": gb / hbm,
                      "frac_of_combined_roof": max(t_hbm, t_tensor) / (r["ms"] / 1e3),
                      "ideal_ms_hbm": t_hbm * 1e3 / n, "ideal_ms_tensor": t_tensor * 1e3 / n,
                      "algorithmic_flops_per_launch": r["flops"] / n, "algorithmic_bytes_per_launch": r["bytes"] / n,
                      "avg_launch_ms": r["ms"] / n, "launches_timed": n,
                      "share_of_step": r["ms"] / n_prof / step_ms}
            if t_hbm >= t_tensor:
                return {"bound": "hbm", "achieved": gb, "peak": hbm, "unit": "GB/s", "frac": gb / hbm,
                        "peak_source": src + " hbm_gbs", **common}
            return {"bound": "tensor", "achieved": tf, "peak": tens, "unit": "TFLOP/s", "frac": tf / tens,
                    "peak_source": src + " bf16_tflops_sustained", **common}

        # MSDeformAttn: algorithmic bytes per launch
        S = sum((H // s) * (W // s) for s in (32, 16, 8))
        alg = msda_alg_bytes(B, S, S, 8, 32, 3, 4)
        fwd_ms = statistics.mean(prof["fwd_ms"]) if prof["fwd_ms"] else None
        bwd_ms = statistics.mean(prof["bwd_ms"]) if prof["bwd_ms"] else None
        msda = None
        if fwd_ms:
            ach = alg / (fwd_ms / 1e3) / 1e9
            msda = {"kernel": "msda_enc_fwd_kernel<8> (MSDeformAttn forward, softmax+locations fused)",
                    "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "traffic": ncu_traffic("msda_enc_fwd_kernel"),
                    "peak_source": src + " hbm_gbs", "algorithmic_bytes_per_launch": alg, "avg_launch_ms": fwd_ms,
                    "launches_timed": len(prof["fwd_ms"]),
                    "share_of_step": fwd_ms * len(prof["fwd_ms"]) / n_prof / step_ms}
            if bwd_ms:
                alg_b = msda_alg_bytes(B, S, S, 8, 32, 3, 4, True)
                msda["backward"] = {"kernel": "msda_enc_bwd_kernel<8> (+ grad_value memset)", "avg_launch_ms": bwd_ms,
                                    "achieved": alg_b / (bwd_ms / 1e3) / 1e9,
                                    "frac": alg_b / (bwd_ms / 1e3) / 1e9 / hbm,
                                    "algorithmic_bytes_per_launch": alg_b,
                                    "traffic": ncu_traffic("msda_enc_bwd_kernel"),
                                    "share_of_step": bwd_ms * len(prof["bwd_ms"]) / n_prof / step_ms}
        gk = gemm_roof("gemm_bf16x3_kernel", "linears / 1x1 convs / mask logits / projections, fwd + input grads")
        gt = gemm_roof("gemm_bf16x3_tn_kernel", "weight gradients, dF of the mask logits")
        xf = gemm_roof("masked_xattn_fwd", "fused masked cross-attention forward, 3xTF32", 6.0)
        xb = gemm_roof("masked_xattn_bwd", "fused masked cross-attention backward (dQ + dK/dV kernels)", 6.0)
        cands = [r for r in (gk, gt, msda) if r]
        roof = dict(max(cands, key=lambda r: r["share_of_step"])) if cands else None
        if roof is not None:
            # the north-star kernels ride inside the headline object (the driver keeps `roofline` whole)
            def brief(r):
                return None if r is None else {k: r[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac",
                                                                 "avg_launch_ms", "share_of_step") if k in r}
            roof["north_star"] = {"msda_fwd": brief(msda), "msda_bwd": None if not msda else msda.get("backward"),
                                  "xattn_fwd": brief(xf), "xattn_bwd": brief(xb)}
        cpu = parity = stock = None
        if world == 1:
            if not a.no_cpu_baseline:
                fn, sample = cpu_port_step_fn(a, p)
                fn()
                t0 = time.perf_counter()
                n = 0
                while n < 8 and time.perf_counter() - t0 < 15:     # bounded sample: ~15-25 s of host work
                    fn(); n += 1
                dt = time.perf_counter() - t0
                cpu = {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                       "sample": sample + f" ({n} steps, {dt:.1f}s)"}
                if msda is not None:
                    try:
                        shapes = [(H // s, W // s) for s in (8, 16, 32)]
                        msda["cpu_reference"] = cpu_msda_baseline(shapes)
                        msda["cpu_reference"]["speedup_vs_cpu_per_image"] = (
                            msda["cpu_reference"]["ms_per_image_per_layer"] / (msda["avg_launch_ms"] / B))
                    except Exception as e:  # noqa: BLE001  (a reporting extra must not cost the bench line)
                        msda["cpu_reference"] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
            if not a.no_parity:
                # one image of this configuration, product vs CPU oracle (tests/parity_full.py): arithmetic error with
                # the oracle's attention masks teacher-forced, and the free-running mask-bit flip rate per layer
                try:
                    sys.path.insert(0, os.path.join(ROOT, "tests"))
                    import parity_full
                    f1 = workload.synthetic_features(1, H, W, backbone=p["backbone"], seed=5)
                    t1 = workload.synthetic_targets(1, H, W, num_classes=p["classes"], seed=5)
                    r = parity_full.compare(pd, dec, f1, t1, dev, num_queries=p["queries"])
                    parity = {"teacher_forced_max_rel_err": max(r["forced_max_rel_logits"], r["forced_max_rel_masks"]),
                              "tolerance": 1e-3, "teacher_forced_flip_rate_max": max(r["forced_flip_rate"]),
                              "teacher_forced_flip_max_dist_from_threshold": r["forced_flip_max_dist"],
                              "mask_flip_rate": statistics.mean(r["free_flip_rate"]),
                              "mask_flip_rate_per_layer": r["free_flip_rate"],
                              "free_running_final_frac_above_1e-3": r["free_frac_above_1e-3_masks"]}
                except Exception as e:  # noqa: BLE001
                    parity = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            if not a.no_stock:
                try:
                    stock = stock_forward_block(pd, dec, p, feats, B)
                except Exception as e:  # noqa: BLE001
                    stock = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        cfg = dict(p["config"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": p["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "impl_notes": {"native_ops": sorted(M.ops.NATIVE_OPS), "tf32": False, "cuda_graph": graph_note,
                           "routes": dict(sorted(M.ops.ROUTES.items())),
                           "criterion_path": getattr(criterion, "last_path", None) if a.loss == "criterion" else None,
                           "weight_operand_cache": None if wcache is None else {
                               "operands": len(wcache.entries), "hits": wcache.hits, "misses": wcache.misses},
                           "arithmetic": "fp32 storage; GEMMs in bf16x3 split arithmetic (fp32 accumulate), attention "
                                         "core in 3xTF32; no single-pass reduced precision"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "parity": parity, "vs_stock_cuda": stock,
            "roofline_msda": msda, "roofline_gemm": gk, "roofline_gemm_tn": gt, "roofline_xattn_fwd": xf,
            "roofline_xattn_bwd": xb,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config == 1:
        if world_info()[1] == 0:                      # one image, one GPU: the other ranks have nothing to do
            run_config1(args)
    else:
        run_ours(args)
