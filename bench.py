#!/usr/bin/env python
"""bench.py -- img/s (forward+backward) of the MP-Former hot path on N B200s, one JSON line.

A "step" is one forward+backward pass of the hot path (MSDeformAttn pixel decoder + masked-attention
transformer decoder, MP-Former COCO-instance R50 head, DN/mask-piloted queries on) over one batch of
synthetic 1024x1024 backbone features.  The backbone (upstream of the path) and the criterion
(downstream) are outside the path: inputs are R50-shaped feature maps, the loss is a
fixed linear functional of every prediction the head returns.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm
  python bench.py --impl reference [...]                          CPU port of the reference path
  torchrun --nproc-per-node N bench.py --gpus N ...               one rank per GPU (weak scaling)

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec fwd+bwd, MP-Former R50 head (MSDeformAttn pixel decoder + masked decoder), 1024x1024"
UNIT = "img/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU (weak scaling)")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--queries", type=int, default=100)
    ap.add_argument("--no-dn", action="store_true", help="drop the mask-piloted (DN) query group")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--criterion", action="store_true",
                    help="BASELINE config 3: SetCriterion + HungarianMatcher (device path, 12544 points, deep "
                         "supervision + dn losses) instead of the linear pseudo-loss; not the default workload")
    return ap.parse_args()


def workload_name(a):
    return (f"R50 feature maps {a.height}x{a.width} -> MSDeformAttnPixelDecoder(6 layers) + "
            f"MultiScaleMaskedTransformerDecoderMaskDN(9 layers, {a.queries} queries, "
            f"{'DN points' if not a.no_dn else 'no DN'}), fwd+bwd, "
            + ("SetCriterion + HungarianMatcher (recipe weights, 12544 points, 10 heads + dn)"
               if getattr(a, "criterion", False) else "linear pseudo-loss"))


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
# reference arm / cpu baseline: the oracle port of the reference's CPU path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_port_step_fn(a):
    """Returns (fn, sample_description): fn() runs ONE image fwd+bwd through the CPU oracle."""
    import torch
    from mp_former_b200 import workload
    from oracle import torch_oracle as O
    torch.set_num_threads(min(32, os.cpu_count() or 1))   # >32 threads slow the gather-heavy port down
    pd, dec = workload.build_head(num_queries=a.queries, device="cpu")
    psd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in pd.state_dict().items()}
    dsd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in dec.state_dict().items()}
    feats = workload.synthetic_features(1, a.height, a.width)
    dn = None if a.no_dn else {"tgt": workload.synthetic_targets(1, a.height, a.width), "scalar": 1,
                               "noise_scale": 0.0}

    loss_of = pseudo_loss
    if getattr(a, "criterion", False):       # same workload as the GPU arm's --criterion: the recipe's SetCriterion
        from oracle import criterion_oracle as CO
        targets = dn["tgt"] if dn is not None else workload.synthetic_targets(1, a.height, a.width)
        _, weighted_sum = workload.build_criterion(device="cpu")

        def loss_of(out):
            return weighted_sum(CO.set_criterion(out, targets, num_classes=80, eos_coef=0.1, losses=["labels", "masks"],
                                                 num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75,
                                                 cost_class=2.0, cost_mask=5.0, cost_dice=5.0, training=True))

    def fn():
        mf, _, ms = O.pixel_decoder_forward(psd, feats)
        out = O.decoder_forward(dsd, ms, mf, num_queries=a.queries, dn_args=dn, dn_label_noise_ratio=0.2)
        loss = loss_of(out)
        loss.backward()
        for sd in (psd, dsd):
            for v in sd.values():
                v.grad = None
        return float(loss.detach())

    return fn, "1 image (same shapes/config) fwd+bwd per step through oracle/torch_oracle.py on host cores"


def cpu_msda_baseline(a, reps=5):
    """The reference's CPU MSDeformAttn path (``ms_deform_attn_core_pytorch``, ref ops/functions/ms_deform_attn_func.py:
    52-72, restated in oracle/torch_oracle.py::msda_core_grid_sample) on the host cores: ONE image, one encoder layer's
    call (S = Lq = all pixels of the three levels, M=8, D=32, L=3, P=4), 2 warm-up + ``reps`` timed calls (~2-4 s).
    Reported next to the GPU kernel's number in `roofline_msda` (north_star: "next to the reference's CPU MSDeformAttn
    path timed on the same box's host cores (core count stated) in the same run")."""
    import torch
    from oracle import torch_oracle as O
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    shapes = [(a.height // s, a.width // s) for s in (8, 16, 32)]
    S = sum(h * w for h, w in shapes)
    M_, D, L, P = 8, 32, 3, 4
    g = torch.Generator().manual_seed(3)
    value = torch.rand(1, S, M_, D, generator=g) * 0.01                      # distributions of ref ops/test.py:36-39
    loc = torch.rand(1, S, M_, L, P, 2, generator=g)
    aw = torch.rand(1, S, M_, L, P, generator=g) + 1e-5
    aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    with torch.no_grad():
        for _ in range(2):
            O.msda_core_grid_sample(value, shapes, loc, aw)
        t0 = time.perf_counter()
        for _ in range(reps):
            O.msda_core_grid_sample(value, shapes, loc, aw)
        ms = (time.perf_counter() - t0) * 1e3 / reps
    alg = 4 * (S * M_ * D + 2 * S * M_ * L * P + S * M_ * L * P + S * M_ * D)   # SURVEY.md §8d, per image and layer
    return {"kernel": "ms_deform_attn_core_pytorch (reference CPU path, per-level F.grid_sample), forward",
            "ms_per_image_per_layer": ms, "achieved": alg / (ms / 1e3) / 1e9, "unit": "GB/s (algorithmic bytes)",
            "algorithmic_bytes": alg, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 image, S=Lq={S}, L=3, M=8, D=32, P=4, {reps} calls"}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn, sample = cpu_port_step_fn(a)
    for _ in range(min(a.warmup, 1)):
        fn()
    steps = a.steps
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        fn()
        done += 1
        if time.perf_counter() - t0 > 240:      # keep the whole arm within minutes
            break
    dt = time.perf_counter() - t0
    v = done / dt
    import torch
    cores = torch.get_num_threads()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
        "steps": done, "warmup": min(a.warmup, 1), "ms_per_step": dt / done * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "images_per_step": 1},
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
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_ours(a):
    import torch
    import torch.distributed as dist
    import mp_former_b200 as M
    from mp_former_b200 import MultiScaleDeformableAttention as MSDA
    from mp_former_b200 import _lib, graphs, native, workload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False          # fp32 path, like the reference's pixel decoder
    torch.backends.cudnn.allow_tf32 = False

    B = a.batch
    pd, dec = workload.build_head(num_queries=a.queries, device=dev, seed=0)

    class Head(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.pixel_decoder, self.predictor = pd, dec

        def forward(self, feats, dn_args):
            mf, _, ms = self.pixel_decoder.forward_features(feats)
            return self.predictor(ms, mf, None, dn_args)

    head = Head()
    params = [p for p in head.parameters()]
    feats = workload.synthetic_features(B, a.height, a.width, seed=rank, device=dev)
    dn_args = None
    if not a.no_dn:
        dn_args = {"tgt": workload.synthetic_targets(B, a.height, a.width, seed=rank, device=dev),
                   "scalar": 1, "noise_scale": 0.0}

    loss_of = pseudo_loss
    if a.criterion:
        targets = dn_args["tgt"] if dn_args is not None else workload.synthetic_targets(
            B, a.height, a.width, seed=rank, device=dev)
        criterion, weighted_sum = workload.build_criterion(device=dev)
        criterion.train(True)

        def loss_of(out):
            return weighted_sum(criterion(out, targets))

    # Gradient all-reduce (the path's only collective): one flat NCCL all-reduce of the head's
    # gradients (~20 M parameters) after the backward, averaged over ranks like DistributedDataParallel.
    def allreduce_grads():
        graphs.allreduce_gradients(params, world)

    def eager_step(f):
        for p in params:
            p.grad = None
        loss = loss_of(head(f, dn_args))
        loss.backward()
        allreduce_grads()
        return loss

    for _ in range(max(a.warmup, 3)):
        eager_step(feats)
    torch.cuda.synchronize()

    # One CUDA graph for forward + loss + backward (mp_former_b200/graphs.py): ~2,600 launches per step would
    # otherwise be issued from Python.  Falls back to eager stepping (and says so) if capture is refused.
    gs, graph_note = None, "disabled (--no-graph)"
    if not a.no_graph:
        try:
            gs = graphs.GraphedStep(lambda inp: loss_of(head(inp, dn_args)), feats, params, warmup=1)
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
        allreduce_grads()
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
        host = workload.synthetic_features(B, a.height, a.width, seed=rank, device="cpu", pin=True)
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
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        hbm = peaks.get("hbm_gbs", 6650.0)
        tens = peaks.get("bf16_tflops_sustained", 1400.0)          # kernels timed inside a long step
        src = "MEASURED_PEAKS.json" if peaks else "fallback (6650 GB/s, 1400 TFLOP/s sustained)"
        step_ms = ms_total / a.steps

        def gemm_roof(kind, what):
            r = gprof.get(kind)
            if not r or r["ms"] <= 0:
                return None
            n = r["launches"]
            tf = r["flops"] / (r["ms"] / 1e3) / 1e12
            gb = r["bytes"] / (r["ms"] / 1e3) / 1e9
            # which ceiling binds this mix of shapes: time at HBM peak for the algorithmic bytes vs time at the
            # tensor peak for the MMAs the split arithmetic issues (3 per product); the larger one is the roofline
            t_hbm = r["bytes"] / (hbm * 1e9)
            t_tensor = 3.0 * r["flops"] / (tens * 1e12)
            common = {"kernel": f"{kind} ({what})", "traffic": None,
                      "traffic_note": "a mix of shapes is timed here, so there is no single per-launch DRAM figure; "
                                      "single-shape ncu --set full captures (dram read+write per launch): encoder "
                                      "FFN1 1.70 GB (profiles/r1o_ncu_gemm_ffn1.txt), 3x3 convolution forward 2.11 GB "
                                      "(profiles/r1x_ncu_conv_fwd.txt), TN weight gradient 1.80 GB "
                                      "(profiles/r1o_ncu_gemm_tn.txt)",
                      "note": "fp32 operands, bf16x3 split arithmetic: 3 MMAs per product (tensor ceiling = peak/3 in "
                              "algorithmic flops); bound = the ceiling with the larger ideal time for the timed launches",
                      "tensor_TFLOPs_algorithmic": tf, "tensor_frac_algorithmic": tf / tens,
                      "tensor_issue_frac": 3 * tf / tens, "achieved_hbm_GBs": gb, "hbm_frac": gb / hbm,
                      "ideal_ms_hbm": t_hbm * 1e3 / n, "ideal_ms_tensor": t_tensor * 1e3 / n,
                      "algorithmic_flops_per_launch": r["flops"] / n, "algorithmic_bytes_per_launch": r["bytes"] / n,
                      "avg_launch_ms": r["ms"] / n, "launches_timed": n,
                      "share_of_step": r["ms"] / n_prof / step_ms}
            if t_hbm >= t_tensor:
                return {"bound": "hbm", "achieved": gb, "peak": hbm, "unit": "GB/s", "frac": gb / hbm,
                        "peak_source": src + " hbm_gbs", **common}
            return {"bound": "tensor", "achieved": tf, "peak": tens, "unit": "TFLOP/s", "frac": tf / tens,
                    "peak_source": src + " bf16_tflops_sustained", **common}

        # MSDeformAttn: algorithmic bytes per launch:
        #   4 * (S*M*D + 2*Lq*M*L*P + Lq*M*L*P + Lq*M*D) per image
        S = sum((a.height // s) * (a.width // s) for s in (32, 16, 8))
        Mh, D, L, P = 8, 32, 3, 4
        alg = 4 * (S * Mh * D + 2 * S * Mh * L * P + S * Mh * L * P + S * Mh * D) * B
        fwd_ms = statistics.mean(prof["fwd_ms"]) if prof["fwd_ms"] else None
        bwd_ms = statistics.mean(prof["bwd_ms"]) if prof["bwd_ms"] else None
        msda = None
        if fwd_ms:
            ach = alg / (fwd_ms / 1e3) / 1e9
            msda = {"kernel": "msda_enc_fwd_kernel<8> (MSDeformAttn forward, softmax+locations fused)",
                    "bound": "hbm", "achieved": ach, "peak": hbm,
                    "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                    "peak_source": src + " hbm_gbs",
                    "algorithmic_bytes_per_launch": alg, "avg_launch_ms": fwd_ms,
                    "launches_timed": len(prof["fwd_ms"]),
                    "share_of_step": fwd_ms * len(prof["fwd_ms"]) / n_prof / step_ms}
            if B == 16 and a.height == 1024 and a.width == 1024:
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at exactly this geometry, from the
                # committed `ncu --set full` capture (a bench run cannot profile itself)
                msda["traffic"] = 749.512192e6 + 334.541312e6
                msda["traffic_source"] = "profiles/r1x_ncu_msda_enc_fwd.txt"
            if bwd_ms:
                alg_b = (4 * (S * Mh * D * 2 + 3 * S * Mh * L * P) + 4 * (S * Mh * D + 3 * S * Mh * L * P)) * B
                msda["backward"] = {"kernel": "msda_enc_bwd_kernel<8> (+ grad_value memset)", "avg_launch_ms": bwd_ms,
                                    "achieved": alg_b / (bwd_ms / 1e3) / 1e9,
                                    "frac": alg_b / (bwd_ms / 1e3) / 1e9 / hbm,
                                    "algorithmic_bytes_per_launch": alg_b,
                                    "share_of_step": bwd_ms * len(prof["bwd_ms"]) / n_prof / step_ms}
        # `roofline` = the dominant kernel of the step (largest share); the others ride along under their names
        gk = gemm_roof("gemm_bf16x3_kernel", "linears / 1x1 convs / mask logits / projections, fwd + input grads")
        gt = gemm_roof("gemm_bf16x3_tn_kernel", "weight gradients, dF of the mask logits")
        cands = [r for r in (gk, gt, msda) if r]
        roof = max(cands, key=lambda r: r["share_of_step"]) if cands else None
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            fn, sample = cpu_port_step_fn(a)
            fn()
            t0 = time.perf_counter()
            n = 0
            while n < 8 and time.perf_counter() - t0 < 15:     # bounded sample: ~15-20 s of host work
                fn(); n += 1
            dt = time.perf_counter() - t0
            cpu = {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": sample + f" ({n} steps, {dt:.1f}s)"}
            if msda is not None:
                try:
                    msda["cpu_reference"] = cpu_msda_baseline(a)
                    msda["cpu_reference"]["speedup_vs_cpu_per_image"] = (
                        msda["cpu_reference"]["ms_per_image_per_layer"] / (msda["avg_launch_ms"] / B))
                except Exception as e:  # noqa: BLE001  (a reporting extra must not cost the bench line)
                    msda["cpu_reference"] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "images_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "l2": "inputs+activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                       "native_ops": sorted(M.ops.NATIVE_OPS), "tf32": False,
                       "arithmetic": "fp32 storage; GEMMs in bf16x3 split arithmetic (fp32 accumulate), attention "
                                     "core in 3xTF32; no single-pass reduced precision",
                       "cuda_graph": graph_note},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof,
            "roofline_msda": msda, "roofline_gemm": gk, "roofline_gemm_tn": gt, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
