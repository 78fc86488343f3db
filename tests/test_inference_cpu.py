"""CPU: inference epilogue.
  * the oracle (oracle/inference_oracle.py) against the outputs of the UNMODIFIED reference MaskFormer.forward in eval
    mode (tests/golden/inference.pt);
  * the host logic of mp_former_b200/inference.py against the same golden, with the one kernel it calls
    (native.instance_masks) replaced -- in this test only -- by a torch emulation;
  * the on-the-fly two-stage resampling formula the kernel implements (csrc/inference.cu: ATen's source index / weight
    arithmetic, applied twice with a crop in between) against torch's F.interpolate chain."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_inference import CFG, IMAGES, inputs  # noqa: E402
from oracle import inference_oracle as IO  # noqa: E402


def _geometry(b):
    d = CFG["size_divisibility"]
    Hp = max((s[0] + d - 1) // d * d for s, _ in IMAGES)
    Wp = max((s[1] + d - 1) // d * d for s, _ in IMAGES)
    return (Hp, Wp), IMAGES[b][0], IMAGES[b][1]


def _canon(res):
    """rows in a canonical order (top-k is unsorted): by class, then score."""
    key = res["pred_classes"].double() * 10 + res["scores"].double()
    o = torch.argsort(key)
    return res["pred_masks"][o], res["scores"][o], res["pred_classes"][o]


def _check_instances(got, ref):
    gm, gs, gc = _canon(got)
    rm, rs, rc = _canon(ref)
    assert torch.equal(gc, rc)
    assert torch.allclose(gs, rs, rtol=1e-5, atol=1e-7)
    assert torch.equal(gm.bool(), rm.bool())


def test_inference_oracle_matches_reference_golden():
    G = torch.load(os.path.join(HERE, "golden", "inference.pt"), weights_only=False)
    outputs, _ = inputs()
    for b in range(len(IMAGES)):
        padded, image, out = _geometry(b)
        full = IO.full_resolution_masks(outputs["pred_masks"][b], padded, image, out)
        _check_instances(IO.instance_inference(outputs["pred_logits"][b], full, CFG["num_classes"], CFG["topk"]),
                         G["instance"][b])
        assert tuple(G["instance"][b]["image_size"]) == tuple(out)
        assert torch.allclose(IO.semantic_inference(outputs["pred_logits"][b], full), G["semantic"][b]["sem_seg"],
                              atol=1e-5)


def _emu_instance_masks(mask_logits, query_index, padded_size, image_size, out_size, mask_dtype=torch.uint8):
    full = IO.full_resolution_masks(mask_logits, padded_size, image_size, out_size)[query_index]
    fg = full > 0
    sums = torch.stack([(full.sigmoid() * fg).flatten(1).sum(1), fg.flatten(1).sum(1).float()], 1)
    return fg.to(mask_dtype), sums


def test_inference_host_logic_matches_reference_golden(monkeypatch):
    from mp_former_b200 import inference, native
    monkeypatch.setattr(native, "instance_masks", _emu_instance_masks)
    G = torch.load(os.path.join(HERE, "golden", "inference.pt"), weights_only=False)
    outputs, _ = inputs()
    for b in range(len(IMAGES)):
        padded, image, out = _geometry(b)
        r = inference.instance_inference(outputs["pred_logits"][b], outputs["pred_masks"][b], padded, image, out,
                                         CFG["num_classes"], CFG["topk"])
        assert r.pred_masks.dtype == torch.float32 and r.image_size == tuple(out) and r.pred_boxes.shape == (CFG["topk"], 4)
        _check_instances(r, G["instance"][b])
        for before, name in ((True, "semantic"), (False, "semantic_after")):
            s = inference.semantic_inference(outputs["pred_logits"][b], outputs["pred_masks"][b], padded, image, out,
                                             postprocess_before_inference=before)
            assert torch.allclose(s, G[name][b]["sem_seg"], atol=1e-5), name
    for b in range(len(IMAGES)):                 # instances of a panoptic model: "thing" classes only (:381-388)
        padded, image, out = _geometry(b)
        r = inference.instance_inference(outputs["pred_logits"][b], outputs["pred_masks"][b], padded, image, out,
                                         CFG["num_classes"], CFG["topk"], thing_ids=set(CFG["thing_ids"]))
        _check_instances(r, G["panoptic"][b])
    # "thing" filter of panoptic models
    r = inference.instance_inference(outputs["pred_logits"][0], outputs["pred_masks"][0], *_geometry(0),
                                     CFG["num_classes"], CFG["topk"], thing_ids={0, 2, 3})
    assert set(r.pred_classes.tolist()) <= {0, 2, 3} and r.pred_masks.shape[0] == r.scores.shape[0] <= CFG["topk"]
    with pytest.raises(AttributeError):
        r.no_such_field


def _src(scale, n_out, n_in):
    s = torch.tensor(scale, dtype=torch.float32) * (torch.arange(n_out, dtype=torch.float32) + 0.5) - 0.5
    s = s.clamp(min=0)
    i0 = s.to(torch.int64).clamp(max=n_in - 1)
    i1 = i0 + (i0 < n_in - 1).to(torch.int64)
    return i0, i1, s - i0.float()


def _resize(x, n_out_h, n_out_w):
    """ATen's upsample_bilinear2d(align_corners=False) arithmetic as csrc/inference.cu evaluates it."""
    h, w = x.shape[-2:]
    y0, y1, ly = _src(float(torch.tensor(h, dtype=torch.float32) / n_out_h), n_out_h, h)
    x0, x1, lx = _src(float(torch.tensor(w, dtype=torch.float32) / n_out_w), n_out_w, w)
    hy, hx = (1 - ly)[:, None], (1 - lx)[None, :]
    ly, lx = ly[:, None], lx[None, :]
    top = hx * x[..., y0[:, None], x0[None, :]] + lx * x[..., y0[:, None], x1[None, :]]
    bot = hx * x[..., y1[:, None], x0[None, :]] + lx * x[..., y1[:, None], x1[None, :]]
    return hy * top + ly * bot


@pytest.mark.parametrize("geom", [((16, 24), (64, 96), (64, 96), (64, 96)), ((16, 24), (64, 96), (50, 70), (75, 105)),
                                  ((64, 64), (256, 256), (200, 256), (480, 613)), ((7, 9), (28, 36), (28, 33), (11, 17))])
def test_two_stage_resampling_formula_equals_interpolate_chain(geom):
    (h, w), padded, image, out = geom
    g = torch.Generator().manual_seed(h * w)
    L = torch.randn(3, h, w, generator=g) * 3
    ref = IO.full_resolution_masks(L, padded, image, out)
    got = _resize(_resize(L, *padded)[:, :image[0], :image[1]], *out)
    assert torch.allclose(got, ref, rtol=2e-6, atol=1e-5)      # a few fp32 ulps (values up to ~10)
    assert ((got > 0) != (ref > 0)).float().mean() < 1e-4


@pytest.mark.parametrize("case", ["panoptic", "panoptic_structured"])
def test_panoptic_inference_matches_reference_golden(case):
    """Oracle (segment by segment) and the product's batched formulation (three counts per kept query from two
    bincounts and a row sum, one host read) against the unmodified reference; library ops only, so it runs here."""
    from make_golden_inference import panoptic_inputs
    from mp_former_b200 import inference
    G = torch.load(os.path.join(HERE, "golden", "inference.pt"), weights_only=False)
    outputs, _ = panoptic_inputs() if case == "panoptic_structured" else inputs()
    args = (CFG["num_classes"], set(CFG["thing_ids"]), CFG["object_mask_threshold"], CFG["overlap_threshold"])
    n_seg = 0
    for b in range(len(IMAGES)):
        padded, image, out = _geometry(b)
        full = IO.full_resolution_masks(outputs["pred_masks"][b], padded, image, out)
        for fn in (IO.panoptic_inference, inference.panoptic_inference):
            seg, info = fn(outputs["pred_logits"][b], full, *args)
            assert info == G[case][b]["segments_info"], fn.__module__
            assert seg.dtype == torch.int32 and torch.equal(seg, G[case][b]["panoptic_seg"]), fn.__module__
        n_seg += len(info)
    assert n_seg >= (5 if case == "panoptic_structured" else 1)
    seg, info = inference.panoptic_inference(outputs["pred_logits"][0], full, CFG["num_classes"], set(), 2.0, 0.5)
    assert info == [] and int(seg.abs().sum()) == 0                              # nothing above the score threshold


def _host_kernel_lib(tmp_path_factory=None):
    """g++ build of tests/host_emul/inference_host.cpp: the kernel's own per-pixel header compiled for the host."""
    import ctypes
    import subprocess
    root = os.path.dirname(HERE)
    out_dir = os.path.join(root, "build", "host_emul")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libinference_host.so")
    src = os.path.join(HERE, "host_emul", "inference_host.cpp")
    hdr = os.path.join(root, "mp_former_b200", "csrc", "inference_math.cuh")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-D__host__=", "-D__device__=",
                               "-D__forceinline__=inline", "-o", lib, src])
    return ctypes.CDLL(lib)


@pytest.mark.parametrize("geom", [((16, 24), (64, 96), (64, 96), (64, 96)), ((16, 24), (64, 96), (50, 70), (75, 105)),
                                  ((64, 64), (256, 256), (200, 256), (480, 613)), ((7, 9), (28, 36), (28, 33), (11, 17)),
                                  ((32, 32), (128, 128), (128, 128), (64, 64))])
def test_kernel_pixel_arithmetic_on_host_equals_interpolate_chain(geom):
    """The header the CUDA kernel includes (inference_math.cuh), compiled for the host: values within a few ulps of
    torch's two interpolates with the crop in between, identical masks except where |logit| ~ 0, same score sums;
    rows address a strided query slice through query_index like the kernel."""
    import ctypes
    lib = _host_kernel_lib()
    (h, w), padded, image, out = geom
    g = torch.Generator().manual_seed(h * 7 + w)
    store = torch.randn(9, h * w + 5, generator=g) * 3              # q_stride > h*w
    logits = store[:, :h * w].reshape(9, h, w)
    rows = torch.tensor([8, 0, 3, 3, 5], dtype=torch.int64)
    R, (oh, ow) = len(rows), out
    masks = torch.zeros(R, oh, ow, dtype=torch.uint8)
    values = torch.zeros(R, oh, ow)
    sums = torch.zeros(R, 2, dtype=torch.float64)
    vp = ctypes.c_void_p
    lib.host_instance_masks(vp(store.data_ptr()), ctypes.c_longlong(store.stride(0)), h, w, vp(rows.data_ptr()), R,
                            padded[0], padded[1], image[0], image[1], oh, ow, vp(masks.data_ptr()),
                            vp(values.data_ptr()), vp(sums.data_ptr()))
    ref = IO.full_resolution_masks(logits, padded, image, out)[rows]
    assert torch.allclose(values, ref, rtol=2e-6, atol=1e-5)
    fg = ref > 0
    assert ((masks != 0) != fg).float().mean() < 1e-4
    ref_sums = torch.stack([(ref.double().sigmoid() * fg).flatten(1).sum(1), fg.flatten(1).sum(1).double()], 1)
    assert torch.allclose(sums, ref_sums, rtol=1e-4, atol=1e-2)
