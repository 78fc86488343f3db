"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the UNMODIFIED reference modules of the hot path straight from
``/root/reference`` so that golden vectors can be generated from the reference's own code
(see ``tests/golden/make_golden.py``).  ``/root/reference`` exists only in the authoring
container, never on the GPU box, so nothing that runs there may call :func:`load_reference`.

Third-party packages the reference imports but that are absent here (detectron2, fvcore and the
compiled ``MultiScaleDeformableAttention`` extension) are replaced by minimal stand-ins that keep
the arithmetic identical:

* ``detectron2.config.configurable``      -> identity decorator (we pass explicit kwargs)
* ``detectron2.layers.Conv2d``            -> ``nn.Conv2d`` + optional ``norm`` / ``activation``
  (same order as detectron2: conv -> norm -> activation)
* ``detectron2.layers.get_norm("GN", c)`` -> ``nn.GroupNorm(32, c)``
* ``fvcore.nn.weight_init.c2_xavier_fill``-> kaiming_uniform_(a=1), zero bias (init only; the
  goldens overwrite every parameter with seeded values anyway)
* ``MultiScaleDeformableAttention``       -> empty module; ``MSDeformAttn.forward`` then takes its
  own ``except:`` branch into ``ms_deform_attn_core_pytorch``
  (reference ``ops/modules/ms_deform_attn.py:116-121``), i.e. the reference's CPU path.
"""
import importlib
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("MPF_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mask2former", "modeling"))


class _Registry(dict):
    def __init__(self, name="registry"):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self[o.__name__] = o
                return o
            return deco
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        return self[name]


class _Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d stand-in: conv -> norm -> activation."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = super().forward(x)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class _ShapeSpec:
    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


def _get_norm(norm, out_channels):
    if norm is None or norm == "":
        return None
    if callable(norm) and not isinstance(norm, str):
        return norm(out_channels)
    if norm == "GN":
        return nn.GroupNorm(32, out_channels)
    raise ValueError(f"stub get_norm: unsupported norm {norm!r}")


def _configurable(init_func=None, *, from_config=None):
    if init_func is not None:
        return init_func

    def deco(f):
        return f
    return deco


def _c2_xavier_fill(module):
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


_LOADED = None


def load_reference():
    """Returns a namespace with the reference's own classes/functions for the hot path."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    # ---- third-party stand-ins -------------------------------------------------------------
    _mod("detectron2")
    _mod("detectron2.config", configurable=_configurable)
    _mod("detectron2.layers", Conv2d=_Conv2d, ShapeSpec=_ShapeSpec, get_norm=_get_norm)
    _mod("detectron2.modeling", SEM_SEG_HEADS_REGISTRY=_Registry("SEM_SEG_HEADS"))
    _mod("detectron2.utils")
    _mod("detectron2.utils.registry", Registry=_Registry)
    _mod("fvcore")
    wi = _mod("fvcore.nn.weight_init", c2_xavier_fill=_c2_xavier_fill)
    _mod("fvcore.nn", weight_init=wi)
    _mod("MultiScaleDeformableAttention")  # empty: reference falls to its CPU path

    # ---- package shells so mask2former/__init__.py (data, D2 engine) is NOT executed -------
    base = os.path.join(REFERENCE_ROOT, "mask2former")
    _pkg("mask2former", base)
    _pkg("mask2former.modeling", os.path.join(base, "modeling"))
    _pkg("mask2former.modeling.pixel_decoder", os.path.join(base, "modeling", "pixel_decoder"))
    _pkg("mask2former.modeling.transformer_decoder",
         os.path.join(base, "modeling", "transformer_decoder"))

    ns = types.SimpleNamespace()
    func = importlib.import_module(
        "mask2former.modeling.pixel_decoder.ops.functions.ms_deform_attn_func")
    ns.ms_deform_attn_core_pytorch = func.ms_deform_attn_core_pytorch
    mods = importlib.import_module("mask2former.modeling.pixel_decoder.ops.modules.ms_deform_attn")
    ns.MSDeformAttn = mods.MSDeformAttn
    pd = importlib.import_module("mask2former.modeling.pixel_decoder.msdeformattn")
    ns.MSDeformAttnPixelDecoder = pd.MSDeformAttnPixelDecoder
    ns.MSDeformAttnTransformerEncoderOnly = pd.MSDeformAttnTransformerEncoderOnly
    pe = importlib.import_module("mask2former.modeling.transformer_decoder.position_encoding")
    ns.PositionEmbeddingSine = pe.PositionEmbeddingSine
    td = importlib.import_module(
        "mask2former.modeling.transformer_decoder.mask2former_transformer_decoder")
    ns.MultiScaleMaskedTransformerDecoder = td.MultiScaleMaskedTransformerDecoder
    ns.MultiScaleMaskedTransformerDecoderMaskDN = td.MultiScaleMaskedTransformerDecoderMaskDN
    ns.CrossAttentionLayer = td.CrossAttentionLayer
    ns.SelfAttentionLayer = td.SelfAttentionLayer
    ns.FFNLayer = td.FFNLayer
    ns.MLP = td.MLP
    ns.ShapeSpec = _ShapeSpec
    _LOADED = ns
    return ns


def _point_sample(input, point_coords, **kwargs):
    """detectron2.projects.point_rend.point_features.point_sample (detectron2 is an unpinned "git master"
    dependency of the reference, INSTALL.md:36-38, and is not vendored): a wrapper around F.grid_sample that takes
    point coordinates in [0, 1] x [0, 1] ([N, P, 2]) instead of [-1, 1] grids -- restated from its published source."""
    import torch.nn.functional as F
    add_dim = False
    if point_coords.dim() == 3:
        add_dim = True
        point_coords = point_coords.unsqueeze(2)
    output = F.grid_sample(input, 2.0 * point_coords - 1.0, **kwargs)
    if add_dim:
        output = output.squeeze(3)
    return output


def load_matcher():
    """The reference's HungarianMatcher (mask2former/modeling/matcher.py), imported unmodified; its only third-party
    import is detectron2's ``point_sample`` (stand-in above).  Groundwork for SURVEY.md §8f rank 1."""
    load_reference()
    _mod("detectron2.projects")
    _mod("detectron2.projects.point_rend")
    _mod("detectron2.projects.point_rend.point_features", point_sample=_point_sample)
    return importlib.import_module("mask2former.modeling.matcher")


def _get_uncertain_point_coords_with_randomness(coarse_logits, uncertainty_func, num_points, oversample_ratio,
                                                importance_sample_ratio):
    """detectron2.projects.point_rend.point_features.get_uncertain_point_coords_with_randomness, restated from its
    published source (PointRend, Kirillov et al. 2020, sec. 3.1): draw ``oversample_ratio * num_points`` uniform
    points per mask, keep the ``importance_sample_ratio * num_points`` most uncertain ones (``torch.topk``) and fill up
    with fresh uniform points.  RNG consumption: one ``torch.rand(R, kN, 2)`` then one ``torch.rand(R, N - bN, 2)``."""
    assert oversample_ratio >= 1
    assert 0 <= importance_sample_ratio <= 1
    num_boxes = coarse_logits.shape[0]
    num_sampled = int(num_points * oversample_ratio)
    point_coords = torch.rand(num_boxes, num_sampled, 2, device=coarse_logits.device, dtype=coarse_logits.dtype)
    point_logits = _point_sample(coarse_logits, point_coords, align_corners=False)
    point_uncertainties = uncertainty_func(point_logits)
    num_uncertain_points = int(importance_sample_ratio * num_points)
    num_random_points = num_points - num_uncertain_points
    idx = torch.topk(point_uncertainties[:, 0, :], k=num_uncertain_points, dim=1)[1]
    shift = num_sampled * torch.arange(num_boxes, dtype=torch.long, device=coarse_logits.device)
    idx = idx + shift[:, None]
    point_coords = point_coords.view(-1, 2)[idx.view(-1), :].view(num_boxes, num_uncertain_points, 2)
    if num_random_points > 0:
        point_coords = torch.cat(
            [point_coords, torch.rand(num_boxes, num_random_points, 2, device=coarse_logits.device)], dim=1)
    return point_coords


def load_criterion():
    """The reference's SetCriterion (mask2former/modeling/criterion.py), imported unmodified.  Third-party stand-ins:
    detectron2's ``get_world_size`` (-> torch.distributed's, 1 when not initialised), ``point_sample`` and
    ``get_uncertain_point_coords_with_randomness`` (above)."""
    load_matcher()
    import torch.distributed as dist

    def get_world_size():
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    _mod("detectron2.utils.comm", get_world_size=get_world_size)
    pf = sys.modules["detectron2.projects.point_rend.point_features"]
    pf.get_uncertain_point_coords_with_randomness = _get_uncertain_point_coords_with_randomness
    _pkg("mask2former.utils", os.path.join(REFERENCE_ROOT, "mask2former", "utils"))
    return importlib.import_module("mask2former.modeling.criterion")


def load_head():
    """The reference's caller of the path, ``MaskFormerHead`` (mask2former/modeling/meta_arch/mask_former_head.py),
    imported unmodified, together with its two builders (pixel_decoder/fpn.py:21-34, transformer_decoder/
    maskformer_transformer_decoder.py:22-28) and their registries -- for the boundary tests."""
    load_reference()
    dl = sys.modules["detectron2.layers"]
    if not hasattr(dl, "DeformConv"):
        dl.DeformConv = object
    base = os.path.join(REFERENCE_ROOT, "mask2former", "modeling")
    _pkg("mask2former.modeling.meta_arch", os.path.join(base, "meta_arch"))
    ns = types.SimpleNamespace()
    head = importlib.import_module("mask2former.modeling.meta_arch.mask_former_head")
    ns.MaskFormerHead = head.MaskFormerHead
    ns.build_pixel_decoder = importlib.import_module("mask2former.modeling.pixel_decoder.fpn").build_pixel_decoder
    td = importlib.import_module("mask2former.modeling.transformer_decoder.maskformer_transformer_decoder")
    ns.build_transformer_decoder = td.build_transformer_decoder
    ns.TRANSFORMER_DECODER_REGISTRY = td.TRANSFORMER_DECODER_REGISTRY
    ns.SEM_SEG_HEADS_REGISTRY = sys.modules["detectron2.modeling"].SEM_SEG_HEADS_REGISTRY
    return ns


class _Instances:
    """detectron2.structures.Instances stand-in: an attribute bag with the image size."""

    def __init__(self, image_size, **kwargs):
        self._image_size = tuple(image_size)
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def image_size(self):
        return self._image_size


class _Boxes:
    def __init__(self, tensor):
        self.tensor = tensor


class _ImageList:
    """detectron2.structures.ImageList stand-in (published behaviour of ``from_tensors``): images are zero-padded at
    the bottom/right to the largest height/width of the batch, rounded up to ``size_divisibility``."""

    def __init__(self, tensor, image_sizes):
        self.tensor, self.image_sizes = tensor, image_sizes

    @staticmethod
    def from_tensors(tensors, size_divisibility=0, pad_value=0.0):
        sizes = [tuple(t.shape[-2:]) for t in tensors]
        H, W = max(s[0] for s in sizes), max(s[1] for s in sizes)
        if size_divisibility > 1:
            H = (H + size_divisibility - 1) // size_divisibility * size_divisibility
            W = (W + size_divisibility - 1) // size_divisibility * size_divisibility
        out = tensors[0].new_full((len(tensors), tensors[0].shape[0], H, W), pad_value)
        for i, t in enumerate(tensors):
            out[i, :, :t.shape[-2], :t.shape[-1]] = t
        return _ImageList(out, sizes)


def _sem_seg_postprocess(result, img_size, output_height, output_width):
    """detectron2.modeling.postprocessing.sem_seg_postprocess (published): crop the padding away, then resize to the
    requested output resolution (bilinear, align_corners=False)."""
    import torch.nn.functional as F
    result = result[:, : img_size[0], : img_size[1]].expand(1, -1, -1, -1)
    return F.interpolate(result, size=(output_height, output_width), mode="bilinear", align_corners=False)[0]


def load_meta_arch():
    """The reference's ``MaskFormer`` meta-architecture (mask2former/maskformer_model.py), imported unmodified, for
    golden vectors of its inference epilogue (:232-279, :300-401).  Detectron2 stand-ins: Instances / Boxes /
    ImageList / sem_seg_postprocess above; registries, ``retry_if_cuda_oom`` (identity) and builders (unused)."""
    load_criterion()
    reg = _Registry("META_ARCH")
    dm = sys.modules["detectron2.modeling"]
    dm.META_ARCH_REGISTRY = reg
    dm.build_backbone = dm.build_sem_seg_head = lambda *a, **k: None
    _mod("detectron2.data", MetadataCatalog=types.SimpleNamespace(get=lambda name: None))
    _mod("detectron2.modeling.backbone", Backbone=nn.Module)
    _mod("detectron2.modeling.postprocessing", sem_seg_postprocess=_sem_seg_postprocess)
    _mod("detectron2.structures", Boxes=_Boxes, ImageList=_ImageList, Instances=_Instances, BitMasks=object)
    _mod("detectron2.utils.memory", retry_if_cuda_oom=lambda f: f)
    _pkg("mask2former.util", os.path.join(REFERENCE_ROOT, "mask2former", "util"))
    return importlib.import_module("mask2former.maskformer_model")


class cuda_is_identity:
    """Context manager: the reference's DN preparation hard-codes ``.cuda()`` / ``.to('cuda')``
    (reference mask2former_transformer_decoder.py:984-985,1029,1052).  To run it on CPU for golden
    generation we make those calls no-ops.  Test-only monkeypatch."""

    def __enter__(self):
        self._cuda = torch.Tensor.cuda
        self._to = torch.Tensor.to
        torch.Tensor.cuda = lambda t, *a, **k: t
        orig_to = self._to

        def _to(t, *a, **k):
            a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
            if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
                k.pop("device")
            if not a and not k:
                return t
            return orig_to(t, *a, **k)
        torch.Tensor.to = _to
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._cuda
        torch.Tensor.to = self._to
        return False
