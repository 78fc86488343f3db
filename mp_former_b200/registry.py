"""Plug-in points.  The reference selects the pixel decoder / transformer decoder by name through
Detectron2 registries (ref: pixel_decoder/fpn.py:21-34 ``build_pixel_decoder`` on
``SEM_SEG_HEADS_REGISTRY``; transformer_decoder/maskformer_transformer_decoder.py:16-28
``build_transformer_decoder`` on ``TRANSFORMER_DECODER_REGISTRY``).

When Detectron2 is importable our classes register into ITS ``SEM_SEG_HEADS_REGISTRY`` under the
reference's class names (see INTEGRATION.md for the decoder registry, which lives inside the
reference package).  Otherwise light-weight local registries with the same ``register()/get()``
surface are used, so the modules are constructible from explicit kwargs without Detectron2.
"""


class _LocalRegistry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            return lambda o: self.register(o)
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]


try:  # pragma: no cover - detectron2 is absent in the build image
    from detectron2.config import configurable  # type: ignore
    from detectron2.modeling import SEM_SEG_HEADS_REGISTRY as _D2_HEADS  # type: ignore
    HAVE_DETECTRON2 = True
except Exception:  # noqa: BLE001
    HAVE_DETECTRON2 = False
    _D2_HEADS = None

    def configurable(init_func=None, *, from_config=None):
        """Identity stand-in for detectron2.config.configurable (explicit kwargs only)."""
        if init_func is not None:
            return init_func
        return lambda f: f


SEM_SEG_HEADS_REGISTRY = _LocalRegistry("SEM_SEG_HEADS")
TRANSFORMER_DECODER_REGISTRY = _LocalRegistry("TRANSFORMER_MODULE")


def register_pixel_decoder(cls):
    SEM_SEG_HEADS_REGISTRY.register(cls)
    return cls


def register_transformer_decoder(cls):
    TRANSFORMER_DECODER_REGISTRY.register(cls)
    return cls


def install_into_detectron2(override=True):
    """Registers the B200 modules into Detectron2's / the reference's registries under the
    reference's names.  Call once after importing ``mask2former``."""
    if not HAVE_DETECTRON2:
        raise RuntimeError("detectron2 is not importable")
    import importlib
    for name, cls in SEM_SEG_HEADS_REGISTRY.items():
        if override and name in _D2_HEADS._obj_map:
            del _D2_HEADS._obj_map[name]
        _D2_HEADS.register(cls)
    ref = importlib.import_module(
        "mask2former.modeling.transformer_decoder.maskformer_transformer_decoder")
    for name, cls in TRANSFORMER_DECODER_REGISTRY.items():
        if override and name in ref.TRANSFORMER_DECODER_REGISTRY._obj_map:
            del ref.TRANSFORMER_DECODER_REGISTRY._obj_map[name]
        ref.TRANSFORMER_DECODER_REGISTRY.register(cls)


def build_pixel_decoder(cfg, input_shape):
    """ref: pixel_decoder/fpn.py:21-34."""
    name = cfg.MODEL.SEM_SEG_HEAD.PIXEL_DECODER_NAME
    model = SEM_SEG_HEADS_REGISTRY.get(name)(**SEM_SEG_HEADS_REGISTRY.get(name).from_config(cfg, input_shape))
    if not callable(getattr(model, "forward_features", None)):
        raise ValueError(f"Only SEM_SEG_HEADS with forward_features method can be used as pixel decoder. "
                         f"Please implement forward_features for {name} to only return mask features.")
    return model


def build_transformer_decoder(cfg, in_channels, mask_classification=True):
    """ref: transformer_decoder/maskformer_transformer_decoder.py:22-28."""
    name = cfg.MODEL.MASK_FORMER.TRANSFORMER_DECODER_NAME
    cls = TRANSFORMER_DECODER_REGISTRY.get(name)
    return cls(**cls.from_config(cfg, in_channels, mask_classification))
