"""mp_former_b200 -- B200-native (sm_100a) implementation of MP-Former's hot path:
the MSDeformAttn pixel decoder and the masked-attention transformer decoder, behind the
reference's own Python module API (see DESIGN.md / INTEGRATION.md).

Importing the package loads the in-tree C-ABI library ``libmpformer_b200.so`` and fails loudly if
it is missing -- there is no CPU or PyTorch fallback for the custom ops.
"""
from . import _lib

_lib.load()

from . import MultiScaleDeformableAttention  # noqa: E402,F401
from . import msdeform_attn  # noqa: E402,F401
from .msdeform_attn import MSDeformAttn, MSDeformAttnFunction  # noqa: E402,F401
from .position_encoding import PositionEmbeddingSine  # noqa: E402,F401
from .pixel_decoder import (MSDeformAttnPixelDecoder, MSDeformAttnTransformerEncoderOnly,  # noqa: E402,F401
                            ShapeSpec)
from .masked_decoder import (MultiScaleMaskedTransformerDecoder,  # noqa: E402,F401
                             MultiScaleMaskedTransformerDecoderMaskDN)
from .matcher import HungarianMatcher  # noqa: E402,F401
from .criterion import SetCriterion  # noqa: E402,F401
from .registry import (SEM_SEG_HEADS_REGISTRY, TRANSFORMER_DECODER_REGISTRY,  # noqa: E402,F401
                       build_pixel_decoder, build_transformer_decoder)

__all__ = [
    "MultiScaleDeformableAttention", "MSDeformAttn", "MSDeformAttnFunction", "PositionEmbeddingSine",
    "MSDeformAttnPixelDecoder", "MSDeformAttnTransformerEncoderOnly", "ShapeSpec",
    "MultiScaleMaskedTransformerDecoder", "MultiScaleMaskedTransformerDecoderMaskDN",
    "HungarianMatcher", "SetCriterion",
    "SEM_SEG_HEADS_REGISTRY", "TRANSFORMER_DECODER_REGISTRY", "build_pixel_decoder",
    "build_transformer_decoder",
]
