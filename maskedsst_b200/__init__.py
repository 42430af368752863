"""maskedsst_b200 -- B200-native (sm_100a) implementation of MaskedSST's ViTSpatialSpectral / SimMIM hot path.

Public surface mirrors the reference modules:
    from maskedsst_b200 import ViTSpatialSpectral, SimMIMSpatialSpectral
(or, as a drop-in, `from src.vit_spatial_spectral import ViTSpatialSpectral`).
All arithmetic runs in libmsst.so (include/msst.h); importing this package without a CUDA device works
(module construction, state_dict handling), calling a kernel without one raises.
"""
from .vit_spatial_spectral import (ViTSpatialSpectral, ViTSpatialSpectral_V1, Transformer, Attention, FeedForward, PreNorm,
                                   BlockwisePatchEmbedding, PatchEmbed, MoveAxis, get_pos_for_spectral_embedding)
from .vit_simmim_original import SimMIMSpatialSpectral, BlockwiseToPixels, MaskGenerator
from .ops import cross_entropy
from .input import RawTiles

__all__ = ["ViTSpatialSpectral", "ViTSpatialSpectral_V1", "SimMIMSpatialSpectral", "BlockwiseToPixels", "MaskGenerator", "Transformer", "Attention",
           "FeedForward", "PreNorm", "BlockwisePatchEmbedding", "PatchEmbed", "MoveAxis", "get_pos_for_spectral_embedding",
           "cross_entropy", "RawTiles"]
