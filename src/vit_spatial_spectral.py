"""Alias of maskedsst_b200.vit_spatial_spectral under the reference's import path (src/vit_spatial_spectral.py)."""
from maskedsst_b200.vit_spatial_spectral import *  # noqa: F401,F403
from maskedsst_b200.vit_spatial_spectral import (ViTSpatialSpectral, ViTSpatialSpectral_V1, AvgPoolMerge, LinearMerge, Transformer, Attention, FeedForward, PreNorm,  # noqa: F401
                                                  BlockwisePatchEmbedding, PatchEmbed, MoveAxis, Mean, Flatten, Squeeze,
                                                  get_pos_for_spectral_embedding, pair)
