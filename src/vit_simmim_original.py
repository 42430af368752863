"""Alias of maskedsst_b200.vit_simmim_original under the reference's import path (src/vit_simmim_original.py)."""
from maskedsst_b200.vit_simmim_original import SimMIMSpatialSpectral, BlockwiseToPixels, MaskGenerator  # noqa: F401
