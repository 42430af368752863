"""Drop-in alias package: the reference's import paths (`from src.vit_spatial_spectral import ViTSpatialSpectral`,
`from src.vit_simmim_original import SimMIMSpatialSpectral`, pretrain.py:18-19, finetune.py:22) resolve to the
B200-native implementation in maskedsst_b200/."""
