"""Whole-tile sliding-window inference.

The reference classifies a 64x64 tile by looping over its 8x8 windows and calling the model once per window with a
host synchronisation each time (inference_example.ipynb cell 13, src/utils.py:477-605: 64 forwards of batch B).  Here all
windows of all tiles go through ONE forward: the patch-embedding kernel reads every window straight out of the tile through
per-sample window offsets (csrc/pixel_source.cuh, `win_y` / `win_x` of msst_raw_input) -- no permute / copy of the input cube,
raw int16 tiles with on-the-fly standardisation included -- and the argmax / accuracy stay on the device.
"""
import torch

from .input import RawTiles


@torch.no_grad()
def predict_tiles(model, tiles, window=None):
    """tiles: fp32 cube [B, C, H, W] (already standardised) or RawTiles over [B, raw_bands, H, W] sensor tiles
    -> logits [B, num_classes, H', W'] over the non-overlapping windows of the model's image_size that fit the tile."""
    w = window or model.image_size
    if isinstance(tiles, RawTiles):
        raw = tiles
        H, W = raw.tiles.shape[2:]
        gh, gw = (H - raw.crop[0]) // w, (W - raw.crop[1]) // w
        x = RawTiles(raw.tiles, raw.means, raw.stds, image_size=w, crop=raw.crop, pad_bands=raw.pad_bands, clip=raw.clip, windows=(gh, gw))
        B = raw.tiles.shape[0]
    else:
        B, C, H, W = tiles.shape
        assert H % w == 0 and W % w == 0, "tile size must be a multiple of the window"
        gh, gw = H // w, W // w
        if tiles.dtype != torch.float32:
            tiles = tiles.float()
        # identity statistics: (v - 0) / 1 in float64 rounds back to the same fp32 value
        x = RawTiles(tiles, torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64), image_size=w, windows=(gh, gw))
    was_training = model.training
    model.eval()
    y = model(x)                                                     # [B*gh*gw, nc, w, w]
    model.train(was_training)
    nc = y.shape[1]
    return y.reshape(B, gh, gw, nc, w, w).permute(0, 3, 1, 4, 2, 5).reshape(B, nc, gh * w, gw * w)


@torch.no_grad()
def tile_accuracy(logits, labels, ignore_index=-1):
    """(overall accuracy, valid-pixel count) on the device; labels [B, H, W] int64."""
    pred = logits.argmax(dim=1)
    valid = labels != ignore_index
    n = valid.sum()
    return ((pred == labels) & valid).sum().float() / n.clamp_min(1), n
