"""Whole-tile sliding-window inference.

The reference classifies a 64x64 tile by looping over its 8x8 windows and calling the model once per window with a
host synchronisation each time (inference_example.ipynb cell 13, src/utils.py:477-605: 64 forwards of batch B).  Here all
windows of all tiles go through ONE forward (windows stacked on the batch axis) and the argmax / accuracy stay on the device.
"""
import torch


@torch.no_grad()
def predict_tiles(model, tiles, window=None):
    """tiles [B, C, H, W] -> logits [B, num_classes, H, W]; non-overlapping windows of the model's image_size."""
    w = window or model.image_size
    B, C, H, W = tiles.shape
    assert H % w == 0 and W % w == 0, "tile size must be a multiple of the window"
    gh, gw = H // w, W // w
    x = tiles.reshape(B, C, gh, w, gw, w).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, C, w, w).contiguous()
    was_training = model.training
    model.eval()
    y = model(x)                                                     # [B*gh*gw, nc, w, w]
    model.train(was_training)
    nc = y.shape[1]
    return y.reshape(B, gh, gw, nc, w, w).permute(0, 3, 1, 4, 2, 5).reshape(B, nc, H, W)


@torch.no_grad()
def tile_accuracy(logits, labels, ignore_index=-1):
    """(overall accuracy, valid-pixel count) on the device; labels [B, H, W] int64."""
    pred = logits.argmax(dim=1)
    valid = labels != ignore_index
    n = valid.sum()
    return ((pred == labels) & valid).sum().float() / n.clamp_min(1), n
