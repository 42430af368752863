"""GPU: SURVEY.md 8(f) rank 1 (device-side mask generation, csrc/maskgen.cu) and rank 3 (whole-tile sliding-window inference
straight out of the tiles, per-window offsets in csrc/pixel_source.cuh)."""
import json

import numpy as np
import pytest
import torch

import maskedsst_b200 as M
from maskedsst_b200.inference import predict_tiles, tile_accuracy
from oracle import maskedsst_oracle as O
from tests.helpers import gold, rel_l2
from tests.test_gpu_parity import make_encoder

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _simmim(spec, ratio, mask_patch, tube):
    m = M.SimMIMSpatialSpectral(encoder=make_encoder(spec), masking_ratio=ratio, mask_patch_size=mask_patch, tube_masking=tube,
                                to_pixels_per_spectral_block=True).to(DEV).train()
    m.mask_backend = "device"
    return m


@pytest.mark.parametrize("kw,ratio,mask_patch,tube", [
    (O.HOUSTON, 0.7, 4, True),      # the shipped pretrain configuration (configs/pretrain_config.yaml:25-29)
    (O.HOUSTON, 0.7, 4, False),     # block masking: one draw per (sample, spectral block)
    (O.ENMAP, 0.7, 4, True),
    (O.ENMAP, 0.5, 2, False),       # 16 cells per draw
])
def test_device_mask_kernel_semantics(kw, ratio, mask_patch, tube):
    """One msst_draw_masks launch == MaskGenerator.get_batch(_tube_masked) + bool_mask_to_indices in distribution and EXACTLY in
    structure: keep-count per draw, upsampling by `scale`, tube sharing across spectral blocks, and the reference's index slicing
    (quirk C3) -- checked by applying the HOST generator's bool_mask_to_indices to the device-drawn bool mask."""
    spec = O.Spec(**kw)
    m = _simmim(spec, ratio, mask_patch, tube)
    g = m.mask_generator
    B, C, T = 37, spec.C, spec.T
    nm = int(ratio * T)
    mask, idx = m.draw_masks(B, DEV)
    assert mask.dtype == torch.bool and mask.shape == (B, T) and idx.shape == (B, nm) and idx.dtype == torch.int64
    mk = mask.cpu().numpy().reshape(B, C, g.rand_size, g.scale, g.rand_size, g.scale)
    cells = mk[:, :, :, 0, :, 0]                                              # [B, C, rand, rand]
    assert np.array_equal(mk, np.broadcast_to(cells[:, :, :, None, :, None], mk.shape))      # constant over every scale x scale cell
    assert (cells.reshape(B, C, -1).sum(-1) == g.mask_count).all()             # exactly mask_count cells per draw
    if tube:
        assert np.array_equal(cells, np.broadcast_to(cells[:, :1], cells.shape))
    else:
        assert any(not np.array_equal(cells[b, 0], cells[b, 1]) for b in range(B))
    want_idx = g.bool_mask_to_indices(mask, B, nm, "cpu")                      # the reference's slicing on the SAME bool mask
    assert torch.equal(idx.cpu(), want_idx)
    # same seed stream position -> different draws; the model consumes them
    mask2, _ = m.draw_masks(B, DEV)
    assert not torch.equal(mask, mask2)
    x = O.synthetic_cube(spec, B, seed=1).to(DEV)
    loss = m(x)
    loss.backward()
    assert torch.isfinite(loss)


def test_device_mask_kernel_uniformity():
    """Every cell is masked with probability mask_count / cells, every SUBSET equally often (chi-square over the C(4,3) = 4 possible
    subsets of the shipped configuration and over the 16-cell / 8-of-16 marginals)."""
    spec = O.Spec(**O.HOUSTON)
    m = _simmim(spec, 0.7, 4, True)
    B = 4096
    counts = np.zeros(4)
    for _ in range(4):
        mask, _ = m.draw_masks(B, DEV)
        cells = mask.cpu().numpy().reshape(B, spec.C, 2, 4, 2, 4)[:, 0, :, 0, :, 0].reshape(B, 4)
        missing = (~cells).argmax(1)                                           # which of the 4 cells stayed visible
        counts += np.bincount(missing, minlength=4)
    n = counts.sum()
    chi2 = float(((counts - n / 4) ** 2 / (n / 4)).sum())
    assert chi2 < 21.1, (counts, chi2)                                         # 3 dof, p = 1e-4
    m2 = _simmim(spec, 0.5, 2, False)
    mask, _ = m2.draw_masks(2048, DEV)
    cells = mask.cpu().numpy().reshape(2048, spec.C, 4, 2, 4, 2)[:, :, :, 0, :, 0].reshape(-1, 16)
    freq = cells.mean(0)
    sd = np.sqrt(0.25 / cells.shape[0])
    assert np.abs(freq - 0.5).max() < 5 * sd, freq


def test_whole_tile_inference_vs_oracle_windows():
    """predict_tiles on a fp32 tile cube AND on raw int16 tiles (standardisation fused): every window equals the CPU oracle's
    encoder_forward on that window; accuracy / valid count computed on the device."""
    spec = O.Spec(**O.HOUSTON, depth=2)
    sd = O.synthetic_state_dict(spec, seed=61)
    m = make_encoder(spec).eval()
    m.load_state_dict(sd)
    m.to(DEV)
    tiles = torch.randn(2, 50, 24, 16)
    full = predict_tiles(m, tiles.to(DEV))
    assert full.shape == (2, 20, 24, 16)
    for (i, j) in [(0, 0), (2, 1)]:
        want = O.encoder_forward(tiles[:, :, 8 * i:8 * i + 8, 8 * j:8 * j + 8].contiguous(), sd, spec)
        assert rel_l2(full[:, :, 8 * i:8 * i + 8, 8 * j:8 * j + 8], want) < 1e-5
    # raw sensor tiles: 48 int16 bands -> 50 model bands, reference statistics from the golden fixture
    g = gold("input_pipeline")
    meta = json.loads(str(g["houston__meta"]))
    raw = O.synthetic_raw_tiles(2, meta["raw_bands"], 64, meta["seed"])
    rt = M.RawTiles(torch.from_numpy(raw).to(DEV), g["houston__means"], g["houston__stds"], image_size=8, pad_bands=meta["pad"])
    logits = predict_tiles(m, rt)                                              # all 64 windows of both tiles, one forward
    assert logits.shape == (2, 20, 64, 64)
    for (i, j) in [(0, 0), (7, 3)]:
        cube = O.input_pipeline(raw, g["houston__means"], g["houston__stds"], 8, crop=(8 * i, 8 * j), pad_bands=meta["pad"])
        assert rel_l2(logits[:, :, 8 * i:8 * i + 8, 8 * j:8 * j + 8], O.encoder_forward(cube, sd, spec)) < 1e-5
    labels = logits.argmax(1)
    labels[0, :4] = -1
    acc, n = tile_accuracy(logits, labels)
    assert float(acc) == 1.0 and int(n) == 2 * 64 * 64 - 4 * 64
