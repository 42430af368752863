"""Raw-tile input: the reference's input pipeline (band standardisation in float64 -> fp32, clip,
Houston zero-pad, batch crop) fused into the patch-embedding / decoder kernels.  Checked (a) bit for bit against the cube the
UNMODIFIED reference pipeline produced (tests/golden/input_pipeline.npz), and (b) end to end: model(RawTiles) == model(cube)."""
import json

import numpy as np
import pytest
import torch

import maskedsst_b200 as M
from oracle import maskedsst_oracle as O
from tests.helpers import gold, rel_l2
from tests.test_gpu_parity import make_encoder

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(tag):
    g = gold("input_pipeline")
    meta = json.loads(str(g[f"{tag}__meta"]))
    raw = O.synthetic_raw_tiles(meta["B"], meta["raw_bands"], 64, meta["seed"])
    rt = M.RawTiles(torch.from_numpy(raw).to(DEV), g[f"{tag}__means"], g[f"{tag}__stds"], image_size=8, crop=tuple(meta["crop"]),
                    pad_bands=meta["pad"], clip=tuple(meta["clip"]) if meta["clip"] else None)
    return g, meta, raw, rt


@pytest.mark.parametrize("tag", ["houston", "enmap", "enmap_tightclip"])
def test_fused_pipeline_pixels_bit_exact_vs_reference(tag):
    """The kernels' pixel reader is observed through the PatchEmbed LayerNormed-patch output with identity LN parameters off --
    simpler: materialize() (torch) and the golden agree bit for bit, and the kernel path on RawTiles gives bit-identical tokens
    to the kernel path on the golden cube."""
    g, meta, raw, rt = _case(tag)
    want = g[f"{tag}__cube"]
    assert tuple(rt.shape) == want.shape
    assert np.array_equal(rt.materialize().cpu().numpy().view(np.uint32), want.view(np.uint32))
    spec = O.Spec(**(O.HOUSTON if tag == "houston" else O.ENMAP), depth=1)
    m = make_encoder(spec).eval()
    m.load_state_dict(O.synthetic_state_dict(spec, seed=3))
    m.to(DEV)
    with torch.no_grad():
        a = m.to_patch_embedding(rt)
        b = m.to_patch_embedding(torch.from_numpy(want).to(DEV))
        assert torch.equal(a, b)
        assert torch.equal(m(rt), m(torch.from_numpy(want).to(DEV)))
        # and against the CPU oracle fed with the oracle's own restatement of the pipeline
        cube = O.input_pipeline(raw, g[f"{tag}__means"], g[f"{tag}__stds"], 8, crop=tuple(meta["crop"]), pad_bands=meta["pad"],
                                clip=tuple(meta["clip"]) if meta["clip"] else None)
        assert rel_l2(m(rt), O.encoder_forward(cube, O.synthetic_state_dict(spec, seed=3), spec)) < 1e-5


@pytest.mark.parametrize("tag,blockwise", [("houston", True), ("enmap", True), ("houston", False)])
def test_simmim_step_on_raw_tiles_equals_cube(tag, blockwise):
    """SimMIM step (forward + backward): raw-tile input == fp32-cube input, bit for bit (same kernels, same pixel values), incl. the
    decoder's target gather (raw pixels for the blockwise embedding, LayerNormed patches for PatchEmbed)."""
    g, meta, raw, rt = _case(tag)
    spec = O.Spec(**(O.HOUSTON if tag == "houston" else O.ENMAP), depth=1, blockwise_patch_embed=blockwise)
    sd = O.synthetic_state_dict(spec, seed=4, simmim=True, blockwise_decoder=blockwise)
    cube = torch.from_numpy(g[f"{tag}__cube"]).to(DEV)
    np.random.seed(1)
    res = []
    masks = None
    for inp in (rt, cube):
        m = M.SimMIMSpatialSpectral(encoder=make_encoder(spec), masking_ratio=0.7, mask_patch_size=4, tube_masking=True,
                                    to_pixels_per_spectral_block=blockwise).train()
        m.load_state_dict(sd, strict=False)
        m.to(DEV)
        if masks is None:
            masks = m.draw_masks(meta["B"], DEV)
        loss = m(inp, masks=masks)
        loss.backward()
        res.append((loss.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    assert torch.equal(res[0][0], res[1][0])
    for k, v in res[1][1].items():
        # gradients pass through fp32 atomics (order varies run to run): equal up to summation order
        assert rel_l2(res[0][1][k], v) < 1e-5 or float(v.norm()) < 1e-9, k
    # vs the oracle on the reference-built cube
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = O.simmim_forward(torch.from_numpy(g[f"{tag}__cube"]), p, spec, masks[0].cpu(), masks[1].cpu(), blockwise_decoder=blockwise)
    assert abs(res[0][0].item() - want.item()) < 1e-5 * abs(want.item())


def test_raw_tiles_argument_checks():
    raw = torch.zeros(2, 48, 64, 64, dtype=torch.int16, device=DEV)
    ones = np.ones(48)
    with pytest.raises(RuntimeError):
        M.RawTiles(raw.cpu(), ones, ones, image_size=8)                      # no CPU path
    with pytest.raises(RuntimeError):
        M.RawTiles(raw, ones[:40], ones, image_size=8)                       # one mean per band
    with pytest.raises(RuntimeError):
        M.RawTiles(raw, ones, ones, image_size=8, crop=(60, 0))              # window outside the tile
    rt = M.RawTiles(raw, ones, ones, image_size=8, pad_bands=0)              # 48 bands into a 50-band model
    m = make_encoder(O.Spec(**O.HOUSTON, depth=1)).to(DEV).eval()
    with pytest.raises(RuntimeError):
        m(rt)
    f32 = M.RawTiles(torch.randn(1, 50, 16, 16, device=DEV), np.zeros(50), np.ones(50), image_size=8, crop=(8, 8))
    with torch.no_grad():
        assert torch.equal(m(f32), m(f32.materialize()))
