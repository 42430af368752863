"""CPU: host-side logic of the drop-in surface -- state_dict layout, checkpoint loading semantics, mask generator
bit-compatibility with the reference, init parity with torch.manual_seed."""
import json
import numpy as np
import torch

import maskedsst_b200 as M
from oracle import maskedsst_oracle as O
from tests.helpers import gold


def make_encoder(spec, dropout=0.0):
    return M.ViTSpatialSpectral(
        image_size=spec.image_size, spatial_patch_size=spec.spatial_patch_size, spectral_patch_size=spec.spectral_patch_size,
        num_classes=spec.num_classes, dim=spec.dim, depth=spec.depth, heads=spec.heads, mlp_dim=spec.mlp_dim,
        dropout=dropout, emb_dropout=dropout, channels=spec.channels, spectral_pos_embed=spec.spectral_pos_embed,
        blockwise_patch_embed=spec.blockwise_patch_embed, spectral_pos=spec.pos(), spectral_only=spec.spectral_only)


def test_state_dict_roundtrip_and_load_checkpoint_semantics(tmp_path):
    """Appendix B: pretrain checkpoint -> strip 'encoder.' -> swap head -> strict load (src/utils.py:276-313)."""
    from src.utils import load_checkpoint, Dotdict
    spec = O.Spec(**O.HOUSTON)
    sd = O.synthetic_state_dict(spec, seed=41, simmim=True)
    path = tmp_path / "pretrain.pth"
    torch.save({"config": Dotdict({"a": 1}), "model_state_dict": sd, "lr_current": 0.008}, path)
    enc = make_encoder(O.Spec(channels=50, num_classes=11))
    head_w = enc.mlp_head[1].weight.detach().clone()
    cfg = Dotdict({"checkpoint_path": str(path), "patch_sub": 0, "image_size": 8})
    load_checkpoint(cfg, enc, "mlp_head", "cpu")
    assert torch.equal(enc.mlp_head[1].weight, head_w)           # fresh head kept
    assert torch.equal(enc.pos_embedding, sd["encoder.pos_embedding"])
    k = "spatial_spectral_transformer.3.layers.2.0.fn.to_qkv.weight"
    assert torch.equal(enc.state_dict()[k], sd["encoder." + k])


def test_state_dict_layout_matches_reference_keys():
    for kw, simmim, bd in [(dict(**O.HOUSTON), False, True), (dict(**O.ENMAP), True, True),
                           (dict(**O.ENMAP, spectral_pos_embed=True), True, True),
                           (dict(**O.HOUSTON, spectral_only=True), False, True),
                           (dict(**O.HOUSTON, blockwise_patch_embed=False), True, False)]:
        spec = O.Spec(**kw)
        m = make_encoder(spec)
        if simmim:
            m = M.SimMIMSpatialSpectral(encoder=m, masking_ratio=0.7, mask_patch_size=4, tube_masking=True,
                                        to_pixels_per_spectral_block=bd)
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        lay = dict(O.state_dict_layout(spec, simmim, bd))
        extra = set(ours) - set(lay)
        assert set(lay) <= set(ours) and all(ours[k] == s for k, s in lay.items())
        assert all(k.startswith(("to_patch.", "patch_to_emb.")) for k in extra)   # PatchEmbed alias keys (Appendix B)
    assert sum(p.numel() for p in make_encoder(O.Spec(**O.ENMAP)).parameters()) == 1_821_564   # notebook cell 8


def test_mask_generator_bit_compatible_with_reference():
    g = gold("maskgen")
    for k in g.files:
        if not k.startswith("mask__"):
            continue
        tag = k[6:]
        f = dict((t[0], t[1:]) for t in tag.split("_"))
        seed, B, C, tube, ratio, mps, img = int(f["s"]), int(f["B"]), int(f["C"]), bool(int(f["t"])), float(f["r"]), int(f["m"]), int(f["i"])
        gen = M.MaskGenerator(input_size=img, mask_patch_size=mps, model_patch_size=1, mask_ratio=ratio)
        np.random.seed(seed)
        fn = gen.get_batch_tube_masked if tube else gen.get_batch
        mask, idx = fn(batch_size=B, channel_tokens=C, num_masked=int(ratio * C * img * img), device="cpu")
        assert mask.dtype == torch.bool and idx.dtype == torch.int64
        assert np.array_equal(mask.numpy(), g["mask__" + tag]) and np.array_equal(idx.numpy(), g["idx__" + tag])


def test_sincos_tables_match_reference():
    from src.pos_embed import get_2d_sincos_pos_embed, get_1d_sincos_pos_embed_from_grid
    g = gold("sincos")
    assert np.allclose(get_2d_sincos_pos_embed(64, 8), g["pos2d_64_8"], atol=1e-12)
    assert np.allclose(get_1d_sincos_pos_embed_from_grid(32, np.array([0, 3, 4, 9, 17])), g["pos1d_32_odd"], atol=1e-12)
    enc = make_encoder(O.Spec(**O.ENMAP, spectral_pos_embed=True))
    assert np.allclose(enc.pos_embed[0].detach().numpy(), g["pos2d_64_8"].astype(np.float32))
    assert np.allclose(enc.channel_embed[0].detach().numpy(), g["pos1d_32_20"].astype(np.float32))


def test_reference_import_paths():
    from src.vit_spatial_spectral import ViTSpatialSpectral, get_pos_for_spectral_embedding, MoveAxis
    from src.vit_simmim_original import SimMIMSpatialSpectral
    assert ViTSpatialSpectral is M.ViTSpatialSpectral and SimMIMSpatialSpectral is M.SimMIMSpatialSpectral
    assert get_pos_for_spectral_embedding(10, list(range(400, 600)), list(range(400, 1000)))[:3] == [0, 1, 2]


def test_device_mask_backend_semantics_on_cpu():
    """mask_backend='device' (torch ops, runs on any device): same counts / structure / index slicing as the host generator."""
    for kw, tube in [(dict(**O.HOUSTON), True), (dict(**O.ENMAP), False)]:
        spec = O.Spec(**kw)
        m = M.SimMIMSpatialSpectral(encoder=make_encoder(spec), masking_ratio=0.7, mask_patch_size=4, tube_masking=tube,
                                    to_pixels_per_spectral_block=True)
        m.mask_backend = "device"
        torch.manual_seed(0)
        B = 6
        mask, idx = m.draw_masks(B, "cpu")
        nm = int(0.7 * spec.T)
        assert mask.shape == (B, spec.T) and mask.dtype == torch.bool and idx.shape == (B, nm) and idx.dtype == torch.int64
        assert int(mask.sum()) == B * 3 * 16 * spec.C                          # ceil(4 * .7) = 3 of 4 cells, 16 tokens each
        cells = mask.view(B, spec.C, 2, 4, 2, 4)
        assert torch.equal(cells, cells[:, :, :, :1, :, :1].expand_as(cells))   # 4x4 blocks are uniform
        if tube:
            assert torch.equal(mask.view(B, spec.C, 64), mask.view(B, spec.C, 64)[:, :1].expand(B, spec.C, 64))
        want = torch.from_numpy(O.MaskGen.indices(mask.numpy(), nm))             # the reference's slicing of nonzero()
        assert torch.equal(idx, want)


def test_v1_module_surface_matches_reference_layout():
    """ViTSpatialSpectral_V1 (reference :600-764) on CPU: state_dict keys / shapes == the reference layout (oracle.state_dict_layout,
    verified against the reference by make_golden.py), alias keys of the SimMIM wrapper == the golden `extra_keys`, and the
    constructor-time errors for the two combinations the reference can only fail on at run time."""
    import numpy as np
    import pytest
    import maskedsst_b200 as M
    from oracle import maskedsst_oracle as O
    from tests.helpers import gold
    for name, kw in [("houston_v1_intermediate", dict(**O.HOUSTON, v1=True)), ("houston_v1_linearmerge", dict(**O.HOUSTON, v1=True, v1_merge="linear", depth=2))]:
        g = gold(name)
        spec = O.Spec(**kw)
        enc = M.ViTSpatialSpectral_V1(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=spec.depth,
                                      heads=8, mlp_dim=64, channels=50, merge=spec.v1_merge)
        assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == dict(O.state_dict_layout(spec, False))
        assert enc.num_spatial_patches == 8 and enc.num_patches == 320            # grid SIDE, not its square (:629)
        sim = M.SimMIMSpatialSpectral(encoder=enc, masking_ratio=0.7, mask_patch_size=4, tube_masking=True, intermediate_losses=True)
        extra = sorted(set(sim.state_dict()) - set(dict(O.state_dict_layout(spec, True, False))))
        assert extra == [str(k) for k in g["extra_keys"]]
        assert sim.state_dict()["patch_to_emb.2.weight"].data_ptr() == enc.to_patch_embedding[2].weight.data_ptr()   # aliases share storage
        with pytest.raises(NotImplementedError):
            M.SimMIMSpatialSpectral(encoder=enc, to_pixels_per_spectral_block=True)
    v2 = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=1, heads=8,
                              mlp_dim=64, channels=50, spectral_pos_embed=False)
    with pytest.raises(NotImplementedError):
        M.SimMIMSpatialSpectral(encoder=v2, intermediate_losses=True)


def test_raw_tiles_and_kernels_refuse_cpu_tensors():
    """No CPU path: RawTiles and the ops raise instead of falling back."""
    import numpy as np
    import pytest
    import torch
    import maskedsst_b200 as M
    with pytest.raises(RuntimeError):
        M.RawTiles(torch.zeros(1, 48, 64, 64, dtype=torch.int16), np.zeros(48), np.ones(48), image_size=8)
    enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=1, heads=8,
                               mlp_dim=64, channels=50, spectral_pos_embed=False)
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 50, 8, 8))
