"""Pins the CPU oracle (oracle/maskedsst_oracle.py) to vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import json
import numpy as np
import pytest
import torch

from oracle import maskedsst_oracle as O
from tests.helpers import gold, rel_l2, grad_rows, check_grad_rows

TOL = 2e-6   # fp32 CPU vs fp32 CPU, different op order


def _params(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


@pytest.mark.parametrize("name,kw,zero_pad,B,seed", [
    ("houston_encoder", dict(**O.HOUSTON), 2, 2, 5),
    ("enmap_encoder", dict(**O.ENMAP), 0, 1, 6),
    ("enmap_encoder_spectralpos", dict(**O.ENMAP, spectral_pos_embed=True), 0, 1, 7),
    ("houston_encoder_spectral_only", dict(**O.HOUSTON, spectral_only=True), 2, 2, 8),
])
def test_encoder_matches_reference(name, kw, zero_pad, B, seed):
    g = gold(name)
    spec = O.Spec(**kw)
    sd = O.synthetic_state_dict(spec, seed=seed)
    assert sum(v.numel() for v in sd.values()) == int(g["nparams"])
    x = O.synthetic_cube(spec, B, seed=seed, zero_pad_bands=zero_pad)
    st = int(g["token_stride"])
    tok = O.embed(O.to_patch(x, spec), sd, spec)
    assert rel_l2(tok[:, ::st], g["tokens"]) < TOL
    feats = O.transformer_forward(O.encoder_tokens(x, sd, spec), sd, spec)
    assert rel_l2(feats[:, ::st], g["features"]) < 5 * TOL
    assert rel_l2(O.encoder_forward(x, sd, spec), g["logits"]) < 5 * TOL


def test_param_count_pin():
    """inference_example.ipynb cell 8: 'Model parameters: 1,821,564' (EnMAP, learned pos-embed)."""
    spec = O.Spec(**O.ENMAP)
    assert sum(int(np.prod(s)) for _, s in O.state_dict_layout(spec, False)) == 1_821_564
    assert int(gold("enmap_encoder")["nparams"]) == 1_821_564
    assert sum(int(np.prod(s)) for _, s in O.state_dict_layout(O.Spec(**O.HOUSTON), False)) == 1_714_728
    assert sum(int(np.prod(s)) for _, s in O.state_dict_layout(spec, True)) == 1_841_060


def test_finetune_ce_matches_reference():
    g = gold("houston_finetune_ce")
    spec = O.Spec(**O.HOUSTON)
    p = _params(O.synthetic_state_dict(spec, seed=9))
    x = O.synthetic_cube(spec, 3, seed=9)
    labels = torch.from_numpy(g["labels"])
    logits = O.encoder_forward(x, p, spec)
    loss = O.cross_entropy(logits, labels)
    loss.backward()
    assert rel_l2(logits, g["logits"]) < 5 * TOL
    assert abs(loss.item() - float(g["loss"])) < 5 * TOL * abs(float(g["loss"]))
    rows = grad_rows([(k, v.grad) for k, v in p.items() if v.grad is not None])
    check_grad_rows(rows, g["grad_names"], g["grad_rows"], tol=2e-5)
    assert rel_l2(p["mlp_head.1.weight"].grad, g["grad_head_w"]) < 2e-5
    assert rel_l2(p["to_patch_embedding.pre_norm.weight"].grad, g["grad_pre_norm_w"]) < 2e-5


@pytest.mark.parametrize("name,kw", [
    ("houston_simmim_tube", dict(**O.HOUSTON)),
    ("enmap_simmim_block", dict(**O.ENMAP)),
    ("houston_simmim_spectralpos", dict(**O.HOUSTON, spectral_pos_embed=True)),
    ("houston_simmim_patchembed", dict(**O.HOUSTON, blockwise_patch_embed=False)),
])
def test_simmim_matches_reference(name, kw):
    g = gold(name)
    meta = json.loads(str(g["meta"]))
    spec = O.Spec(**kw)
    p = _params(O.synthetic_state_dict(spec, seed=meta["seed"], simmim=True, blockwise_decoder=meta["blockwise_decoder"]))
    x = O.synthetic_cube(spec, meta["B"], seed=meta["seed"], zero_pad_bands=meta["zero_pad"])
    # the oracle's own host mask generator must reproduce the reference's draw bit for bit
    np.random.seed(meta["seed"])
    nm = int(meta["ratio"] * spec.T)
    mask, idx = O.MaskGen(spec.image_size, meta["mask_patch"], spec.spatial_patch_size, meta["ratio"]).batch(
        meta["B"], spec.C, nm, meta["tube"])
    assert np.array_equal(mask.numpy(), g["mask"]) and np.array_equal(idx.numpy(), g["idx"])
    loss = O.simmim_forward(x, p, spec, mask, idx, blockwise_decoder=meta["blockwise_decoder"])
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 5 * TOL * abs(float(g["loss"]))
    rows = grad_rows([(k, v.grad) for k, v in p.items() if v.grad is not None])
    names = [str(n) for n in g["grad_names"]]
    # PatchEmbed case: the reference reports the shared tensors under alias names (to_patch.1.*, patch_to_emb.*)
    alias = {"to_patch.1.weight": "encoder.to_patch_embedding.to_patch.1.weight",
             "to_patch.1.bias": "encoder.to_patch_embedding.to_patch.1.bias",
             "patch_to_emb.0.weight": "encoder.to_patch_embedding.embed.0.weight",
             "patch_to_emb.0.bias": "encoder.to_patch_embedding.embed.0.bias",
             "patch_to_emb.1.weight": "encoder.to_patch_embedding.embed.1.weight",
             "patch_to_emb.1.bias": "encoder.to_patch_embedding.embed.1.bias"}
    for a, b in alias.items():
        if b in rows:
            rows[a] = rows[b]
    # projections are keyed by name: recompute alias projections under the alias key
    for a, b in alias.items():
        if b in p and p[b].grad is not None:
            rows[a] = grad_rows([(a, p[b].grad)])[a]
    check_grad_rows(rows, names, g["grad_rows"], tol=2e-5)
    for k in g.files:
        if k.startswith("grad__"):
            assert rel_l2(p[k[6:]].grad, g[k]) < 2e-5, k


V1_CASES = [("houston_v1_intermediate", dict(**O.HOUSTON, v1=True)),
            ("houston_v1_linearmerge", dict(**O.HOUSTON, v1=True, v1_merge="linear", depth=2))]


@pytest.mark.parametrize("name,kw", V1_CASES)
def test_v1_matches_reference(name, kw):
    """Legacy ViTSpatialSpectral_V1 (+ SimMIM with the shared decoder, `intermediate_losses`)."""
    g = gold(name)
    meta = json.loads(str(g["meta"]))
    spec = O.Spec(**kw)
    x = O.synthetic_cube(spec, meta["B"], seed=meta["seed"], zero_pad_bands=meta["zero_pad"])
    assert rel_l2(O.encoder_forward(x, O.synthetic_state_dict(spec, seed=meta["seed"]), spec), g["logits"]) < 5 * TOL
    p = _params(O.synthetic_state_dict(spec, seed=meta["seed"] + 100, simmim=True, blockwise_decoder=False))
    np.random.seed(meta["seed"])
    mask, idx = O.MaskGen(spec.image_size, 4, 1, 0.7).batch(meta["B"], spec.C, int(0.7 * spec.T), True)
    assert np.array_equal(mask.numpy(), g["mask"]) and np.array_equal(idx.numpy(), g["idx"])
    loss = O.simmim_forward(x, p, spec, mask, idx, blockwise_decoder=False, intermediate_losses=meta["intermediate"])
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 5 * TOL * abs(float(g["loss"]))
    rows = grad_rows([(k, v.grad) for k, v in p.items() if v.grad is not None])
    check_grad_rows(rows, [str(n) for n in g["grad_names"]], g["grad_rows"], tol=2e-5)
    assert rel_l2(p["mask_token"].grad, g["grad__mask_token"]) < 2e-5
    assert rel_l2(p["to_pixels.weight"].grad, g["grad__to_pixels_weight"]) < 2e-5
    pg = p["encoder.pos_embedding"].grad[0, [0, 1, 2, spec.T]]
    assert float(pg[0].abs().max()) == 0.0          # row 0 is dead on the SimMIM path (C11)
    assert rel_l2(pg, g["grad__pos_embedding_rows"]) < 2e-5


@pytest.mark.parametrize("tag", ["houston", "enmap", "enmap_tightclip"])
def test_input_pipeline_matches_reference_bit_exact(tag):
    """Standardize* (float64) -> ToTensor (fp32) -> clip / zero-pad -> batch crop: the oracle restatement reproduces the
    reference's cube bit for bit."""
    g = gold("input_pipeline")
    meta = json.loads(str(g[f"{tag}__meta"]))
    raw = O.synthetic_raw_tiles(meta["B"], meta["raw_bands"], 64, meta["seed"])
    cube = O.input_pipeline(raw, g[f"{tag}__means"], g[f"{tag}__stds"], 8, crop=tuple(meta["crop"]), pad_bands=meta["pad"],
                            clip=tuple(meta["clip"]) if meta["clip"] else None)
    assert cube.shape == g[f"{tag}__cube"].shape
    assert np.array_equal(cube.numpy().view(np.uint32), g[f"{tag}__cube"].view(np.uint32))


def test_maskgen_matches_reference():
    g = gold("maskgen")
    for k in g.files:
        if not k.startswith("mask__"):
            continue
        tag = k[6:]
        f = dict((t[0], t[1:]) for t in tag.split("_"))
        seed, B, C, tube, ratio, mps, img = int(f["s"]), int(f["B"]), int(f["C"]), bool(int(f["t"])), float(f["r"]), int(f["m"]), int(f["i"])
        np.random.seed(seed)
        mask, idx = O.MaskGen(img, mps, 1, ratio).batch(B, C, int(ratio * C * img * img), tube)
        assert np.array_equal(mask.numpy(), g["mask__" + tag])
        assert np.array_equal(idx.numpy(), g["idx__" + tag])


def test_sincos_matches_reference():
    g = gold("sincos")
    assert np.allclose(O.sincos_2d(64, 8), g["pos2d_64_8"], atol=1e-12)
    assert np.allclose(O.sincos_2d(32, 4), g["pos2d_32_4"], atol=1e-12)
    assert np.allclose(O.sincos_1d(32, np.arange(20)), g["pos1d_32_20"], atol=1e-12)
    assert np.allclose(O.sincos_1d(32, np.array([0, 3, 4, 9, 17])), g["pos1d_32_odd"], atol=1e-12)


def test_adam_rules_match_torch_optim():
    """oracle.adam_step vs torch.optim.AdamW / Adam (the update rules of src/utils.py:36-44, finetune.py:133)."""
    torch.manual_seed(0)
    for decoupled, wd, Opt in [(True, 0.05, torch.optim.AdamW), (False, 0.005, torch.optim.Adam)]:
        p0 = torch.randn(257, dtype=torch.float64)
        p = torch.nn.Parameter(p0.clone())
        opt = Opt([p], lr=0.008, weight_decay=wd)
        q, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        for step in range(1, 4):
            g = torch.randn(257, dtype=torch.float64)
            p.grad = g.clone()
            opt.step()
            q, m, v = O.adam_step(q, g, m, v, step, lr=0.008, weight_decay=wd, decoupled=decoupled)
            assert torch.allclose(q, p.detach(), rtol=1e-12, atol=1e-14)
