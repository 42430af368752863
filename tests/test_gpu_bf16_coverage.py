"""GPU, bf16 mode (the BENCHMARKED mode): the model-level cases round 1 only covered in fp32 -- every SimMIM golden, the finetune CE
step, spectral-only, the legacy V1 encoder with intermediate losses, the second-tier heads, a batch large enough for multi-wave
persistent scheduling (B = 256 against the CPU oracle), and a 30-step loss trajectory against the fp32 mode.
Tolerances (north star): logits and loss rel-err <= 1e-2.  Gradients: per-tensor rel-l2 <= 5e-2, cosine >= 0.998 (they carry the
bf16 rounding of the activations AND of the upstream gradients; BASELINE's north_star states no gradient tolerance)."""
import json

import numpy as np
import pytest
import torch

import maskedsst_b200 as M
from maskedsst_b200.optim import FusedAdam
from oracle import maskedsst_oracle as O
from tests.helpers import gold, rel_l2
from tests.test_gpu_parity import make_encoder, _head_variant_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check_grads(m, ref_params, tol=5e-2, cos_min=0.998):
    worst, n = 0.0, 0
    seen = set()
    for k, v in m.named_parameters(remove_duplicate=False):
        if k not in ref_params or id(v) in seen or v.grad is None:
            continue
        seen.add(id(v))
        gw = ref_params[k].grad
        if gw is None or float(gw.norm()) < 1e-10:
            continue
        e = rel_l2(v.grad, gw)
        cos = float((v.grad.cpu().double().flatten() @ gw.double().flatten()) / (v.grad.double().norm().cpu() * gw.double().norm()))
        worst = max(worst, e)
        n += 1
        assert e < tol and cos > cos_min, (k, e, cos)
    assert n > 20
    return worst


@pytest.mark.parametrize("name,kw", [
    ("houston_simmim_tube", dict(**O.HOUSTON)),
    ("enmap_simmim_block", dict(**O.ENMAP)),
    ("houston_simmim_spectralpos", dict(**O.HOUSTON, spectral_pos_embed=True)),
    ("houston_simmim_patchembed", dict(**O.HOUSTON, blockwise_patch_embed=False)),
])
def test_simmim_bf16_step_all_goldens(name, kw):
    g = gold(name)
    meta = json.loads(str(g["meta"]))
    spec = O.Spec(**kw)
    sd = O.synthetic_state_dict(spec, seed=meta["seed"], simmim=True, blockwise_decoder=meta["blockwise_decoder"])
    enc = make_encoder(spec)
    enc.precision = "bf16"
    m = M.SimMIMSpatialSpectral(encoder=enc, masking_ratio=meta["ratio"], mask_patch_size=meta["mask_patch"],
                                tube_masking=meta["tube"], to_pixels_per_spectral_block=meta["blockwise_decoder"]).train()
    m.load_state_dict(sd, strict=False)
    m.to(DEV)
    x = O.synthetic_cube(spec, meta["B"], seed=meta["seed"], zero_pad_bands=meta["zero_pad"])
    mask, idx = torch.from_numpy(g["mask"]), torch.from_numpy(g["idx"])
    loss = m(x.to(DEV), masks=(mask.to(DEV), idx.to(DEV)))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-2 * abs(float(g["loss"]))       # golden = the unmodified reference
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.simmim_forward(x, p, spec, mask, idx, blockwise_decoder=meta["blockwise_decoder"]).backward()
    worst = _check_grads(m, p)
    print(name, "bf16 loss rel", abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])), "worst grad rel-l2", worst)


def test_finetune_ce_bf16_step_vs_reference_golden():
    g = gold("houston_finetune_ce")
    spec = O.Spec(**O.HOUSTON)
    sd = O.synthetic_state_dict(spec, seed=9)
    m = make_encoder(spec).train()
    m.load_state_dict(sd)
    m.precision = "bf16"
    m.to(DEV)
    x = O.synthetic_cube(spec, 3, seed=9)
    labels = torch.from_numpy(g["labels"])
    logits = m(x.to(DEV))
    loss = M.cross_entropy(logits, labels.to(DEV), ignore_index=-1)
    loss.backward()
    assert rel_l2(logits, g["logits"]) < 1e-2
    assert abs(loss.item() - float(g["loss"])) < 1e-2 * abs(float(g["loss"]))
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.cross_entropy(O.encoder_forward(x, p, spec), labels).backward()
    _check_grads(m, p)


def test_spectral_only_bf16_vs_reference_golden():
    g = gold("houston_encoder_spectral_only")
    spec = O.Spec(**O.HOUSTON, spectral_only=True)
    m = make_encoder(spec).eval()
    m.load_state_dict(O.synthetic_state_dict(spec, seed=8), strict=True)
    m.precision = "bf16"
    m.to(DEV)
    x = O.synthetic_cube(spec, 2, seed=8, zero_pad_bands=2).to(DEV)
    with torch.no_grad():
        assert rel_l2(m(x), g["logits"]) < 1e-2


@pytest.mark.parametrize("name,kw", [("houston_v1_intermediate", dict(**O.HOUSTON, v1=True)),
                                     ("houston_v1_linearmerge", dict(**O.HOUSTON, v1=True, v1_merge="linear", depth=2))])
def test_v1_bf16_vs_reference_golden(name, kw):
    g = gold(name)
    meta = json.loads(str(g["meta"]))
    spec = O.Spec(**kw)

    def make():
        return M.ViTSpatialSpectral_V1(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=spec.num_classes,
                                       dim=spec.dim, depth=spec.depth, heads=spec.heads, mlp_dim=spec.mlp_dim, channels=spec.channels,
                                       merge=meta["merge"], precision="bf16")
    enc = make().eval()
    enc.load_state_dict(O.synthetic_state_dict(spec, seed=meta["seed"]), strict=True)
    enc.to(DEV)
    x = O.synthetic_cube(spec, meta["B"], seed=meta["seed"], zero_pad_bands=meta["zero_pad"]).to(DEV)
    with torch.no_grad():
        assert rel_l2(enc(x), g["logits"]) < 1e-2
    m = M.SimMIMSpatialSpectral(encoder=make(), masking_ratio=0.7, mask_patch_size=4, tube_masking=True,
                                intermediate_losses=meta["intermediate"]).train()
    m.load_state_dict(O.synthetic_state_dict(spec, seed=meta["seed"] + 100, simmim=True, blockwise_decoder=False), strict=False)
    m.to(DEV)
    np.random.seed(meta["seed"])
    loss = m(x)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-2 * abs(float(g["loss"]))
    assert rel_l2(m.mask_token.grad, g["grad__mask_token"]) < 5e-2
    assert rel_l2(m.to_pixels.weight.grad, g["grad__to_pixels_weight"]) < 5e-2


@pytest.mark.parametrize("kind", ["pixelwise", "spectral_mlp_head"])
def test_head_variants_bf16_vs_oracle(kind):
    spec = O.Spec(**O.HOUSTON, depth=1)
    torch.manual_seed(0)
    m = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=1, heads=8,
                             mlp_dim=64, channels=50, spectral_pos_embed=False, pixelwise=kind == "pixelwise",
                             spectral_mlp_head=kind == "spectral_mlp_head", precision="bf16")
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.to(DEV).train()
    x = O.synthetic_cube(spec, 3, seed=2)
    got = m(x.to(DEV))
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = _head_variant_oracle(x, p, spec, kind)
    assert got.shape == want.shape and rel_l2(got, want) < 1e-2
    got.square().sum().backward()
    want.square().sum().backward()
    for k, v in m.named_parameters():
        if p[k].grad is not None and float(p[k].grad.norm()) > 1e-9:
            assert rel_l2(v.grad, p[k].grad) < 5e-2, k


def test_simmim_bf16_batch256_vs_cpu_oracle():
    """B = 256 (81,920 tokens: 640 tiles of 128 slots = 4-5 tiles per persistent CTA in each stack, tile changes, the h double
    buffer, the dh rounds of the fused backward) through the full model, against the CPU oracle (~2 s): loss and gradients."""
    spec = O.Spec(**O.HOUSTON)
    sd = O.synthetic_state_dict(spec, seed=31, simmim=True)
    enc = make_encoder(spec)
    enc.precision = "bf16"
    m = M.SimMIMSpatialSpectral(encoder=enc, masking_ratio=0.7, mask_patch_size=4, tube_masking=True, to_pixels_per_spectral_block=True).train()
    m.load_state_dict(sd)
    m.to(DEV)
    B = 256
    x = O.synthetic_cube(spec, B, seed=31, zero_pad_bands=2)
    np.random.seed(31)
    mask, idx = O.MaskGen(8, 4, 1, 0.7).batch(B, spec.C, int(0.7 * spec.T), tube=True)
    loss = m(x.to(DEV), masks=(mask.to(DEV), idx.to(DEV)))
    loss.backward()
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = O.simmim_forward(x, p, spec, mask, idx)
    want.backward()
    assert abs(loss.item() - want.item()) < 1e-2 * abs(want.item())
    worst = _check_grads(m, p)
    print("B=256 bf16: loss rel", abs(loss.item() - want.item()) / abs(want.item()), "worst grad rel-l2", worst)


def test_bf16_vs_fp32_mode_loss_trajectory():
    """30 optimiser steps, dropout 0, same initial weights / data / masks: the bf16 (tensor-core) mode must track the fp32 (FFMA,
    parity) mode -- per-step loss within 2 %, final parameters within 2 % rel-l2 of the total parameter movement."""
    spec = O.Spec(**O.HOUSTON, depth=2)
    sd = O.synthetic_state_dict(spec, seed=41, simmim=True)
    x = O.synthetic_cube(spec, 32, seed=41, zero_pad_bands=2).to(DEV)
    np.random.seed(41)
    masks = [tuple(t.to(DEV) for t in O.MaskGen(8, 4, 1, 0.7).batch(32, spec.C, int(0.7 * spec.T), tube=True)) for _ in range(4)]
    runs = {}
    for prec in ("fp32", "bf16"):
        enc = make_encoder(spec)
        enc.precision = prec
        m = M.SimMIMSpatialSpectral(encoder=enc, masking_ratio=0.7, mask_patch_size=4, tube_masking=True, to_pixels_per_spectral_block=True).train()
        m.load_state_dict(sd)
        m.to(DEV)
        opt = FusedAdam(m.parameters(), lr=0.002, weight_decay=0.05, clamp=1.0)
        p0 = opt.param_arena.clone()
        losses = []
        for s in range(30):
            opt.zero_grad()
            loss = m(x, masks=masks[s % 4])
            loss.backward()
            opt.step()
            losses.append(float(loss))
        runs[prec] = (losses, opt.param_arena.clone(), p0)
    lf, lb = runs["fp32"][0], runs["bf16"][0]
    assert lf[-1] < 0.8 * lf[0]                                   # it actually trains
    for a, b in zip(lf, lb):
        assert abs(a - b) < 2e-2 * abs(a), (lf, lb)
    move = (runs["fp32"][1] - runs["fp32"][2]).norm()
    assert float((runs["bf16"][1] - runs["fp32"][1]).norm() / move) < 0.1
