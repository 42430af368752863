"""Generate golden vectors from the UNMODIFIED reference modules (build container only).

    python tests/golden/make_golden.py            # needs /root/reference, writes tests/golden/*.npz

The reference has no tests or known-answer vectors, so its own CPU fp32 outputs on
seeded synthetic weights/inputs are the parity pin.  Weights and inputs come from
oracle.maskedsst_oracle.synthetic_state_dict / synthetic_cube (numpy PCG64 -> reproducible on the GPU
box, where /root/reference does not exist); only the reference's OUTPUTS are stored.
Gradients are stored as per-tensor (l2 norm, sum, signed random projection) triples plus a few full
small tensors, to keep fixtures small.
"""
import os, sys, json
import numpy as np
np.float = float  # src/pos_embed.py:52 uses the alias removed in numpy>=1.24; shim, no edit
REF = os.environ.get("MSST_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the reference's `src` is a namespace package (no __init__.py), so a regular `src` package anywhere on sys.path (this
# repo's drop-in alias!) would win: keep the repo root and the script directory off the path until the reference is imported
sys.path[:] = [REF] + [p for p in sys.path if os.path.abspath(p or os.getcwd()) not in (ROOT, os.path.dirname(os.path.abspath(__file__)))]
import torch
from src.vit_spatial_spectral import ViTSpatialSpectral, ViTSpatialSpectral_V1   # reference
from src.vit_simmim_original import SimMIMSpatialSpectral         # reference
assert os.path.abspath(list(sys.modules["src"].__path__)[0]).startswith(REF)
sys.path.append(ROOT)
from oracle import maskedsst_oracle as O

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def ref_encoder(spec, dropout=0.0):
    return ViTSpatialSpectral(
        image_size=spec.image_size, spatial_patch_size=spec.spatial_patch_size,
        spectral_patch_size=spec.spectral_patch_size, num_classes=spec.num_classes, dim=spec.dim,
        depth=spec.depth, heads=spec.heads, mlp_dim=spec.mlp_dim, dropout=dropout, emb_dropout=dropout,
        channels=spec.channels, spectral_pos_embed=spec.spectral_pos_embed,
        blockwise_patch_embed=spec.blockwise_patch_embed, spectral_pos=spec.pos(),
        spectral_only=spec.spectral_only)


def proj_vec(shape, key):
    rng = np.random.Generator(np.random.PCG64(abs(hash_str(key)) % (2 ** 31)))
    return rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=shape)


def hash_str(s):
    h = 0
    for ch in s:
        h = (h * 131 + ord(ch)) % 2147483647
    return h


def grad_summary(named):
    names, rows = [], []
    for k, g in named:
        g = g.detach().double().numpy()
        names.append(k)
        rows.append([np.sqrt((g * g).sum()), g.sum(), (g * proj_vec(g.shape, k)).sum()])
    return names, np.asarray(rows, dtype=np.float64)


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in kw.items()})


def case_encoder(name, spec, B, zero_pad, seed):
    sd = O.synthetic_state_dict(spec, seed=seed, simmim=False)
    m = ref_encoder(spec).eval()
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == \
        [(k, s) for k, s in O.state_dict_layout(spec, False)] or \
        dict((k, tuple(v.shape)) for k, v in m.state_dict().items()) == dict(O.state_dict_layout(spec, False))
    m.load_state_dict(sd, strict=True)
    x = O.synthetic_cube(spec, B, seed=seed, zero_pad_bands=zero_pad)
    with torch.no_grad():
        tok = m.to_patch_embedding(x)
        feats = m.forward_features(x)
        logits = m(x)
    nparams = sum(p.numel() for p in m.parameters())
    st = 1 if spec.T <= 320 else 8   # keep fixtures small: every 8th token for the EnMAP-sized cases
    save(name, tokens=tok.numpy()[:, ::st], features=feats.numpy()[:, ::st], token_stride=np.int64(st),
         logits=logits.numpy(), nparams=np.int64(nparams),
         meta=json.dumps(dict(spec=spec.__dict__ | {"spectral_pos": None}, B=B, zero_pad=zero_pad, seed=seed)))
    return m


def case_ce(name, spec, B, seed):
    """finetune step: CE(ignore_index=-1) on logits, grads of every encoder parameter."""
    sd = O.synthetic_state_dict(spec, seed=seed, simmim=False)
    m = ref_encoder(spec).train()   # dropout p = 0 -> deterministic
    m.load_state_dict(sd, strict=True)
    x = O.synthetic_cube(spec, B, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    labels = torch.from_numpy(rng.integers(-1, spec.num_classes, (B, spec.image_size, spec.image_size)).astype(np.int64))
    logits = m(x)
    loss = torch.nn.CrossEntropyLoss(ignore_index=-1)(logits, labels)
    loss.backward()
    names, rows = grad_summary([(k, p.grad) for k, p in m.named_parameters()])
    save(name, logits=logits.detach().numpy(), loss=np.float64(loss.item()), labels=labels.numpy(),
         grad_names=np.array(names), grad_rows=rows,
         grad_head_w=m.mlp_head[1].weight.grad.numpy(), grad_pre_norm_w=m.to_patch_embedding.pre_norm.weight.grad.numpy(),
         meta=json.dumps(dict(B=B, seed=seed)))


def case_simmim(name, spec, B, seed, tube, blockwise_decoder=True, mask_patch=4, ratio=0.7, zero_pad=0):
    sd = O.synthetic_state_dict(spec, seed=seed, simmim=True, blockwise_decoder=blockwise_decoder)
    enc = ref_encoder(spec)
    m = SimMIMSpatialSpectral(encoder=enc, masking_ratio=ratio, mask_patch_size=mask_patch, tube_masking=tube,
                              to_pixels_per_spectral_block=blockwise_decoder).train()
    ref_keys = dict((k, tuple(v.shape)) for k, v in m.state_dict().items())
    # with PatchEmbed the wrapper registers alias keys (to_patch.1.*, patch_to_emb.*) sharing encoder tensors
    lay = dict(O.state_dict_layout(spec, True, blockwise_decoder))
    assert all(k in ref_keys and ref_keys[k] == s for k, s in lay.items()), "layout mismatch"
    extra = sorted(set(ref_keys) - set(lay))
    m.load_state_dict(sd, strict=False)
    # capture the mask pair the reference draws (host numpy RNG, seeded) by wrapping the generator
    np.random.seed(seed)
    captured = {}
    gen = m.mask_generator
    for fn in ("get_batch", "get_batch_tube_masked"):
        orig = getattr(gen, fn)
        def wrap(*a, _orig=orig, **k):
            bm, ix = _orig(*a, **k)
            captured["mask"], captured["idx"] = bm.clone(), ix.clone()
            return bm, ix
        setattr(gen, fn, wrap)
    x = O.synthetic_cube(spec, B, seed=seed, zero_pad_bands=zero_pad)
    loss = m(x)
    loss.backward()
    seen = set()
    named = []
    for k, p in m.named_parameters():
        if id(p) in seen or p.grad is None:
            continue
        seen.add(id(p)); named.append((k, p.grad))
    names, rows = grad_summary(named)
    full = {}
    for k, p in m.named_parameters():
        if k in ("mask_token", "encoder.pos_embedding", "encoder.channel_embed", "encoder.pos_embed",
                 "to_pixels.layers.0.weight", "to_pixels.weight",
                 "encoder.to_patch_embedding.blockwise_embed.0.weight",
                 "encoder.spatial_spectral_transformer.1.layers.0.1.fn.net.0.weight",
                 "encoder.spatial_spectral_transformer.3.layers.3.0.fn.to_out.0.bias"):
            full["grad__" + k] = p.grad.numpy()
    save(name, loss=np.float64(loss.item()), mask=captured["mask"].numpy(), idx=captured["idx"].numpy(),
         grad_names=np.array(names), grad_rows=rows, extra_keys=np.array(extra),
         meta=json.dumps(dict(B=B, seed=seed, tube=tube, blockwise_decoder=blockwise_decoder,
                              mask_patch=mask_patch, ratio=ratio, zero_pad=zero_pad)), **full)


def ref_encoder_v1(spec, merge="avgpool"):
    return ViTSpatialSpectral_V1(
        image_size=spec.image_size, spatial_patch_size=spec.spatial_patch_size, spectral_patch_size=spec.spectral_patch_size,
        num_classes=spec.num_classes, dim=spec.dim, depth=spec.depth, heads=spec.heads, mlp_dim=spec.mlp_dim,
        channels=spec.channels, merge=merge)


def case_v1(name, spec, B, seed, zero_pad, intermediate):
    """Legacy ViTSpatialSpectral_V1: encoder logits, then a SimMIM step (shared decoder; `intermediate_losses`)."""
    sd = O.synthetic_state_dict(spec, seed=seed, simmim=False)
    enc = ref_encoder_v1(spec, spec.v1_merge).eval()
    assert dict((k, tuple(v.shape)) for k, v in enc.state_dict().items()) == dict(O.state_dict_layout(spec, False))
    enc.load_state_dict(sd, strict=True)
    x = O.synthetic_cube(spec, B, seed=seed, zero_pad_bands=zero_pad)
    with torch.no_grad():
        logits = enc(x)
    ssd = O.synthetic_state_dict(spec, seed=seed + 100, simmim=True, blockwise_decoder=False)
    m = SimMIMSpatialSpectral(encoder=ref_encoder_v1(spec, spec.v1_merge), masking_ratio=0.7, mask_patch_size=4, tube_masking=True,
                              intermediate_losses=intermediate, to_pixels_per_spectral_block=False).train()
    ref_keys = dict((k, tuple(v.shape)) for k, v in m.state_dict().items())
    lay = dict(O.state_dict_layout(spec, True, False))
    assert all(k in ref_keys and ref_keys[k] == s for k, s in lay.items()), "layout mismatch"
    extra = sorted(set(ref_keys) - set(lay))
    m.load_state_dict(ssd, strict=False)
    np.random.seed(seed)
    captured = {}
    orig = m.mask_generator.get_batch_tube_masked
    def wrap(*a, **k):
        bm, ix = orig(*a, **k)
        captured["mask"], captured["idx"] = bm.clone(), ix.clone()
        return bm, ix
    m.mask_generator.get_batch_tube_masked = wrap
    loss = m(x)
    loss.backward()
    seen, named = set(), []
    for k, p in m.named_parameters():
        if id(p) in seen or p.grad is None:
            continue
        seen.add(id(p)); named.append((k, p.grad))
    names, rows = grad_summary(named)
    save(name, logits=logits.numpy(), loss=np.float64(loss.item()), mask=captured["mask"].numpy(), idx=captured["idx"].numpy(),
         grad_names=np.array(names), grad_rows=rows, extra_keys=np.array(extra),
         grad__mask_token=m.mask_token.grad.numpy(), grad__to_pixels_weight=m.to_pixels.weight.grad.numpy(),
         grad__pos_embedding_rows=m.encoder.pos_embedding.grad.numpy()[0, [0, 1, 2, spec.T]],
         meta=json.dumps(dict(B=B, seed=seed, zero_pad=zero_pad, intermediate=intermediate, merge=spec.v1_merge)))


def _stub_missing(roots=("rasterio", "spectral", "torchmetrics", "wandb")):
    """The reference's data modules import I/O libraries that are not installed here; only their Standardize* / ToTensor
    classes (pure numpy / torch) are exercised, so absent libraries are replaced by empty stub modules."""
    import importlib.abc, importlib.machinery, types
    class _Stub(types.ModuleType):
        __path__ = []
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {})
    class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, name, path=None, target=None):
            if name.split(".")[0] in roots:
                return importlib.machinery.ModuleSpec(name, self, is_package=True)
        def create_module(self, spec):
            return _Stub(spec.name)
        def exec_module(self, module):
            pass
    sys.meta_path.append(_Finder())


def case_input_pipeline():
    """The reference's input pipeline on synthetic raw int16 tiles: Standardize* (numpy float64) -> ToTensor (fp32) ->
    [EnMAP: torch.clip] / [Houston: F.pad 48 -> 50] -> one crop window per batch (pretrain.py:99-107)."""
    _stub_missing()
    import torch.nn.functional as F
    from src.data_enmap import StandardizeEnMAP, ToTensor
    from src.data_houston2018 import StandardizeHouston2018
    out = {}
    for tag, std, nb, pad, clip, crop, seed in [("houston", StandardizeHouston2018(), 48, 2, None, (3, 41), 21),
                                                ("enmap", StandardizeEnMAP(use_clipped=True), None, 0, (-200, 10000), (56, 0), 22),
                                                ("enmap_tightclip", StandardizeEnMAP(use_clipped=False), None, 0, (-1.5, 2.0), (17, 30), 23)]:
        means = std.means_clipped if getattr(std, "use_clipped", False) else std.means
        stds = std.stds_clipped if getattr(std, "use_clipped", False) else std.stds
        nb = nb or len(means)
        raw = O.synthetic_raw_tiles(2, nb, 64, seed)
        imgs = []
        for i in range(raw.shape[0]):                     # per-sample transform chain, as the datasets apply it
            img = ToTensor()(std(raw[i]))
            if clip is not None:
                img = torch.clip(img, min=clip[0], max=clip[1])
            if pad:
                img = F.pad(img, (0, 0, 0, 0, 0, pad), "constant", 0)
            imgs.append(img)
        batch = torch.stack(imgs)
        x, y = crop
        cube = batch[:, :, x: x + 8, y: y + 8]
        out[f"{tag}__means"], out[f"{tag}__stds"] = np.asarray(means, dtype=np.float64), np.asarray(stds, dtype=np.float64)
        out[f"{tag}__cube"] = cube.numpy()
        out[f"{tag}__meta"] = json.dumps(dict(raw_bands=nb, pad=pad, clip=clip, crop=crop, seed=seed, B=2))
    save("input_pipeline", **out)


def case_maskgen():
    """MaskGenerator draws for several seeds/shapes (host numpy RNG)."""
    from src.vit_simmim_original import MaskGenerator
    out = {}
    for seed, B, C, tube, ratio, mps, img in [(5, 4, 5, True, .7, 4, 8), (6, 3, 20, False, .7, 4, 8),
                                              (7, 5, 5, False, .5, 2, 8), (8, 2, 4, True, .6, 4, 16)]:
        g = MaskGenerator(input_size=img, mask_patch_size=mps, model_patch_size=1, mask_ratio=ratio)
        np.random.seed(seed)
        nm = int(ratio * C * img * img)
        fn = g.get_batch_tube_masked if tube else g.get_batch
        bm, ix = fn(batch_size=B, channel_tokens=C, num_masked=nm, device="cpu")
        tag = f"s{seed}_B{B}_C{C}_t{int(tube)}_r{ratio}_m{mps}_i{img}"
        out["mask__" + tag] = bm.numpy(); out["idx__" + tag] = ix.numpy()
    save("maskgen", **out)


def case_sincos():
    from src.pos_embed import get_2d_sincos_pos_embed, get_1d_sincos_pos_embed_from_grid
    save("sincos", pos2d_64_8=get_2d_sincos_pos_embed(64, 8), pos2d_32_4=get_2d_sincos_pos_embed(32, 4),
         pos1d_32_20=get_1d_sincos_pos_embed_from_grid(32, np.arange(20)),
         pos1d_32_odd=get_1d_sincos_pos_embed_from_grid(32, np.array([0, 3, 4, 9, 17])))


if __name__ == "__main__":
    H = O.Spec(**O.HOUSTON)
    E = O.Spec(**O.ENMAP)
    if sys.argv[1:] == ["v1"]:    # only the cases added after the first fixture set (keeps the older .npz files byte-identical)
        case_v1("houston_v1_intermediate", O.Spec(**O.HOUSTON, v1=True), B=2, seed=14, zero_pad=2, intermediate=True)
        case_v1("houston_v1_linearmerge", O.Spec(**O.HOUSTON, v1=True, v1_merge="linear", depth=2), B=2, seed=15, zero_pad=0,
                intermediate=False)
        sys.exit(0)
    if sys.argv[1:] == ["input"]:
        case_input_pipeline()
        sys.exit(0)
    m = case_encoder("houston_encoder", H, B=2, zero_pad=2, seed=5)
    case_encoder("enmap_encoder", E, B=1, zero_pad=0, seed=6)
    case_encoder("enmap_encoder_spectralpos", O.Spec(**O.ENMAP, spectral_pos_embed=True), B=1, zero_pad=0, seed=7)
    case_encoder("houston_encoder_spectral_only", O.Spec(**O.HOUSTON, spectral_only=True), B=2, zero_pad=2, seed=8)
    case_ce("houston_finetune_ce", H, B=3, seed=9)
    case_simmim("houston_simmim_tube", H, B=3, seed=5, tube=True, zero_pad=2)
    case_simmim("enmap_simmim_block", E, B=2, seed=11, tube=False)
    case_simmim("houston_simmim_spectralpos", O.Spec(**O.HOUSTON, spectral_pos_embed=True), B=2, seed=12, tube=True)
    case_simmim("houston_simmim_patchembed", O.Spec(**O.HOUSTON, blockwise_patch_embed=False), B=2, seed=13,
                tube=True, blockwise_decoder=False)
    case_maskgen()
    case_sincos()
