"""B200-native ViTSpatialSpectral: same nn.Module surface, constructor arguments, attribute names and
state_dict keys as the reference's src/vit_spatial_spectral.py (so its .pth checkpoints load unchanged),
with every arithmetic step executed by hand-written sm_100a kernels (libmsst.so, include/msst.h).

Reference map (file:line in the upstream repo):
  PreNorm :22-29, FeedForward :32-44, Attention :47-78, Transformer :81-104   -> parameter containers here;
      the arithmetic of a whole Transformer stack is one msst_transformer_fwd/bwd call
  BlockwisePatchEmbedding :178-229, PatchEmbed :232-253                       -> msst_patch_embed_fwd/bwd
  ViTSpatialSpectral :256-564                                                  -> ViTSpatialSpectral below
  ViTSpatialSpectral_V1 :600-764 (legacy; the only encoder `intermediate_losses` works with) -> ViTSpatialSpectral_V1
The spatial->spectral re-tiling copies of the reference (:409-431) do not exist here: the residual stream stays in
(b, c, s) row order and the spectral stack addresses its sequences with a stride (msst.h, attention section).
"""
from functools import reduce
from operator import mul

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import PREC_FP32, PREC_BF16
from .pos_embed import get_1d_sincos_pos_embed_from_grid, get_2d_sincos_pos_embed


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


_PREC = {"fp32": PREC_FP32, "32-true": PREC_FP32, "bf16": PREC_BF16, "bf16-mixed": PREC_BF16}


# ---------------------------------------------------------------------------------------------------
# layout-only helper modules (no parameters; kept so Sequential indices / state_dict keys match)
# ---------------------------------------------------------------------------------------------------
class Retile(nn.Module):
    """Parameter-free re-tiling between (b, c, s) token order and per-sequence batches -- the reference's
    einops Rearrange layers at indices 0/2/4 of `spatial_spectral_transformer` (:409-431).  Only used when
    the Sequential is called directly; ViTSpatialSpectral.transformer_forward never copies."""

    def __init__(self, kind, c, s):
        super().__init__()
        self.kind, self.c, self.s = kind, c, s

    def forward(self, x):
        c, s, d = self.c, self.s, x.shape[-1]
        if self.kind == "to_spatial":        # b (c s) d -> (b c) s d
            return x.reshape(-1, s, d)
        if self.kind == "spatial_to_spectral":  # (b c) s d -> (b s) c d
            return x.reshape(-1, c, s, d).transpose(1, 2).reshape(-1, c, d)
        if self.kind == "to_spectral":       # b (c s) d -> (b s) c d
            return x.reshape(-1, c, s, d).transpose(1, 2).reshape(-1, c, d)
        if self.kind == "from_spectral":     # (b s) c d -> b (c s) d
            return x.reshape(-1, s, c, d).transpose(1, 2).reshape(-1, c * s, d)
        raise ValueError(self.kind)

    def extra_repr(self):
        return self.kind


class ToPatch(nn.Module):
    """'b (c p0) (h p1) (w p2) -> b c (h w) (p0 p1 p2)' (:197-202) / '-> b (c h w) (p0 p1 p2)' (:235-241)."""

    def __init__(self, p0, p1, flat):
        super().__init__()
        self.p0, self.p1, self.flat = p0, p1, flat

    def forward(self, x):
        b, ch, h, w = x.shape
        c, gh, gw = ch // self.p0, h // self.p1, w // self.p1
        x = x.reshape(b, c, self.p0, gh, self.p1, gw, self.p1).permute(0, 1, 3, 5, 2, 4, 6)
        x = x.reshape(b, c, gh * gw, self.p0 * self.p1 * self.p1)
        return x.reshape(b, c * gh * gw, -1) if self.flat else x


def _patches_to_img(patches, C, G, p0, p1):
    """inverse of ToPatch for [B,C,S,P] (or [B,T,P]) patches -> [B, C*p0, G*p1, G*p1]."""
    b = patches.shape[0]
    x = patches.reshape(b, C, G, G, p0, p1, p1).permute(0, 1, 4, 2, 5, 3, 6)
    return x.reshape(b, C * p0, G * p1, G * p1).contiguous()


class MoveAxis(nn.Module):
    def __init__(self, axes):
        super().__init__()
        self.axes = axes

    def forward(self, x):
        return torch.moveaxis(x, *self.axes)


class Mean(nn.Module):
    def __init__(self, axis):
        super().__init__()
        self.axis = axis

    def forward(self, x):
        return x.mean(axis=self.axis)


class Flatten(nn.Module):
    def __init__(self, start_dim, end_dim):
        super().__init__()
        self.start_dim, self.end_dim = start_dim, end_dim

    def forward(self, x):
        return x.flatten(start_dim=self.start_dim, end_dim=self.end_dim)


class Squeeze(nn.Module):
    def forward(self, x):
        return x.squeeze()


class HeadRearrange(nn.Module):
    """'b h w (p1 p2 nc) -> b (h p1) (w p2) nc' (:485-490)."""

    def __init__(self, p1, nc):
        super().__init__()
        self.p1, self.nc = p1, nc

    def forward(self, x):
        b, h, w, _ = x.shape
        x = x.reshape(b, h, w, self.p1, self.p1, self.nc).permute(0, 1, 3, 2, 4, 5)
        return x.reshape(b, h * self.p1, w * self.p1, self.nc)


class KLayerNorm(nn.LayerNorm):
    """nn.LayerNorm parameters, msst_layernorm_fwd/bwd arithmetic."""

    def forward(self, x):
        return ops.layer_norm(x, self.weight, self.bias, self.eps)


class KLinear(nn.Linear):
    """nn.Linear parameters, msst_linear_* arithmetic (fp32 FFMA GEMM)."""

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


# ---------------------------------------------------------------------------------------------------
# transformer parameter containers (reference module tree => identical state_dict keys)
# ---------------------------------------------------------------------------------------------------
class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner_dim = dim_head * heads
        if heads == 1 and dim_head == dim:
            raise NotImplementedError("maskedsst_b200: heads == 1 with dim_head == dim (no output projection) is not built")
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))


class Transformer(nn.Module):
    """L x { x = attn(LN(x)) + x ; x = ff(LN(x)) + x } (:100-104), executed as one fused stack."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.dim, self.heads, self.dim_head, self.mlp_dim, self.p = dim, heads, dim_head, mlp_dim, dropout
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout)),
            ]))
        self.precision = "fp32"
        self.site_base = ops.SITE_LAYER_BASE

    def layer_params(self):
        out = []
        for attn, ff in self.layers:
            out.append((attn.norm.weight, attn.norm.bias, attn.fn.to_qkv.weight, attn.fn.to_out[0].weight,
                        attn.fn.to_out[0].bias, ff.norm.weight, ff.norm.bias, ff.fn.net[0].weight, ff.fn.net[0].bias,
                        ff.fn.net[3].weight, ff.fn.net[3].bias))
        return out

    def run(self, rows, n_seq, N, inner):
        """rows [n_seq*N, D] in the strided row order described in msst.h."""
        p = self.p if self.training else 0.0
        return ops.transformer_stack(rows, self.layer_params(), n_seq=n_seq, N=N, inner=inner, heads=self.heads,
                                     dim_head=self.dim_head, mlp_dim=self.mlp_dim, drop_p=p,
                                     seed=ops.next_seed() if p > 0 else 0, site_base=self.site_base,
                                     prec=_PREC[self.precision])

    def forward(self, x):   # x [n, N, D] -- every sequence contiguous (reference calling convention)
        n, N, D = x.shape
        x = x if x.is_contiguous() else x.contiguous()
        return self.run(x.reshape(n * N, D), n, N, 1).reshape(n, N, D)


# ---------------------------------------------------------------------------------------------------
# patch embeddings
# ---------------------------------------------------------------------------------------------------
class BlockwisePatchEmbedding(nn.Module):
    def __init__(self, num_channels, transformer_dim, patch_depth, patch_height, patch_width):
        super().__init__()
        assert num_channels % patch_depth == 0, \
            f"Number of channels {num_channels=} not divisible by patch_depth {patch_depth=}"
        assert patch_height == patch_width, "maskedsst_b200: square spatial patches only"
        self.patch_depth, self.patch_height, self.patch_width = patch_depth, patch_height, patch_width
        self.transformer_dim = transformer_dim
        self.patch_dim = reduce(mul, [patch_depth, patch_height, patch_width])
        self.num_blocks = num_channels // patch_depth
        self.pre_norm = nn.LayerNorm(self.patch_dim)
        self.post_norm = nn.LayerNorm(self.transformer_dim)
        self.to_patch = ToPatch(patch_depth, patch_height, flat=False)
        self.blockwise_embed = nn.ModuleList([nn.Linear(self.patch_dim, self.transformer_dim) for _ in range(self.num_blocks)])

    def kernel_params(self):
        W = torch.stack([l.weight for l in self.blockwise_embed])
        b = torch.stack([l.bias for l in self.blockwise_embed])
        return self.pre_norm.weight, self.pre_norm.bias, W, b, self.post_norm.weight, self.post_norm.bias

    def _embed_img(self, img, pos=None, mask_token=None, mask=None, drop_p=0.0, want_ln=False):
        G = img.shape[-1] // self.patch_height
        D = self.transformer_dim
        if pos is None:
            pos = torch.zeros(self.num_blocks * G * G, D, device=img.device, dtype=torch.float32)
        geom = (self.num_blocks, G, self.patch_depth, self.patch_height, D)
        return ops.patch_embed(img, *self.kernel_params(), pos, mask_token, mask, geom=geom, drop_p=drop_p,
                               seed=ops.next_seed() if drop_p > 0 else 0, want_ln=want_ln)

    def embed(self, patches):
        """patches [B, C, S, P] (output of to_patch) -> tokens [B, T, D] (:210-222)."""
        G = int(round(patches.shape[2] ** 0.5))
        return self._embed_img(_patches_to_img(patches, self.num_blocks, G, self.patch_depth, self.patch_height))

    def forward(self, x):
        return self._embed_img(x)


class _LNPatches(nn.Module):
    """nn.Sequential(Rearrange, LayerNorm) of the reference PatchEmbed.to_patch (:235-243): index 1 holds the LN."""

    def __init__(self, patch_dim, p0, p1):
        super().__init__()
        self.add_module("0", ToPatch(p0, p1, flat=True))
        self.add_module("1", KLayerNorm(patch_dim))

    def __getitem__(self, i):
        return getattr(self, str(i))

    def forward(self, x):
        return self[1](self[0](x))


class PatchEmbed(nn.Module):
    """Non-blockwise alternative (:232-253): one Linear shared by all spectral blocks."""

    def __init__(self, dim, patch_dim, patch_depth, patch_height, patch_width):
        super().__init__()
        self.patch_depth, self.patch_height, self.dim = patch_depth, patch_height, dim
        self.to_patch = _LNPatches(patch_dim, patch_depth, patch_height)
        self.embed = nn.Sequential(KLinear(patch_dim, dim), KLayerNorm(dim))

    def kernel_params(self):
        ln = self.to_patch[1]
        return ln.weight, ln.bias, self.embed[0].weight[None], self.embed[0].bias[None], self.embed[1].weight, self.embed[1].bias

    def _embed_img(self, img, pos=None, mask_token=None, mask=None, drop_p=0.0, want_ln=False):
        G = img.shape[-1] // self.patch_height
        C = img.shape[1] // self.patch_depth
        if pos is None:
            pos = torch.zeros(C * G * G, self.dim, device=img.device, dtype=torch.float32)
        geom = (C, G, self.patch_depth, self.patch_height, self.dim)
        return ops.patch_embed(img, *self.kernel_params(), pos, mask_token, mask, geom=geom, drop_p=drop_p,
                               seed=ops.next_seed() if drop_p > 0 else 0, want_ln=want_ln)

    def forward(self, x):
        return self._embed_img(x)


# ---------------------------------------------------------------------------------------------------
# the encoder
# ---------------------------------------------------------------------------------------------------
class ViTSpatialSpectral(nn.Module):
    def __init__(self, *, image_size, spatial_patch_size, spectral_patch_size, num_classes, dim, depth, heads, mlp_dim,
                 spectral_pos_embed=True, pool="mean", blockwise_patch_embed=True, channels=3, dim_head=64, dropout=0.0,
                 emb_dropout=0.0, spectral_pos=tuple(range(20)), spectral_only=False, spectral_mlp_head=False,
                 pixelwise=False, pos_embed_len=None, precision="fp32"):
        super().__init__()
        image_height, image_width = pair(image_size)
        image_depth = channels
        self.patch_height, self.patch_width = pair(spatial_patch_size)
        self.patch_depth = spectral_patch_size
        self.image_size = image_size
        self.pixels_per_patch = reduce(mul, [self.patch_depth, self.patch_height, self.patch_width])
        self.spectral_pos = np.array(spectral_pos)
        self.spectral_pos_embed = spectral_pos_embed
        self.blockwise_patch_embed = blockwise_patch_embed
        self.spectral_only = spectral_only
        self.spectral_mlp_head = spectral_mlp_head
        self.pixelwise = pixelwise
        self.num_classes = num_classes

        assert (image_height % self.patch_height == 0 and image_width % self.patch_width == 0
                and image_depth % self.patch_depth == 0), \
            f"Image dimensions must be divisible by the patch size. {image_height=}, {self.patch_height=}, " \
            f"{image_width=}, {self.patch_width=}, {image_depth=}, {self.patch_depth=}"
        assert image_height == image_width and self.patch_height == self.patch_width, \
            "maskedsst_b200: square images / patches only (the reference assumes it too, :324-325)"

        self.num_spatial_patches_sqrt = image_height // self.patch_height
        self.num_spatial_patches = self.num_spatial_patches_sqrt ** 2
        self.num_spectral_patches = image_depth // self.patch_depth
        self.num_patches = self.num_spatial_patches * self.num_spectral_patches
        assert pool in {"mean"}, "pool type must be either cls (cls token) or mean (mean pooling)"

        # construction order == the reference's, so torch.manual_seed(s) yields the same initial weights
        if self.blockwise_patch_embed:
            self.to_patch_embedding = BlockwisePatchEmbedding(channels, dim, self.patch_depth, self.patch_height, self.patch_width)
        else:
            self.to_patch_embedding = PatchEmbed(dim, self.pixels_per_patch, self.patch_depth, self.patch_height, self.patch_width)

        if self.spectral_pos_embed:
            channel_embed_dim = dim // 3
            pos_embed_dim = dim - channel_embed_dim
            self.pos_embed = nn.Parameter(torch.zeros(1, self.num_spatial_patches, pos_embed_dim))
            p_embed = get_2d_sincos_pos_embed(pos_embed_dim, self.num_spatial_patches_sqrt, cls_token=False)
            self.pos_embed.data.copy_(torch.from_numpy(p_embed).float().unsqueeze(0))
            assert len(self.spectral_pos) == self.num_spectral_patches, \
                f"{self.spectral_pos.shape=}, {self.num_spectral_patches=}"
            self.channel_embed = nn.Parameter(torch.zeros(1, self.num_spectral_patches, channel_embed_dim))
            chan_embed = get_1d_sincos_pos_embed_from_grid(channel_embed_dim, self.spectral_pos)
            self.channel_embed.data.copy_(torch.from_numpy(chan_embed).float().unsqueeze(0))
        else:
            n_rows = pos_embed_len if pos_embed_len is not None else self.num_patches + 1
            self.pos_embedding = nn.Parameter(torch.randn(1, n_rows, dim))

        self.dropout = nn.Dropout(emb_dropout)

        c, s = self.num_spectral_patches, self.num_spatial_patches
        if self.spectral_only:
            self.spatial_spectral_transformer = nn.Sequential(
                Retile("to_spectral", c, s), Transformer(dim, depth, heads, dim_head, mlp_dim, dropout),
                Retile("from_spectral", c, s))
        else:
            self.spatial_spectral_transformer = nn.Sequential(
                Retile("to_spatial", c, s), Transformer(dim, depth, heads, dim_head, mlp_dim, dropout),
                Retile("spatial_to_spectral", c, s), Transformer(dim, depth, heads, dim_head, mlp_dim, dropout),
                Retile("from_spectral", c, s))
            self.spatial_spectral_transformer[3].site_base = ops.SITE_LAYER_BASE + 8 * depth

        self.pool = pool
        self.to_latent = nn.Identity()
        self.dim = dim

        num_out_pixels = self.patch_width * self.patch_height
        if self.spectral_mlp_head:
            self.mlp_head = nn.Sequential(
                KLayerNorm(dim * c), KLinear(dim * c, num_classes * num_out_pixels),
                HeadRearrange(self.patch_height, num_classes), MoveAxis((-1, 1)))
        elif self.pixelwise:
            self.mlp_head = nn.Sequential(
                KLayerNorm(dim), Flatten(start_dim=1, end_dim=-1), KLinear(dim * s, num_classes),
                _PixelwiseRearrange(self.patch_height, num_classes), MoveAxis((-1, 1)), Squeeze())
        else:
            self.mlp_head = nn.Sequential(
                nn.LayerNorm(dim), nn.Linear(dim, num_classes * num_out_pixels),
                HeadRearrange(self.patch_height, num_classes), MoveAxis((-1, 1)))
        self.precision = precision

    # ---- precision mode -----------------------------------------------------------------------------
    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, value):
        if value not in _PREC:
            raise ValueError(f"precision must be one of {sorted(_PREC)}")
        self._precision = value
        for m in self.spatial_spectral_transformer:
            if isinstance(m, Transformer):
                m.precision = value

    # ---- positional rows ------------------------------------------------------------------------------
    def get_pos_embeddings(self):
        """[1, T, D]: cat(spatial table broadcast over blocks, spectral table broadcast over positions) (:501-516)."""
        c, s = self.num_spectral_patches, self.num_spatial_patches
        pe = self.pos_embed.unsqueeze(1).expand(-1, c, -1, -1)
        ce = self.channel_embed.unsqueeze(2).expand(-1, -1, s, -1)
        return torch.cat((pe, ce), dim=-1).reshape(1, c * s, self.dim)

    def _pos_rows(self):
        if self.spectral_pos_embed:
            return self.get_pos_embeddings()[0]
        return self.pos_embedding[0, : self.num_patches]

    # ---- forward --------------------------------------------------------------------------------------
    def transformer_forward(self, x):
        """x [B, T, D] (t = c*S + s) -> [B, T, D]; spatial stack then spectral stack, no re-tiling copies."""
        B, T, D = x.shape
        c, s = self.num_spectral_patches, self.num_spatial_patches
        assert T == c * s
        rows = (x if x.is_contiguous() else x.contiguous()).reshape(B * T, D)
        seq = self.spatial_spectral_transformer
        if self.spectral_only:
            rows = seq[1].run(rows, B * s, c, s)
        else:
            rows = seq[1].run(rows, B * c, s, 1)
            rows = seq[3].run(rows, B * s, c, s)
        return rows.reshape(B, T, D)

    def forward_features(self, img):
        p = self.dropout.p if self.training else 0.0
        x = self.to_patch_embedding._embed_img(img, pos=self._pos_rows(), drop_p=p)
        return self.transformer_forward(x)

    def forward(self, img):
        x = self.forward_features(img)
        B = x.shape[0]
        c, g = self.num_spectral_patches, self.num_spatial_patches_sqrt
        if self.spectral_mlp_head:
            x = x.reshape(B, c, g, g, self.dim).permute(0, 2, 3, 1, 4).reshape(B, g, g, c * self.dim)
            return self.mlp_head(self.to_latent(x))
        if self.pixelwise:
            x = x.reshape(B, c, g, g, self.dim).mean(dim=1)   # second-tier variant: pooling via torch
            return self.mlp_head(self.to_latent(x))
        ln, lin = self.mlp_head[0], self.mlp_head[1]
        return ops.head(x, ln.weight, ln.bias, lin.weight, lin.bias,
                        geom=(B, c, g, self.patch_height, self.dim, self.num_classes))


# ---------------------------------------------------------------------------------------------------
# legacy encoder (reference :600-764) -- kept so V1 checkpoints / `intermediate_losses` runs have a drop-in
# ---------------------------------------------------------------------------------------------------
class AvgPoolMerge(nn.Module):
    """(x1 + x2) / 2 (:567-576).  Constructed by V1 but never called by its forward (:735-744)."""

    def forward(self, x1, x2):
        return torch.stack((x1, x2), dim=1).mean(axis=1)


class LinearMerge(nn.Module):
    """Linear(2 dim -> dim) over cat(x1, x2) (:579-588); parameter container for checkpoint compatibility."""

    def __init__(self, dim):
        super().__init__()
        self.fc = KLinear(2 * dim, dim)

    def forward(self, x1, x2):
        return self.fc(torch.concat((x1, x2), dim=2))


class ViTSpatialSpectral_V1(nn.Module):
    """ViTSpatialSpectral_V1 (:600-764): one shared Linear patch embedding held in an nn.Sequential
    (keys to_patch_embedding.{1,2,3}.*), learned positional table, the same spatial -> spectral stacks, default head.
    NB `num_spatial_patches` is the grid SIDE here (:629), not its square as in ViTSpatialSpectral."""

    def __init__(self, *, image_size, spatial_patch_size, spectral_patch_size, num_classes, dim, depth, heads, mlp_dim,
                 pool="mean", channels=3, dim_head=64, dropout=0.0, emb_dropout=0.0, merge="avgpool", precision="fp32"):
        super().__init__()
        image_height, image_width = pair(image_size)
        self.patch_height, self.patch_width = pair(spatial_patch_size)
        self.patch_depth = spectral_patch_size
        self.image_size = image_size
        assert (image_height % self.patch_height == 0 and image_width % self.patch_width == 0
                and channels % self.patch_depth == 0), "Image dimensions must be divisible by the patch size."
        assert image_height == image_width and self.patch_height == self.patch_width, \
            "maskedsst_b200: square images / patches only"
        self.num_spatial_patches = image_height // self.patch_height
        self.num_spectral_patches = channels // self.patch_depth
        self.num_patches = self.num_spatial_patches ** 2 * self.num_spectral_patches
        patch_dim = self.patch_depth * self.patch_height * self.patch_width
        self.pixels_per_patch = patch_dim
        assert pool in {"mean"}, "pool type must be either cls (cls token) or mean (mean pooling)"
        self.to_patch_embedding = nn.Sequential(ToPatch(self.patch_depth, self.patch_height, flat=True),
                                                KLayerNorm(patch_dim), KLinear(patch_dim, dim), KLayerNorm(dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, self.num_patches + 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        c, s = self.num_spectral_patches, self.num_spatial_patches ** 2
        self.spatial_spectral_transformer = nn.Sequential(
            Retile("to_spatial", c, s), Transformer(dim, depth, heads, dim_head, mlp_dim, dropout),
            Retile("spatial_to_spectral", c, s), Transformer(dim, depth, heads, dim_head, mlp_dim, dropout),
            Retile("from_spectral", c, s))
        self.spatial_spectral_transformer[3].site_base = ops.SITE_LAYER_BASE + 8 * depth
        if merge == "avgpool":
            self.merge = AvgPoolMerge()
        elif merge == "linear":
            self.merge = LinearMerge(dim)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.dim = dim
        self.num_classes = num_classes
        self.mlp_head = nn.Sequential(
            nn.LayerNorm(dim), nn.Linear(dim, num_classes * self.patch_width * self.patch_height),
            HeadRearrange(self.patch_height, num_classes), MoveAxis((-1, 1)))
        self.precision = precision

    precision = ViTSpatialSpectral.precision

    def _embed_img(self, img, pos, mask_token=None, mask=None, drop_p=0.0):
        """to_patch_embedding (+ pos, + mask-token select, + emb-dropout) as one msst_patch_embed launch."""
        seq = self.to_patch_embedding
        G, C_ = self.num_spatial_patches, self.num_spectral_patches
        geom = (C_, G, self.patch_depth, self.patch_height, self.dim)
        return ops.patch_embed(img, seq[1].weight, seq[1].bias, seq[2].weight[None], seq[2].bias[None], seq[3].weight,
                               seq[3].bias, pos, mask_token, mask, geom=geom, drop_p=drop_p,
                               seed=ops.next_seed() if drop_p > 0 else 0, want_ln=False)

    def transformer_forward(self, x):
        """-> (x, x, x): the reference returns the final representation three times (:735-744)."""
        B, T, D = x.shape
        c, s = self.num_spectral_patches, self.num_spatial_patches ** 2
        assert T == c * s
        rows = (x if x.is_contiguous() else x.contiguous()).reshape(B * T, D)
        seq = self.spatial_spectral_transformer
        rows = seq[3].run(seq[1].run(rows, B * c, s, 1), B * s, c, s)
        x = rows.reshape(B, T, D)
        return x, x, x

    def forward_features(self, img):
        p = self.dropout.p if self.training else 0.0
        x = self._embed_img(img, self.pos_embedding[0, : self.num_patches], drop_p=p)
        return self.transformer_forward(x)

    def forward(self, img):
        x, _, _ = self.forward_features(img)
        ln, lin = self.mlp_head[0], self.mlp_head[1]
        return ops.head(x, ln.weight, ln.bias, lin.weight, lin.bias,
                        geom=(x.shape[0], self.num_spectral_patches, self.num_spatial_patches, self.patch_height, self.dim,
                              self.num_classes))


class _PixelwiseRearrange(nn.Module):
    """'b (p1 p2 nc) -> b p1 p2 nc' (:471-476)."""

    def __init__(self, p1, nc):
        super().__init__()
        self.p1, self.nc = p1, nc

    def forward(self, x):
        return x.reshape(x.shape[0], self.p1, self.p1, self.nc)


def get_pos_for_spectral_embedding(spectral_patch_depth, wavelengths, reference_wavelengths):
    """For each spectral block of `wavelengths`, the index of the closest block (by mean wavelength) of
    `reference_wavelengths` (reference :767-801).  Host-side numpy."""
    def block_means(w):
        w = np.array(w)
        total = len(w)
        if total % spectral_patch_depth != 0:
            total += spectral_patch_depth - total % spectral_patch_depth
        return [w[i: i + spectral_patch_depth].mean() for i in range(0, total, spectral_patch_depth)]

    ref = np.array(block_means(reference_wavelengths))
    return [np.argmin(np.abs(ref - m)) for m in block_means(wavelengths)]
