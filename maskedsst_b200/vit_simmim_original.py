"""B200-native SimMIMSpatialSpectral: same surface / state_dict as the reference's src/vit_simmim_original.py
(:139-340), with mask-token substitution fused into the patch-embedding kernel and gather + to_pixels + L1 fused
into msst_simmim_decode_l1_*.  Mask generation stays on the host and is bit-compatible with the reference's
MaskGenerator (numpy global RNG, :343-416), including the mask/index mismatch quirk."""
import numpy as np
import torch
from torch import nn

from . import ops
from .vit_spatial_spectral import ViTSpatialSpectral, ViTSpatialSpectral_V1


class BlockwiseToPixels(nn.Module):
    """One Linear(dim -> pixels_per_patch) per spectral block (reference :9-40); parameter container --
    the arithmetic runs inside msst_simmim_decode_l1_*."""

    def __init__(self, dim, num_spectral_blocks, pixels_per_patch, precision):
        super().__init__()
        self.pixels_per_patch = pixels_per_patch
        self.layers = nn.ModuleList([nn.Linear(dim, pixels_per_patch) for _ in range(num_spectral_blocks)])
        if precision == "16-mixed":
            self.dtype = torch.float16
        elif precision == "32-true":
            self.dtype = torch.float32

    def kernel_params(self):
        return torch.stack([l.weight for l in self.layers]), torch.stack([l.bias for l in self.layers])


class MaskGenerator:
    """Host mask generator, draws from numpy's global RNG in the same order as the reference (:343-416)."""

    def __init__(self, input_size=16, mask_patch_size=4, model_patch_size=1, mask_ratio=0.6):
        self.input_size, self.mask_patch_size = input_size, mask_patch_size
        self.model_patch_size, self.mask_ratio = model_patch_size, mask_ratio
        assert self.input_size % self.mask_patch_size == 0
        assert self.mask_patch_size % self.model_patch_size == 0
        self.rand_size = self.input_size // self.mask_patch_size
        self.scale = self.mask_patch_size // self.model_patch_size
        self.token_count = self.rand_size ** 2
        self.mask_count = int(np.ceil(self.token_count * self.mask_ratio))

    def __call__(self):
        chosen = np.random.permutation(self.token_count)[: self.mask_count]
        cells = np.zeros(self.token_count, dtype=int)
        cells[chosen] = 1
        cells = cells.reshape(self.rand_size, self.rand_size)
        return cells.repeat(self.scale, axis=0).repeat(self.scale, axis=1)

    def bool_mask_to_indices(self, masked_bool_mask, batch, num_masked, device):
        """Column ids of the set bits in row-major order, cut into consecutive runs of num_masked -- NOT per row
        (reference :372-382; rows >= 1 therefore take ids that belong to neighbouring samples, quirk C3)."""
        m = masked_bool_mask.cpu().numpy() if torch.is_tensor(masked_bool_mask) else np.asarray(masked_bool_mask)
        cols = np.nonzero(m)[1]
        if cols.shape[0] < batch * num_masked:
            raise RuntimeError("mask has fewer set positions than batch * num_masked")
        idx = cols[: batch * num_masked].reshape(batch, num_masked).astype(np.int64)
        return torch.from_numpy(idx).to(device)

    def _finish(self, cells, batch_size, num_masked, device):
        flat = torch.from_numpy(np.ascontiguousarray(cells.reshape(batch_size, -1)).astype(bool))
        return flat.to(device), self.bool_mask_to_indices(flat, batch_size, num_masked, device)

    def _draw(self, n):
        """n masks [n, size, size] (bool).  Same RNG call sequence as n calls of __call__ (one np.random.permutation each,
        so np.random.seed reproduces the reference's draws); everything after the draw is vectorised."""
        chosen = np.stack([np.random.permutation(self.token_count)[: self.mask_count] for _ in range(n)])
        cells = np.zeros((n, self.token_count), dtype=bool)
        cells[np.arange(n)[:, None], chosen] = True
        cells = cells.reshape(n, self.rand_size, self.rand_size)
        return cells.repeat(self.scale, axis=1).repeat(self.scale, axis=2)

    def get_batch(self, batch_size, channel_tokens, num_masked, device):
        cells = self._draw(batch_size * channel_tokens)
        return self._finish(cells.reshape(batch_size, channel_tokens, *cells.shape[1:]), batch_size, num_masked, device)

    def get_batch_tube_masked(self, batch_size, channel_tokens, num_masked, device):
        cells = np.broadcast_to(self._draw(batch_size)[:, None], (batch_size, channel_tokens, self.input_size // self.model_patch_size,
                                                                 self.input_size // self.model_patch_size))
        return self._finish(cells, batch_size, num_masked, device)


class SimMIMSpatialSpectral(nn.Module):
    def __init__(self, *, encoder, masking_ratio=0.5, mask_patch_size=1, tube_masking=False, intermediate_losses=False,
                 to_pixels_per_spectral_block=False, precision="32-true"):
        super().__init__()
        assert masking_ratio > 0 and masking_ratio < 1, "masking ratio must be kept between 0 and 1"
        self.v1 = isinstance(encoder, ViTSpatialSpectral_V1)
        if not (self.v1 or isinstance(encoder, ViTSpatialSpectral)):
            raise NotImplementedError("maskedsst_b200: SimMIMSpatialSpectral supports ViTSpatialSpectral(_V1) encoders")
        if intermediate_losses and not self.v1:
            # the reference reads `encoded_spatial` / `encoded_spectral`, which only the V1 branch assigns (:297-313)
            raise NotImplementedError("intermediate_losses needs a ViTSpatialSpectral_V1 encoder (the reference raises "
                                      "UnboundLocalError in forward for any other encoder)")
        if to_pixels_per_spectral_block and self.v1:
            # :319-327 builds the block-id table with V1's num_spatial_patches (= grid side, not its square), so the
            # reference's gather runs out of bounds for every index >= C * side
            raise NotImplementedError("to_pixels_per_spectral_block is not usable with ViTSpatialSpectral_V1 (the reference "
                                      "raises IndexError in forward)")
        self.masking_ratio = masking_ratio
        self.mask_patch_size = mask_patch_size
        self.intermediate_losses = intermediate_losses
        self.to_pixels_per_spectral_block = to_pixels_per_spectral_block
        self.tube_masking = tube_masking
        if self.mask_patch_size != 1:
            self.mask_generator = MaskGenerator(input_size=encoder.image_size, mask_patch_size=mask_patch_size,
                                                model_patch_size=encoder.patch_height, mask_ratio=self.masking_ratio)
        # "host": the reference's numpy generator, bit-compatible given np.random.seed (default);
        # "device": same distribution and the same mask/index semantics (incl. quirk C3), drawn with torch on the GPU --
        # no host work and no synchronisation per step.
        self.mask_backend = "host"
        self.encoder = encoder
        if self.v1:   # reference :172-178 -- the slice is a new Sequential sharing the LN/Linear/LN modules (alias keys)
            encoder_dim = encoder.pos_embedding.shape[-1]
            self.to_patch, self.patch_to_emb = encoder.to_patch_embedding[0], encoder.to_patch_embedding[1:]
            self.pixel_values_per_patch = self.patch_to_emb[1].weight.shape[-1]
        else:
            encoder_dim = encoder.dim
            # reference :181-182 -- with PatchEmbed these are sub-modules and add alias keys to the state_dict
            self.to_patch = encoder.to_patch_embedding.to_patch
            self.patch_to_emb = encoder.to_patch_embedding.embed
            self.pixel_values_per_patch = encoder.pixels_per_patch
        self.mask_token = nn.Parameter(torch.randn(encoder_dim))
        if self.to_pixels_per_spectral_block:
            self.to_pixels = BlockwiseToPixels(encoder_dim, encoder.num_spectral_patches, self.pixel_values_per_patch,
                                               precision=precision)
        else:
            self.to_pixels = nn.Linear(encoder_dim, self.pixel_values_per_patch)

    # ---- masks --------------------------------------------------------------------------------------------
    def draw_masks(self, batch, device):
        """(bool mask [B,T], indices [B,nm]) exactly as the reference draws them (:252-282)."""
        enc = self.encoder
        num_patches = enc.num_patches
        num_masked = int(self.masking_ratio * num_patches)
        if self.mask_patch_size == 1:
            idx = torch.rand(batch, num_patches, device=device).topk(k=num_masked, dim=-1).indices
            mask = torch.zeros((batch, num_patches), device=device).scatter_(-1, idx, 1).bool()
            return mask, idx
        if self.mask_backend == "device":
            return self._draw_masks_device(batch, num_masked, device)
        fn = self.mask_generator.get_batch_tube_masked if self.tube_masking else self.mask_generator.get_batch
        return fn(batch_size=batch, channel_tokens=enc.num_spectral_patches, num_masked=num_masked, device=device)

    def _draw_masks_device(self, batch, num_masked, device):
        """MaskGenerator.get_batch / get_batch_tube_masked (+ bool_mask_to_indices) on the device in ONE kernel launch
        (csrc/maskgen.cu, msst_draw_masks): a uniformly random subset of mask_count cells per draw, upsampled by `scale`; the
        index list is the row-major list of set positions cut into consecutive runs of num_masked (the reference's slicing,
        quirk C3).  On a CPU device (host-logic tests only) the same semantics run as torch ops."""
        if torch.device(device).type != "cuda":
            return self._draw_masks_torch(batch, num_masked, device)
        import ctypes as C
        from . import _lib
        g, enc = self.mask_generator, self.encoder
        T = enc.num_patches
        if batch * g.mask_count * g.scale * g.scale * enc.num_spectral_patches < batch * num_masked:
            raise RuntimeError("mask has fewer set positions than batch * num_masked")
        mask = torch.empty(batch, T, dtype=torch.uint8, device=device)
        idx = torch.empty(batch, num_masked, dtype=torch.int64, device=device)
        dims = _lib.MaskGenDims(batch, enc.num_spectral_patches, g.rand_size, g.scale, g.mask_count, num_masked, int(bool(self.tube_masking)),
                                ops.next_seed(), ops._seed_dev())
        _lib.check(_lib.lib().msst_draw_masks(C.byref(dims), mask.data_ptr(), idx.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return mask.view(torch.bool), idx

    def _draw_masks_torch(self, batch, num_masked, device):
        g, enc = self.mask_generator, self.encoder
        C, T = enc.num_spectral_patches, enc.num_patches
        n = batch if self.tube_masking else batch * C
        chosen = torch.rand(n, g.token_count, device=device).argsort(dim=1)[:, : g.mask_count]
        cells = torch.zeros(n, g.token_count, dtype=torch.bool, device=device).scatter_(1, chosen, True)
        cells = cells.view(n, g.rand_size, g.rand_size).repeat_interleave(g.scale, 1).repeat_interleave(g.scale, 2)
        size = g.rand_size * g.scale
        cells = cells[:, None].expand(batch, C, size, size) if self.tube_masking else cells.view(batch, C, size, size)
        mask = cells.reshape(batch, T)
        per_row = g.mask_count * g.scale * g.scale * C                     # set positions per sample (static)
        ar = torch.arange(T, device=device)
        cols = torch.where(mask, ar, ar + T).sort(dim=1).values[:, :per_row]   # ascending set positions of every row
        if batch * num_masked > batch * per_row:
            raise RuntimeError("mask has fewer set positions than batch * num_masked")
        idx = cols.reshape(-1)[: batch * num_masked].view(batch, num_masked)
        return mask, idx

    # ---- forward ------------------------------------------------------------------------------------------
    def forward(self, img, masks=None):
        """img [B, channels, H, W] -> scalar loss.  `masks` = optional externally supplied
        (masked_bool_mask [B,T], masked_indices [B,nm]); by default they are drawn like the reference."""
        enc = self.encoder
        B = img.shape[0]
        mask, idx = masks if masks is not None else self.draw_masks(B, img.device)
        if self.v1:
            return self._forward_v1(img, mask, idx)
        blockwise = enc.blockwise_patch_embed
        # tokens = where(mask, mask_token + pos, embed(patches) + pos); no emb-dropout on this path (C5)
        out = enc.to_patch_embedding._embed_img(img, pos=enc._pos_rows(), mask_token=self.mask_token, mask=mask,
                                                drop_p=0.0, want_ln=not blockwise)
        tokens, pln = (out, None) if blockwise else out
        encoded = enc.transformer_forward(tokens)
        if self.to_pixels_per_spectral_block:
            W, b = self.to_pixels.kernel_params()
        else:
            W, b = self.to_pixels.weight[None], self.to_pixels.bias[None]
        geom = (B, enc.num_spectral_patches, enc.num_spatial_patches_sqrt, enc.patch_depth, enc.patch_height, enc.dim,
                idx.shape[1])
        # target: raw pixels (blockwise embedding) or the LayerNormed patches (PatchEmbed), C6
        return ops.simmim_decode_l1(encoded, idx, img if blockwise else None, None if blockwise else pln, W, b, geom=geom)

    def _forward_v1(self, img, mask, idx):
        """ViTSpatialSpectral_V1 branch (:172-178, :232-234, :297-338): positional rows [1, T], raw-pixel target, shared
        decoder; with intermediate_losses the (identical) term is accumulated three times."""
        enc = self.encoder
        B, T = img.shape[0], enc.num_patches
        tokens = enc._embed_img(img, enc.pos_embedding[0, 1: T + 1], mask_token=self.mask_token, mask=mask)
        encoded, _, _ = enc.transformer_forward(tokens)
        geom = (B, enc.num_spectral_patches, enc.num_spatial_patches, enc.patch_depth, enc.patch_height, enc.dim, idx.shape[1])
        loss = ops.simmim_decode_l1(encoded, idx, img, None, self.to_pixels.weight[None], self.to_pixels.bias[None], geom=geom)
        if self.intermediate_losses:
            loss = (loss + loss) + loss
        return loss
