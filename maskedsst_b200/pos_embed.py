"""Sin-cos positional tables (init values of pos_embed / channel_embed).
Same values as the reference's src/pos_embed.py:16-63 (MAE-style), computed in float64 without the removed
`np.float` alias."""
import numpy as np


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    """[M] positions -> [M, embed_dim] = [sin(pos*w_k) | cos(pos*w_k)], w_k = 10000^(-2k/embed_dim)."""
    assert embed_dim % 2 == 0
    half = embed_dim // 2
    omega = 1.0 / 10000 ** (np.arange(half, dtype=np.float64) / half)
    out = np.einsum("m,d->md", np.asarray(pos, dtype=np.float64).reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_sincos_pos_embed_from_grid(embed_dim, grid):
    assert embed_dim % 2 == 0
    a = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0])
    b = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])
    return np.concatenate([a, b], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    """[grid_size^2 (+1), embed_dim]; 'w goes first' like the reference (first half encodes the column)."""
    coords = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(coords, coords), axis=0).reshape(2, 1, grid_size, grid_size)
    emb = get_2d_sincos_pos_embed_from_grid(embed_dim, grid)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb
