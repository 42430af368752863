"""Alias of maskedsst_b200.pos_embed under the reference's import path (src/pos_embed.py)."""
from maskedsst_b200.pos_embed import (get_1d_sincos_pos_embed_from_grid, get_2d_sincos_pos_embed,  # noqa: F401
                                      get_2d_sincos_pos_embed_from_grid)
