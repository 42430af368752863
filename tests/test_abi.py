"""CPU: libmsst.so loads and exports exactly the symbols include/msst.h declares (no compute calls)."""
import ctypes
import os
import re

from maskedsst_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "msst.h")).read()
    return sorted(set(re.findall(r"MSST_API[^;(]*?\b(msst_\w+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in msst.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with include/msst.h"


def test_version_and_error_channel():
    L = _lib.lib()
    assert L.msst_version() >= 100
    # argument validation happens before any CUDA call: a bad shape returns MSST_ERR_ARG on a CPU-only box too
    d = _lib.AttnDims(4, 8, 1, 8, 48, 0.0, 0, 0, 0, None)   # dim_head 48 unsupported
    assert L.msst_attention_fwd(ctypes.byref(d), None, None, None, None) == 1
    assert b"dim_head" in L.msst_last_error()


def test_no_cpu_fallback():
    import pytest, torch
    import maskedsst_b200 as M
    enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=1,
                               heads=8, mlp_dim=64, channels=50, spectral_pos_embed=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.zeros(1, 50, 8, 8))
