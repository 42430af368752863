"""GPU: the fused FeedForward kernel (csrc/mlp_block_tc.cu) through the C ABI against an fp64 torch restatement of
Linear -> GELU(erf) -> Linear -> + residual -> LayerNorm (reference FeedForward.forward, src/vit_spatial_spectral.py:35-44,
Transformer.forward :103, PreNorm :25-29), and -- with dropout on -- against the two-GEMM path it replaces (same masks)."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import pytest
import torch

from maskedsst_b200 import _lib
from maskedsst_b200._lib import check
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(h2, xmid, w1, w2, b1, b2, lnw, lnb, with_ln, drop_p=0.0, seed=0):
    R, D = h2.shape
    M = w1.shape[0]
    nan = float("nan")
    u = torch.full((R, M), nan, device=DEV, dtype=torch.bfloat16)
    g = torch.full((R, M), nan, device=DEV, dtype=torch.bfloat16)
    y = torch.full((R, D), nan, device=DEV)
    h1 = torch.full((R, D), nan, device=DEV, dtype=torch.bfloat16) if with_ln else None
    st = torch.full((R, 2), nan, device=DEV) if with_ln else None
    p = lambda t: None if t is None else t.data_ptr()
    check(_lib.lib().msst_mlp_block_fwd(p(h2), p(xmid), p(w1), p(w2), p(b1), p(b2), p(u), p(g), p(y), p(lnw) if with_ln else None,
                                        p(lnb) if with_ln else None, p(h1), p(st), R, D, M, float(drop_p), seed, 18, 19, None,
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return u, g, y, h1, st


@pytest.mark.parametrize("R,D,with_ln", [(128, 96, True), (1000, 96, True), (327, 64, True), (5, 32, False), (128 * 300 + 17, 96, True), (640, 96, False)])
def test_mlp_block_vs_fp64(R, D, with_ln):
    torch.manual_seed(0)
    M = 64
    h2 = torch.randn(R, D).bfloat16()
    xmid = torch.randn(R, D)
    w1 = (torch.randn(M, D) * D ** -0.5).bfloat16()
    w2 = (torch.randn(D, M) * M ** -0.5).bfloat16()
    b1, b2 = torch.randn(M) * 0.1, torch.randn(D) * 0.1
    lnw, lnb = torch.rand(D) + 0.5, torch.randn(D) * 0.1
    u_ref = h2.double() @ w1.double().t() + b1.double()
    g_ref = torch.nn.functional.gelu(u_ref)
    y_ref = xmid.double() + g_ref.bfloat16().double() @ w2.double().t() + b2.double()      # the kernel feeds the bf16-rounded g to the 2nd GEMM
    h1_ref = torch.nn.functional.layer_norm(y_ref, (D,), lnw.double(), lnb.double())
    dev = [t.to(DEV) for t in (h2, xmid, w1, w2, b1, b2, lnw, lnb)]
    u, g, y, h1, st = _run(*dev, with_ln)
    assert rel_l2(u, u_ref) < 4e-3 and rel_l2(g, g_ref) < 4e-3
    assert rel_l2(y, y_ref) < 2e-3
    if with_ln:
        assert rel_l2(h1, h1_ref) < 5e-3
        assert rel_l2(st[:, 0], y_ref.mean(-1)) < 1e-3
        assert rel_l2(st[:, 1], (y_ref.var(-1, unbiased=False) + 1e-5).rsqrt()) < 1e-3


def test_fused_mlp_equals_two_gemm_path_with_dropout():
    """Whole transformer stack forward + backward, dropout 0.1: MSST_MLP_FUSED=1 (default) vs 0 -- identical dropout masks, so the
    outputs and every gradient agree to rounding (the unfused path rounds g / h1' at the same points)."""
    code = r'''
import sys, torch
sys.path.insert(0, ".")
import maskedsst_b200 as M
from maskedsst_b200 import ops
torch.manual_seed(0)
enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=3, heads=8, mlp_dim=64,
                           channels=50, spectral_pos_embed=False, dropout=0.1, precision="bf16").cuda().train()
tf = enc.spatial_spectral_transformer[1]
rows = torch.randn(7 * 5 * 64, 96, device="cuda", requires_grad=True)
out = ops.transformer_stack(rows, tf.layer_params(), n_seq=35, N=64, inner=1, heads=8, dim_head=64, mlp_dim=64, drop_p=0.1, seed=99, prec=1)
w = torch.randn_like(out)
(out * w).sum().backward()
torch.save({"out": out.detach().cpu(), "dx": rows.grad.cpu(), **{f"g{i}": p.grad.cpu() for i, p in enumerate(tf.parameters())}}, sys.argv[1])
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for flag in ("1", "0"):
            path = os.path.join(td, f"m{flag}.pt")
            r = subprocess.run([sys.executable, "-c", code, path], cwd=root, env=dict(os.environ, MSST_MLP_FUSED=flag), capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout + r.stderr
            res[flag] = torch.load(path)
    for k, v in res["0"].items():
        assert torch.isfinite(res["1"][k]).all()
        assert rel_l2(res["1"][k], v) < 3e-3, (k, rel_l2(res["1"][k], v))
