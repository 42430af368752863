"""GPU: the fused FeedForward kernels (csrc/mlp_block_tc.cu, forward and backward) through the C ABI against an fp64 torch restatement of
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


@pytest.mark.parametrize("R,D", [(128, 96), (1000, 96), (327, 64), (5, 32), (128 * 300 + 17, 96), (128 * 149, 96)])
def test_mlp_block_bwd_vs_fp64(R, D):
    """msst_mlp_block_bwd (one kernel: du on the SM, dW1 / dW2 / db1 accumulated in TMEM, dh by TMA store) against fp64 autograd of
    Linear -> GELU(erf) -> Linear fed with the same bf16 operands; the gradients ACCUMULATE into non-zero buffers."""
    torch.manual_seed(1)
    M = 64
    h2 = torch.randn(R, D).bfloat16()
    dyb = torch.randn(R, D).bfloat16()
    w1 = (torch.randn(M, D) * D ** -0.5).bfloat16()
    w2 = (torch.randn(D, M) * M ** -0.5).bfloat16()
    b1 = torch.randn(M) * 0.1
    u = (h2.double() @ w1.double().t() + b1.double()).bfloat16()
    g = torch.nn.functional.gelu(u.double()).bfloat16()
    # fp64 reference on the saved (bf16-rounded) u / g, du rounded to bf16 before the two contractions that consume it (as the kernel does)
    ud = u.double()
    gp = 0.5 * (1 + torch.erf(ud / 2 ** 0.5)) + ud * torch.exp(-0.5 * ud * ud) / (2 * torch.pi) ** 0.5
    du = (dyb.double() @ w2.double()) * gp
    dub = du.bfloat16().double()
    ref = {"dw2": dyb.double().t() @ g.double(), "dw1": dub.t() @ h2.double(), "db1": du.sum(0), "dh": dub @ w1.double()}
    dev = lambda t: t.to(DEV)
    init = {"dw1": torch.randn(M, D), "dw2": torch.randn(D, M), "db1": torch.randn(M)}
    out = {k: dev(v.clone()) for k, v in init.items()}
    dh = torch.full((R, D), float("nan"), device=DEV)
    args = [dev(t) for t in (dyb, u, g, h2, w2.t().contiguous(), w1.t().contiguous())]
    check(_lib.lib().msst_mlp_block_bwd(*[a.data_ptr() for a in args], out["dw1"].data_ptr(), out["dw2"].data_ptr(), out["db1"].data_ptr(), dh.data_ptr(),
                                        R, D, M, 0.0, 0, 18, None, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.isfinite(dh).all()
    assert rel_l2(dh, ref["dh"]) < 2e-3, rel_l2(dh, ref["dh"])
    for k in ("dw1", "dw2", "db1"):
        got = out[k].cpu().double() - init[k].double()
        assert rel_l2(got, ref[k]) < 2e-3, (k, rel_l2(got, ref[k]))


def test_mlp_block_bwd_dropout_matches_unfused_kernels():
    """dropout on: the fused backward regenerates the hidden-dropout mask of gemm_tn<4> (same site / quad indexing): du-dependent outputs agree
    with the GEMM path run through msst_linear_bwd_data / msst_linear_bwd_weight."""
    torch.manual_seed(2)
    R, D, M, p, seed, site = 128 * 37 + 5, 96, 64, 0.25, 1234, 18
    h2 = torch.randn(R, D, device=DEV).bfloat16(); dyb = torch.randn(R, D, device=DEV).bfloat16()
    u = torch.randn(R, M, device=DEV).bfloat16(); g = torch.randn(R, M, device=DEV).bfloat16()
    w1 = (torch.randn(M, D, device=DEV) * D ** -0.5).bfloat16(); w2 = (torch.randn(D, M, device=DEV) * M ** -0.5).bfloat16()
    w1t, w2t = w1.t().contiguous(), w2.t().contiguous()
    lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    dw1, dw2, db1 = torch.zeros(M, D, device=DEV), torch.zeros(D, M, device=DEV), torch.zeros(M, device=DEV)
    dh = torch.empty(R, D, device=DEV)
    check(lib.msst_mlp_block_bwd(dyb.data_ptr(), u.data_ptr(), g.data_ptr(), h2.data_ptr(), w2t.data_ptr(), w1t.data_ptr(), dw1.data_ptr(), dw2.data_ptr(),
                                 db1.data_ptr(), dh.data_ptr(), R, D, M, p, seed, site, None, st))
    # unfused: du = (dyb . W2) * gelu'(u) * drop  (bf16)  ->  dW1 = du^T h2, dh = du . W1
    du = torch.empty(R, M, device=DEV, dtype=torch.bfloat16)
    dims = _lib.LinearDims(R, D, M, 0, p, seed, site, _lib.PREC_BF16, None, 0)          # dx[M_rows, K] = dy[M_rows, N] . W^T-copy[K, N]^T
    check(lib.msst_linear_bwd_data(C.byref(dims), dyb.data_ptr(), w2t.data_ptr(), u.data_ptr(), None, du.data_ptr(), st))
    dw1_ref = torch.zeros(M, D, device=DEV)
    dims_w = _lib.LinearDims(R, M, D, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 0)
    check(lib.msst_linear_bwd_weight(C.byref(dims_w), du.data_ptr(), h2.data_ptr(), dw1_ref.data_ptr(), None, st))
    torch.cuda.synchronize()
    keep = (du.float() != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.01
    assert rel_l2(dw1, dw1_ref) < 1e-3, rel_l2(dw1, dw1_ref)
    assert rel_l2(db1, du.float().sum(0)) < 5e-3
    assert rel_l2(dh, du.double() @ w1.double()) < 1e-3
    assert rel_l2(dw2, dyb.double().t() @ g.double()) < 1e-3


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
