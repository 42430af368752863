"""GPU: the fused projection + attention kernels (csrc/attn_block_tc.cu) called through the C ABI, against an fp64 torch
reference of  to_qkv -> softmax(q k^T dh^-0.5) v  (reference Attention.forward, src/vit_spatial_spectral.py:67-77) and its
autograd; plus bit-agreement of the dropout masks with the unfused tcgen05 kernels."""
import ctypes as C

import pytest
import torch

from maskedsst_b200 import _lib, ops
from maskedsst_b200._lib import check, PREC_BF16
from tests.helpers import rel_l2
from tests.test_gpu_components import ref_attention

pytestmark = pytest.mark.gpu
DEV = "cuda"

GEOMS = [
    (10, 64, 1, 8, 96),       # spatial stack, full tiles
    (7, 64, 1, 8, 96),        # odd number of 64-slot groups: the last tile is half empty
    (128, 5, 64, 8, 96),      # Houston spectral stack: 12 sequences of 5 per group, rows strided by 64
    (128, 20, 64, 8, 96),     # EnMAP spectral stack
    (640, 5, 64, 8, 96),      # more tiles than one wave of a small grid would hold per CTA: multi-item loops
    (6, 22, 2, 4, 64),        # ragged packing, 4 heads, D = 64
    (37, 16, 1, 2, 32),       # contiguous short sequences with a partial last group, D = 32
    (3, 1, 1, 1, 32),         # degenerate
    (2368, 64, 1, 8, 96),     # 1184 tiles = 8 per CTA: steady-state pipeline, tile changes, h double buffer
]


def _call_fwd(h, w, n_seq, N, inner, H, drop_p=0.0, seed=0, site=0):
    R, D = h.shape
    I = H * 64
    dims = _lib.AttnDims(n_seq, N, inner, H, 64, float(drop_p), seed, site, PREC_BF16, None)
    out = torch.full((R, I), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = torch.full((R, H), float("nan"), device=DEV, dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    check(_lib.lib().msst_attn_block_fwd(C.byref(dims), D, h.data_ptr(), w.data_ptr(), out.data_ptr(), lse.data_ptr(), st))
    return out, lse, dims


def _call_bwd(dims, h, w, d_out, lse):
    R, D = h.shape
    I3 = w.shape[0]
    wt = w.t().contiguous()
    dqkv = torch.full((R, I3), float("nan"), device=DEV, dtype=torch.bfloat16)
    dh = torch.full((R, D), float("nan"), device=DEV, dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    check(_lib.lib().msst_attn_block_bwd(C.byref(dims), D, h.data_ptr(), w.data_ptr(), wt.data_ptr(), d_out.data_ptr(), lse.data_ptr(),
                                         dqkv.data_ptr(), dh.data_ptr(), st))
    return dqkv, dh


@pytest.mark.parametrize("n_seq,N,inner,H,D", GEOMS)
def test_attn_block_fwd_bwd_vs_fp64(n_seq, N, inner, H, D):
    torch.manual_seed(0)
    R, I = n_seq * N, H * 64
    h = torch.randn(R, D).bfloat16()
    w = (torch.randn(3 * I, D) * D ** -0.5).bfloat16()
    g_out = torch.randn(R, I).bfloat16()
    # reference: the kernel rounds q, k, v to bf16 before the attention contractions
    hd = h.double().requires_grad_(True)
    qkv = hd @ w.double().t()
    qkv_r = qkv + (qkv.detach().float().bfloat16().double() - qkv.detach())      # straight-through bf16 rounding
    qkv_r.retain_grad()
    want = ref_attention(qkv_r, n_seq, N, inner, H, 64)
    (want * g_out.double()).sum().backward()

    hg, wg = h.to(DEV), w.to(DEV)
    out, lse, dims = _call_fwd(hg, wg, n_seq, N, inner, H)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert rel_l2(out, want) < 8e-3
    # lse against the reference scores
    q, k, _ = qkv_r.detach().split(I, dim=-1)
    dqkv, dh = _call_bwd(dims, hg, wg, g_out.to(DEV), lse)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all() and torch.isfinite(dh).all()
    assert rel_l2(dqkv, qkv_r.grad) < 2e-2
    assert rel_l2(dh, hd.grad) < 2e-2


@pytest.mark.parametrize("n_seq,N,inner,H", [(16, 64, 1, 8), (256, 5, 64, 8), (128, 20, 64, 4)])
def test_attn_block_matches_unfused_kernels_with_dropout(n_seq, N, inner, H):
    """Same (seed, site) -> the fused kernels regenerate exactly the attention-dropout mask of the unfused tcgen05 kernels:
    outputs agree to bf16 rounding of q/k/v (identical here: the unfused path is fed the bf16 qkv the fused one computes)."""
    torch.manual_seed(2)
    D, I = 96, H * 64
    R = n_seq * N
    h = torch.randn(R, D).bfloat16().to(DEV)
    w = (torch.randn(3 * I, D) * D ** -0.5).bfloat16().to(DEV)
    g_out = torch.randn(R, I).bfloat16().to(DEV)
    out, lse, dims = _call_fwd(h, w, n_seq, N, inner, H, drop_p=0.25, seed=77, site=5)
    dqkv, dh = _call_bwd(dims, h, w, g_out, lse)
    # unfused: qkv by an fp32-accumulated matmul rounded to bf16 (what the tcgen05 GEMM produces), then the attention kernels
    qkv = (h.float() @ w.float().t()).bfloat16().requires_grad_(True)
    ref = ops.attention(qkv, n_seq=n_seq, N=N, inner=inner, heads=H, dim_head=64, drop_p=0.25, seed=77, site=5)
    (ref.float() * g_out.float()).sum().backward()
    assert rel_l2(out, ref) < 4e-3            # a different mask would give O(1) differences
    assert rel_l2(dqkv, qkv.grad) < 8e-3
    want_dh = qkv.grad.float() @ w.float()
    assert rel_l2(dh, want_dh) < 8e-3
    # determinism
    out2, _, _ = _call_fwd(h, w, n_seq, N, inner, H, drop_p=0.25, seed=77, site=5)
    assert torch.equal(out, out2)


# ---- forward with the fused tail: out-projection + residual + FeedForward pre-norm in the same kernel ----
def _call_out_fwd(h, w, w_out, b_out, x, ln_w, ln_b, n_seq, N, inner, H, drop_p=0.0, seed=0, site=0, site_out=1, want_o=True):
    R, D = h.shape
    I = H * 64
    dims = _lib.AttnDims(n_seq, N, inner, H, 64, float(drop_p), seed, site, PREC_BF16, None)
    out = torch.full((R, I), float("nan"), device=DEV, dtype=torch.bfloat16) if want_o else None
    lse = torch.full((R, H), float("nan"), device=DEV, dtype=torch.float32)
    xmid = torch.full((R, D), float("nan"), device=DEV, dtype=torch.float32)
    h2 = torch.full((R, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    stats = torch.full((R, 2), float("nan"), device=DEV, dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    check(_lib.lib().msst_attn_block_out_fwd(C.byref(dims), D, h.data_ptr(), w.data_ptr(), out.data_ptr() if want_o else None, lse.data_ptr(),
                                             w_out.data_ptr(), b_out.data_ptr(), x.data_ptr(), xmid.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(),
                                             h2.data_ptr(), stats.data_ptr(), site_out, st))
    return out, lse, xmid, h2, stats


def _tail_inputs(R, D, I):
    w_out = (torch.randn(D, I) * I ** -0.5).bfloat16().to(DEV)
    b_out = torch.randn(D).to(DEV)
    x = torch.randn(R, D).to(DEV)
    ln_w = (1 + 0.2 * torch.randn(D)).to(DEV)
    ln_b = (0.2 * torch.randn(D)).to(DEV)
    return w_out, b_out, x, ln_w, ln_b


@pytest.mark.parametrize("n_seq,N,inner,H,D", GEOMS)
def test_attn_block_out_fwd_vs_fp64(n_seq, N, inner, H, D):
    """xmid = x + o Wo^T + b and h2 = LN(xmid) from the fused kernel against fp64 torch on the kernel's own bf16 o
    (reference Attention.to_out + residual + PreNorm, src/vit_spatial_spectral.py:77,102,25-29); o / lse must be those of the
    plain forward kernel bit for bit."""
    torch.manual_seed(3)
    R, I = n_seq * N, H * 64
    h = torch.randn(R, D).bfloat16().to(DEV)
    w = (torch.randn(3 * I, D) * D ** -0.5).bfloat16().to(DEV)
    w_out, b_out, x, ln_w, ln_b = _tail_inputs(R, D, I)
    o_ref, lse_ref, _ = _call_fwd(h, w, n_seq, N, inner, H)
    out, lse, xmid, h2, stats = _call_out_fwd(h, w, w_out, b_out, x, ln_w, ln_b, n_seq, N, inner, H)
    torch.cuda.synchronize()
    assert torch.equal(out, o_ref) and torch.equal(lse, lse_ref)
    for t in (xmid, h2.float(), stats):
        assert torch.isfinite(t).all()
    want = x.double() + out.double() @ w_out.double().t() + b_out.double()
    assert rel_l2(xmid, want) < 2e-6
    mean = xmid.double().mean(-1)
    var = xmid.double().var(-1, unbiased=False)
    assert rel_l2(stats[:, 0], mean) < 1e-5 and rel_l2(stats[:, 1], (var + 1e-5).rsqrt()) < 1e-5
    want_h2 = torch.nn.functional.layer_norm(xmid.double(), (D,), ln_w.double(), ln_b.double(), 1e-5)
    assert rel_l2(h2, want_h2) < 4e-3
    # inference form: no o output
    _, lse3, xmid3, h23, _ = _call_out_fwd(h, w, w_out, b_out, x, ln_w, ln_b, n_seq, N, inner, H, want_o=False)
    assert torch.equal(xmid3, xmid) and torch.equal(h23, h2) and torch.equal(lse3, lse)


@pytest.mark.parametrize("n_seq,N,inner,H", [(16, 64, 1, 8), (256, 5, 64, 8), (2368, 64, 1, 8)])
def test_attn_block_out_matches_unfused_gemm_with_dropout(n_seq, N, inner, H):
    """Same (seed, site) -> the fused tail regenerates the out-projection dropout mask of the stand-alone GEMM epilogue."""
    torch.manual_seed(4)
    D, I = 96, H * 64
    R = n_seq * N
    h = torch.randn(R, D).bfloat16().to(DEV)
    w = (torch.randn(3 * I, D) * D ** -0.5).bfloat16().to(DEV)
    w_out, b_out, x, ln_w, ln_b = _tail_inputs(R, D, I)
    out, lse, xmid, h2, stats = _call_out_fwd(h, w, w_out, b_out, x, ln_w, ln_b, n_seq, N, inner, H, drop_p=0.25, seed=77, site=5, site_out=9)
    o_ref, _, _ = _call_fwd(h, w, n_seq, N, inner, H, drop_p=0.25, seed=77, site=5)
    assert torch.equal(out, o_ref)
    y = torch.empty(R, D, device=DEV)
    d = _lib.LinearDims(R, D, I, 0, 0.25, 77, 9, PREC_BF16, None, 1)
    st = torch.cuda.current_stream().cuda_stream
    check(_lib.lib().msst_linear_fwd(C.byref(d), out.data_ptr(), w_out.data_ptr(), b_out.data_ptr(), x.data_ptr(), y.data_ptr(), None, st))
    torch.cuda.synchronize()
    assert rel_l2(xmid, y) < 2e-6                 # a different mask would give O(1) differences
    want_h2 = torch.nn.functional.layer_norm(xmid.double(), (D,), ln_w.double(), ln_b.double(), 1e-5)
    assert rel_l2(h2, want_h2) < 4e-3
    out2, _, xmid2, h22, _ = _call_out_fwd(h, w, w_out, b_out, x, ln_w, ln_b, n_seq, N, inner, H, drop_p=0.25, seed=77, site=5, site_out=9)
    assert torch.equal(xmid, xmid2) and torch.equal(h2, h22)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("n_seq,N,inner,H,D", [
    (900, 64, 1, 1, 96),      # one head: every item ends a tile, all epilogue phases share one hosting item
    (900, 64, 1, 2, 96),
    (900, 64, 1, 3, 64),
    (1500, 64, 1, 4, 96),     # H - 1 = NCH: the LayerNorm phase lands on the tile's last head
    (2048, 20, 64, 2, 96),    # strided sequences (4-D boxes), few heads
    (4736, 64, 1, 8, 32),     # 16 tiles per CTA, one 32-column chunk
])
def test_attn_block_out_many_tiles_per_cta(n_seq, N, inner, H, D):
    """Several tiles per CTA with few heads: the epilogue phases of tile k run under the items of tile k + 1 and share hosting items;
    checks the hand-off protocol (a protocol error shows up as a hang -> timeout, or as wrong rows) against the unfused GEMM."""
    torch.manual_seed(5)
    I, R = H * 64, n_seq * N
    h = torch.randn(R, D).bfloat16().to(DEV)
    w = (torch.randn(3 * I, D) * D ** -0.5).bfloat16().to(DEV)
    w_out, b_out, x, ln_w, ln_b = _tail_inputs(R, D, I)
    for rep in range(3):
        out, lse, xmid, h2, stats = _call_out_fwd(h, w, w_out, b_out, x, ln_w, ln_b, n_seq, N, inner, H, drop_p=0.1, seed=11 + rep, site=3, site_out=4)
    o_ref, lse_ref, _ = _call_fwd(h, w, n_seq, N, inner, H, drop_p=0.1, seed=13, site=3)
    assert torch.equal(out, o_ref) and torch.equal(lse, lse_ref)
    y = torch.empty(R, D, device=DEV)
    d = _lib.LinearDims(R, D, I, 0, 0.1, 13, 4, PREC_BF16, None, 1)
    st = torch.cuda.current_stream().cuda_stream
    check(_lib.lib().msst_linear_fwd(C.byref(d), out.data_ptr(), w_out.data_ptr(), b_out.data_ptr(), x.data_ptr(), y.data_ptr(), None, st))
    torch.cuda.synchronize()
    assert rel_l2(xmid, y) < 2e-6
    want_h2 = torch.nn.functional.layer_norm(xmid.double(), (D,), ln_w.double(), ln_b.double(), 1e-5)
    assert rel_l2(h2, want_h2) < 4e-3
    mean = xmid.double().mean(-1)
    assert rel_l2(stats[:, 0], mean) < 1e-5
