"""Fused FeedForward kernel (csrc/mlp_block_tc.cu) vs the two GEMM launches it replaces (gemm_tn_kernel<3> + <6>), CUDA-event timings
with an L2 flush between launches, at the bench's shape (B*T rows, D = 96, M = 64)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskedsst_b200 import _lib   # noqa: E402

lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
B = int(os.environ.get("B", 1024))
R, D, M = B * 320, int(os.environ.get("D", 96)), 64
torch.manual_seed(0)
h2 = torch.randn(R, D, device="cuda").bfloat16(); xmid = torch.randn(R, D, device="cuda")
w1 = (torch.randn(M, D, device="cuda") * D ** -0.5).bfloat16(); w2 = (torch.randn(D, M, device="cuda") * M ** -0.5).bfloat16()
b1 = torch.zeros(M, device="cuda"); b2 = torch.zeros(D, device="cuda"); lnw = torch.ones(D, device="cuda"); lnb = torch.zeros(D, device="cuda")
u = torch.empty(R, M, device="cuda", dtype=torch.bfloat16); g = torch.empty_like(u)
y = torch.empty(R, D, device="cuda"); h1 = torch.empty(R, D, device="cuda", dtype=torch.bfloat16); stats = torch.empty(R, 2, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timeit(f, n=10):
    for _ in range(3):
        f()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1000


for p in (0.0, 0.1):
    fused = lambda: _lib.check(lib.msst_mlp_block_fwd(h2.data_ptr(), xmid.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(), b2.data_ptr(), u.data_ptr(),
                                                      g.data_ptr(), y.data_ptr(), lnw.data_ptr(), lnb.data_ptr(), h1.data_ptr(), stats.data_ptr(), R, D, M, p, 7, 18, 19,
                                                      None, st))
    d1 = _lib.LinearDims(R, M, D, 1, p, 7, 18, _lib.PREC_BF16, None, 0)
    d2 = _lib.LinearDims(R, D, M, 0, p, 7, 19, _lib.PREC_BF16, None, 1)
    gemm1 = lambda: _lib.check(lib.msst_linear_fwd(C.byref(d1), h2.data_ptr(), w1.data_ptr(), b1.data_ptr(), None, g.data_ptr(), u.data_ptr(), st))
    gemm2 = lambda: _lib.check(lib.msst_linear_fwd(C.byref(d2), g.data_ptr(), w2.data_ptr(), b2.data_ptr(), xmid.data_ptr(), y.data_ptr(), None, st))
    t = [timeit(f) for f in (gemm1, gemm2, fused)]
    byts = R * (D * 2 + D * 4 + 2 * M * 2 + D * 4 + D * 2 + 8)
    print(f"B={B} drop {p}: mlp1 gemm {t[0]:6.1f} + mlp2 gemm (no LN epilogue) {t[1]:6.1f} = {t[0] + t[1]:6.1f} us | fused (with LN) {t[2]:6.1f} us = {byts / t[2] / 1e3:5.0f} GB/s "
          f"({byts / t[2] / 1e3 / 6540.5:.2f} of the measured HBM peak)", flush=True)

# ---- backward: gemm_wgrad(W2) + gemm_tn<4> (du, db1) + gemm_wgrad(W1) + gemm_tn<1> (dh) vs mlp_block_bwd_kernel ----
dyb = torch.randn(R, D, device="cuda").bfloat16()
w1t, w2t = w1.t().contiguous(), w2.t().contiguous()
u.copy_(torch.randn(R, M, device="cuda")); g.copy_(torch.randn(R, M, device="cuda"))
dw1 = torch.zeros(M, D, device="cuda"); dw2 = torch.zeros(D, M, device="cuda"); db1 = torch.zeros(M, device="cuda")
dh = torch.empty(R, D, device="cuda"); du = torch.empty(R, M, device="cuda", dtype=torch.bfloat16)
for p in (0.0, 0.1):
    fused = lambda: _lib.check(lib.msst_mlp_block_bwd(dyb.data_ptr(), u.data_ptr(), g.data_ptr(), h2.data_ptr(), w2t.data_ptr(), w1t.data_ptr(), dw1.data_ptr(),
                                                      dw2.data_ptr(), db1.data_ptr(), dh.data_ptr(), R, D, M, p, 7, 18, None, st))
    dwa = _lib.LinearDims(R, D, M, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 0)
    dwb = _lib.LinearDims(R, M, D, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 0)
    dda = _lib.LinearDims(R, D, M, 0, p, 7, 18, _lib.PREC_BF16, None, 0)
    ddb = _lib.LinearDims(R, M, D, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 1)
    k1 = lambda: _lib.check(lib.msst_linear_bwd_weight(C.byref(dwa), dyb.data_ptr(), g.data_ptr(), dw2.data_ptr(), None, st))
    k2 = lambda: _lib.check(lib.msst_linear_bwd_data(C.byref(dda), dyb.data_ptr(), w2t.data_ptr(), u.data_ptr(), None, du.data_ptr(), st))
    k3 = lambda: _lib.check(lib.msst_linear_bwd_weight(C.byref(dwb), du.data_ptr(), h2.data_ptr(), dw1.data_ptr(), None, st))
    k4 = lambda: _lib.check(lib.msst_linear_bwd_data(C.byref(ddb), du.data_ptr(), w1t.data_ptr(), None, None, dh.data_ptr(), st))
    t = [timeit(f) for f in (k1, k2, k3, k4, fused)]
    byts = R * (2 * D * 2 + 2 * M * 2 + D * 4)
    print(f"B={B} drop {p}: bwd wgrad-w2 {t[0]:6.1f} + dgrad-w2/gelu' {t[1]:6.1f} + wgrad-w1 {t[2]:6.1f} + dgrad-w1 {t[3]:6.1f} = {sum(t[:4]):6.1f} us | fused {t[4]:6.1f} us = "
          f"{byts / t[4] / 1e3:5.0f} GB/s ({byts / t[4] / 1e3 / 6540.5:.2f} of the measured HBM peak)", flush=True)
