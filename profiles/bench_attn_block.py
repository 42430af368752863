"""Fused projection + attention kernels (csrc/attn_block_tc.cu) vs the unfused pipeline at the bench's stack shapes:
CUDA-event timings of  [QKV GEMM + attention fwd]  vs  attn_block_fwd  and  [attention bwd + Wqkv data-gradient GEMM]  vs
attn_block_bwd.  MSST_AB_DBG=<n> prints the clock64 timeline of CTA 0 at the n-th launch of each fused kernel."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskedsst_b200 import _lib                      # noqa: E402

lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
B = int(os.environ.get("B", 1024))
Cb = int(os.environ.get("C", 5))
T = Cb * 64
R, H, I, D = B * T, 8, 512, 96
torch.manual_seed(0)
h = torch.randn(R, D, device="cuda").bfloat16()
w = (torch.randn(3 * I, D, device="cuda") * D ** -0.5).bfloat16()
wt = w.t().contiguous()
qkv = torch.empty(R, 3 * I, device="cuda", dtype=torch.bfloat16)
o = torch.empty(R, I, device="cuda", dtype=torch.bfloat16); lse = torch.empty(R, H, device="cuda")
do = torch.randn(R, I, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv)
dh = torch.empty(R, D, device="cuda")
w_out = (torch.randn(D, I, device="cuda") * I ** -0.5).bfloat16(); b_out = torch.randn(D, device="cuda")
xres = torch.randn(R, D, device="cuda"); xmid = torch.empty(R, D, device="cuda"); h2 = torch.empty(R, D, device="cuda", dtype=torch.bfloat16)
ln_w = torch.ones(D, device="cuda"); ln_b = torch.zeros(D, device="cuda"); stats2 = torch.empty(R, 2, device="cuda")
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


for name, (n_seq, N, inner) in (("spatial", (B * Cb, 64, 1)), ("spectral", (B * 64, Cb, 64))):
    for p in (0.0, 0.1):
        ad = _lib.AttnDims(n_seq, N, inner, H, 64, p, 1234, 16, _lib.PREC_BF16, None)
        ld = _lib.LinearDims(R, 3 * I, D, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 0)
        ld2 = _lib.LinearDims(R, 3 * I, D, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 1)
        f_gemm = lambda: _lib.check(lib.msst_linear_fwd(C.byref(ld), h.data_ptr(), w.data_ptr(), None, None, qkv.data_ptr(), None, st))
        f_attn = lambda: _lib.check(lib.msst_attention_fwd(C.byref(ad), qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), st))
        f_fused = lambda: _lib.check(lib.msst_attn_block_fwd(C.byref(ad), D, h.data_ptr(), w.data_ptr(), o.data_ptr(), lse.data_ptr(), st))
        b_attn = lambda: _lib.check(lib.msst_attention_bwd(C.byref(ad), qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), do.data_ptr(), dqkv.data_ptr(), st))
        b_dgrad = lambda: _lib.check(lib.msst_linear_bwd_data(C.byref(ld2), dqkv.data_ptr(), wt.data_ptr(), None, None, dh.data_ptr(), st))
        b_fused = lambda: _lib.check(lib.msst_attn_block_bwd(C.byref(ad), D, h.data_ptr(), w.data_ptr(), wt.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                                             dqkv.data_ptr(), dh.data_ptr(), st))
        # tail of the attention block: out-projection + dropout + residual + LN2, stand-alone GEMM launch vs folded into the fused forward
        ld3 = _lib.LinearDims(R, D, I, 0, p, 1234, 17, _lib.PREC_BF16, None, 1)
        f_oproj = lambda: _lib.check(lib.msst_linear_fwd(C.byref(ld3), o.data_ptr(), w_out.data_ptr(), b_out.data_ptr(), xres.data_ptr(), xmid.data_ptr(), None, st))
        f_tail = lambda: _lib.check(lib.msst_attn_block_out_fwd(C.byref(ad), D, h.data_ptr(), w.data_ptr(), o.data_ptr(), lse.data_ptr(), w_out.data_ptr(),
                                                                b_out.data_ptr(), xres.data_ptr(), xmid.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(),
                                                                h2.data_ptr(), stats2.data_ptr(), 17, st))
        items = (n_seq * N // 128) * H
        f_gemm(); f_attn()
        t = [timeit(f) for f in (f_gemm, f_attn, f_fused, b_attn, b_dgrad, b_fused)]
        t2 = [timeit(f) for f in (f_oproj, f_tail)]
        print(f"{name:8s} B={B} C={Cb} drop {p}: fwd + tail  fused {t[2]:6.1f} + out-proj GEMM (no LN) {t2[0]:6.1f} = {t[2] + t2[0]:6.1f} us   fused with tail {t2[1]:6.1f} us"
              f" ({t2[1] * 1.9e3 * 148 / items:5.0f} clk/item)", flush=True)
        print(f"{name:8s} B={B} C={Cb} drop {p}: fwd  qkv-gemm {t[0]:6.1f} + attn {t[1]:6.1f} = {t[0] + t[1]:6.1f} us   fused {t[2]:6.1f} us ({t[2] * 1.9e3 * 148 / items:5.0f} clk/item)"
              f" | bwd  attn {t[3]:6.1f} + dgrad {t[4]:6.1f} = {t[3] + t[4]:6.1f} us   fused {t[5]:6.1f} us ({t[5] * 1.9e3 * 148 / items:5.0f} clk/item)", flush=True)
