"""Launches each roofline kernel of bench.py exactly once (after one warm-up each) at the bench's shapes, for
    ncu --set full --clock-control none -k regex:'attn_block|gemm_wgrad' -s 5 -c 5 --csv --page raw --log-file gpurun_out/ncu_targets_raw.csv python profiles/ncu_targets.py
profiles/ncu_traffic.py turns that CSV into profiles/r02_ncu_traffic.json (dram bytes per launch keyed by kernel + shape + R)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskedsst_b200 import _lib   # noqa: E402

ORDER = ["attn_block_fwd_kernel:spatial", "attn_block_bwd_kernel:spatial", "attn_block_fwd_kernel:spectral", "attn_block_bwd_kernel:spectral",
         "gemm_wgrad_kernel:wqkv"]

if __name__ == "__main__":
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    B, Cb = int(os.environ.get("B", 1024)), 5
    R, H, I, D = B * Cb * 64, 8, 512, 96
    torch.manual_seed(0)
    h = torch.randn(R, D, device="cuda").bfloat16()
    w = (torch.randn(3 * I, D, device="cuda") * D ** -0.5).bfloat16()
    wt = w.t().contiguous()
    o = torch.empty(R, I, device="cuda", dtype=torch.bfloat16); lse = torch.empty(R, H, device="cuda")
    do = torch.randn(R, I, device="cuda").bfloat16(); dqkv = torch.empty(R, 3 * I, device="cuda", dtype=torch.bfloat16)
    dh = torch.empty(R, D, device="cuda"); dW = torch.zeros(3 * I, D, device="cuda")
    w_out = (torch.randn(D, I, device="cuda") * I ** -0.5).bfloat16(); b_out = torch.randn(D, device="cuda")
    xres = torch.randn(R, D, device="cuda"); xmid = torch.empty(R, D, device="cuda"); h2 = torch.empty(R, D, device="cuda", dtype=torch.bfloat16)
    ln_w = torch.ones(D, device="cuda"); ln_b = torch.zeros(D, device="cuda"); stats2 = torch.empty(R, 2, device="cuda")
    fns = []
    for n_seq, N, inner in ((B * Cb, 64, 1), (B * 64, Cb, 64)):
        ad = _lib.AttnDims(n_seq, N, inner, H, 64, 0.1, 1234, 16, _lib.PREC_BF16, None)
        fns.append(lambda ad=ad: _lib.check(lib.msst_attn_block_out_fwd(C.byref(ad), D, h.data_ptr(), w.data_ptr(), o.data_ptr(), lse.data_ptr(), w_out.data_ptr(),
                                                                        b_out.data_ptr(), xres.data_ptr(), xmid.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(),
                                                                        h2.data_ptr(), stats2.data_ptr(), 17, st)))
        fns.append(lambda ad=ad: _lib.check(lib.msst_attn_block_bwd(C.byref(ad), D, h.data_ptr(), w.data_ptr(), wt.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                                                    dqkv.data_ptr(), dh.data_ptr(), st)))
    ld = _lib.LinearDims(R, 3 * I, D, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 0)
    fns.append(lambda: _lib.check(lib.msst_linear_bwd_weight(C.byref(ld), dqkv.data_ptr(), h.data_ptr(), dW.data_ptr(), None, st)))
    for f in fns:      # warm-up pass (skipped by ncu -s 5)
        f()
    torch.cuda.synchronize()
    for f in fns:      # measured pass, in ORDER
        f()
    torch.cuda.synchronize()
    print("R =", R)
