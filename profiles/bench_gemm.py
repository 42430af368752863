"""Micro-benchmark of the tcgen05 GEMM shapes of one transformer layer (CUDA events, L2-cold via big tensors)."""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maskedsst_b200 import _lib
B = int(os.environ.get("B", 1024)); R = B * 320
dev = "cuda"
lib = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
def run(name, M, N, K, out_fp32, residual=False, bias=False, act=0, aux=False, pre=False, reps=5):
    x = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    y = torch.empty(M, N, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    res = torch.randn(M, N, device=dev) if residual else None
    b = torch.randn(N, device=dev) if bias else None
    a = torch.randn(M, N, device=dev).bfloat16() if aux else None
    pr = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if pre else None
    d = _lib.LinearDims(M, N, K, act, 0.0, 0, 0, _lib.PREC_BF16, None, out_fp32)
    p = lambda t: None if t is None else t.data_ptr()
    def call():
        if aux:
            _lib.check(lib.msst_linear_bwd_data(C.byref(_lib.LinearDims(M, K, N, 0, 0.0, 0, 0, _lib.PREC_BF16, None, out_fp32)), p(x), p(W), p(a), None, p(y), st))
        else:
            _lib.check(lib.msst_linear_fwd(C.byref(d), p(x), p(W), p(b), p(res), p(y), p(pr), st))
    if aux:   # dgrad call computes dx[M, K'] with K' = N here: rebuild operands so that out is [M,N]: dy [M,K], Wt [N,K]
        pass
    for _ in range(2): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    byts = M * K * 2 + N * K * 2 + M * N * (4 if out_fp32 else 2) + (M * N * 4 if residual else 0) + (M * N * 2 if aux else 0) + (M * N * 2 if pre else 0)
    print(f"{name:28s} M={M} N={N:5d} K={K:5d} {us:8.1f} us  {byts/us/1e3:7.1f} GB/s  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)
print("debug", os.environ.get("MSST_GEMM_DEBUG", "0"))
run("qkv fwd", R, 1536, 96, 0)
run("outproj fwd (+b,res)", R, 96, 512, 1, residual=True, bias=True)
run("mlp1 fwd (gelu,pre)", R, 64, 96, 0, bias=True, act=1, pre=True)
run("mlp2 fwd (+b,res)", R, 96, 64, 1, residual=True, bias=True)
run("dgrad w1 (fp32 out)", R, 96, 64, 1)
run("dgrad wo (bf16 out)", R, 512, 96, 0)
run("dgrad wqkv (fp32 out)", R, 96, 1536, 1)
run("dgrad w2 bf16", R, 64, 96, 0)
