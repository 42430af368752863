"""Attention backward: parity vs an fp64 torch reference and CUDA-event timing at the bench's spatial-stack shape.
Run once per variant (the selector is read once per process):
    MSST_ATTN_BWD_TC=0 python profiles/bench_attn_bwd.py      # mma.sync kernel
    MSST_ATTN_BWD_TC=1 python profiles/bench_attn_bwd.py      # tcgen05 / TMEM kernel
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskedsst_b200 import _lib, ops                      # noqa: E402
from tests.test_gpu_components import ref_attention       # noqa: E402
from tests.helpers import rel_l2                          # noqa: E402

print("MSST_ATTN_BWD_TC =", os.environ.get("MSST_ATTN_BWD_TC", "(default)"))
torch.manual_seed(0)
for n_seq, N, inner, H in [(10, 64, 1, 8), (33, 64, 1, 3), (5, 32, 1, 2), (7, 16, 1, 8), (1, 64, 1, 1), (128, 5, 64, 8), (192, 20, 64, 8), (6, 22, 2, 4), (5, 22, 1, 2), (1, 1, 1, 1)]:
    R, I = n_seq * N, H * 64
    qkv = torch.randn(R, 3 * I).bfloat16(); w = torch.randn(R, I).bfloat16()
    a = qkv.double().requires_grad_(True)
    want = ref_attention(a, n_seq, N, inner, H, 64); (want * w.double()).sum().backward()
    b = qkv.cuda().requires_grad_(True)
    got = ops.attention(b, n_seq=n_seq, N=N, inner=inner, heads=H, dim_head=64)
    (got.float() * w.cuda().float()).sum().backward()
    e = [rel_l2(b.grad[:, i * I:(i + 1) * I], a.grad[:, i * I:(i + 1) * I]) for i in range(3)]
    print(f"n_seq {n_seq:4d} N {N:3d} inner {inner:2d} H {H}: fwd {rel_l2(got, want):.2e}  dq {e[0]:.2e} dk {e[1]:.2e} dv {e[2]:.2e}", flush=True)

# dropout: checksums must agree between the two backward kernels (same regenerated mask)
n_seq, N, H = 6, 64, 2
R, I = n_seq * N, H * 64
qkv = torch.randn(R, 3 * I).bfloat16().cuda()
b = qkv.clone().requires_grad_(True)
o = ops.attention(b, n_seq=n_seq, N=N, heads=H, dim_head=64, drop_p=0.3, seed=11, site=3)
o.float().sum().backward()
print("dropout run: grad finite", bool(torch.isfinite(b.grad).all()), "checksums",
      [round(float(b.grad.float()[:, i * I:(i + 1) * I].abs().sum()), 3) for i in range(3)])

# timing at the bench shapes: spatial stack (B*5 sequences of 64 contiguous rows) and spectral stack (B*64 sequences of 5 rows, stride 64)
lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
B = int(os.environ.get("B", 1024))
R, H, I = B * 320, 8, 512
qkv = torch.randn(R, 3 * I, device="cuda").bfloat16()
o = torch.empty(R, I, device="cuda", dtype=torch.bfloat16); lse = torch.empty(R, H, device="cuda")
do = torch.randn(R, I, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv)
for name, (n_seq, N, inner) in (("spatial", (B * 5, 64, 1)), ("spectral", (B * 64, 5, 64))):
    for p in (0.0, 0.1):
        ad = _lib.AttnDims(n_seq, N, inner, H, 64, p, 1234, 16, _lib.PREC_BF16, None)
        _lib.check(lib.msst_attention_fwd(C.byref(ad), qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), st))
        fb = lambda: _lib.check(lib.msst_attention_bwd(C.byref(ad), qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), do.data_ptr(), dqkv.data_ptr(), st))
        ff = lambda: _lib.check(lib.msst_attention_fwd(C.byref(ad), qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), st))
        for tag, f, gb in (("fwd", ff, R * 4 * I * 2 + R * H * 4), ("bwd", fb, R * 7 * I * 2 + R * H * 4)):
            for _ in range(3):
                f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10):
                f()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            print(f"{name} {tag} B={B} drop {p}: {us:.1f} us  {gb / us / 1e3:.0f} GB/s  checksum {float((o if tag == 'fwd' else dqkv).float().abs().mean()):.6f}", flush=True)
