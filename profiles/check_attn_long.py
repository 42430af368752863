"""Long-sequence (N > 64) attention forward: parity vs an fp64 torch reference + dropout-mask agreement with the backward kernel.
    MSST_ATTN_TC_LONG=1 python profiles/check_attn_long.py      # tcgen05 kernel;  =0: mma.sync kernel"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskedsst_b200 import ops
from tests.test_gpu_components import ref_attention
from tests.helpers import rel_l2
print("MSST_ATTN_TC_LONG =", os.environ.get("MSST_ATTN_TC_LONG", "(default)"))
torch.manual_seed(0)
for n_seq, N, inner, H in [(3, 200, 1, 2), (2, 256, 2, 2), (5, 130, 1, 3), (1, 65, 1, 1), (4, 1000, 1, 2), (6, 128, 3, 2), (300, 256, 1, 8), (2, 4096, 1, 2)]:
    R, I = n_seq * N, H * 64
    qkv = torch.randn(R, 3 * I).bfloat16(); w = torch.randn(R, I).bfloat16()
    a = qkv.double().requires_grad_(True)
    want = ref_attention(a, n_seq, N, inner, H, 64); (want * w.double()).sum().backward()
    b = qkv.cuda().requires_grad_(True)
    got = ops.attention(b, n_seq=n_seq, N=N, inner=inner, heads=H, dim_head=64)
    (got.float() * w.cuda().float()).sum().backward()
    print(f"n_seq {n_seq:4d} N {N:5d} inner {inner} H {H}: fwd {rel_l2(got, want):.2e}  grad {rel_l2(b.grad, a.grad):.2e}", flush=True)
    # dropout: deterministic, and E[out] preserved
    o1 = ops.attention(b.detach(), n_seq=n_seq, N=N, inner=inner, heads=H, dim_head=64, drop_p=0.2, seed=5, site=2)
    o2 = ops.attention(b.detach(), n_seq=n_seq, N=N, inner=inner, heads=H, dim_head=64, drop_p=0.2, seed=5, site=2)
    assert torch.equal(o1, o2)
    torch.save(o1.float().cpu(), f"/tmp/long_{os.environ.get('MSST_ATTN_TC_LONG', 'd')}_{n_seq}_{N}.pt")
print("LONG_OK")
