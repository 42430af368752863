"""Stress run of the tcgen05 attention kernels: many shapes x many launches, results checked against the mma.sync kernels' invariants
(finite, deterministic across repeats).  A hang shows up as the caller's `timeout` firing."""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskedsst_b200 import ops

torch.manual_seed(0)
shapes = [(5, 64, 1), (7, 64, 1), (148 * 2 + 3, 64, 1), (1023, 64, 1), (64 * 3, 5, 64), (64 * 37, 5, 64), (64 * 9, 20, 64), (64 * 2, 22, 64),
          (16 * 5, 7, 16), (3, 33, 1), (2048, 16, 1), (64 * 300, 5, 64), (5120, 64, 1)]
for (n_seq, N, inner), H, p in itertools.product(shapes, (1, 8), (0.0, 0.1)):
    R, I = n_seq * N, H * 64
    qkv = torch.randn(R, 3 * I, device="cuda").bfloat16().requires_grad_(True)
    w = torch.randn(R, I, device="cuda").bfloat16()
    ref = None
    for rep in range(4):
        qkv.grad = None
        o = ops.attention(qkv, n_seq=n_seq, N=N, inner=inner, heads=H, dim_head=64, drop_p=p, seed=7, site=3)
        (o.float() * w.float()).sum().backward()
        cur = (o.detach().clone(), qkv.grad.clone())
        assert torch.isfinite(cur[0]).all() and torch.isfinite(cur[1]).all(), (n_seq, N, inner, H, p)
        if ref is None:
            ref = cur
        else:
            assert torch.equal(ref[0], cur[0]) and torch.equal(ref[1], cur[1]), ("non-deterministic", n_seq, N, inner, H, p, rep)
    torch.cuda.synchronize()
print("STRESS_OK", len(shapes) * 4)
