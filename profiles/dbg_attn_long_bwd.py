import sys, os, torch
sys.path.insert(0, "/root/repo")
from maskedsst_b200 import ops
torch.manual_seed(0)
N, H = 4096, 8
n_seq = 256
qkv = torch.randn(n_seq * N, 3 * H * 64, device="cuda").bfloat16().requires_grad_(True)
w = torch.randn(n_seq * N, H * 64, device="cuda").bfloat16()
for _ in range(3):
    qkv.grad = None
    ops.attention(qkv, n_seq=n_seq, N=N, heads=H, dim_head=64).backward(w)
torch.cuda.synchronize()
