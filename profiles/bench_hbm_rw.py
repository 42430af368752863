"""Write-only vs read+write HBM bandwidth of the box (torch memset / copy of 1 GiB, CUDA events): the ceiling a write-heavy kernel can reach."""
import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
def t(f, n=10):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
us = t(lambda: x.zero_()); print(f"memset 1 GiB: {us:.1f} us  {(1<<30)/us/1e3:.0f} GB/s (write only)")
us = t(lambda: y.copy_(x)); print(f"copy 1 GiB: {us:.1f} us  {2*(1<<30)/us/1e3:.0f} GB/s (read+write)")
