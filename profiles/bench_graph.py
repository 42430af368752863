"""Small-batch regime: eager step vs CUDA-graph replay (GraphedStep) of the SimMIM pre-training step, bf16, Houston shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maskedsst_b200 as M
from maskedsst_b200.optim import FusedAdam
from maskedsst_b200.graph import GraphedStep

def build(capturable):
    torch.manual_seed(5)
    enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=4, heads=8,
                               mlp_dim=64, dropout=0.1, emb_dropout=0.1, channels=50, spectral_pos_embed=False, precision="bf16")
    m = M.SimMIMSpatialSpectral(encoder=enc, masking_ratio=0.7, mask_patch_size=4, tube_masking=True, to_pixels_per_spectral_block=True).cuda().train()
    m.mask_backend = "device"
    return m, FusedAdam(m.parameters(), lr=0.008, weight_decay=0.05, clamp=1.0, capturable=capturable)

for B in (16, 64, 256, 1024):
    x = torch.randn(B, 50, 8, 8, device="cuda")
    m, opt = build(False)
    def eager():
        opt.zero_grad(); loss = m(x); loss.backward(); opt.step(); return loss
    for _ in range(5): eager()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 30
    for _ in range(n): eager()
    torch.cuda.synchronize(); te = (time.perf_counter() - t0) / n
    m2, opt2 = build(True)
    g = GraphedStep(m2, opt2, (x,))
    for _ in range(5): g(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): g(x)
    torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / n
    print(f"B={B:5d}  eager {te*1e3:7.3f} ms ({B/te:9.0f} samples/s)   graph {tg*1e3:7.3f} ms ({B/tg:9.0f} samples/s)   loss {float(g(x)):.5f}", flush=True)
