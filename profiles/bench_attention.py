"""BASELINE.json configs[4]: spatial-spectral attention fwd / fwd+bwd micro-benchmark, sweeping the token count
(spatial N = (image/patch)^2, spectral N = blocks) and the head dim; inputs randn (seed 5), H = 8, number of sequences
chosen so that total tokens ~ 2^20.  Reports TFLOP/s (4*N*dh FLOP per token-head forward, x3.5 for fwd+bwd: 2 GEMMs fwd,
5 bwd) against the dense bf16 peak of MEASURED_PEAKS.json, and GB/s of q/k/v/o traffic.
bf16 = tensor-core kernels (dh = 64 only), fp32 = FFMA parity kernels (dh 32/64/128).

    python profiles/bench_attention.py > profiles/r01_attention_microbench.txt     (on a B200)
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maskedsst_b200 import ops

peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
PEAK = peaks.get("bf16_tflops", 1590.0)
H = 8
torch.manual_seed(5)


def run(N, dh, dtype, inner=1, total_tokens=1 << 20, reps=5):
    n_seq = max(inner, (total_tokens // N) // inner * inner)
    R, I = n_seq * N, H * dh
    qkv = torch.randn(R, 3 * I, device="cuda").to(dtype).requires_grad_(True)
    w = torch.randn(R, I, device="cuda").to(dtype)
    kw = dict(n_seq=n_seq, N=N, inner=inner, heads=H, dim_head=dh)
    def fwd():
        with torch.no_grad():
            return ops.attention(qkv, **kw)
    def fwdbwd():
        qkv.grad = None
        ops.attention(qkv, **kw).backward(w)
    out = []
    for fn in (fwd, fwdbwd):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / reps)
    flop_f = 4.0 * N * dh * H * R
    es = qkv.element_size()
    print(f"{str(dtype)[6:]:9s} N={N:5d} dh={dh:3d} inner={inner:3d} seqs={n_seq:7d} | fwd {out[0]:9.1f} us {flop_f/out[0]/1e6:7.1f} TF/s "
          f"({100*flop_f/out[0]/1e6/PEAK:5.2f}% of {PEAK:.0f}) {R*4*I*es/out[0]/1e3:7.0f} GB/s | fwd+bwd {out[1]:9.1f} us "
          f"{3.5*flop_f/out[1]/1e6:7.1f} TF/s ({100*3.5*flop_f/out[1]/1e6/PEAK:5.2f}%)", flush=True)


if __name__ == "__main__":
    print(f"# attention micro-benchmark, H={H}, ~2^20 tokens, peak = {PEAK} TFLOP/s dense bf16 ({'measured' if peaks else 'fallback'})")
    for N in (16, 64, 256, 1024, 4096):            # spatial sequences: image 4..64, patch 1
        run(N, 64, torch.bfloat16)
    for N in (5, 20, 22, 40):                      # spectral sequences (strided rows, inner = 64)
        run(N, 64, torch.bfloat16, inner=64)
    for dh in (32, 64, 128):                       # head-dim sweep: fp32 parity kernels
        run(64, dh, torch.float32, total_tokens=1 << 18)
        run(1024, dh, torch.float32, total_tokens=1 << 18)
