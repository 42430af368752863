"""Timeline (MSST_AB_DBG=1) of one attn_block forward launch with the fused tail at the bench's spatial / spectral stack shape."""
import ctypes as C
import os
import sys

import torch

os.environ.setdefault("MSST_AB_DBG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskedsst_b200 import _lib                      # noqa: E402

lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
B, Cb = int(os.environ.get("B", 1024)), 5
R, H, I, D = B * Cb * 64, 8, 512, 96
tail = int(os.environ.get("TAIL", 1))
spectral = int(os.environ.get("SPECTRAL", 0))
h = torch.randn(R, D, device="cuda").bfloat16()
w = (torch.randn(3 * I, D, device="cuda") * D ** -0.5).bfloat16()
o = torch.empty(R, I, device="cuda", dtype=torch.bfloat16); lse = torch.empty(R, H, device="cuda")
w_out = (torch.randn(D, I, device="cuda") * I ** -0.5).bfloat16(); b_out = torch.randn(D, device="cuda")
xres = torch.randn(R, D, device="cuda"); xmid = torch.empty(R, D, device="cuda"); h2 = torch.empty(R, D, device="cuda", dtype=torch.bfloat16)
ln_w = torch.ones(D, device="cuda"); ln_b = torch.zeros(D, device="cuda"); stats2 = torch.empty(R, 2, device="cuda")
n_seq, N, inner = (B * 64, Cb, 64) if spectral else (B * Cb, 64, 1)
ad = _lib.AttnDims(n_seq, N, inner, H, 64, 0.1, 1234, 16, _lib.PREC_BF16, None)
if tail:
    _lib.check(lib.msst_attn_block_out_fwd(C.byref(ad), D, h.data_ptr(), w.data_ptr(), o.data_ptr(), lse.data_ptr(), w_out.data_ptr(), b_out.data_ptr(),
                                           xres.data_ptr(), xmid.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), h2.data_ptr(), stats2.data_ptr(), 17, st))
else:
    _lib.check(lib.msst_attn_block_fwd(C.byref(ad), D, h.data_ptr(), w.data_ptr(), o.data_ptr(), lse.data_ptr(), st))
torch.cuda.synchronize()
