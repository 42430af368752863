"""GPU: tcgen05 bf16 GEMMs (TN forward/dgrad kernel with fused epilogues, MN-major split-K wgrad kernel) through the
C-ABI vs fp32/fp64 torch matmul of the SAME bf16-rounded operands.  Tolerance: bf16 output rounding (2^-8 relative
per element) -> rel-l2 <= 4e-3 for bf16 outputs, <= 1e-5 for fp32 outputs (fp32 accumulation of exact products)."""
import ctypes as C
import pytest
import torch

from maskedsst_b200 import _lib
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


@pytest.mark.parametrize("M,N,K", [(128, 96, 64), (300, 1536, 96), (1000, 96, 512), (257, 64, 96), (4096, 512, 96),
                                   (64, 96, 1536), (131, 40, 72), (20000, 1536, 96)])
@pytest.mark.parametrize("out_fp32", [0, 1])
def test_gemm_tn_plain(M, N, K, out_fp32):
    torch.manual_seed(0)
    x = torch.randn(M, K, device=DEV).bfloat16()
    W = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    y = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    d = _lib.LinearDims(M, N, K, 0, 0.0, 0, 0, _lib.PREC_BF16, None, out_fp32)
    _lib.check(_lib.lib().msst_linear_fwd(C.byref(d), _p(x), _p(W), None, None, _p(y), None, _st()))
    torch.cuda.synchronize()
    want = x.double() @ W.double().T
    assert torch.isfinite(y.float()).all()
    assert rel_l2(y, want) < (2e-6 if out_fp32 else 4e-3)


def test_gemm_tn_epilogues():
    torch.manual_seed(1)
    M, N, K = 777, 64, 96
    x = torch.randn(M, K, device=DEV).bfloat16()
    W = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    lib = _lib.lib()
    # bias + GELU, bf16 out, pre-activation out
    y = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    pre = torch.empty_like(y)
    d = _lib.LinearDims(M, N, K, 1, 0.0, 0, 0, _lib.PREC_BF16, None, 0)
    _lib.check(lib.msst_linear_fwd(C.byref(d), _p(x), _p(W), _p(bias), None, _p(y), _p(pre), _st()))
    u = x.double() @ W.double().T + bias.double()
    assert rel_l2(pre, u) < 4e-3
    assert rel_l2(y, torch.nn.functional.gelu(u)) < 4e-3
    # bias + residual, fp32 out
    y2 = torch.empty(M, N, device=DEV)
    d = _lib.LinearDims(M, N, K, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 1)
    _lib.check(lib.msst_linear_fwd(C.byref(d), _p(x), _p(W), _p(bias), _p(res), _p(y2), None, _st()))
    assert rel_l2(y2, u + res.double()) < 2e-6
    # dropout: kept elements scaled by 1/(1-p), same mask as msst_dropout_apply at the same (seed, site)
    y3 = torch.empty(M, N, device=DEV)
    d = _lib.LinearDims(M, N, K, 0, 0.25, 77, 5, _lib.PREC_BF16, None, 1)
    _lib.check(lib.msst_linear_fwd(C.byref(d), _p(x), _p(W), _p(bias), None, _p(y3), None, _st()))
    ones = torch.ones(M * N, device=DEV); fac = torch.empty_like(ones)
    _lib.check(lib.msst_dropout_apply(_p(ones), _p(fac), M * N, 0.25, 77, 5, None, _st()))
    assert rel_l2(y3, u * fac.view(M, N).double()) < 2e-6
    assert abs((fac == 0).float().mean().item() - 0.25) < 0.02
    # data gradient with GELU' epilogue: dx = (dy @ Wt^T) * gelu'(aux)
    dy = torch.randn(M, N, device=DEV).bfloat16()
    Wt = W.T.contiguous()                      # [K, N]: dx[M,K] = dy[M,N] . Wt[K,N]^T
    aux = torch.randn(M, K, device=DEV).bfloat16()
    dx = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
    d = _lib.LinearDims(M, N, K, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 0)
    _lib.check(lib.msst_linear_bwd_data(C.byref(d), _p(dy), _p(Wt), _p(aux), None, _p(dx), _st()))
    a = aux.double()
    gp = 0.5 * (1 + torch.erf(a / 2 ** 0.5)) + a * torch.exp(-0.5 * a * a) / (2 * torch.pi) ** 0.5
    assert rel_l2(dx, (dy.double() @ W.double()) * gp) < 4e-3


@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (1000, 1536, 96), (5000, 96, 512), (333, 64, 96), (4099, 96, 64),
                                   (100000, 1536, 96)])
def test_gemm_wgrad(M, N, K):
    torch.manual_seed(2)
    dy = torch.randn(M, N, device=DEV).bfloat16()
    x = torch.randn(M, K, device=DEV).bfloat16()
    dW = torch.zeros(N, K, device=DEV)
    d = _lib.LinearDims(M, N, K, 0, 0.0, 0, 0, _lib.PREC_BF16, None, 1)
    _lib.check(_lib.lib().msst_linear_bwd_weight(C.byref(d), _p(dy), _p(x), _p(dW), None, _st()))
    torch.cuda.synchronize()
    want = dy.double().T @ x.double()
    assert rel_l2(dW, want) < 1e-5
    # accumulates
    _lib.check(_lib.lib().msst_linear_bwd_weight(C.byref(d), _p(dy), _p(x), _p(dW), None, _st()))
    assert rel_l2(dW, 2 * want) < 1e-5
