"""GPU component tests (fp32 mode) through the C-ABI ops: attention geometries, linear/LayerNorm gradients,
dropout statistics, fused Adam vs torch.optim."""
import math
import numpy as np
import pytest
import torch

import maskedsst_b200 as M
from maskedsst_b200 import ops
from maskedsst_b200.optim import FusedAdam
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ref_attention(qkv, n_seq, N, inner, H, dh):
    """plain fp64 softmax attention over the strided row layout (CPU)."""
    R = n_seq * N
    I = H * dh
    rows = torch.empty(n_seq, N, dtype=torch.long)
    for s in range(n_seq):
        base = (s // inner) * N * inner + (s % inner)
        rows[s] = base + torch.arange(N) * inner
    q, k, v = qkv.double().split(I, dim=-1)
    g = lambda t: t[rows].reshape(n_seq, N, H, dh).permute(0, 2, 1, 3)
    p = torch.softmax(g(q) @ g(k).transpose(-1, -2) * dh ** -0.5, dim=-1)
    o = (p @ g(v)).permute(0, 2, 1, 3).reshape(n_seq, N, I)
    out = torch.zeros(R, I, dtype=torch.float64)
    out[rows.reshape(-1)] = o.reshape(-1, I)
    return out


@pytest.mark.parametrize("n_seq,N,inner,H,dh", [
    (10, 64, 1, 8, 64),     # spatial transformer geometry
    (128, 5, 64, 8, 64),    # Houston spectral: 12 sequences packed per tile, stride 64
    (128, 20, 64, 8, 64),   # EnMAP spectral
    (6, 22, 2, 4, 32),      # ragged packing, dh 32
    (3, 200, 1, 2, 64),     # N > 64: online softmax over 4 key tiles (last one ragged)
    (2, 256, 2, 2, 128),    # long + strided + dh 128
    (1, 1, 1, 1, 64),       # degenerate
])
def test_attention_fwd_bwd(n_seq, N, inner, H, dh):
    torch.manual_seed(0)
    R, I = n_seq * N, H * dh
    qkv = torch.randn(R, 3 * I)
    w = torch.randn(R, I, dtype=torch.float64)
    a = qkv.clone().double().requires_grad_(True)
    want = ref_attention(a, n_seq, N, inner, H, dh)
    (want * w).sum().backward()
    b = qkv.to(DEV).requires_grad_(True)
    got = ops.attention(b, n_seq=n_seq, N=N, inner=inner, heads=H, dim_head=dh)
    (got * w.to(DEV).float()).sum().backward()
    assert rel_l2(got, want) < 1e-5
    assert rel_l2(b.grad, a.grad) < 2e-5


def test_linear_and_layernorm_grads():
    torch.manual_seed(1)
    x = torch.randn(1000, 96)
    W = torch.randn(200, 96) / 10
    bias = torch.randn(200)
    lw, lb = torch.rand(96) + 0.5, torch.randn(96)
    ps = [t.clone().double().requires_grad_(True) for t in (x, W, bias, lw, lb)]
    want = torch.nn.functional.layer_norm(ps[0], (96,), ps[3], ps[4]) @ ps[1].T + ps[2]
    want.square().sum().backward()
    gs = [t.clone().to(DEV).requires_grad_(True) for t in (x, W, bias, lw, lb)]
    got = ops.linear(ops.layer_norm(gs[0], gs[3], gs[4]), gs[1], gs[2])
    got.square().sum().backward()
    assert rel_l2(got, want) < 1e-5
    for g, p in zip(gs, ps):
        assert rel_l2(g.grad, p.grad) < 2e-5


@pytest.mark.parametrize("M_,N_,K_", [(1, 1, 1), (77, 10, 13), (513, 96, 64), (64, 1536, 96)])
def test_linear_odd_shapes(M_, N_, K_):
    torch.manual_seed(2)
    x, W, b = torch.randn(M_, K_), torch.randn(N_, K_), torch.randn(N_)
    ps = [t.clone().double().requires_grad_(True) for t in (x, W, b)]
    (ps[0] @ ps[1].T + ps[2]).sin().sum().backward()
    gs = [t.clone().to(DEV).requires_grad_(True) for t in (x, W, b)]
    y = ops.linear(gs[0], gs[1], gs[2])
    y.sin().sum().backward()
    assert rel_l2(y, ps[0] @ ps[1].T + ps[2]) < 1e-5
    for g, p in zip(gs, ps):
        assert rel_l2(g.grad, p.grad) < 2e-5


def test_dropout_statistics_and_determinism():
    """Training-mode dropout (p = 0.1 everywhere in the shipped configs): masks are Philox-regenerated, so (a) the same
    seed gives the same output, (b) another seed differs, (c) keep-rate ~ 1-p, (d) gradients are consistent with the
    forward mask (finite differences along a direction)."""
    enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=2,
                               heads=8, mlp_dim=64, channels=50, spectral_pos_embed=False, dropout=0.1, emb_dropout=0.1).to(DEV)
    x = torch.randn(4, 50, 8, 8, device=DEV)
    enc.train()
    y1 = enc(x); y2 = enc(x)
    assert not torch.equal(y1, y2)
    enc.eval()
    assert torch.equal(enc(x), enc(x))
    # keep-rate through the stand-alone entry point
    from maskedsst_b200 import _lib
    import ctypes as C
    n = 1 << 20
    a = torch.ones(n, device=DEV); b = torch.empty_like(a)
    _lib.check(_lib.lib().msst_dropout_apply(a.data_ptr(), b.data_ptr(), n, 0.1, 1234, 7, None,
                                            torch.cuda.current_stream().cuda_stream))
    keep = (b != 0).float().mean().item()
    assert abs(keep - 0.9) < 3e-3
    assert abs(b.mean().item() - 1.0) < 5e-3
    assert torch.all((b == 0) | ((b - 1 / 0.9).abs() < 1e-6))
    # fixed-seed stack: forward determinism and gradient/forward mask consistency
    tf = enc.spatial_spectral_transformer[1]
    rows = torch.randn(4 * 5 * 64, 96, device=DEV, requires_grad=True)
    kw = dict(n_seq=20, N=64, inner=1, heads=8, dim_head=64, mlp_dim=64, drop_p=0.1, seed=99)
    f = lambda r: ops.transformer_stack(r, tf.layer_params(), **kw)
    out = f(rows)
    assert torch.equal(out, f(rows))
    d = torch.randn_like(rows)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    eps = 1e-2
    with torch.no_grad():
        fd = ((f(rows + eps * d) - f(rows - eps * d)) * w).sum().item() / (2 * eps)
    an = (rows.grad * d).sum().item()
    assert abs(fd - an) < 2e-2 * max(1.0, abs(an))


@pytest.mark.parametrize("decoupled,wd,clamp", [(True, 0.05, 1.0), (False, 0.005, 0.0)])
def test_fused_adam_matches_torch_optim(decoupled, wd, clamp):
    torch.manual_seed(3)
    shapes = [(96,), (1536, 96), (7,), (1, 321, 96), (10, 96)]
    ref_p = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref_p]
    Opt = torch.optim.AdamW if decoupled else torch.optim.Adam
    groups_ref = [{"params": ref_p[:2], "lr": 0.008}, {"params": ref_p[2:], "lr": 0.0005}]
    groups_our = [{"params": our_p[:2], "lr": 0.008}, {"params": our_p[2:], "lr": 0.0005}]
    ref = Opt(groups_ref, lr=0.008, weight_decay=wd)
    ours = FusedAdam(groups_our, lr=0.008, weight_decay=wd, decoupled=decoupled, clamp=clamp)
    for step in range(5):
        ours.zero_grad(set_to_none=False)     # gradients stay arena views: they are written by hand below
        for rp, op in zip(ref_p, our_p):
            g = torch.randn(rp.shape) * 3
            rp.grad = g.clamp(-clamp, clamp) if clamp > 0 else g.clone()
            op.grad.copy_(g.to(DEV))
        ours.mark_all_touched()          # gradients were written without autograd
        ref.step(); ours.step()
    for rp, op in zip(ref_p, our_p):
        assert torch.allclose(op.detach().cpu(), rp.detach(), rtol=2e-6, atol=2e-7)
    assert our_p[1].data_ptr() == ours.param_arena.data_ptr() + 4 * ours.offset_of(our_p[1])[0]


@pytest.mark.parametrize("decoupled", [True, False])
def test_fused_adam_skips_parameters_without_gradient(decoupled):
    """torch.optim.Adam(W) skips parameters whose .grad is None: no moments, no weight decay (the reference's unused
    encoder.mlp_head during SimMIM pre-training keeps its LayerNorm gamma at exactly 1).  The arena gives every parameter a
    zero gradient view, so FusedAdam tracks which parameters autograd touched."""
    torch.manual_seed(4)
    shapes = [(96,), (64, 96), (33,), (96,), (10, 7)]
    used = [True, True, False, True, False]
    ref_p = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref_p]
    Opt = torch.optim.AdamW if decoupled else torch.optim.Adam
    ref = Opt(ref_p, lr=0.01, weight_decay=0.05)
    ours = FusedAdam(our_p, lr=0.01, weight_decay=0.05, decoupled=decoupled)
    for step in range(4):
        ref.zero_grad(); ours.zero_grad()
        coef = [torch.randn(s) for s in shapes]
        sum((p * c).sum() for p, c, u in zip(ref_p, coef, used) if u).backward()
        sum((p * c.to(DEV)).sum() for p, c, u in zip(our_p, coef, used) if u).backward()
        ref.step(); ours.step()
    for rp, op, u in zip(ref_p, our_p, used):
        assert torch.allclose(op.detach().cpu(), rp.detach(), rtol=2e-6, atol=2e-7)
    assert ref_p[2].grad is None   # the premise: torch never touched them
    # a gradient that left the arena after backward is an error, not a silent no-op
    our_p[0].grad = None
    with pytest.raises(RuntimeError, match="arena"):
        ours.step()
    our_p[0].grad = torch.zeros_like(our_p[0])
    with pytest.raises(RuntimeError, match="arena"):
        ours.step()


def test_sequential_call_matches_fast_path():
    """Calling the nn.Sequential directly (re-tiling copies, the reference's calling convention) equals
    transformer_forward (stride addressing, no copies)."""
    enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=2,
                               heads=8, mlp_dim=64, channels=50, spectral_pos_embed=False).to(DEV).eval()
    x = torch.randn(3, 320, 96, device=DEV)
    with torch.no_grad():
        a = enc.transformer_forward(x)
        b = enc.spatial_spectral_transformer(x)
    assert rel_l2(a, b) < 1e-6


def test_larger_image_long_sequences():
    """image_size 16 -> S = 256 spatial tokens per sequence (N > 64 path inside the full model)."""
    from oracle import maskedsst_oracle as O
    spec = O.Spec(image_size=16, channels=30, num_classes=5, depth=1)
    sd = O.synthetic_state_dict(spec, seed=51)
    m = M.ViTSpatialSpectral(image_size=16, spatial_patch_size=1, spectral_patch_size=10, num_classes=5, dim=96, depth=1,
                             heads=8, mlp_dim=64, channels=30, spectral_pos_embed=False).eval()
    m.load_state_dict(sd); m.to(DEV)
    x = O.synthetic_cube(spec, 2, seed=51)
    with torch.no_grad():
        assert rel_l2(m(x.to(DEV)), O.encoder_forward(x, sd, spec)) < 1e-5


def test_cuda_graph_step_matches_eager():
    """A captured training step (GraphedStep) replays to the same parameters as eager steps (dropout off), and with
    dropout on it draws different masks per replay."""
    from maskedsst_b200.graph import GraphedStep
    from oracle import maskedsst_oracle as O

    def build(p):
        torch.manual_seed(0)
        enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=1,
                                   heads=8, mlp_dim=64, channels=50, spectral_pos_embed=False, dropout=p, emb_dropout=p)
        return enc.to(DEV).train()

    x = torch.randn(6, 50, 8, 8, device=DEV)
    y = torch.randint(-1, 20, (6, 8, 8), device=DEV)
    loss_fn = lambda m, xi, yi: M.cross_entropy(m(xi), yi, ignore_index=-1)
    a, b = build(0.0), build(0.0)
    oa = FusedAdam(a.parameters(), lr=0.01, weight_decay=0.05, clamp=1.0)
    ob = FusedAdam(b.parameters(), lr=0.01, weight_decay=0.05, clamp=1.0, capturable=True)
    p0 = ob.param_arena.clone()
    g = GraphedStep(b, ob, (x, y), loss_fn=loss_fn, warmup=2)        # warm-up steps are rolled back: capturing changes no training state
    assert torch.equal(ob.param_arena, p0) and int(ob._step_t) == 0 and float(ob.exp_avg.abs().sum()) == 0.0
    for _ in range(3):
        lg = g(x, y)                                                 # replayed steps 1..3
        oa.zero_grad(); le = loss_fn(a, x, y); le.backward(); oa.step()
    torch.cuda.synchronize()
    assert abs(float(lg) - float(le.detach())) < 1e-5 * abs(float(le.detach()))
    assert rel_l2(ob.param_arena, oa.param_arena) < 1e-5
    assert ob.state_dict()["step"] == 3                              # the device counter, not the stale host one
    g.recapture()                                                    # e.g. after a scheduler LR change: must not move the state either
    assert rel_l2(ob.param_arena, oa.param_arena) < 1e-5 and ob.state_dict()["step"] == 3
    # dropout: two replays on the same input give different losses (fresh masks), and training state stays finite
    c = build(0.2)
    oc = FusedAdam(c.parameters(), lr=0.0, weight_decay=0.0, capturable=True)   # lr 0: parameters frozen, only masks change
    gc = GraphedStep(c, oc, (x, y), loss_fn=loss_fn, warmup=1)
    l1 = float(gc(x, y)); l2 = float(gc(x, y))
    assert l1 != l2 and np.isfinite(l1) and np.isfinite(l2)
