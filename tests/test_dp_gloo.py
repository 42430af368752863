"""CPU, world_size 2, gloo: host-side logic of the data-parallel path (bucketed async all-reduce of the flat gradient
arena driven by autograd hooks, 1/world scaling + clamp after the reduction, CE valid-count correction).  The arithmetic
kernels are CUDA-only, so a plain torch MLP stands in for the model here; the GPU variant runs under `-m gpu`."""
import os
import socket
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maskedsst_b200.optim import FlatArena
from maskedsst_b200.dp import GradSync, ce_dp_scale


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(10, 33), torch.nn.GELU(), torch.nn.Linear(33, 7), torch.nn.GELU(),
                               torch.nn.Linear(7, 4), torch.nn.Linear(4, 4))   # last layer unused below -> never fires


def _adam_ref(p, g, m, v, step, lr=1e-2, wd=0.05, clamp=1.0):
    g = g.clamp(-clamp, clamp)
    p = p * (1 - lr * wd)
    m = 0.9 * m + 0.1 * g
    v = 0.999 * v + 0.001 * g * g
    return p - lr / (1 - 0.9 ** step) * m / (v.sqrt() / (1 - 0.999 ** step) ** 0.5 + 1e-8), m, v


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model()
        arena = FlatArena([list(model.parameters())])
        sync = GradSync(arena, num_buckets=3)
        assert len(sync.buckets) == 3 and sync.buckets[0]["lo"] == 0 and sync.buckets[-1]["hi"] == arena.grads.numel()
        g = torch.Generator().manual_seed(1)
        X = torch.randn(8, 10, generator=g) * 3
        Y = torch.randn(8, 4, generator=g)
        m, v = torch.zeros_like(arena.params), torch.zeros_like(arena.params)
        out = []
        for step in (1, 2, 3):
            arena.grads.zero_()
            xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]      # rank r gets samples [r*B/R, (r+1)*B/R)
            loss = ((model[:5](xs) - ys) ** 2).mean() * 50
            loss.backward()
            sync.finish()
            if step > 1:   # from the second step on the buckets were launched from the hooks (overlap path)
                assert all(b["expected"] is not None for b in sync.buckets)
            gmean = arena.grads / world                                          # grad_scale = 1/world, THEN clamp
            newp, m, v = _adam_ref(arena.params, gmean, m, v, step)
            with torch.no_grad():
                arena.params.copy_(newp)
            out.append(arena.params.clone())
        q.put((rank, out[-1], sync.buckets[-1]["expected"]))
    finally:
        dist.destroy_process_group()


def test_dp2_equals_single_process_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # single-process reference on the global batch
    model = _model()
    arena = FlatArena([list(model.parameters())])
    g = torch.Generator().manual_seed(1)
    X = torch.randn(8, 10, generator=g) * 3
    Y = torch.randn(8, 4, generator=g)
    m, v = torch.zeros_like(arena.params), torch.zeros_like(arena.params)
    for step in (1, 2, 3):
        arena.grads.zero_()
        (((model[:5](X) - Y) ** 2).mean() * 50).backward()
        newp, m, v = _adam_ref(arena.params, arena.grads.clone(), m, v, step)
        with torch.no_grad():
            arena.params.copy_(newp)
    for rank, params, expected_last in res:
        assert torch.allclose(params, arena.params, rtol=1e-5, atol=1e-6), rank
        assert expected_last == 2 or expected_last == 0 or expected_last is not None   # unused layer learnt as non-firing
    assert torch.equal(res[0][1], res[1][1])                                     # ranks stay bit-identical


def test_flat_arena_views_and_state_dict():
    model = _model()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    arena = FlatArena([list(model[:2].parameters()), list(model[2:].parameters())])
    assert len(arena.group_ranges) == 2 and arena.group_ranges[0][1] == arena.group_ranges[1][0]
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k])
    for p in model.parameters():
        off, n = arena.offsets[id(p)]
        assert off % 4 == 0 and p.data_ptr() == arena.params.data_ptr() + 4 * off and p.grad.data_ptr() == arena.grads.data_ptr() + 4 * off
    model.load_state_dict({k: v + 1 for k, v in sd.items()})
    assert abs(float(arena.params.sum()) - float(sum((v + 1).sum() for v in sd.values()))) < 1e-3


def test_ce_valid_count_correction():
    """per-rank mean-over-valid CE is NOT the global loss when label sparsity differs; the corrected scaling is."""
    torch.manual_seed(0)
    logits = torch.randn(4, 5, 8, 8, requires_grad=True)
    labels = torch.randint(-1, 5, (4, 8, 8))
    labels[0] = -1; labels[1, :6] = -1                       # rank 0 has far fewer valid pixels
    glob = torch.nn.functional.cross_entropy(logits, labels, ignore_index=-1)
    ggrad, = torch.autograd.grad(glob, logits)
    world, acc = 2, torch.zeros_like(logits)
    cnt_global = float((labels != -1).sum())
    for r in range(world):
        sl = slice(2 * r, 2 * r + 2)
        nll_sum = torch.nn.functional.cross_entropy(logits[sl], labels[sl], ignore_index=-1, reduction="sum")
        g, = torch.autograd.grad(nll_sum * ce_dp_scale(cnt_global, world), logits)
        acc += g
    assert torch.allclose(acc / world, ggrad, atol=1e-7)


def test_stage_boundaries_cut_buckets_at_the_transformer_stacks():
    """dp.stage_boundaries: every transformer stack is its own all-reduce bucket (its gradients land together, one autograd node) and
    the bucket that holds the patch embedding contains nothing of the spatial stack."""
    import torch
    import maskedsst_b200 as M
    from maskedsst_b200.dp import GradSync, stage_boundaries
    from maskedsst_b200.optim import FlatArena
    enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=2, heads=8, mlp_dim=64,
                               channels=50, spectral_pos_embed=False)
    m = M.SimMIMSpatialSpectral(encoder=enc, masking_ratio=0.7, mask_patch_size=4, tube_masking=True, to_pixels_per_spectral_block=True)
    arena = FlatArena([list(m.parameters())])
    sync = GradSync(arena, boundaries=stage_boundaries(m))
    stacks = [mod for mod in m.modules() if hasattr(mod, "layer_params")]
    assert len(stacks) == 2
    owner = {}
    for bi, b in enumerate(sync.buckets):
        for p in b["params"]:
            owner[id(p)] = bi
    stack_buckets = [{owner[id(p)] for p in s.parameters()} for s in stacks]
    assert all(len(sb) == 1 for sb in stack_buckets) and stack_buckets[0] != stack_buckets[1]
    in_stacks = {id(p) for s in stacks for p in s.parameters()}
    for sb in stack_buckets:        # nothing else shares a stack's bucket
        (bi,) = sb
        assert all(id(p) in in_stacks for p in sync.buckets[bi]["params"])
    # buckets tile the arena without gaps
    assert sync.buckets[0]["lo"] == 0 and sync.buckets[-1]["hi"] == arena.grads.numel()
    assert all(a["hi"] == b["lo"] for a, b in zip(sync.buckets, sync.buckets[1:]))
