"""GPU, 2 ranks over NCCL (needs >= 2 visible GPUs; skipped otherwise): a data-parallel SimMIM step equals the
single-rank step on the global batch (fp32 mode, dropout off)."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _build(dev):
    import maskedsst_b200 as M
    from oracle import maskedsst_oracle as O
    spec = O.Spec(**O.HOUSTON, depth=2)
    enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=20, dim=96, depth=2,
                               heads=8, mlp_dim=64, channels=50, spectral_pos_embed=False)
    m = M.SimMIMSpatialSpectral(encoder=enc, masking_ratio=0.7, mask_patch_size=4, tube_masking=True, to_pixels_per_spectral_block=True)
    m.load_state_dict(O.synthetic_state_dict(spec, seed=5, simmim=True))
    x = O.synthetic_cube(spec, 8, seed=5, zero_pad_bands=2)
    np.random.seed(5)
    mask, idx = m.draw_masks(8, "cpu")
    return m.to(dev).train(), x, mask, idx


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from maskedsst_b200.optim import FusedAdam
    from maskedsst_b200.dp import GradSync
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = f"cuda:{rank}"
        m, x, mask, idx = _build(dev)
        opt = FusedAdam(m.parameters(), lr=0.008, weight_decay=0.05, clamp=1.0, grad_scale=1.0 / world)
        sync = GradSync(opt.arena, num_buckets=3)
        sl = slice(rank * 4, rank * 4 + 4)
        for _ in range(2):
            opt.zero_grad()
            loss = m(x[sl].to(dev), masks=(mask[sl].to(dev), idx[sl].to(dev)))
            loss.backward()
            sync.finish()
            opt.step()
        torch.cuda.synchronize()
        q.put((rank, opt.param_arena.cpu()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp2_nccl_equals_single_rank():
    from maskedsst_b200.optim import FusedAdam
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in procs)
    [p.join(60) for p in procs]
    m, x, mask, idx = _build("cuda:0")
    opt = FusedAdam(m.parameters(), lr=0.008, weight_decay=0.05, clamp=1.0)
    for _ in range(2):
        opt.zero_grad()
        # the global-batch loss: mean over 8 samples == mean of the two 4-sample means
        loss = m(x.to("cuda:0"), masks=(mask.to("cuda:0"), idx.to("cuda:0")))
        loss.backward()
        opt.step()
    ref = opt.param_arena.cpu()
    assert torch.equal(res[0], res[1])
    assert float((res[0] - ref).norm() / ref.norm()) < 1e-5
