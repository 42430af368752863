#!/usr/bin/env python
"""bench.py -- SimMIM pre-training step throughput (samples/s) of the B200-native MaskedSST hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision bf16|fp32] [--impl reference]

Workload (BASELINE.json configs[1]): one SimMIM masked-pretraining step = forward + backward + AdamW (+ the reference's
grad clamp) of ViTSpatialSpectral(dim 96, depth 4+4, heads 8, mlp 64, dropout 0.1) on synthetic Houston2018-shaped
cubes [B, 50, 8, 8] (bands 48-49 zero), tube masking ratio 0.7 / mask patch 4, blockwise decoder, lr .008 wd .05
(configs/config.yaml, configs/pretrain_config.yaml of the reference).  Random-init weights, synthetic data.

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs (cubes + masks) resident in HBM;
`e2e` = the same step through the public nn.Module API with HOST inputs: pinned cube -> H2D, host mask generation
(the reference's numpy generator), loss read back D2H every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_TRAIN = {"houston": 3.754e9, "enmap": 15.49e9}   # 3 x forward GEMM+attention FLOPs


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 1024 ours / 32 reference)")
    ap.add_argument("--precision", default=os.environ.get("MSST_BENCH_PRECISION", "bf16"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dataset", default="houston", choices=["houston", "enmap"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "finetune"],
                    help="pretrain = BASELINE configs[1] (default); finetune = CE classification step with Adam (configs[2])")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference algorithm, torch CPU fp32, all host threads
# ------------------------------------------------------------------------------------------------------------
def cpu_step_time(dataset, batch, steps, warmup, dropout, budget_s=25.0):
    import numpy as np
    import torch
    from oracle import maskedsst_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = O.Spec(**(O.HOUSTON if dataset == "houston" else O.ENMAP))
    sd = O.synthetic_state_dict(spec, seed=5, simmim=True)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=0.008, weight_decay=0.05)
    gen = O.MaskGen(spec.image_size, 4, 1, 0.7)
    x = O.synthetic_cube(spec, batch, seed=5, zero_pad_bands=2 if dataset == "houston" else 0)
    nm = int(0.7 * spec.T)
    np.random.seed(5)
    times = []
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        mask, idx = gen.batch(batch, spec.C, nm, tube=True)
        opt.zero_grad()
        loss = O.simmim_forward(x, params, spec, mask, idx, drop=dropout)
        loss.backward()
        for p in params.values():            # pretrain.py:71-73 elementwise clamp
            if p.grad is not None:
                p.grad.clamp_(-1, 1)
        opt.step()
        loss.item()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if it >= warmup and time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    times.sort()
    med = times[len(times) // 2]
    return med, cores, len(times)


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    batch = args.batch or 32
    med, cores, n = cpu_step_time(args.dataset, batch, args.steps, min(args.warmup, 2), args.dropout, budget_s=120.0)
    v = batch / med
    line = {
        "impl": "reference", "metric": "simmim_pretrain_samples_per_sec", "value": v, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": n, "warmup": min(args.warmup, 2), "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SimMIM pretrain step (fwd+bwd+AdamW+clamp), ViTSpatialSpectral {args.dataset} shape, CPU",
                   "per_gpu_batch": batch, "dropout": args.dropout},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{n} timed steps of batch {batch} (median), torch CPU fp32 oracle port of the reference"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import maskedsst_b200 as M
    from maskedsst_b200 import _lib
    from maskedsst_b200.optim import FusedAdam

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or 1024
    ds = args.dataset
    channels, ncls = (50, 20) if ds == "houston" else (200, 8)
    torch.manual_seed(5)
    np.random.seed(5 + rank)
    enc = M.ViTSpatialSpectral(image_size=8, spatial_patch_size=1, spectral_patch_size=10, num_classes=ncls, dim=96, depth=4,
                               heads=8, mlp_dim=64, dropout=args.dropout, emb_dropout=args.dropout, channels=channels,
                               spectral_pos_embed=False, blockwise_patch_embed=True, spectral_pos=list(range(channels // 10)),
                               precision=args.precision)
    model = M.SimMIMSpatialSpectral(encoder=enc, masking_ratio=0.7, mask_patch_size=4, tube_masking=True,
                                    to_pixels_per_spectral_block=True).to(dev).train()
    if args.workload == "finetune":   # finetune.py:116-135: Adam + L2, head lr .005 / rest .0005, CE(ignore_index=-1)
        from maskedsst_b200.dp import cross_entropy_dp
        model = enc.to(dev).train()
        head = [p for n, p in model.named_parameters() if "mlp_head" in n]
        rest = [p for n, p in model.named_parameters() if "mlp_head" not in n]
        opt = FusedAdam([{"params": head, "lr": 0.005}, {"params": rest, "lr": 0.0005}], lr=0.0005, weight_decay=0.005,
                        decoupled=False, grad_scale=1.0 / world)
    else:
        opt = FusedAdam(model.parameters(), lr=0.008, weight_decay=0.05, decoupled=True, clamp=1.0, grad_scale=1.0 / world)

    # synthetic inputs: a pool of pinned host batches (per step a different one) + their masks
    pool = 4
    g = torch.Generator().manual_seed(1234 + rank)
    host_x = [torch.randn(B, channels, 8, 8, generator=g).pin_memory() for _ in range(pool)]
    if ds == "houston":
        for t in host_x:
            t[:, 48:] = 0
    dev_x = [t.to(dev) for t in host_x]
    if args.workload == "finetune":
        host_y = [torch.randint(-1, ncls, (B, 8, 8), generator=g).pin_memory() for _ in range(pool)]
        dev_y = [t.to(dev) for t in host_y]
    else:
        dev_masks = [model.draw_masks(B, dev) for _ in range(pool)]

    from maskedsst_b200.dp import GradSync
    sync = GradSync(opt.arena, num_buckets=3) if world > 1 else None

    def allreduce_grads():   # bucketed all-reduce overlapped with backward (hooks); this only drains it
        if sync is not None:
            sync.finish()

    def step_resident(i):
        opt.zero_grad()
        if args.workload == "finetune":
            loss = cross_entropy_dp(model(dev_x[i % pool]), dev_y[i % pool], ignore_index=-1)
        else:
            loss = model(dev_x[i % pool], masks=dev_masks[i % pool])
        loss.backward()
        allreduce_grads()
        opt.step()
        return loss

    def step_e2e(i):
        x = host_x[i % pool].to(dev, non_blocking=True)      # H2D from pinned memory, inside the timed region
        opt.zero_grad()
        if args.workload == "finetune":
            loss = cross_entropy_dp(model(x), host_y[i % pool].to(dev, non_blocking=True), ignore_index=-1)
        else:
            loss = model(x)                                  # public API: model(img); masks drawn inside (see mask_backend)
        loss.backward()
        allreduce_grads()
        opt.step()
        return loss.item()                                   # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_resident(i)
    barrier()
    clocks = Clocks(local_rank)
    clocks.start()
    n0 = _lib.lib().msst_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss = step_resident(i)
    ev1.record()
    barrier()
    launches = _lib.lib().msst_launch_count() - n0
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    final_loss = float(loss.item())

    # e2e leg: (a) masks drawn on the device (module option mask_backend="device"), (b) the reference-compatible host generator
    e2e_steps = max(3, min(args.steps, 10))
    e2e_res = {}
    for backend in ("device", "host"):
        model.mask_backend = backend
        for i in range(2):
            step_e2e(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            step_e2e(i)
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_res[backend] = float(t.item())
    e2e_s = e2e_res["device"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    value = world * B * args.steps / (ms / 1e3)
    roof = roofline(args, B, dev, model, peaks)
    line = {
        "metric": "simmim_pretrain_samples_per_sec" if args.workload == "pretrain" else "finetune_samples_per_sec",
        "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": ("SimMIM pretrain step (fwd+bwd+AdamW+clamp)" if args.workload == "pretrain" else
                                "finetune step (fwd+CE+bwd+Adam, 2 lr groups)") + f", ViTSpatialSpectral {ds} shape "
                               f"[B,{channels},8,8], dim 96 depth 4+4 heads 8 mlp 64, dropout {args.dropout}, tube mask 0.7/4",
                   "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2_policy": "per-step working set (activations ~%.1f GB) exceeds the 126 MB L2; input pool of %d batches"
                                % (B * (320 if ds == "houston" else 1280) * 37e3 / 1e9, pool),
                   "model_tflops_per_s": value * FLOP_PER_SAMPLE_TRAIN[ds] / 1e12, "final_loss": final_loss},
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "samples/s",
                "h2d_bytes_per_step": host_x[0].numel() * 4, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "masks": "drawn on the device inside model(img) (mask_backend='device')",
                "value_host_masks": world * B * e2e_steps / e2e_res["host"],
                "host_masks_note": "reference-compatible numpy generator on the host: +B*(T + 8*nm) H2D bytes per step"},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
    }
    if not args.no_cpu_baseline and world == 1 and args.workload == "pretrain":
        med, cores, n = cpu_step_time(ds, 32, 8, 1, args.dropout, budget_s=20.0)
        line["cpu_baseline"] = {"value": 32 / med, "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": f"{n} timed steps of batch 32 (median) of the same step, torch CPU fp32 oracle port"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _time_launch(fn, reps=10):
    import torch
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / reps


def roofline(args, B, dev, model, peaks):
    """Roofline of the DOMINANT kernel of the step (largest share of the ncu launch list, profiles/r01s3_launches_*.summary.txt):
    the attention backward kernel (bf16 mode: attn_bwd_tc_kernel, tcgen05 / TMEM / TMA), timed alone with CUDA events on the
    launching stream at the step's spatial-stack shape (B*C sequences of 64 tokens, 8 heads, dh 64).  HBM-bound: algorithmic bytes
    per launch = R*(3*I + I)*e [q,k,v,dO in] + R*3*I*e [dq,dk,dv out] + R*H*4 [lse].  `other_kernels`: the same kernel at the
    spectral-stack shape, the attention forward at both shapes and the largest GEMM (QKV projection), measured the same way."""
    import ctypes as C
    import torch
    from maskedsst_b200 import _lib
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    bf16 = args.precision == "bf16"
    prec = _lib.PREC_BF16 if bf16 else _lib.PREC_FP32
    dt = torch.bfloat16 if bf16 else torch.float32
    es = 2 if bf16 else 4
    Cb, T = (5, 320) if args.dataset == "houston" else (20, 1280)
    R, H, dh, I = B * T, 8, 64, 512
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback"
    out = {}
    others = []
    try:
        qkv = torch.randn(R, 3 * I, device=dev).to(dt)
        o = torch.empty(R, I, device=dev, dtype=dt)
        lse = torch.empty(R, H, device=dev)
        do = torch.randn(R, I, device=dev).to(dt)
        dqkv = torch.empty_like(qkv)
        bwd_bytes = R * 4 * I * es + R * 3 * I * es + R * H * 4
        fwd_bytes = R * 3 * I * es + R * I * es + R * H * 4
        kname = "attn_bwd_tc_kernel, tcgen05/TMEM/TMA" if bf16 else "fp32 FFMA kernel"
        for shape, (n_seq, N, inner) in (("spatial", (B * Cb, 64, 1)), ("spectral", (B * 64, Cb, 64))):
            ad = _lib.AttnDims(n_seq, N, inner, H, dh, float(args.dropout), 1234, 16, prec, None)
            fwd = lambda: _lib.check(lib.msst_attention_fwd(C.byref(ad), qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), st))
            bwd = lambda: _lib.check(lib.msst_attention_bwd(C.byref(ad), qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), do.data_ptr(),
                                                            dqkv.data_ptr(), st))
            sec_f = _time_launch(fwd)
            sec_b = _time_launch(bwd)
            if shape == "spatial":
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this kernel at
                # this shape (profiles/r01s3_ncu_full_kernels.txt): 1.371 GB + 0.983 GB
                traffic = 2353589000 if (bf16 and R == 327680) else None
                out = {"kernel": f"attention backward ({kname}), spatial stack shape", "bound": "hbm", "achieved": bwd_bytes / sec_b / 1e9,
                       "peak": hbm_peak, "unit": "GB/s", "frac": bwd_bytes / sec_b / 1e9 / hbm_peak, "traffic": traffic,
                       "algorithmic_bytes": bwd_bytes, "us_per_launch": sec_b * 1e6, "peak_source": src,
                       "tflops": 2.0 * 5 * 64 * 64 * 64 * (R // 64) * H / sec_b / 1e12}
            else:
                others.append({"kernel": f"attention backward ({kname}), spectral stack shape (B*64 sequences of {Cb} tokens, row stride 64)",
                               "bound": "hbm", "achieved": bwd_bytes / sec_b / 1e9, "peak": hbm_peak, "unit": "GB/s",
                               "frac": bwd_bytes / sec_b / 1e9 / hbm_peak, "traffic": 2346758000 if (bf16 and R == 327680) else None,
                               "algorithmic_bytes": bwd_bytes, "us_per_launch": sec_b * 1e6})
            others.append({"kernel": f"attention forward ({'attn_fwd_tc_kernel, tcgen05/TMEM/TMA' if bf16 else 'fp32 FFMA kernel'}), {shape} stack shape",
                           "bound": "hbm", "achieved": fwd_bytes / sec_f / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": fwd_bytes / sec_f / 1e9 / hbm_peak, "traffic": None, "algorithmic_bytes": fwd_bytes,
                           "us_per_launch": sec_f * 1e6})
        del qkv, o, lse, do, dqkv
    except Exception as e:   # noqa
        out = {"error": str(e)}
    try:
        D, N = 96, 1536
        x = torch.randn(R, D, device=dev).to(dt)
        W = (torch.randn(N, D, device=dev) / 10).to(dt)
        y = torch.empty(R, N, device=dev, dtype=dt)
        dims = _lib.LinearDims(R, N, D, 0, 0.0, 0, 0, prec, None, 0)
        sec = _time_launch(lambda: _lib.check(lib.msst_linear_fwd(C.byref(dims), x.data_ptr(), W.data_ptr(), None, None, y.data_ptr(), None, st)))
        gb = R * D * es + R * N * es + N * D * es
        others.append({
            "kernel": "qkv projection GEMM [R,96]x[96,1536] (gemm_tn_kernel<7>: tcgen05 + TMA loads + TMA-store epilogue)" if bf16 else "qkv projection GEMM (fp32 FFMA)",
            "bound": "hbm", "achieved": gb / sec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": gb / sec / 1e9 / hbm_peak,
            "traffic": 1014752768 if (bf16 and R == 327680) else None, "algorithmic_bytes": gb, "us_per_launch": sec * 1e6,
            "tflops": 2.0 * R * N * D / sec / 1e12})
    except Exception as e:   # noqa
        others.append({"error": str(e)})
    out["other_kernels"] = others
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
