"""CUDA-graph capture of a whole training step (zero_grad -> forward -> backward -> fused optimiser).

At the reference's batch sizes (64 pretrain, 32 / 2 finetune) a step is ~160 short kernels and the host (Python autograd,
ctypes calls) is the bottleneck; replaying one captured graph removes it.  What makes the step capturable:
  * every libmsst entry point is asynchronous on the current stream, allocates nothing and never synchronises;
  * dropout masks come from (seed + *seed_dev, site, index): the graph increments a device counter once per replay;
  * FusedAdam(capturable=True) keeps its step counter on the device;
  * SimMIM masks are drawn on the device (mask_backend = "device").
Learning rates are baked in at capture time: call `recapture()` after a scheduler step changes them.
"""
import torch

from . import ops


class GraphedStep:
    def __init__(self, model, optimizer, example_inputs, loss_fn=None, warmup=3):
        """loss_fn(model, *inputs) -> scalar loss; default: model(*inputs).  example_inputs: tuple of CUDA tensors whose
        shapes/dtypes are fixed for the lifetime of the graph."""
        if getattr(optimizer, "_step_t", None) is None:
            raise ValueError("GraphedStep needs FusedAdam(..., capturable=True)")
        if hasattr(model, "mask_backend"):
            model.mask_backend = "device"
        self.model, self.opt = model, optimizer
        self.loss_fn = loss_fn or (lambda m, *xs: m(*xs))
        self.static_in = tuple(torch.empty_like(t) for t in example_inputs)
        for s, t in zip(self.static_in, example_inputs):
            s.copy_(t)
        self.seed_t = torch.zeros((), dtype=torch.int64, device=self.static_in[0].device)
        self.warmup = warmup
        self.graph = None
        self.recapture()

    def _step(self):
        self.seed_t += 1
        self.opt.zero_grad()
        loss = self.loss_fn(self.model, *self.static_in)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def recapture(self):
        ops.set_device_seed(self.seed_t)
        try:
            # the warm-up steps are real optimiser steps on the (stale) static inputs: snapshot the training state and put it
            # back afterwards, so (re)capturing -- e.g. after every scheduler LR change -- never alters parameters, moments or
            # the step counters
            opt = self.opt
            snap = (opt.param_arena.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt._step, opt._step_t.clone(),
                    None if opt.bf16_arena is None else opt.bf16_arena.clone(), self.seed_t.clone())
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self.warmup):       # also runs every one-time cudaFuncSetAttribute outside the capture
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            with torch.no_grad():
                opt.param_arena.copy_(snap[0]); opt.exp_avg.copy_(snap[1]); opt.exp_avg_sq.copy_(snap[2])
                opt._step = snap[3]; opt._step_t.copy_(snap[4])
                if snap[5] is not None:
                    opt.bf16_arena.copy_(snap[5])
                self.seed_t.copy_(snap[6])
                opt.zero_grad()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_loss = self._step()
        finally:
            ops.set_device_seed(None)

    def __call__(self, *inputs):
        """Copies the inputs into the graph's static buffers, replays the step, returns the (device) loss tensor."""
        for s, t in zip(self.static_in, inputs):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_loss
