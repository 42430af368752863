"""Fused Adam / AdamW over a flat parameter arena (kernel (5), msst_adam_step).

Update rules are torch.optim.AdamW / torch.optim.Adam as the reference configures them (src/utils.py:36-44:
AdamW lr .008 wd .05 over ALL parameters; finetune.py:116-135: Adam with L2 weight decay and two lr groups), plus
the reference's elementwise gradient clamp (pretrain.py:71-73, `clip_grad_norm: True` == clamp to [-1, 1]) and the
data-parallel 1/world_size scaling folded into the same kernel.

At construction all parameters of a group are re-homed into one contiguous fp32 arena (the nn.Parameters become
views, so names / shapes / state_dict are unchanged) and their .grad tensors into a matching gradient arena:
one launch per group per step instead of ~10 launches per tensor x 178 tensors, and one flat buffer to all-reduce.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=True, clamp=0.0,
                 grad_scale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled, clamp=clamp)
        super().__init__(params, defaults)
        self.grad_scale = grad_scale
        self._flatten()

    # ---- arena construction ---------------------------------------------------------------------------------
    def _flatten(self):
        all_params = [p for g in self.param_groups for p in g["params"]]
        if not all_params:
            raise ValueError("FusedAdam: no parameters")
        dev = all_params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam needs CUDA parameters (move the model to the GPU first; there is no CPU path)")
        for p in all_params:
            if p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError("FusedAdam: all parameters must be fp32 on one CUDA device")
        # every tensor starts on a 16-byte boundary (float4 kernel, and bucket boundaries for all-reduce)
        offsets, total = [], 0
        self._group_ranges = []
        for g in self.param_groups:
            start = total
            for p in g["params"]:
                offsets.append(total)
                total += (p.numel() + 3) // 4 * 4
            self._group_ranges.append((start, total))
        self.param_arena = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad_arena = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.bf16_arena = None
        self._offsets = {}
        with torch.no_grad():
            for p, off in zip(all_params, offsets):
                n = p.numel()
                self.param_arena[off: off + n].copy_(p.detach().reshape(-1))
                old_grad = p.grad
                p.data = self.param_arena[off: off + n].view(p.shape)
                p.grad = self.grad_arena[off: off + n].view(p.shape)
                if old_grad is not None:
                    p.grad.copy_(old_grad)
                self._offsets[id(p)] = (off, n)
        self._step = 0

    def offset_of(self, p):
        return self._offsets[id(p)]

    def enable_bf16_copy(self):
        """Keeps a bf16 mirror of the arena up to date from inside the update kernel (operands of the tcgen05 GEMMs)."""
        if self.bf16_arena is None:
            self.bf16_arena = self.param_arena.to(torch.bfloat16)
        return self.bf16_arena

    # ---- torch.optim API --------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=False):
        """One memset; gradients stay views of the arena (set_to_none would break the flat layout)."""
        self.grad_arena.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._step += 1
        stream = torch.cuda.current_stream().cuda_stream
        lib = _lib.lib()
        for g, (a, b) in zip(self.param_groups, self._group_ranges):
            if b == a:
                continue
            args = _lib.AdamArgs(float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                 float(g["weight_decay"]), int(bool(g["decoupled"])), float(g["clamp"]),
                                 float(self.grad_scale), self._step)
            bf = None if self.bf16_arena is None else self.bf16_arena[a:b].data_ptr()
            check(lib.msst_adam_step(C.byref(args), self.param_arena[a:b].data_ptr(), self.grad_arena[a:b].data_ptr(),
                                     self.exp_avg[a:b].data_ptr(), self.exp_avg_sq[a:b].data_ptr(), bf, b - a, stream))
        return loss

    # optimizer state is three flat tensors + the step counter
    def state_dict(self):
        return {"step": self._step, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self._step = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)


def FusedAdamW(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **kw):
    return FusedAdam(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=True, **kw)
