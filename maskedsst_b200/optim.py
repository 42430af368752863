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


class FlatArena:
    """Re-homes parameters (and their .grad) into one contiguous fp32 buffer each; the nn.Parameters become views,
    so names / shapes / state_dict are unchanged.  Every tensor starts on a 16-byte boundary (float4 kernels and
    all-reduce bucket boundaries).  Device-agnostic (the data-parallel host logic is tested on CPU with gloo)."""

    def __init__(self, groups):
        flat = [p for g in groups for p in g]
        dev = flat[0].device
        for p in flat:
            if p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError("FlatArena: all parameters must be fp32 on one device")
        self.param_list = flat
        self.offsets, self.group_ranges, total = {}, [], 0
        for g in groups:
            start = total
            for p in g:
                self.offsets[id(p)] = (total, p.numel())
                total += (p.numel() + 3) // 4 * 4
            self.group_ranges.append((start, total))
        self.params = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grads = torch.zeros(total, device=dev, dtype=torch.float32)
        # direct-gradient protocol (ops._grad_buffers): `clean` = the gradient arena has been zeroed and no optimiser step has
        # consumed it since; `_claimed` = parameters whose arena view a backward kernel already accumulates into
        self.clean = False
        self._claimed = set()
        with torch.no_grad():
            for p in flat:
                off, n = self.offsets[id(p)]
                self.params[off: off + n].copy_(p.detach().reshape(-1))
                old_grad = p.grad
                p.data = self.params[off: off + n].view(p.shape)
                p.grad = self.grads[off: off + n].view(p.shape)
                if old_grad is not None:
                    p.grad.copy_(old_grad)


    def grad_view(self, p):
        off, n = self.offsets[id(p)]
        return self.grads[off: off + n].view(p.shape)

    def claim(self, p):
        """The arena view that a backward kernel may accumulate the gradient of parameter `p` into, or None (not ours, arena
        not freshly zeroed, or already handed out in this accumulation window -- a second writer plus autograd's own sum of
        the two returned views would double count)."""
        if not self.clean:
            return None
        ent = self.offsets.get(id(p))
        if ent is None or id(p) in self._claimed:
            return None
        self._claimed.add(id(p))
        off, n = ent
        return self.grads[off: off + n].view(p.shape)

    def mark_zeroed(self):
        self.clean = True
        self._claimed = set()

    def mark_consumed(self):
        self.clean = False


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=True, clamp=0.0,
                 grad_scale=1.0, capturable=False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled, clamp=clamp)
        super().__init__(params, defaults)
        self.grad_scale = grad_scale
        self._flatten()
        # capturable: the step counter lives on the device so step() can be recorded into a CUDA graph
        # (learning rates are still baked in at capture time: re-capture after a scheduler changes them)
        self._step_t = torch.zeros((), dtype=torch.int32, device=self.param_arena.device) if capturable else None

    # ---- arena construction ---------------------------------------------------------------------------------
    def _flatten(self):
        groups = [g["params"] for g in self.param_groups]
        if not any(groups):
            raise ValueError("FusedAdam: no parameters")
        dev = groups[0][0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam needs CUDA parameters (move the model to the GPU first; there is no CPU path)")
        self.arena = FlatArena(groups)
        self.param_arena, self.grad_arena = self.arena.params, self.arena.grads
        self._group_ranges = self.arena.group_ranges
        self._offsets = self.arena.offsets
        self.exp_avg = torch.zeros_like(self.param_arena)
        self.exp_avg_sq = torch.zeros_like(self.param_arena)
        self.bf16_arena = None
        self._step = 0
        # torch.optim.Adam(W) skips parameters whose .grad is None (e.g. encoder.mlp_head.* during SimMIM pre-training,
        # unused V1 merge layers): no moment update and, above all, NO weight decay.  In the arena every parameter has a
        # (zero) gradient view, so the set of parameters that actually received a gradient since zero_grad() is tracked
        # with post-accumulate hooks and only their arena ranges are updated.
        self._touched = {}
        self._range_cache = {}
        self._hooks = [p.register_post_accumulate_grad_hook(self._mark) for g in groups for p in g if p.requires_grad]
        from . import ops
        ops.register_grad_arena(self.arena)

    def _mark(self, p):
        self._touched[id(p)] = p
        # zero_grad() leaves p.grad = None, so autograd ADOPTS whatever tensor arrives first: an arena view when the fused
        # backward kernels wrote the gradient in place (nothing to do), otherwise a fresh tensor that is moved home here --
        # before GradSync's hook (registered later) may launch the all-reduce of this arena range
        g = p.grad
        off, n = self._offsets[id(p)]
        if g.data_ptr() != self.grad_arena.data_ptr() + 4 * off:
            view = self.grad_arena[off: off + n].view(p.shape)
            view.copy_(g)
            p.grad = view

    def mark_all_touched(self):
        """For callers that write gradients into the arena without autograd."""
        self._touched.update({id(p): p for g in self.param_groups for p in g["params"]})

    def _ranges(self, gi):
        """Contiguous arena ranges (within group gi) of the parameters that received a gradient."""
        key = (gi, frozenset(self._touched.keys()))
        r = self._range_cache.get(key)
        if r is None:
            r = []
            for p in self.param_groups[gi]["params"]:
                if id(p) not in self._touched:
                    continue
                off, n = self._offsets[id(p)]
                end = off + (n + 3) // 4 * 4
                if r and r[-1][1] == off:
                    r[-1][1] = end
                else:
                    r.append([off, end])
            if len(self._range_cache) > 64:
                self._range_cache.clear()
            self._range_cache[key] = r
        return r

    def _check_homes(self):
        """nn.Module.zero_grad() / p.grad = None re-allocates gradients outside the arena: training would silently stop."""
        gbase, pbase = self.grad_arena.data_ptr(), self.param_arena.data_ptr()
        for g in self.param_groups:
            ps = g["params"]
            for p in (ps[0], ps[-1]) if ps else ():
                off, n = self._offsets[id(p)]
                if (p.grad is not None and p.grad.data_ptr() != gbase + 4 * off) or p.data_ptr() != pbase + 4 * off:
                    raise RuntimeError("FusedAdam: a parameter or its .grad no longer lives in the flat arena (was "
                                       "model.to() called, or p.grad assigned by hand?).  Use optimizer.zero_grad().")
        if any(p.grad is None for p in self._touched.values()):
            raise RuntimeError("FusedAdam: a parameter received a gradient and then lost it (model.zero_grad() / p.grad = None between "
                               "backward and step?): its arena range would be stepped with stale data.  Use optimizer.zero_grad().")

    def offset_of(self, p):
        return self._offsets[id(p)]

    def enable_bf16_copy(self):
        """Keeps a bf16 mirror of the arena up to date from inside the update kernel (operands of the tcgen05 GEMMs)."""
        if self.bf16_arena is None:
            self.bf16_arena = self.param_arena.to(torch.bfloat16)
        return self.bf16_arena

    # ---- torch.optim API --------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        """One memset of the gradient arena.  With set_to_none (default, as torch.optim) p.grad becomes None: the next backward's
        kernels then accumulate straight into the arena and autograd adopts those views as p.grad -- no AccumulateGrad launch per
        parameter; a parameter that receives no gradient keeps p.grad = None and is skipped by step(), like torch.optim.  With
        set_to_none=False the gradients stay (zeroed) arena views and autograd adds into them."""
        self.grad_arena.zero_()
        self._touched = {}
        if set_to_none:
            for g in self.param_groups:
                for p in g["params"]:
                    p.grad = None
            self.arena.mark_zeroed()
        else:
            for g in self.param_groups:
                for p in g["params"]:
                    if p.grad is None:
                        p.grad = self.arena.grad_view(p)
            self.arena.mark_consumed()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._step += 1
        step_dev = None
        if self._step_t is not None:
            self._step_t += 1
            step_dev = self._step_t.data_ptr()
        stream = torch.cuda.current_stream().cuda_stream
        lib = _lib.lib()
        self._check_homes()
        self.arena.mark_consumed()
        for gi, g in enumerate(self.param_groups):
            for a, b in self._ranges(gi):
                args = _lib.AdamArgs(float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                     float(g["weight_decay"]), int(bool(g["decoupled"])), float(g["clamp"]),
                                     float(self.grad_scale), self._step, step_dev)
                bf = None if self.bf16_arena is None else self.bf16_arena[a:b].data_ptr()
                check(lib.msst_adam_step(C.byref(args), self.param_arena[a:b].data_ptr(), self.grad_arena[a:b].data_ptr(),
                                         self.exp_avg[a:b].data_ptr(), self.exp_avg_sq[a:b].data_ptr(), bf, b - a, stream))
        return loss

    # optimizer state is three flat tensors + the step counter (NOT a torch.optim state_dict: the moments are arena-shaped)
    def state_dict(self):
        step = int(self._step_t.item()) if self._step_t is not None else self._step   # graph replays advance only the device counter
        return {"format": "maskedsst_b200.FusedAdam/1", "step": step, "arena_numel": self.param_arena.numel(),
                "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        for k in ("step", "exp_avg", "exp_avg_sq", "param_groups"):
            if k not in sd:
                raise KeyError(f"FusedAdam.load_state_dict: missing key '{k}' (this is not a torch.optim.Adam state_dict: "
                               "the moments are stored as flat arena tensors)")
        if sd["exp_avg"].numel() != self.exp_avg.numel() or sd["exp_avg_sq"].numel() != self.exp_avg_sq.numel() or \
                len(sd["param_groups"]) != len(self.param_groups):
            raise ValueError("FusedAdam.load_state_dict: arena size / group count mismatch (different model or parameter order)")
        self._step = int(sd["step"])
        if self._step_t is not None:
            self._step_t.fill_(self._step)
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)


def FusedAdamW(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **kw):
    return FusedAdam(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=True, **kw)
