"""Data-parallel gradient exchange: one process per GPU, the ONE exchange step of the path is a
mean all-reduce of the flat gradient arena, issued per bucket on a side stream as soon as the bucket's last gradient
has been accumulated (reverse layer order: decoder/head -> spectral stack -> spatial stack -> patch embedding), so it
overlaps the rest of backward.  The 1/world scaling and the reference's elementwise clamp (which acts on the FULL-batch
gradient, pretrain.py:71-73) are applied after the reduction inside the fused optimiser kernel
(FusedAdam(grad_scale=1/world, clamp=1)).  The reference itself has no distributed code.

Backend: torch.distributed (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def stage_boundaries(model):
    """Parameters that start a new all-reduce bucket so that a bucket never straddles two backward stages: every transformer stack
    (one autograd node: all of its gradients land together when the stack's backward returns) gets its own bucket, and so does
    whatever follows it in parameter order.  With equal-size buckets the bucket holding the patch embedding also holds part of the
    spatial stack, so that stack's gradient is exchanged only after the LAST backward kernel; cut at the stage edges it overlaps the
    patch-embedding backward and only the (small) embedding bucket is exposed."""
    order = list(model.parameters())
    pos = {id(p): i for i, p in enumerate(order)}
    cuts = set()
    for mod in model.modules():
        if hasattr(mod, "layer_params") and callable(mod.layer_params):
            idx = sorted(pos[id(p)] for p in mod.parameters())
            if idx:
                cuts.add(idx[0])
                if idx[-1] + 1 < len(order):
                    cuts.add(idx[-1] + 1)
    return [order[i] for i in sorted(cuts) if i > 0]


class GradSync:
    def __init__(self, arena, process_group=None, num_buckets=3, overlap=True, boundaries=None):
        """arena: maskedsst_b200.optim.FlatArena (FusedAdam(...).arena).  boundaries: optional parameters that each START a new bucket
        (see stage_boundaries); default: num_buckets roughly equal arena ranges."""
        self.arena = arena
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap
        self.is_cuda = arena.grads.is_cuda
        self.comm_stream = torch.cuda.Stream(device=arena.grads.device) if self.is_cuda else None
        # buckets = contiguous arena ranges cut at parameter boundaries, roughly equal sizes
        params = arena.param_list
        total = arena.grads.numel()
        target = max(1, total // max(1, num_buckets))
        self.buckets, lo, members = [], 0, []
        starts = None if boundaries is None else {id(p) for p in boundaries}
        for i, p in enumerate(params):
            off, n = arena.offsets[id(p)]
            members.append(p)
            end = (off + (n + 3) // 4 * 4)
            last = i == len(params) - 1
            if starts is not None:
                cut = last or id(params[i + 1]) in starts
            else:
                cut = last or (end - lo >= target and len(self.buckets) < num_buckets - 1)
            if cut:
                self.buckets.append({"lo": lo, "hi": total if last else end, "params": members, "expected": None,
                                     "pending": 0, "work": None, "launched": False})
                lo, members = end, []
        self._bucket_of = {id(p): b for b in self.buckets for p in b["params"]}
        self._fired = set()
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params if p.requires_grad]
        self._arm()

    def _arm(self):
        for b in self.buckets:
            b["pending"] = b["expected"] if b["expected"] is not None else -1   # -1: unknown until the first backward
            b["work"], b["launched"] = None, False
        self._fired = set()

    def _on_grad(self, p):
        if self.world == 1:
            return
        self._fired.add(id(p))
        b = self._bucket_of[id(p)]
        if b["pending"] > 0:
            b["pending"] -= 1
            if b["pending"] == 0 and self.overlap:
                self._launch(b)

    def _launch(self, b):
        if b["launched"]:
            return
        b["launched"] = True
        view = self.arena.grads[b["lo"]: b["hi"]]
        if self.is_cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                b["work"] = dist.all_reduce(view, group=self.group, async_op=True)
        else:
            b["work"] = dist.all_reduce(view, group=self.group, async_op=True)

    def finish(self):
        """Call after backward, before optimizer.step(): launches whatever has not been launched, waits for all buckets
        (the compute stream waits on the communication stream; the host does not block on GPUs)."""
        if self.world == 1:
            return
        for b in reversed(self.buckets):
            self._launch(b)
        for b in self.buckets:
            if b["work"] is not None:
                b["work"].wait()
        if self.is_cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        for b in self.buckets:   # learn which parameters take part in this graph (unused ones never fire, e.g. mlp_head in SimMIM)
            if b["expected"] is None:
                b["expected"] = sum(1 for p in b["params"] if id(p) in self._fired)
        self._arm()

    def remove(self):
        for h in self._hooks:
            h.remove()


def ce_dp_scale(cnt_local_sum_over_ranks, world):
    """nn.CrossEntropyLoss(ignore_index) averages over VALID pixels; with per-rank label sparsity the global-batch loss is
    sum_r nll_r / sum_r cnt_r.  Each rank therefore backpropagates nll_r * world / cnt_global, so that the 1/world mean of
    the all-reduced gradients equals the gradient of the global loss."""
    return world / cnt_local_sum_over_ranks


def cross_entropy_dp(logits, labels, ignore_index=-1, group=None):
    """Data-parallel-exact CE: all-reduces the valid-pixel count (one float)."""
    from . import ops
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ops.cross_entropy(logits, labels, ignore_index)
    nll_sum, cnt = ops.cross_entropy_sum_count(logits, labels, ignore_index)
    cnt_g = cnt.detach().clone()
    dist.all_reduce(cnt_g, group=group)
    return nll_sum * ce_dp_scale(cnt_g, dist.get_world_size(group))
