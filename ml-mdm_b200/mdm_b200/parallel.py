"""Data parallelism of the path: one process per GPU, batch sharded by rank, and ONE gradient
all-reduce per optimisation step (reference: DistributedDataParallel around diffusion_model.model,
clis/train_parallel.py:147-154; no other collective exists in the reference).

The engine writes all parameter gradients of a step into one flat fp32 arena
(models/native.py), so the all-reduce is a single NCCL call over NVLink/NVSwitch on ~1.9 GB instead
of DDP's 25 MB buckets; `no_sync`-style accumulation is obtained by simply not calling it.
torch DDP also works unchanged on `pipeline.model` because gradients arrive through autograd.
"""
import torch
import torch.distributed as dist


def flat_grads(module):
    """The flat gradient arena if every parameter gradient currently lives in it, else None."""
    native = getattr(module, "_native", None)
    if native is None or native.grad_arena is None:
        return None
    arena = native.grad_arena
    lo, hi = arena.data_ptr(), arena.data_ptr() + arena.numel() * arena.element_size()
    for p in native.params:
        if p.requires_grad and (p.grad is None or not (lo <= p.grad.data_ptr() < hi)):
            return None
    return arena


def allreduce_gradients(module, group=None, average=True):
    """Sum (and average) gradients across ranks with one collective when possible."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    arena = flat_grads(module)
    if arena is not None:
        if average and dist.get_backend(group) == "nccl":   # NCCL averages in the collective: no extra 1.8 GB pass
            dist.all_reduce(arena, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=group)
            if average:
                arena.div_(world)
        return 1
    bucket = [p.grad for p in module.parameters() if p.grad is not None]
    return allreduce_tensors(bucket, group=group, average=average)


class GradientOverlap:
    """Gradient all-reduce overlapped with backward (what DDP's bucket hooks do for the reference,
    clis/train_parallel.py:147-154). The engine reports, while it is still enqueuing backward kernels,
    every further `bucket_mb` of the flat gradient arena that has received its last write
    (mdm_net_set_grad_ready, include/mdm_b200.h); each report starts an asynchronous all-reduce of that
    slice, which NCCL orders after the kernels enqueued so far and runs on its own stream next to the
    rest of backward.

        overlap = GradientOverlap(pipeline.get_model().vision_model)   # once
        overlap.arm(); loss.backward(); overlap.finish()               # every synchronising step
    (a backward without arm() is left alone, which is the `no_sync` accumulation case)

    finish() reduces whatever was not reported (everything, on the first step, while the engine learns
    which closure touches which parameter) and makes the current stream wait for all collectives."""

    def __init__(self, module, group=None, average=True, bucket_mb=64, sm_reserve=0):
        """sm_reserve: SMs the engine's persistent GEMM kernels leave free for the collective's CTAs
        (mdm_set_sm_reserve); pair it with NCCL_MAX_CTAS=<sm_reserve> in the environment before the process group is
        created, otherwise NCCL's CTAs and the one-CTA-per-SM GEMM queue behind each other."""
        self.module, self.group, self.average = module, group, average
        self.sm_reserve = int(sm_reserve)
        self.works, self.low = [], None
        self.armed = False
        self.laid_out = False
        self.segments = []  # (lo, hi) element ranges reduced during the last step, in issue order
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        native = module.native() if hasattr(module, "native") else getattr(module, "_native", None)
        self.native = native
        if self.enabled and native is not None:
            native.set_grad_ready(self._on_ready, int(bucket_mb) << 20)
            if self.sm_reserve > 0:
                native.lib.mdm_set_sm_reserve(self.sm_reserve)

    def close(self):
        if self.native is not None:
            self.native.set_grad_ready(None)
            if self.sm_reserve > 0:
                self.native.lib.mdm_set_sm_reserve(0)

    def _reduce(self, t, async_op):
        world = dist.get_world_size(self.group)
        if self.average and dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if self.average:
            if w is not None:
                w.wait()
            t.div_(world)
        return w

    def arm(self):
        """The next backward's gradients are to be reduced (call right before loss.backward())."""
        self.armed = True
        self.works, self.low, self.segments = [], None, []

    def _on_ready(self, lo_ptr, hi_ptr, arena=None):
        if not self.armed:
            return
        if arena is None:
            arena = self.native.active_arena
            if arena is not self.native.grad_arena:
                return  # gradients are being accumulated into existing .grad tensors: reduce at the end
        base, n = arena.data_ptr(), arena.numel()
        lo = min(max((lo_ptr - base) // arena.element_size(), 0), n)
        hi = min(max((hi_ptr - base) // arena.element_size(), 0), n)
        if self.low is None:
            self.segments = []
        if hi > lo:
            self.works.append(self._reduce(arena[lo:hi], async_op=True))
            self.segments.append((lo, hi))
        self.low = lo if self.low is None else min(self.low, lo)

    def finish(self, arena=None):
        """Reduce the part of the arena no report covered, then join the collectives."""
        if not self.enabled:
            return 0
        self.armed = False
        if arena is None:
            arena = flat_grads(self.module)
        if arena is None:  # gradients are not in the arena (accumulation): plain path
            self.works, self.low = [], None
            return allreduce_gradients(self.module, group=self.group, average=self.average)
        rest = arena.numel() if self.low is None else self.low
        if self.low is None:
            self.segments = []
        if rest > 0:
            self.works.append(self._reduce(arena[:rest], async_op=True))
            self.segments.append((0, rest))
        for w in self.works:
            if w is not None:
                w.wait()
        n = len(self.works)
        self.works, self.low = [], None
        if not self.laid_out and self.native is not None and hasattr(self.native, "optimize_arena_layout"):
            # after the first backward the engine knows when each gradient becomes final
            self.laid_out = bool(self.native.optimize_arena_layout())
        return n


def allreduce_tensors(tensors, group=None, average=True):
    """Fallback for gradients that are not in the arena (e.g. accumulation mode): flatten once."""
    if not tensors:
        return 0
    world = dist.get_world_size(group)
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(world)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return 1


def shard_batch(sample: dict, rank: int, world: int) -> dict:
    """Rank r takes rows [r*B/W, (r+1)*B/W) of every batched tensor (reference: dataset.partition,
    reader.py:192-193)."""
    out = {}
    for k, v in sample.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] % world == 0:
            n = v.shape[0] // world
            out[k] = v[rank * n:(rank + 1) * n]
        else:
            out[k] = v
    return out
