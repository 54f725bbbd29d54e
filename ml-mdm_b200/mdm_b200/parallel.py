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
        dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=group)
        if average:
            arena.div_(world)
        return 1
    bucket = [p.grad for p in module.parameters() if p.grad is not None]
    return allreduce_tensors(bucket, group=group, average=average)


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
