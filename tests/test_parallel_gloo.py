"""world_size-2 CPU (gloo) test of the data-parallel host logic: batch sharding and the single
gradient all-reduce (mdm_b200/parallel.py)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "ml-mdm_b200"))
    from mdm_b200 import parallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sample = {"images": torch.arange(8.0).view(8, 1), "tag": "x"}
        sh = parallel.shard_batch(sample, rank, world)
        assert sh["images"].flatten().tolist() == [4.0 * rank + i for i in range(4)] and sh["tag"] == "x"

        # a fake "native" module: gradients living in one flat arena, as models/native.py arranges
        class Fake:
            pass

        lin = torch.nn.Linear(3, 2)
        arena = torch.zeros(sum(p.numel() for p in lin.parameters()))
        off = 0
        for p in lin.parameters():
            p.grad = arena[off:off + p.numel()].view_as(p)
            off += p.numel()
        arena.fill_(float(rank + 1))
        nat = Fake()
        nat.grad_arena, nat.params = arena, list(lin.parameters())
        lin._native = nat
        calls = parallel.allreduce_gradients(lin)
        assert calls == 1
        assert torch.allclose(arena, torch.full_like(arena, 1.5))  # mean of 1 and 2
        assert torch.allclose(lin.weight.grad, torch.full((2, 3), 1.5))
        # gradients outside the arena take the flatten path
        lin2 = torch.nn.Linear(2, 2)
        for p in lin2.parameters():
            p.grad = torch.full_like(p, float(rank))
        parallel.allreduce_gradients(lin2)
        assert torch.allclose(lin2.bias.grad, torch.full((2,), 0.5))
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
