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
        # overlapped reduction: the engine reports final address ranges from the top of the arena downwards
        # (simulated here); finish() must cover the rest exactly once
        class FakeNative(Fake):
            def set_grad_ready(self, fn, min_bytes=0):
                self.fn = fn

        lin3 = torch.nn.Linear(5, 4)
        n3 = sum(p.numel() for p in lin3.parameters())
        arena3 = torch.zeros(n3)
        off = 0
        for p in lin3.parameters():
            p.grad = arena3[off:off + p.numel()].view_as(p)
            off += p.numel()
        nat3 = FakeNative()
        nat3.grad_arena, nat3.active_arena, nat3.params = arena3, arena3, list(lin3.parameters())
        lin3._native = nat3
        ov = parallel.GradientOverlap(lin3, bucket_mb=1)
        assert ov.enabled and nat3.fn is not None
        base = arena3.data_ptr()
        for step_i in range(2):
            arena3.copy_(torch.arange(n3, dtype=torch.float32) * (rank + 1))
            if step_i == 1:
                ov.arm()
                nat3.fn(base + 4 * 16, base + 4 * n3)   # [16, n3) final
                nat3.fn(base + 4 * 6, base + 4 * 16)    # [6, 16) final
            # step 0: no reports at all (the engine's learning step)
            ncalls = ov.finish()
            assert ncalls == (3 if step_i == 1 else 1), ncalls
            assert torch.allclose(arena3, torch.arange(n3, dtype=torch.float32) * 1.5), arena3
        assert ov.segments == [(16, n3), (6, 16), (0, 6)], ov.segments
        # an un-armed backward (accumulation step) must not start any collective
        nat3.fn(base, base + 4 * n3)
        assert ov.works == [] and ov.low is None
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
