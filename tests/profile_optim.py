"""Device time of the fused post-backward sweep on the full cc12m_64x64 model (461 M parameters) next to the
reference's separate calls (clip_grad_norm_ + torch Adam + ModelEma.update + zero_grad) on the same tensors
(development aid; the numbers are quoted in DESIGN.md)."""
import copy
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from mdm_b200 import optim  # noqa: E402


class Ema:
    def __init__(self, model, decay=0.9999, warmup_steps=0):
        self.module = copy.deepcopy(model)
        self.decay, self.warmup_steps, self.counter = decay, warmup_steps, 0

    def update(self, model):  # ml_mdm/models/model_ema.py:25-34
        decay = (self.counter >= self.warmup_steps) * self.decay
        self.counter += 1
        with torch.no_grad():
            msd = model.state_dict()
            for k, ema_v in self.module.state_dict().items():
                ema_v.mul_(decay).add_(msd[k].detach(), alpha=1.0 - decay)


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main(iters=5):
    dev = torch.device("cuda", 0)
    pipe, _ = bench.build_pipeline("cc12m_64x64", dev)
    pipe.train()
    vm = pipe.get_model().vision_model
    sample = {k: v.to(dev) for k, v in bench.synthetic_host_batch("cc12m_64x64", 4, 1).items()}
    loss, *_ = pipe.get_loss(sample)
    loss.mean().backward()
    n = sum(p.numel() for p in vm.parameters())
    ema = Ema(vm)
    opt = optim.FusedAdam(vm, lr=1e-4)
    fused = timed(lambda: opt.step(max_grad_norm=2.0, ema_model=ema), iters)
    # the reference's sequence on the same tensors (gradients are whatever the arena holds: traffic is what counts)
    ref_opt = torch.optim.Adam(vm.parameters(), lr=1e-4, eps=1e-8)

    def separate():
        torch.nn.utils.clip_grad_norm_(vm.parameters(), 2.0)
        ref_opt.step()
        ema.update(vm)
        for p in vm.parameters():
            p.grad.zero_()

    sep = timed(separate, iters)
    gb = n * 44 / 1e9  # 4 B norm read + 20 B read + 20 B written per parameter
    print(f"params {n/1e6:.1f} M: fused sweep {fused:.2f} ms ({gb/fused*1e3/1e3:.2f} TB/s of {gb:.1f} GB algorithmic) | "
          f"clip_grad_norm_ + torch Adam + EMA + zero_grad {sep:.2f} ms -> x{sep/fused:.1f}", flush=True)


if __name__ == "__main__":
    main()
