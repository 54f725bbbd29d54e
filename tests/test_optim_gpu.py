"""B200: the fused clip + Adam/AdamW + EMA + zero-grad sweep (mdm_grad_norm, mdm_adam_ema_sweep through the C ABI)
against the golden fixture of torch's Adam/AdamW + clip_grad_norm_ + the reference ModelEma, and
mdm_b200.optim.FusedAdam / trainer.train_batch against the reference's separate calls on a tiny U-Net.

Tolerance: fp32 formulas restated op for op, differing only in fused-multiply-add contraction and the reduction
order of the norm: |delta| <= 2e-6 * |ref| + 1e-7 per element (a few ulps), norm 1e-6 relative."""
import ctypes as C
import copy
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "ml-mdm_b200"))
sys.path.insert(0, HERE)
from mdm_b200 import _lib, optim  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(HERE, "golden", "optim_steps.npz"))
VARIANTS = {"adam": (False, 0.0), "adamw": (True, 0.01), "adam_l2": (False, 0.01)}
NT = 4
RTOL, ATOL = 2e-6, 1e-7


def close(a, b):
    return bool(((a.double().cpu() - b.double()).abs() <= RTOL * b.double().abs() + ATOL).all())


def gold_list(tag, key):
    return [torch.from_numpy(GOLD[f"{tag}/{key}/{i}"].copy()) for i in range(NT)]


@pytest.mark.parametrize("tag", list(VARIANTS))
def test_sweep_kernels_vs_torch_golden(tag):
    adamw, wd = VARIANTS[tag]
    dev = "cuda"
    lib = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ps = [p.to(dev) for p in gold_list(tag, "p0")]
    emas = [p.clone() for p in ps]
    pad = lambda n: (n + 63) // 64 * 64
    offs, total = [], 0
    for p in ps:
        offs.append(total)
        total += pad(p.numel())
    arena = torch.zeros(total, device=dev)
    m = torch.zeros(total, device=dev)
    v = torch.zeros(total, device=dev)
    rows = []
    for p, e, off in zip(ps, emas, offs):
        rows += optim.chunk_rows(p.data_ptr(), arena.data_ptr() + 4 * off, m.data_ptr() + 4 * off,
                                 v.data_ptr() + 4 * off, e.data_ptr(), p.numel(), chunk=300)  # several chunks, odd tails
    table = torch.tensor(rows, dtype=torch.int64).reshape(-1, 6).to(dev)
    scratch = torch.zeros(optim.GRAD_NORM_SCRATCH, device=dev, dtype=torch.float64)
    norm = torch.zeros(1, device=dev)
    for step in range(3):
        for g, p, off in zip(gold_list(tag, f"g{step}"), ps, offs):
            arena[off:off + p.numel()] = g.to(dev).flatten()
        _lib.check(lib.mdm_grad_norm(C.c_void_p(arena.data_ptr()), C.c_int64(total), C.c_float(1.0),
                                     C.c_void_p(scratch.data_ptr()), C.c_int32(optim.GRAD_NORM_SCRATCH),
                                     C.c_void_p(norm.data_ptr()), st), "mdm_grad_norm")
        cfg = optim.AdamCfg()
        cfg.lr, cfg.beta1, cfg.beta2, cfg.eps, cfg.weight_decay = 3e-3, 0.9, 0.999, 1e-8, wd
        cfg.adamw, cfg.step, cfg.grad_scale, cfg.max_norm, cfg.zero_grad = int(adamw), step + 1, 1.0, 2.0, 1
        cfg.ema_decay = float(step >= 1) * 0.9
        _lib.check(lib.mdm_adam_ema_sweep(C.c_void_p(table.data_ptr()), C.c_int32(table.shape[0]), C.byref(cfg),
                                          C.c_void_p(norm.data_ptr()), st), "mdm_adam_ema_sweep")
        torch.cuda.synchronize()
        ref_norm = float(GOLD[f"{tag}/norm{step}"])
        assert abs(float(norm) - ref_norm) <= 1e-6 * ref_norm
        assert float(arena.abs().max()) == 0.0, "gradients must be left zeroed"
        ms = [m[off:off + p.numel()].view_as(p) for p, off in zip(ps, offs)]
        vs = [v[off:off + p.numel()].view_as(p) for p, off in zip(ps, offs)]
        for name, mine in (("p", ps), ("ema", emas), ("m", ms), ("v", vs)):
            for i, (a, b) in enumerate(zip(mine, gold_list(tag, f"{name}{step + 1}"))):
                assert close(a, b), f"{tag} step {step} {name}[{i}]: max diff {float((a.cpu() - b).abs().max()):.3e}"


class _Ema:  # the attributes of ml_mdm.models.model_ema.ModelEma that the sweep uses
    def __init__(self, model, decay, warmup_steps):
        self.module = copy.deepcopy(model)
        self.decay, self.warmup_steps, self.counter = decay, warmup_steps, 0

    def update(self, model):  # model_ema.py:25-34
        decay = (self.counter >= self.warmup_steps) * self.decay
        self.counter += 1
        with torch.no_grad():
            msd = model.state_dict()
            for k, ema_v in self.module.state_dict().items():
                ema_v.mul_(decay).add_(msd[k].detach(), alpha=1.0 - decay)


def test_fused_adam_on_tiny_unet_matches_separate_calls():
    """Two optimisation steps of a tiny U-Net: FusedAdam.step(clip, ema) on the engine's gradient arena vs
    clip_grad_norm_ + torch Adam + ModelEma.update on a copy that is handed the very same gradients (Adam divides
    by sqrt(v): feeding both sides one set of gradients keeps round-off-sized gradients from deciding signs)."""
    import net_cases as nc
    import tiny_configs as tc
    model_a, _, _ = nc.build("unet")
    model_a = model_a.cuda()
    model_b = copy.deepcopy(model_a)
    x, t, lm, mask = tc.seeded_inputs(3, 2, 16, 6, nlevels=1)
    x, t, lm, mask = x.cuda(), t.cuda(), lm.cuda(), mask.cuda()
    opt_a = optim.FusedAdam(model_a, lr=2e-3)
    opt_b = torch.optim.Adam(model_b.parameters(), lr=2e-3, eps=1e-8)
    ema_a, ema_b = _Ema(model_a, 0.9, 1), _Ema(model_b, 0.9, 1)
    for step in range(2):
        out = model_a(x, t, lm, mask, {})
        ((out - 0.3 * x) ** 2).mean().backward()
        for pa, pb in zip(model_a.parameters(), model_b.parameters()):
            pb.grad = pa.grad.detach().clone()
        total_b = torch.nn.utils.clip_grad_norm_(model_b.parameters(), 0.05)
        opt_b.step()
        ema_b.update(model_b)
        opt_b.zero_grad()
        opt_a.step(max_grad_norm=0.05, ema_model=ema_a)
        assert all(float(p.grad.abs().max()) == 0.0 for p in model_a.parameters()), "arena must be left zeroed"
        opt_a.zero_grad()
        torch.cuda.synchronize()
        assert abs(float(opt_a.last_grad_norm) - float(total_b)) <= 2e-6 * float(total_b)
    sa, sb = model_a.state_dict(), model_b.state_dict()
    for k in sb:
        assert close(sa[k], sb[k].cpu()), f"{k}: {float((sa[k] - sb[k]).abs().max()):.3e}"
    for (k, ea), (_, eb) in zip(ema_a.module.state_dict().items(), ema_b.module.state_dict().items()):
        assert close(ea, eb.cpu()), f"ema {k}: {float((ea - eb).abs().max()):.3e}"
    assert ema_a.counter == ema_b.counter == 2
    st = opt_a.state_dict()["state"]
    assert len(st) == len(list(model_a.parameters())) and float(st[0]["step"]) == 2.0
    # and the engine picks the updated weights up: a third forward differs from the first
    out3 = model_a(x, t, lm, mask, {})
    assert float((out3 - out).abs().max()) > 0
