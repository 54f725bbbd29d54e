"""CPU: host-side pieces either side of the denoising path against the unmodified reference:
train_batch control flow (trainer.py:13-97), checkpoint files (unet.py:794-832, model_ema.py:36-55) and the
gradient-adoption contract of the flat arena. The reference tree exists only in the build container: tests that
need it skip elsewhere."""
import argparse
import copy
import os
import sys

import pytest
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "ml-mdm_b200"))
import refharness as rh  # noqa: E402


# ------------------------------------------------------------------ train_batch
class _Vision(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.a = nn.Linear(6, 8)
        self.b = nn.Linear(8, 6)

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


class _Inner(nn.Module):
    def __init__(self):
        super().__init__()
        self.vision_model = _Vision()


class _Pipe(nn.Module):
    """What train_batch touches of a Diffusion pipeline: .model.vision_model and .get_loss(sample)."""

    def __init__(self, weighted):
        super().__init__()
        self.model = _Inner()
        self.weighted = weighted

    def get_loss(self, sample):
        x = sample["x"]
        pred = self.model.vision_model(x)
        losses = ((pred - sample["y"]) ** 2).mean(dim=1)
        weights = sample["w"] if self.weighted else None
        return losses, torch.zeros(x.shape[0]), x, pred, sample["y"], weights


def _run(train_batch, ModelEma, weighted):
    pipe = _Pipe(weighted)
    opt = torch.optim.Adam(pipe.model.vision_model.parameters(), lr=1e-2, eps=1e-8)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    ema = ModelEma(pipe.model.vision_model, decay=0.9, warmup_steps=1)
    args = argparse.Namespace(fp16=False, gradient_clip_norm=0.05)
    g = torch.Generator().manual_seed(1)
    out = []
    for step in range(5):
        s = {"x": torch.randn(4, 6, generator=g), "y": torch.randn(4, 6, generator=g), "w": torch.rand(4, generator=g)}
        if step == 2:
            s["x"][0, 0] = float("nan")          # NaN loss: trainer.py:69-74
        accumulate = step == 3                  # no optimizer step / zero_grad on this one
        r = train_batch(pipe, s, opt, sched, None, args, accumulate_gradient=accumulate,
                        num_grad_accumulations=2 if accumulate else 1, ema_model=ema)
        out.append((r[0], sched.get_last_lr()[0]))
    return pipe, ema, out


@pytest.mark.parametrize("weighted", [False, True])
def test_train_batch_mirrors_reference_control_flow(weighted):
    if not rh.available():
        pytest.skip("reference tree not mounted")
    rh.load()
    from ml_mdm import trainer as ref_trainer
    from ml_mdm.models.model_ema import ModelEma

    from mdm_b200 import trainer as my_trainer

    pa, ea, oa = _run(ref_trainer.train_batch, ModelEma, weighted)
    pb, eb, ob = _run(my_trainer.train_batch, ModelEma, weighted)
    for (la, lra), (lb, lrb) in zip(oa, ob):
        assert (la == lb or (la != la and lb != lb)) and lra == lrb
    for (k, a), (_, b) in zip(pa.state_dict().items(), pb.state_dict().items()):
        assert torch.equal(a, b), k
    for (k, a), (_, b) in zip(ea.module.state_dict().items(), eb.module.state_dict().items()):
        assert torch.equal(a, b), k
    assert ea.counter == eb.counter


def _run_fp16(train_batch, ModelEma):
    """args.fp16 branch (trainer.py:29-61) with a disabled GradScaler (bf16 autocast needs no loss scaling; CPU has no
    CUDA scaler): loss_factor, pre-backward accumulation divide, NaN path without optimizer/scheduler step."""
    pipe = _Pipe(True)
    opt = torch.optim.Adam(pipe.model.vision_model.parameters(), lr=1e-2, eps=1e-8)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    ema = ModelEma(pipe.model.vision_model, decay=0.9, warmup_steps=1)
    args = argparse.Namespace(fp16=True, gradient_clip_norm=0.05)
    scaler = torch.amp.GradScaler("cuda", enabled=False)
    g = torch.Generator().manual_seed(1)
    out = []
    for step in range(6):
        s = {"x": torch.randn(4, 6, generator=g), "y": torch.randn(4, 6, generator=g), "w": torch.rand(4, generator=g)}
        if step == 2:
            s["x"][0, 0] = float("nan")
        accumulate = step == 3
        r = train_batch(pipe, s, opt, sched, None, args, grad_scaler=scaler, accumulate_gradient=accumulate,
                        num_grad_accumulations=2 if step in (3, 4) else 1, ema_model=ema, loss_factor=0.5)
        out.append((r[0], sched.get_last_lr()[0]))
    return pipe, ema, out


def test_train_batch_fp16_branch_mirrors_reference_control_flow():
    if not rh.available():
        pytest.skip("reference tree not mounted")
    rh.load()
    import warnings

    from ml_mdm import trainer as ref_trainer
    from ml_mdm.models.model_ema import ModelEma

    from mdm_b200 import trainer as my_trainer

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # torch.cuda.amp.autocast on a CPU-only host warns and disables itself
        pa, ea, oa = _run_fp16(ref_trainer.train_batch, ModelEma)
    pb, eb, ob = _run_fp16(my_trainer.train_batch, ModelEma)
    for (la, lra), (lb, lrb) in zip(oa, ob):
        assert (la == lb or (la != la and lb != lb)) and lra == lrb
    for (k, a), (_, b) in zip(pa.state_dict().items(), pb.state_dict().items()):
        assert torch.equal(a, b), k
    for (k, a), (_, b) in zip(ea.module.state_dict().items(), eb.module.state_dict().items()):
        assert torch.equal(a, b), k
    assert ea.counter == eb.counter


# ------------------------------------------------------------------ checkpoints
@pytest.mark.parametrize("kind", ["unet", "nested"])
def test_checkpoint_files_interchange_with_reference(kind, tmp_path):
    if not rh.available():
        pytest.skip("reference tree not mounted")
    import tiny_configs as tc
    from mdm_b200 import config as mc
    from mdm_b200.models import NestedUNet, UNet
    ref = rh.load()
    from ml_mdm.models.model_ema import ModelEma

    ucfg_d = copy.deepcopy(tc.TINY_UNET if kind == "unet" else tc.TINY_NESTED)
    arch = "unet" if kind == "unet" else "nested_unet"
    ref_model, _ = rh.build(copy.deepcopy(ucfg_d), {}, arch, tc.LM_DIM)
    cfg = mc.unet_config_from_dict(copy.deepcopy(ucfg_d))
    cfg.conditioning_feature_dim = tc.LM_DIM
    mine = (UNet if kind == "unet" else NestedUNet)(3, 3, cfg)
    # reference -> ours
    with torch.no_grad():
        for p in ref_model.parameters():
            p.normal_(0, 0.3)
    f1 = str(tmp_path / "vis_model_ref.pth")
    ref_model.save(f1, other_items={"batch_num": 41})
    items = mine.load(f1)
    assert items["batch_num"] == 41
    for (k, a), (k2, b) in zip(ref_model.state_dict().items(), mine.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k
    # ours -> reference (model file and the EMA file pair written by the reference's ModelEma around our module)
    with torch.no_grad():
        for p in mine.parameters():
            p.mul_(1.5)
    f2 = str(tmp_path / "vis_model_mine.pth")
    mine.save(f2, other_items={"batch_num": 42, "args": {"lr": 1e-4}})
    items = ref_model.load(f2)
    assert items["batch_num"] == 42 and items["args"] == {"lr": 1e-4}
    for (k, a), (_, b) in zip(mine.state_dict().items(), ref_model.state_dict().items()):
        assert torch.equal(a, b), k
    ema = ModelEma(mine, decay=0.5)
    ema.update(mine)
    f3 = str(tmp_path / "ema.pth")
    ema.save(f3, other_items={"batch_num": 42})
    ema_ref = ModelEma(ref_model, decay=0.5)
    ema_ref.load(f3)
    for (k, a), (_, b) in zip(ema.module.state_dict().items(), ema_ref.module.state_dict().items()):
        assert torch.equal(a, b), k
    ck = torch.load(f2, map_location="cpu")
    assert set(ck) == {"state_dict", "batch_num", "args"}


# ------------------------------------------------------------------ gradient adoption
def test_arena_views_are_adopted_not_cloned():
    """`.grad` must alias the flat arena after backward (one all-reduce, FusedAdam, GradientOverlap depend on it).
    autograd adopts an incoming gradient only if nothing else references it."""
    from mdm_b200.models.native import ARENA_ALIGN, _pad, arena_views

    ps = [nn.Parameter(torch.randn(3, 4)), nn.Parameter(torch.randn(5)), nn.Parameter(torch.randn(2, 2, 3))]
    offs, total = [], 0
    for p in ps:
        offs.append(total)
        total += _pad(p.numel())
    assert all(o % ARENA_ALIGN == 0 for o in offs)
    arena = torch.zeros(total)

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, *params):
            return x.sum() + 0 * sum(p.sum() for p in params)

        @staticmethod
        def backward(ctx, g):
            arena.fill_(2.0)
            return (None, *arena_views(arena, ps, offs))

    Fn.apply(torch.randn(3, requires_grad=True), *ps).backward()
    lo, hi = arena.data_ptr(), arena.data_ptr() + 4 * arena.numel()
    assert all(lo <= p.grad.data_ptr() < hi for p in ps)
    assert all(p.grad.shape == p.shape and float(p.grad.min()) == 2.0 for p in ps)
