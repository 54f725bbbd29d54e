"""Generates tests/golden/optim_steps.npz from torch's own Adam/AdamW + clip_grad_norm_ and the UNMODIFIED
reference ModelEma (imported from /root/reference, which exists only in the build container):
3 optimisation steps on 4 small tensors, clipping active, EMA warm-up of one step.
Run: python tests/golden/make_golden_optim.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/ml-mdm-matryoshka"
sys.path.insert(0, REF)
# ModelEma imports ml_mdm.utils.fix_old_checkpoints -> keep the import light
for name in ("ml_mdm", "ml_mdm.utils", "ml_mdm.utils.fix_old_checkpoints"):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
sys.modules["ml_mdm.utils"].fix_old_checkpoints = sys.modules["ml_mdm.utils.fix_old_checkpoints"]
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_model_ema", os.path.join(REF, "ml_mdm/models/model_ema.py"))
ref_ema = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_ema)

SHAPES = [(3,), (64,), (1000,), (257, 5)]


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s, generator=g)) for s in SHAPES])


def run(adamw, wd):
    torch.manual_seed(0)
    model = Tiny()
    ema = ref_ema.ModelEma(model, decay=0.9, warmup_steps=1)
    opt = (torch.optim.AdamW if adamw else torch.optim.Adam)(model.parameters(), lr=3e-3, eps=1e-8, weight_decay=wd)
    g = torch.Generator().manual_seed(9)
    out = {"p0": [p.detach().clone().numpy() for p in model.ps]}
    for step in range(3):
        grads = [torch.randn(s, generator=g) * (3.0 if step == 1 else 0.05) for s in SHAPES]  # step 1 clips hard
        for p, gr in zip(model.ps, grads):
            p.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), 2.0)
        opt.step()
        ema.update(model)
        opt.zero_grad()
        out[f"g{step}"] = [x.numpy() for x in grads]
        out[f"norm{step}"] = total.numpy()
        out[f"p{step + 1}"] = [p.detach().clone().numpy() for p in model.ps]
        out[f"ema{step + 1}"] = [p.detach().clone().numpy() for p in ema.module.ps]
        out[f"m{step + 1}"] = [opt.state[p]["exp_avg"].clone().numpy() for p in model.ps]
        out[f"v{step + 1}"] = [opt.state[p]["exp_avg_sq"].clone().numpy() for p in model.ps]
    return out


flat = {}
for tag, (adamw, wd) in {"adam": (False, 0.0), "adamw": (True, 0.01), "adam_l2": (False, 0.01)}.items():
    for k, v in run(adamw, wd).items():
        if isinstance(v, list):
            for i, a in enumerate(v):
                flat[f"{tag}/{k}/{i}"] = a
        else:
            flat[f"{tag}/{k}"] = v
np.savez_compressed(os.path.join(HERE, "optim_steps.npz"), **flat)
print("wrote optim_steps.npz with", len(flat), "arrays")
