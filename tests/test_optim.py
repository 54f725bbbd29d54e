"""CPU: the optimiser-sweep oracle (oracle/optim_ref.py) against the golden fixture produced by torch's own
Adam/AdamW + clip_grad_norm_ and the unmodified reference ModelEma (tests/golden/make_golden_optim.py), and the
host-side chunk table logic of mdm_b200.optim."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "ml-mdm_b200"))
from oracle import optim_ref  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "optim_steps.npz"))
VARIANTS = {"adam": (False, 0.0), "adamw": (True, 0.01), "adam_l2": (False, 0.01)}
NT = 4


def gold_list(tag, key):
    return [torch.from_numpy(GOLD[f"{tag}/{key}/{i}"].copy()) for i in range(NT)]


@pytest.mark.parametrize("tag", list(VARIANTS))
def test_oracle_reproduces_torch_adam_clip_and_reference_ema(tag):
    adamw, wd = VARIANTS[tag]
    ps = gold_list(tag, "p0")
    emas = [p.clone() for p in ps]      # ModelEma deep-copies the model
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    for step in range(3):
        grads = gold_list(tag, f"g{step}")
        decay = float(step >= 1) * 0.9   # warmup_steps = 1 (model_ema.py:26)
        total = optim_ref.sweep_(ps, grads, ms, vs, emas, step + 1, 3e-3, (0.9, 0.999), 1e-8, wd, adamw,
                                 max_norm=2.0, ema_decay=decay)
        assert float(total) == float(GOLD[f"{tag}/norm{step}"])
        for name, mine in (("p", ps), ("ema", emas), ("m", ms), ("v", vs)):
            for a, b in zip(mine, gold_list(tag, f"{name}{step + 1}")):
                assert torch.equal(a, b), f"{tag} step {step} {name}: max diff {float((a - b).abs().max())}"
        assert all(float(g.abs().max()) == 0.0 for g in grads)  # zero_grad


def test_chunk_rows_cover_a_tensor_exactly():
    from mdm_b200 import optim
    rows = optim.chunk_rows(1000, 2000, 3000, 4000, 0, 70000, chunk=32768)
    assert [r[5] for r in rows] == [32768, 32768, 4464]
    assert rows[1][:5] == (1000 + 4 * 32768, 2000 + 4 * 32768, 3000 + 4 * 32768, 4000 + 4 * 32768, 0)
    rows = optim.chunk_rows(16, 32, 48, 64, 80, 5)
    assert rows == [(16, 32, 48, 64, 80, 5)]
    # struct mirrors stay in step with include/mdm_b200.h
    import ctypes as C
    assert C.sizeof(optim.OptChunk) == 48 and C.sizeof(optim.AdamCfg) == 72
