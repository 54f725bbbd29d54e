"""Ad-hoc timing of the full-size denoiser (development aid; bench.py is the contract)."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "ml-mdm_b200"))
import torch
from mdm_b200 import config as mc
from mdm_b200.models import UNet, NestedUNet
from mdm_b200 import _lib

def main(cfg_name="cc12m_64x64", B=8, steps=3, train=True):
    ucfg, dcfg, nested = mc.load_yaml_configs(os.path.join(HERE, "..", "ml-mdm_b200", "mdm_b200", "configs", cfg_name + ".yaml"))
    torch.manual_seed(0)
    m = (NestedUNet if nested else UNet)(3, 3, ucfg)
    with torch.no_grad():
        for p in m.parameters():
            if float(p.abs().max()) == 0:
                p.normal_(0, 0.02)
    m = m.cuda()
    res = {"cc12m_64x64": [64], "cc12m_256x256": [256, 64], "cc12m_1024x1024": [1024, 256, 64]}[cfg_name]
    xs = [torch.randn(B, 3, r, r, device="cuda") for r in res]
    t = torch.randint(0, 1000, (B,), device="cuda")
    lm = torch.randn(B, 128, 2048, device="cuda")
    mask = torch.ones(B, 128, device="cuda")
    inp = xs if nested else xs[0]
    for it in range(steps + 2):
        if it == 2:
            torch.cuda.synchronize(); l0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
        if train:
            out = m(inp, t, lm, mask, {})
            outs = out if nested else [out]
            loss = sum((o * o).mean() for o in outs)
            loss.backward()
            m.zero_grad(set_to_none=True)
        else:
            with torch.no_grad():
                out = m(inp, t, lm, mask, {})
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    gf = {"cc12m_64x64": 385.4, "cc12m_256x256": 610.6, "cc12m_1024x1024": 1040.1}[cfg_name] * B * (3 if train else 1)
    print(f"{cfg_name} B={B} train={train}: {ms:.1f} ms/step, {gf/ms:.1f} TFLOP/s, launches/step {(_lib.launch_count()-l0)//steps}, pool {m.native().workspace_bytes()}", flush=True)

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "cc12m_64x64"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    train = (sys.argv[3] != "infer") if len(sys.argv) > 3 else True
    main(name, B, 3, train)
