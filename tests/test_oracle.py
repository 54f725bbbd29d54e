"""Pins the oracle (oracle/*.py) against the golden fixtures generated from the unmodified reference
(tests/golden/make_golden.py) and, when /root/reference is present, against the reference live."""
import copy
import os
import types

import numpy as np
import pytest
import torch

import refharness as rh
import tiny_configs as tc
from oracle import diffusion_ref as dref
from oracle import unet_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ns(d):
    if isinstance(d, dict):
        return types.SimpleNamespace(**{k: ns(v) for k, v in d.items()})
    return d


def tiny(kind):
    nested = kind == "nested"
    ucfg = copy.deepcopy(tc.TINY_NESTED if nested else tc.TINY_UNET)
    if nested:
        ucfg["initialize_inner_with_pretrained"] = None
    net = unet_ref.OracleNet(ns(ucfg), tc.LM_DIM)
    names = open(os.path.join(GOLD, f"tiny_{kind}_params.txt")).read().split()
    gold = np.load(os.path.join(GOLD, f"tiny_{kind}.npz"))
    return net, names, gold, nested


def params_for(kind, names, requires_grad=False):
    # shapes come from the golden fixture-independent module tree: rebuild through the product container
    from mdm_b200 import config as mc
    from mdm_b200.models import NestedUNet, UNet

    ucfg = copy.deepcopy(tc.TINY_NESTED if kind == "nested" else tc.TINY_UNET)
    cfg = mc.unet_config_from_dict(ucfg)
    cfg.conditioning_feature_dim = tc.LM_DIM
    m = (NestedUNet if kind == "nested" else UNet)(3, 3, cfg)
    assert [k for k, _ in m.named_parameters()] == names
    sd = tc.seeded_state_dict(m.state_dict(), 7)
    return {k: v.clone().requires_grad_(requires_grad) for k, v in sd.items()}


def close(a, b, tol=2e-5):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert err <= tol, err


@pytest.mark.parametrize("kind", ["unet", "nested"])
def test_forward_matches_reference_golden(kind):
    net, names, gold, nested = tiny(kind)
    P = params_for(kind, names)
    x, t, lm, mask = tc.seeded_inputs(3, 2, 32 if nested else 16, 6, nlevels=2 if nested else 1)
    with torch.no_grad():
        out = net.forward(P, x, t, lm, mask, {})
    for i, o in enumerate(out if nested else [out]):
        close(o, gold[f"fwd_out{i}"])


@pytest.mark.parametrize("kind", ["unet", "nested"])
def test_loss_and_gradients_match_reference_golden(kind):
    net, names, gold, nested = tiny(kind)
    P = params_for(kind, names, requires_grad=True)
    x, t, lm, mask = tc.seeded_inputs(3, 2, 32 if nested else 16, 6, nlevels=2 if nested else 1)
    images = (x[0] if nested else x).clamp(-1, 1)
    gam = dref.gammas_f32("DEEPFLOYD", 1000)
    torch.manual_seed(1234)  # same CPU generator draws as Diffusion.get_loss (samplers.py:236-241)
    time = torch.randint(0, 1000, (2,))
    eps = [torch.randn_like(images)]
    scales = [4, 1] if nested else [1]
    if nested:
        eps.append(torch.empty(2, 3, 8, 8).normal_())
    assert np.array_equal(time.numpy(), gold["loss_time"])
    loss, x_t, outs = dref.training_loss(net, P, images, eps, time, lm, mask, gam, scales, dref.V_PREDICTION, dref.DDPM,
                                         shifted=nested, power=1)
    close(x_t[0], gold["loss_xt"], 1e-6)
    close(loss, gold["loss"])
    loss.mean().backward()
    norms = np.array([float(P[k].grad.norm()) for k in names])
    ref = gold["grad_norms"]
    floor = 1e-3 * np.median(ref)
    assert np.max(np.abs(norms - ref) / np.maximum(ref, floor)) < 1e-3
    for k in gold.files:
        if k.startswith("grad__"):
            close(P[k[6:]].grad, gold[k], 2e-4)


@pytest.mark.parametrize("kind", ["unet", "nested"])
def test_reverse_steps_and_sampling_match_reference_golden(kind):
    net, names, gold, nested = tiny(kind)
    P = params_for(kind, names)
    x, t, lm, mask = tc.seeded_inputs(3, 2, 32 if nested else 16, 6, nlevels=2 if nested else 1)
    xs = list(x) if nested else [x]
    scales = [4, 1] if nested else [1]
    gam = dref.gammas_f32("DEEPFLOYD", 1000)
    tabs = [dref.shift_table(gam, s, 1) if nested else gam for s in scales]
    times = torch.full((2,), 499, dtype=torch.long)
    with torch.no_grad():
        o = net.forward(P, xs if nested else xs[0], times, lm, mask, {})
        o = list(o) if nested else [o]
        for i, (xi, p, tab) in enumerate(zip(xs, o, tabs)):
            x0, x_s = dref.reverse_step(xi, p, tab[500], tab[480], dref.V_PREDICTION, True, 1.0, 0.0, True)
            close(x0, gold[f"ddim_x0_{i}"])
            close(x_s, gold[f"ddim_xs_{i}"])
        torch.manual_seed(99)
        for i, (xi, p, tab) in enumerate(zip(xs, o, tabs)):
            _, x_s = dref.reverse_step(xi, p, tab[500], tab[499], dref.V_PREDICTION, True, 1.0, None, True)
            close(x_s, gold[f"ddpm_xs_{i}"])
        # classifier-free guidance, rows [uncond; cond]
        lm2 = torch.cat([torch.zeros_like(lm), lm])
        mask2 = torch.cat([mask, mask])
        o2 = net.forward(P, [torch.cat([a, a]) for a in xs] if nested else torch.cat([xs[0], xs[0]]),
                         torch.cat([times, times]), lm2, mask2, {})
        o2 = list(o2) if nested else [o2]
        for i, (xi, p, tab) in enumerate(zip(xs, o2, tabs)):
            u, c = p.chunk(2)
            _, x_s = dref.reverse_step(xi, u + 3.0 * (c - u), tab[500], tab[480], dref.V_PREDICTION, True, 1.0, 0.0, True)
            close(x_s, gold[f"cfg_xs_{i}"])
        torch.manual_seed(7)
        init = [xs[0]] + ([torch.empty(2, 3, 8, 8).normal_()] if nested else [])
        final = dref.sample_loop(net, P, init, lm, mask, gam, scales, dref.V_PREDICTION, 1000, 4, 0.0, shifted=nested)
        close(final[0], gold["sample4"], 5e-5)


def test_schedule_tables_and_timesteps_bit_exact():
    g = np.load(os.path.join(GOLD, "schedules.npz"))
    for st in ["DEEPFLOYD", "DDPM", "COSINE"]:
        tab = dref.gammas_f32(st, 1000)
        assert np.array_equal(tab.numpy().view(np.uint32), g[f"gammas_{st}"].view(np.uint32)), st
        assert np.array_equal(dref.vdm_weights(tab).numpy().view(np.uint32), g[f"vdm_{st}"].view(np.uint32)), st
    base = dref.gammas_f32("DEEPFLOYD", 1000)
    for p, scales in [(1, [4, 1]), (2, [16, 4, 1])]:
        for s in scales:
            got = dref.shift_table(base, s, p).numpy()
            assert np.array_equal(got.view(np.uint32), g[f"shift_p{p}_s{s}"].view(np.uint32)), (p, s)
    for n in [1, 2, 5, 50, 100, 250, 999, 1000]:
        assert np.array_equal(dref.set_timesteps(1000, n), g[f"timesteps_{n}"])
    # closed-form known answers derived from the reference code
    assert float(base[0]) == 1.0
    assert float(base[1]) == pytest.approx(0.9999586939811707, abs=0)
    assert float(base[500]) == pytest.approx(0.49384358525276184, abs=0)
    ts = dref.set_timesteps(1000, 50)
    assert len(ts) == 51 and list(ts[:4]) == [981, 962, 942, 922] and list(ts[-4:]) == [59, 39, 20, 0]


@pytest.mark.skipif(not rh.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_matches_live_reference_on_shipped_64_config():
    """cc12m_64x64 at full width (461 M parameters), B=1, S=16: oracle vs the reference modules."""
    torch.manual_seed(0)
    y = rh.load_yaml("cc12m_64x64.yaml")
    model, _ = rh.build(y["unet_config"], y["diffusion_config"], "unet", 2048)
    with torch.no_grad():
        for p in model.parameters():
            if float(p.abs().max()) == 0:
                p.normal_(0, 0.02)  # zero-initialised layers would hide most of the network
    ucfg = copy.deepcopy(y["unet_config"])
    net = unet_ref.OracleNet(ns(ucfg), 2048)
    x = torch.randn(1, 3, 64, 64)
    t = torch.tensor([417])
    lm = torch.randn(1, 16, 2048)
    mask = torch.ones(1, 16)
    with torch.no_grad():
        ref = model(x, t, lm, mask, {})
        out = net.forward(dict(model.state_dict()), x, t, lm, mask, {})
    close(out, ref, 1e-5)


@pytest.mark.skipif(not rh.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_matches_live_reference_on_shipped_256_config():
    """cc12m_256x256 at full width (2-level nest, 476.6 M parameters), B=1, S=8: oracle vs the reference's
    NestedUNet — pins the nesting adapters, the inner/outer skip wiring and the 4x resolution ratio at the real
    channel widths (the tiny nested fixture pins them at toy widths)."""
    torch.manual_seed(1)
    y = rh.load_yaml("cc12m_256x256.yaml")
    model, _ = rh.build(y["unet_config"], y["diffusion_config"], "nested_unet", 2048)
    with torch.no_grad():
        for p in model.parameters():
            if float(p.abs().max()) == 0:
                p.normal_(0, 0.02)
    ucfg = copy.deepcopy(y["unet_config"])
    ucfg["initialize_inner_with_pretrained"] = None
    net = unet_ref.OracleNet(ns(ucfg), 2048)
    xs = [torch.randn(1, 3, 256, 256), torch.randn(1, 3, 64, 64)]
    t = torch.tensor([233])
    lm = torch.randn(1, 8, 2048)
    mask = torch.ones(1, 8)
    with torch.no_grad():
        ref = model(xs, t, lm, mask, {})
        out = net.forward(dict(model.state_dict()), xs, t, lm, mask, {})
    assert len(out) == len(ref) == 2
    for o, r in zip(out, ref):
        assert o.shape == r.shape
        close(o, r, 1e-5)


@pytest.mark.skipif(not rh.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_matches_live_reference_on_shipped_1024_config():
    """cc12m_1024x1024 at full width (3-level nest, 481 M parameters), B=1, S=4: oracle vs the reference."""
    torch.manual_seed(2)
    y = rh.load_yaml("cc12m_1024x1024.yaml")
    model, _ = rh.build(y["unet_config"], y["diffusion_config"], "nested2_unet", 2048)
    with torch.no_grad():
        for p in model.parameters():
            if float(p.abs().max()) == 0:
                p.normal_(0, 0.02)

    def strip(d):  # nested dicts: no pretrained initialisation anywhere
        if isinstance(d, dict):
            if "initialize_inner_with_pretrained" in d:
                d["initialize_inner_with_pretrained"] = None
            for v in d.values():
                strip(v)
        return d

    net = unet_ref.OracleNet(ns(strip(copy.deepcopy(y["unet_config"]))), 2048)
    xs = [torch.randn(1, 3, 1024, 1024), torch.randn(1, 3, 256, 256), torch.randn(1, 3, 64, 64)]
    t = torch.tensor([77])
    lm = torch.randn(1, 4, 2048)
    mask = torch.ones(1, 4)
    with torch.no_grad():
        ref = model(xs, t, lm, mask, {})
        out = net.forward(dict(model.state_dict()), xs, t, lm, mask, {})
    assert len(out) == len(ref) == 3
    for o, r in zip(out, ref):
        assert o.shape == r.shape
        close(o, r, 1e-5)


@pytest.mark.skipif(not rh.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_mixed_ratio_loss_matches_live_reference():
    """NestedDiffusion.get_loss with mixed_ratio='2:1' (what configs/models/cc12m_256x256.yaml:108 sets): only the
    leading int(2/3 * B) samples run the high-resolution level, predictions are zero-padded, the per-level loss is
    divided by the fraction and masked (diffusion.py:262-274, 378-382; nested_unet.py:180,193-204,209). The oracle
    replays the reference's CPU generator draws; loss and every gradient are compared."""
    B = 3
    dcfg = copy.deepcopy(tc.TINY_NESTED_DIFFUSION)
    dcfg["mixed_ratio"] = "2:1"
    ucfg = copy.deepcopy(tc.TINY_NESTED)
    model, pipe = rh.build(ucfg, dcfg, "nested_unet", tc.LM_DIM)
    sd = tc.seeded_state_dict(model.state_dict(), 7)
    model.load_state_dict(sd)
    x, t, lm, mask = tc.seeded_inputs(3, B, 32, 6, nlevels=2)
    images = x[0].clamp(-1, 1)
    torch.manual_seed(4321)
    pipe.train()
    loss, time, x_t, pred, tgt, _ = pipe.get_loss({"images": images, "lm_outputs": lm, "lm_mask": mask})
    loss.mean().backward()
    torch.manual_seed(4321)
    time_r = torch.randint(0, 1000, (B,))
    eps = [torch.randn_like(images), None]
    eps[1] = torch.empty(B, 3, 8, 8).normal_()
    assert torch.equal(time_r, time)
    ocfg = copy.deepcopy(tc.TINY_NESTED)
    ocfg["initialize_inner_with_pretrained"] = None
    net = unet_ref.OracleNet(ns(ocfg), tc.LM_DIM)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    gam = dref.gammas_f32("DEEPFLOYD", 1000)
    mr = dref.mixed_ratio_fractions("2:1")
    assert int(mr[0] * B) == 2 and mr[1] == 1.0
    oloss, ox_t, _ = dref.training_loss(net, P, images, eps, time, lm, mask, gam, [4, 1], dref.V_PREDICTION, dref.DDPM,
                                        shifted=True, power=1, mixed_ratio=mr)
    close(ox_t[0], x_t, 1e-6)
    close(oloss, loss.detach())
    oloss.mean().backward()
    for k, p in model.named_parameters():
        close(P[k].grad, p.grad, 2e-4)


@pytest.mark.skipif(not rh.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("mode", ["DYNAMIC", "DYNAMIC_IF", "CLIP", "NONE"])
def test_oracle_clip_sample_matches_live_reference(mode):
    """Sampler.clip_sample incl. dynamic thresholding (samplers.py:461-508): bit-identical restatement."""
    ref = rh.load()
    cfg = ref.samplers.SamplerConfig()
    smp = ref.samplers.Sampler(cfg)
    cfg.threshold_function = getattr(ref.samplers.ThresholdType, mode)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 3, 32, 32, generator=g) * torch.tensor([0.4, 1.3, 9.0]).view(3, 1, 1, 1)
    for scale in (1.0, 4.0):
        assert torch.equal(smp.clip_sample(x, scale), dref.clip_sample(x, scale, mode if mode != "NONE" else False))
