"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through tests/refharness.py) on CPU in fp32.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs on seeded inputs are committed here:
  schedules.npz     gamma tables (float32 bits) for every schedule type / shift used by the configs,
                    vdm loss weights, set_timesteps() for several N
  tiny_unet.npz     tiny UNet: forward, get_loss (loss, model output, per-parameter gradient norms,
                    a few full gradients), one DDIM step, one DDPM step input/output, a 4-step DDIM sample
  tiny_nested.npz   the same for a 2-level NestedUNet (shifted schedule, double loss)
  keys_*.txt        state_dict key order + shapes of the three shipped configs
Parameters and inputs come from numpy PCG64 seeds (tests/tiny_configs.py), so only outputs are stored.
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import refharness as rh  # noqa: E402
import tiny_configs as tc  # noqa: E402

torch.set_num_threads(8)
ref = rh.load()
S = ref.samplers


def schedules():
    out = {}
    for st in ["DEEPFLOYD", "DDPM", "COSINE"]:
        cfg = S.SamplerConfig(num_diffusion_steps=1000, schedule_type=S.ScheduleType[st])
        smp = S.Sampler(cfg)
        out[f"gammas_{st}"] = smp.gammas.numpy()
        out[f"vdm_{st}"] = smp.vdm_loss_weights.numpy()
    for power, scales in [(1, [4, 1]), (2, [16, 4, 1])]:
        cfg = S.SamplerConfig(num_diffusion_steps=1000, schedule_type=S.ScheduleType.DEEPFLOYD, schedule_shifted=True,
                              schedule_shifted_power=power)
        smp = S.NestedSampler(cfg)
        for s in scales:
            out[f"shift_p{power}_s{s}"] = smp.get_schedule_shifted(smp.gammas, s).numpy()
    # (rescale_schedule > 1 raises in the reference itself: Sampler.__init__ reads self._config before
    #  assigning it, samplers.py:183-190,259 -- every shipped config uses 1.0)
    smp = S.Sampler(S.SamplerConfig(num_diffusion_steps=1000, schedule_type=S.ScheduleType.DEEPFLOYD))
    for n in [1, 2, 5, 50, 100, 250, 999, 1000]:
        out[f"timesteps_{n}"] = smp.set_timesteps(n)
    np.savez_compressed(os.path.join(HERE, "schedules.npz"), **out)
    print("schedules.npz", len(out))


def model_case(kind):
    nested = kind == "nested"
    ucfg = copy.deepcopy(tc.TINY_NESTED if nested else tc.TINY_UNET)
    dcfg = copy.deepcopy(tc.TINY_NESTED_DIFFUSION if nested else tc.TINY_DIFFUSION)
    model, pipe = rh.build(ucfg, dcfg, "nested_unet" if nested else "unet", tc.LM_DIM)
    sd = tc.seeded_state_dict(model.state_dict(), 7)
    model.load_state_dict(sd)
    res = 32 if nested else 16
    nlev = 2 if nested else 1
    x, t, lm, mask = tc.seeded_inputs(3, 2, res, 6, nlevels=nlev)
    out = {}
    with torch.no_grad():
        o = model(x, t, lm, mask, {})
    for i, oi in enumerate(o if nested else [o]):
        out[f"fwd_out{i}"] = oi.numpy()

    # ---- get_loss with a pinned CPU RNG
    images = (x[0] if nested else x).clamp(-1, 1)
    torch.manual_seed(1234)
    pipe.train()
    loss, time, x_t, pred, tgt, _ = pipe.get_loss({"images": images, "lm_outputs": lm, "lm_mask": mask})
    loss.mean().backward()
    out["loss"] = loss.detach().numpy()
    out["loss_time"] = time.numpy()
    out["loss_xt"] = x_t.detach().numpy()
    out["loss_pred"] = pred.detach().numpy()
    out["loss_tgt"] = tgt.detach().numpy()
    names = [k for k, _ in model.named_parameters()]
    out["grad_norms"] = np.array([float(p.grad.norm()) for _, p in model.named_parameters()], dtype=np.float64)
    out["grad_absmax"] = np.array([float(p.grad.abs().max()) for _, p in model.named_parameters()], dtype=np.float64)
    pick = [n for n in names if n.endswith(("conv_in.weight", "mid_blocks.0.attn.0.kv_cond.weight", "temb_layer2.bias",
                                            "up_blocks.1.resnets.0.conv3.weight", "cond_emb.weight", "in_adapter.bias"))]
    for n in pick:
        out["grad__" + n] = dict(model.named_parameters())[n].grad.numpy()
    model.zero_grad()

    # ---- one reverse step (DDIM eta=0 and DDPM with pinned noise), then a 4-step DDIM sample
    pipe.eval()
    smp = pipe.sampler
    m = pipe.get_model()
    xin = x if nested else x
    with torch.no_grad():
        x0, xs, _ = smp.get_xt_minus_1(m, 500, [xi.clone() for xi in x] if nested else x.clone(), lm, mask, {},
                                       time_step_last=480, ddim_eta=0.0, return_details=True)
        for i, (a, b) in enumerate(zip(x0, xs) if nested else [(x0, xs)]):
            out[f"ddim_x0_{i}"], out[f"ddim_xs_{i}"] = a.numpy(), b.numpy()
        torch.manual_seed(99)
        xs = smp.get_xt_minus_1(m, 500, [xi.clone() for xi in x] if nested else x.clone(), lm, mask, {},
                                time_step_last=499, ddim_eta=None)
        for i, b in enumerate(xs if nested else [xs]):
            out[f"ddpm_xs_{i}"] = b.numpy()
        # CFG step: doubled conditioning rows [uncond; cond]
        lm2 = torch.cat([torch.zeros_like(lm), lm])
        mask2 = torch.cat([mask, mask])
        xs = smp.get_xt_minus_1(m, 500, [xi.clone() for xi in x] if nested else x.clone(), lm2, mask2, {},
                                time_step_last=480, ddim_eta=0.0, guidance_scale=3.0)
        for i, b in enumerate(xs if nested else [xs]):
            out[f"cfg_xs_{i}"] = b.numpy()
        # nested sampling starts from the full-resolution tensor; the low-resolution start is drawn inside
        # with normal_() (samplers.py:669-676) -> pinned by the seed below
        torch.manual_seed(7)
        final = smp.sample(m, x[0].clone() if nested else x.clone(), lm, mask, {}, num_inference_steps=4,
                           ddim_eta=0.0, resample_steps=True)
        out["sample4"] = final.numpy()
    np.savez_compressed(os.path.join(HERE, f"tiny_{kind}.npz"), **out)
    with open(os.path.join(HERE, f"tiny_{kind}_params.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    print(f"tiny_{kind}.npz", {k: v.shape for k, v in out.items() if k.startswith(("fwd", "loss", "sample"))})


def shipped_keys():
    for y, arch in [("cc12m_64x64", "unet"), ("cc12m_256x256", "nested_unet"), ("cc12m_1024x1024", "nested2_unet")]:
        cfg = rh.load_yaml(y + ".yaml")
        model, _ = rh.build(cfg["unet_config"], cfg["diffusion_config"], arch, 2048)
        with open(os.path.join(HERE, f"keys_{y}.txt"), "w") as f:
            for k, v in model.state_dict().items():
                f.write(f"{k} {'x'.join(str(d) for d in v.shape)}\n")
        print(y, sum(p.numel() for p in model.parameters()))


if __name__ == "__main__":
    schedules()
    model_case("unet")
    model_case("nested")
    shipped_keys()
