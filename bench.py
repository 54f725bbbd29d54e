#!/usr/bin/env python
"""bench.py -- denoising steps/sec (fwd+bwd) of the Matryoshka denoising path on N x B200.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1     # CPU arm (oracle port)

One step = Diffusion.get_loss(sample) + loss.mean().backward() on a synthetic batch of the
BASELINE.json workload (default configs[1]: cc12m_64x64 U-Net training, batch 64 per GPU, random T5
embeddings with S=128 tokens), i.e. q-sample, full U-Net forward, loss, full backward with every
parameter gradient, plus the one gradient all-reduce when N > 1.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ml-mdm_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# forward GFLOP per sample (2*MAC of every conv/linear + 4*B*C*T*S per attention), measured by hooking
# the reference modules with S=128; training = 3x.
FWD_GFLOP = {"cc12m_64x64": 385.4, "cc12m_256x256": 610.6, "cc12m_1024x1024": 1040.1}
ATTN_FWD_GFLOP = 19.9  # of which attention (QK^T and PV); runs in the fused attention kernels, not the GEMM engine
RES = {"cc12m_64x64": [64], "cc12m_256x256": [256, 64], "cc12m_1024x1024": [1024, 256, 64]}
DEFAULT_BATCH = {"cc12m_64x64": 64, "cc12m_256x256": 32, "cc12m_1024x1024": 1}
TOKENS = 128


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1402.9), d.get("hbm_gbs", 6576.1), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.path = None, None
        try:
            self.path = tempfile.mktemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "200"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path).read().strip().split("\n"):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_pipeline(cfg_name, device):
    from mdm_b200 import config as mc
    from mdm_b200.diffusion import Diffusion, NestedDiffusion
    from mdm_b200.models import NestedUNet, UNet

    ucfg, dcfg, nested = mc.load_yaml_configs(os.path.join(ROOT, "ml-mdm_b200", "mdm_b200", "configs", cfg_name + ".yaml"))
    torch.manual_seed(4321)
    model = (NestedUNet if nested else UNet)(3, 3, ucfg)
    with torch.no_grad():  # the reference zero-initialises ~1/3 of its layers; a trained net has none at zero
        for p in model.parameters():
            if float(p.detach().abs().max()) == 0:
                p.normal_(0, 0.02)
    pipe = (NestedDiffusion if nested else Diffusion)(model, dcfg).to(device)
    return pipe, nested


def synthetic_host_batch(cfg_name, B, seed):
    g = torch.Generator().manual_seed(seed)
    R = RES[cfg_name][0]
    return {
        "images": (torch.rand(B, 3, R, R, generator=g) * 2 - 1).pin_memory(),
        "lm_outputs": torch.randn(B, TOKENS, 2048, generator=g).pin_memory(),
        "lm_mask": torch.ones(B, TOKENS).pin_memory(),
    }


def run_ours(args):
    import torch.distributed as dist
    from mdm_b200 import _lib, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", init_method="env://")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg_name = args.config
    B = args.batch or DEFAULT_BATCH[cfg_name]
    pipe, nested = build_pipeline(cfg_name, dev)
    pipe.train()
    vm = pipe.get_model().vision_model
    host = synthetic_host_batch(cfg_name, B, 1234 + rank)
    resident = {k: v.to(dev) for k, v in host.items()}

    # Gradient all-reduce: one NCCL call over the flat arena after backward. MDM_OVERLAP=1 instead reduces
    # slices of the arena while backward is still running (parallel.GradientOverlap); measured at N=2 that is
    # not faster yet (782 vs 792 sample-steps/s: NCCL's copy CTAs and the one-CTA-per-SM persistent GEMM
    # compete for SMs), so it is opt-in.
    overlap = (parallel.GradientOverlap(vm, bucket_mb=int(os.environ.get("MDM_BUCKET_MB", "64")))
               if world > 1 and os.environ.get("MDM_OVERLAP") is not None else None)

    def step(sample):
        loss, *_ = pipe.get_loss(sample)
        if overlap is not None:
            overlap.arm()
        loss.mean().backward()
        if overlap is not None:
            overlap.finish()
        elif world > 1:
            parallel.allreduce_gradients(vm)
        return loss

    def zero():
        vm.zero_grad(set_to_none=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also sizes the engine's memory pool)
    for _ in range(max(args.warmup, 3)):
        step(resident)
        zero()
    barrier()
    # ---- timed: inputs resident in HBM
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(resident)
        zero()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    clk = clocks.stop() if clocks is not None else None
    # ---- end to end: host (pinned) inputs, H2D inside the timed region, D2H of the loss
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loss_host = torch.empty(B).pin_memory()
    barrier()
    f0.record()
    for _ in range(args.steps):
        sample = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss = step(sample)
        loss_host.copy_(loss.detach(), non_blocking=True)
        zero()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    # ---- roofline of the dominant kernel (tcgen05 GEMM/conv engine): its launches are bracketed with
    # CUDA events on the launching stream for two extra steps
    roof = None
    import ctypes as C
    lib = _lib.lib()
    if rank == 0:
        lib.mdm_profile_gemm(1)
    for _ in range(2):  # every rank runs these steps (they contain the gradient all-reduce)
        step(resident)
        zero()
    barrier()
    if rank == 0:
        tot = C.c_double()
        cnt = C.c_longlong()
        lib.mdm_profile_read(C.byref(tot), C.byref(cnt))
        lib.mdm_profile_gemm(0)
        gemm_ms = tot.value / 2
        peak_tf, peak_bw, how = measured_peaks()
        fused_attn = os.environ.get("MDM_UNFUSED_ATTENTION") is None
        flops = (FWD_GFLOP[cfg_name] - (ATTN_FWD_GFLOP if fused_attn else 0.0)) * 3 * B * 1e9
        ach = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        roof = {"bound": "tensor",
                "kernel": "gemm_tc_persistent_kernel / gemm_tc_kernel (tcgen05 implicit-GEMM 3x3 conv + linear layers; "
                          "the fused attention kernels are timed separately and excluded from these FLOPs)",
                "achieved": round(ach, 1), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4),
                "traffic": 349.2e6,
                "traffic_note": "bytes; ncu --set full of the largest conv launch (3x3 256->256 @ 64x64, batch 64; "
                                "profiles/r01_ncu_conv256_v2_summary.txt): dram read 135.5 MB + write 213.7 MB vs 403.8 MB "
                                "algorithmic (fp16 in, fp32 out) -> inputs read once, no re-reads",
                "peak_source": how, "launches_per_step": int(cnt.value // 2),
                "kernel_ms_per_step": round(gemm_ms, 3), "share_of_step": round(gemm_ms / (ms / args.steps), 3),
                "algorithmic_flops_per_step": flops}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    gb = B * world
    out = {
        "metric": f"denoising steps/sec (fwd+bwd), {cfg_name} U-Net, per-sample steps summed over all GPUs",
        "value": round(gb * args.steps / (ms * 1e-3), 2),
        "unit": "sample-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate/residual stream (tcgen05 kind::f16)",
        "data": "synthetic (uniform images, random T5 embeddings S=128, random-init weights incl. the reference's zero-init layers)",
        "config": {"workload": f"{cfg_name} training fwd+bwd, batch {B}/GPU (BASELINE.json configs[1])" if cfg_name == "cc12m_64x64"
                   else f"{cfg_name} training fwd+bwd, batch {B}/GPU",
                   "global_batch": gb, "tokens": TOKENS, "batch_steps_per_sec": round(args.steps / (ms * 1e-3), 3),
                   "parallelism": f"dp{world}", "l2": "per-step working set (activation stash, GBs) far exceeds the 126 MB L2",
                   "tflops_per_gpu": round(FWD_GFLOP[cfg_name] * 3 * B / (ms / args.steps), 1)},
        "e2e": {"value": round(gb * args.steps / (ms_e2e * 1e-3), 2), "unit": "sample-steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
        "engine_pool_bytes": vm.native().workspace_bytes()[0],
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_arm(cfg_name, steps=1, warmup=1)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_arm(cfg_name, steps, warmup, batch=1):
    """The reference's CPU path for this workload, timed on this host's cores: the oracle port
    (oracle/unet_ref.py + oracle/diffusion_ref.py; the Python reference itself cannot travel to the GPU
    box). One step = get_loss + backward on a bounded sample of `batch` images."""
    import types
    import yaml

    from oracle import diffusion_ref as dref
    from oracle import unet_ref

    # oneDNN/MKL stop scaling (and then collapse) beyond ~32 threads on these shapes; measured on the
    # 128-thread GPU host: conv 256->256@64x64 4.0 ms at 32 threads, 18 ms at 64.
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    y = yaml.safe_load(open(os.path.join(ROOT, "ml-mdm_b200", "mdm_b200", "configs", cfg_name + ".yaml")))

    def ns(d):
        return types.SimpleNamespace(**{k: ns(v) for k, v in d.items()}) if isinstance(d, dict) else d

    ucfg = ns(y["unet_config"])
    c = ucfg
    while c is not None:
        if hasattr(c, "initialize_inner_with_pretrained"):
            c.initialize_inner_with_pretrained = None
        c = getattr(c, "inner_config", None)
    net = unet_ref.OracleNet(ucfg, 2048)
    from mdm_b200 import config as mc
    from mdm_b200.models import NestedUNet, UNet
    cfg2, _, nested = mc.load_yaml_configs(os.path.join(ROOT, "ml-mdm_b200", "mdm_b200", "configs", cfg_name + ".yaml"))
    torch.manual_seed(4321)
    shapes = (NestedUNet if nested else UNet)(3, 3, cfg2)  # parameter container only (shapes + init)
    P = {}
    for k, p in shapes.named_parameters():
        v = p.detach().clone()
        if float(v.abs().max()) == 0:
            v.normal_(0, 0.02)
        P[k] = v.requires_grad_(True)
    del shapes
    R = RES[cfg_name][0]
    scales = net.nest_ratio + [1] if nested else [1]
    sc = y["diffusion_config"]["sampler_config"]
    gam = dref.gammas_f32(sc.get("schedule_type", "DDPM"), sc.get("num_diffusion_steps", 1000))
    g = torch.Generator().manual_seed(1234)
    images = torch.rand(batch, 3, R, R, generator=g) * 2 - 1
    lm = torch.randn(batch, TOKENS, 2048, generator=g)
    mask = torch.ones(batch, TOKENS)

    def one():
        time_ = torch.randint(0, 1000, (batch,), generator=g)
        eps = [torch.randn(batch, 3, R // (scales[0] // s), R // (scales[0] // s), generator=g) for s in scales]
        loss, _, _ = dref.training_loss(net, P, images, eps, time_, lm, mask, gam, scales, dref.V_PREDICTION, dref.DDPM,
                                        shifted=bool(sc.get("schedule_shifted", False)),
                                        power=sc.get("schedule_shifted_power", 1))
        loss.mean().backward()
        for p in P.values():
            p.grad = None

    if warmup:  # one forward-only pass: pages in the weights and sizes oneDNN's primitives
        with torch.no_grad():
            net.forward(P, [torch.zeros(1, 3, R // (scales[0] // s), R // (scales[0] // s)) for s in scales] if nested
                        else torch.zeros(1, 3, R, R), torch.zeros(1, dtype=torch.long), lm[:1], mask[:1], {})
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return {"value": round(batch * steps / dt, 4), "unit": "sample-steps/s", "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) of get_loss+backward at batch {batch} (fp32, torch CPU oneDNN/MKL ops, {cores} threads)",
            "seconds": round(dt, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name = args.config
    base = cpu_arm(cfg_name, steps=max(1, min(args.steps, 3)), warmup=max(1, min(args.warmup, 1)))
    out = {
        "impl": "reference",
        "metric": f"denoising steps/sec (fwd+bwd), {cfg_name} U-Net, per-sample steps summed over all GPUs",
        "value": base["value"], "unit": "sample-steps/s", "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)),
        "warmup": 1, "ms_per_step": round(1000.0 * 1 / base["value"], 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (same generator as the GPU arm)",
        "config": {"workload": f"{cfg_name} training fwd+bwd (BASELINE.json configs[1]), bounded sample of batch 1 per step"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cc12m_64x64", choices=sorted(FWD_GFLOP))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: the mdm_b200 path has no CPU fallback (use --impl reference for the CPU arm)")
        run_ours(args)


if __name__ == "__main__":
    main()
