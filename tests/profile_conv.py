"""Micro-benchmark of the implicit-GEMM 3x3 conv / weight-gradient launches through the C ABI (mdm_gemm_raw), for the
narrow layers of the 256- and 1024-px levels (development aid; no reference computation).
usage: python tests/profile_conv.py fwd|wgrad nimg H W Cin Cout [iters] [kfactor]"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "ml-mdm_b200"))
sys.path.insert(0, HERE)
import ctypes as C  # noqa: E402

from mdm_b200 import _lib  # noqa: E402

DEV = "cuda"


def st():
    return torch.cuda.current_stream().cuda_stream


def main(kind, nimg, H, W, Cin, Cout, iters=20, kf=1):
    x = (torch.randn(nimg, H, W, Cin, device=DEV) * 0.5).half()
    PW = 16 if W >= 16 else 8
    p = _lib.GemmParams()
    if kind == "fwd":
        wp = (torch.randn(Cout, 9, Cin, device=DEV) * 0.1).half()
        PH = 128 // PW
        sa = _lib.tmap(x.data_ptr(), (Cin, W, H, nimg), (1, Cin, W * Cin, H * W * Cin), (64, PW, PH, 1))
        bn = min(256, (Cout + 15) // 16 * 16)
        sb = _lib.tmap(wp.data_ptr(), (Cin, Cout, 9, 1), (1, 9 * Cin, Cin, 9 * Cin * Cout), (64, bn, 1, 1))
        p.kind = 1
        p.N, p.K, p.block_n = Cout, Cin, bn
        p.H, p.W, p.PW, p.PH = H, W, PW, PH
        p.tiles_w, p.tiles_h, p.nimg = (W + PW - 1) // PW, (H + PH - 1) // PH, nimg
        p.taps, p.kblocks_c = 9, (Cin + 63) // 64
        p.num_kblocks = 9 * p.kblocks_c
        p.alpha, p.ldc = 1.0, Cout
        out = torch.zeros(nimg, H, W, Cout, device=DEV)
        p.out_f32 = out.data_ptr()
        a_mn = b_mn = 0
        nbytes = x.numel() * 2 + out.numel() * 4
        flops = 2.0 * nimg * H * W * Cout * Cin * 9
    else:
        dy = (torch.randn(nimg, H, W, Cout, device=DEV) * 0.5).half()
        PH = 64 * kf // PW
        sa = _lib.tmap(dy.data_ptr(), (Cout, W, H, nimg), (1, Cout, W * Cout, H * W * Cout), (64, PW, PH, 1))
        sb = _lib.tmap(x.data_ptr(), (Cin, W, H, nimg), (1, Cin, W * Cin, H * W * Cin), (64, PW, PH, 1))
        p.kind, p.kfactor = 2, kf
        p.M, p.N = Cout, Cin
        p.block_n = min(256, (Cin + 15) // 16 * 16)
        p.H, p.W, p.PW, p.PH = H, W, PW, PH
        p.tiles_w, p.tiles_h, p.nimg = (W + PW - 1) // PW, (H + PH - 1) // PH, nimg
        p.taps, p.nz1 = 9, 9
        p.num_kblocks = nimg * p.tiles_w * p.tiles_h
        p.nsplit = max(1, min(296 // 9, p.num_kblocks // 2))
        p.alpha, p.ldc, p.c_z1_stride, p.atomic = 1.0, 9 * Cin, Cin, 1
        out = torch.zeros(Cout, 9, Cin, device=DEV)
        p.out_f32 = out.data_ptr()
        a_mn = b_mn = 1
        nbytes = (x.numel() + dy.numel()) * 2
        flops = 2.0 * nimg * H * W * Cout * Cin * 9
    for _ in range(3):
        _lib.gemm_raw(sa, sb, a_mn, b_mn, p, st())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _lib.gemm_raw(sa, sb, a_mn, b_mn, p, st())
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"{kind} n={nimg} {H}x{W} {Cin}->{Cout} kf={kf} env[{os.environ.get('MDM_PERSIST_MIN_N','-')},"
          f"{os.environ.get('MDM_SMEM_NARROW_KB','-')}]: {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s (algorithmic)  "
          f"{flops / us / 1e6:7.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), int(a[6]) if len(a) > 6 else 20,
         int(a[7]) if len(a) > 7 else 1)
