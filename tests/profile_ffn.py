"""Micro-benchmark of the plain (Linear) tcgen05 GEMM launches of the transformer blocks through the C ABI
(mdm_gemm_raw), one epilogue variant per line, warm (same buffers back to back) and cold (L2 flushed between
launches, each launch timed alone).  Development aid; no reference computation.
usage: python tests/profile_ffn.py [M]"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "ml-mdm_b200"))
from mdm_b200 import _lib  # noqa: E402

DEV = "cuda"


def st():
    return torch.cuda.current_stream().cuda_stream


FLUSH = None


def flush():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    FLUSH.add_(1)


def bench(name, M, N, K, b_mn, bn, bias=True, res=False, f32=False, f16=False, act=False, gsrc=False, iters=20):
    A = (torch.randn(M, K, device=DEV) * 0.5).half()
    B = (torch.randn(K, N, device=DEV) * 0.05).half() if b_mn else (torch.randn(N, K, device=DEV) * 0.05).half()
    sa = _lib.tmap(A.data_ptr(), (K, M, 1, 1), (1, K, K * M, K * M), (64, 128, 1, 1))
    if b_mn:
        sb = _lib.tmap(B.data_ptr(), (N, K, 1, 1), (1, N, K * N, K * N), (64, 64, 1, 1))
    else:
        sb = _lib.tmap(B.data_ptr(), (K, N, 1, 1), (1, K, K * N, K * N), (64, bn, 1, 1))
    p = _lib.GemmParams()
    p.kind = 0
    p.M, p.N, p.K, p.block_n = M, N, K, bn
    p.nz1 = p.nz2 = p.nsplit = 1
    p.num_kblocks = (K + 63) // 64
    p.alpha, p.ldc = 1.0, N
    keep = []
    if bias:
        t = torch.randn(N, device=DEV); keep.append(t); p.bias = t.data_ptr()
    if res:
        t = torch.randn(M, N, device=DEV); keep.append(t); p.residual = t.data_ptr()
    if f32:
        t = torch.zeros(M, N, device=DEV); keep.append(t); p.out_f32 = t.data_ptr()
    if f16:
        t = torch.zeros(M, N, device=DEV, dtype=torch.float16); keep.append(t); p.out_f16 = t.data_ptr()
    if act:
        t = torch.zeros(M, N, device=DEV, dtype=torch.float16); keep.append(t); p.out_act_f16 = t.data_ptr(); p.act = 1
    if gsrc:
        t = torch.randn(M, N, device=DEV).half(); keep.append(t); p.gelu_grad_src = t.data_ptr()
    for _ in range(3):
        _lib.gemm_raw(sa, sb, 0, b_mn, p, st())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _lib.gemm_raw(sa, sb, 0, b_mn, p, st())
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) * 1e3 / iters
    cold = 0.0
    for _ in range(5):
        flush()
        e0.record()
        _lib.gemm_raw(sa, sb, 0, b_mn, p, st())
        e1.record()
        torch.cuda.synchronize()
        cold += e0.elapsed_time(e1) * 1e3 / 5
    fl = 2.0 * M * N * K
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    print(f"{name:34s} M={M} N={N} K={K} {'nn' if b_mn else 'nt'} bn={bn:3d} tiles={tiles:5d}: warm {warm:7.1f} us "
          f"{fl / warm / 1e6:6.0f} TF/s | cold {cold:7.1f} us {fl / cold / 1e6:6.0f} TF/s", flush=True)


def main(M):
    C = 768
    if "--ncu" in sys.argv:  # the two TMA-operand epilogues, a few launches each, for an `ncu --set full` capture
        bench("proj fwd: res + f32 (train)", M, C, C, 0, 192, res=True, f32=True, iters=1)
        bench("fc2 dgrad: f16 * gelu'(src) (train)", M, 4 * C, C, 1, 192, bias=False, f16=True, gsrc=True, iters=1)
        return
    for bn in (192, 256, 128):
        bench("fc1 fwd: bias only, f16", M, 4 * C, C, 0, bn, f16=True)
        bench("fc1 fwd: +GELU, act only", M, 4 * C, C, 0, bn, act=True)
        bench("fc1 fwd: +GELU, f16 + act (train)", M, 4 * C, C, 0, bn, f16=True, act=True)
        bench("fc2 dgrad: f16", M, 4 * C, C, 1, bn, bias=False, f16=True)
        bench("fc2 dgrad: f16 * gelu'(src) (train)", M, 4 * C, C, 1, bn, bias=False, f16=True, gsrc=True)
    for bn in (192, 256, 128):
        bench("proj fwd: f16", M, C, C, 0, bn, f16=True)
        bench("proj fwd: f32", M, C, C, 0, bn, f32=True)
        bench("proj fwd: res + f32 (train)", M, C, C, 0, bn, res=True, f32=True)
        bench("proj dgrad: f16 (train)", M, C, C, 1, bn, bias=False, f16=True)
        bench("fc2 fwd: res + f32 (train)", M, C, 4 * C, 0, bn, res=True, f32=True)
        bench("fc2 fwd: f16", M, C, 4 * C, 0, bn, f16=True)
        bench("fc1 dgrad: f32 (train)", M, C, 4 * C, 1, bn, bias=False, f32=True)
        bench("qkv fwd: f16 (train)", M, 3 * C, C, 0, bn, f16=True)
        bench("qkv dgrad: f32 (train)", M, C, 3 * C, 1, bn, bias=False, f32=True)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16384)
