"""Device timing of the fused attention operator at the two shapes of cc12m_64x64 (development aid;
also the target of the ncu captures under profiles/)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ml-mdm_b200"))
from mdm_b200 import _lib  # noqa: E402


def bench(B, T, S, Cc, heads=8, iters=10):
    dev = "cuda"
    qkv = (torch.randn(B, T, 3 * Cc, device=dev) * 0.7).half()
    kv = (torch.randn(B, S, 2 * Cc, device=dev) * 0.7).half()
    dO = (torch.randn(B, T, Cc, device=dev) * 0.5).half()
    h16 = torch.empty(B, T, Cc, device=dev, dtype=torch.float16)
    os16 = torch.empty_like(h16)
    stats = torch.empty(B, heads, 2, T, 2, device=dev)
    Dterm = torch.empty(B, heads, 2, T, device=dev)
    dq32 = torch.empty(B, T, Cc, device=dev)
    dqkv = torch.zeros(B, T, 3 * Cc, device=dev, dtype=torch.float16)
    dkv = torch.zeros(B, S, 2 * Cc, device=dev, dtype=torch.float16)
    lib = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())

    def fwd():
        _lib.check(lib.mdm_op_attention_fwd(P(qkv), P(kv), None, B, T, S, Cc, heads, P(h16), P(os16), P(stats), st), "fwd")

    def bwd():
        _lib.check(lib.mdm_op_attention_bwd(P(qkv), P(kv), None, P(dO), P(h16), P(os16), P(stats), B, T, S, Cc, heads,
                                            P(Dterm), P(dq32), P(dqkv), P(dkv), st), "bwd")

    out = {}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / iters * 1e3
    d = Cc // heads
    flops_f = 4.0 * B * T * (T + S) * Cc
    print(f"B={B} T={T} S={S} C={Cc} d={d}: fwd {out['fwd']:.0f} us ({flops_f / out['fwd'] / 1e6:.0f} TFLOP/s)  "
          f"bwd {out['bwd']:.0f} us ({2.5 * flops_f / out['bwd'] / 1e6:.0f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    bench(64, 1024, 128, 512, iters=iters)
    bench(64, 256, 128, 768, iters=iters)
