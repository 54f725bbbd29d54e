"""Micro-benchmark of the MMA-bound launches of the one-tile tcgen05 kernel (long contractions on wide tiles) for the
CTA-pair experiments: one process per knob setting (MDM_GEMM_PAIR, MDM_SMEM_BUDGET_KB are read once).  Development
aid; no reference computation except the cuBLAS line."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gemm_cases as gc  # noqa: E402
import profile_conv as pc  # noqa: E402

if __name__ == "__main__":
    tag = f"pair={os.environ.get('MDM_GEMM_PAIR', '0')} budget={os.environ.get('MDM_SMEM_BUDGET_KB', '-')}"
    print("==", tag, flush=True)
    for (M, N, K, bn) in [(8192, 8192, 8192, 256), (8192, 768, 6912, 256), (8192, 768, 6912, 192), (16384, 256, 2304, 256),
                          (16384, 3072, 3200, 256), (16384, 768, 3200, 256)]:
        ms, tf, msr, tfr = gc.bench_one(M, N, K, bn=bn)
        print(f"plain {M}x{N}x{K} bn={bn}: {ms * 1e3:8.1f} us {tf:6.0f} TFLOP/s | cuBLAS {tfr:6.0f}", flush=True)
    pc.main("fwd", 64, 16, 16, 768, 768)
    pc.main("fwd", 64, 32, 32, 512, 512)
    pc.main("fwd", 64, 64, 64, 256, 256)
    pc.main("wgrad", 64, 16, 16, 768, 768)
    pc.main("wgrad", 64, 32, 32, 512, 512)
