"""Fused attention operator (mdm_op_attention_fwd / _bwd) vs a torch fp32 restatement of
SelfAttention.attention for both branches (reference models/unet.py:276-307)."""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ml-mdm_b200"))
from mdm_b200 import _lib  # noqa: E402


def ref_attention(qkv, kv, mask, heads):
    """qkv (B,T,3C), kv (B,S,2C) or None, mask (B,S) or None -> (B,T,C), fp32/64 torch."""
    B, T, C3 = qkv.shape
    Cc = C3 // 3
    d = Cc // heads
    q, k, v = qkv.split(Cc, dim=2)

    def heads_(x):
        return x.reshape(B, -1, heads, d).permute(0, 2, 1, 3)

    def attend(q, k, v, m):
        w = heads_(q) @ heads_(k).transpose(-1, -2) / math.sqrt(d)
        if m is not None:
            w = w.masked_fill(m[:, None, None, :] == 0, float("-inf"))
        return (torch.softmax(w, -1) @ heads_(v)).permute(0, 2, 1, 3).reshape(B, T, Cc)

    out = attend(q, k, v, None)
    oself = out
    if kv is not None:
        kc, vc = kv.split(Cc, dim=2)
        out = out + attend(q, kc, vc, mask)
    return out, oself


def run(B, T, S, Cc, heads=8, masked=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    qkv = (torch.randn(B, T, 3 * Cc, generator=g) * 0.7).half()
    kv = (torch.randn(B, S, 2 * Cc, generator=g) * 0.7).half() if S > 0 else None
    mask = None
    if masked and S > 0:
        mask = torch.ones(B, S)
        for i in range(B):
            mask[i, (S // 2 + i):] = 0
    dO = (torch.randn(B, T, Cc, generator=g) * 0.5).half()
    # reference in fp64 on the same fp16 inputs
    qr = qkv.double().requires_grad_(True)
    kr = kv.double().requires_grad_(True) if kv is not None else None
    out, oself = ref_attention(qr, kr, mask, heads)
    (out * dO.double()).sum().backward()

    dev = "cuda"
    qc, dOc = qkv.to(dev), dO.to(dev)
    kc = kv.to(dev) if kv is not None else None
    mc = mask.to(dev) if mask is not None else None
    h16 = torch.empty(B, T, Cc, device=dev, dtype=torch.float16)
    os16 = torch.empty_like(h16)
    stats = torch.empty(B, heads, 2, T, 2, device=dev)
    lib = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(lib.mdm_op_attention_fwd(P(qc), P(kc), P(mc), B, T, S, Cc, heads, P(h16), P(os16), P(stats), st), "attn fwd")
    Dterm = torch.empty(B, heads, 2, T, device=dev)
    dq32 = torch.empty(B, T, Cc, device=dev)
    dqkv = torch.zeros(B, T, 3 * Cc, device=dev, dtype=torch.float16)
    dkv = torch.zeros(B, max(S, 1), 2 * Cc, device=dev, dtype=torch.float16)
    _lib.check(lib.mdm_op_attention_bwd(P(qc), P(kc), P(mc), P(dOc), P(h16), P(os16), P(stats), B, T, S, Cc, heads,
                                        P(Dterm), P(dq32), P(dqkv), P(dkv) if S > 0 else None, st), "attn bwd")
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))

    errs = {"out": rel(h16, out.detach()), "oself": rel(os16, oself.detach()), "dqkv": rel(dqkv, qr.grad)}
    if S > 0:
        errs["dkv"] = rel(dkv, kr.grad)
    return errs


CASES = [
    ("t256_d96_s128", lambda: run(2, 256, 128, 768)),
    ("t1024_d64_s128", lambda: run(1, 1024, 128, 512)),
    ("t16_d8_s6_masked", lambda: run(2, 16, 6, 64, masked=True)),
    ("t200_d32_s77_masked", lambda: run(2, 200, 77, 256, masked=True)),
    ("t256_d64_nocross", lambda: run(2, 256, 0, 512)),
    ("t384_d96_s130", lambda: run(1, 384, 130, 768)),
]
TOL = 4e-3  # fp16 P / dS tiles and fp16 outputs

if __name__ == "__main__":
    bad = 0
    for name, fn in CASES:
        try:
            e = fn()
            ok = all(v <= TOL for v in e.values())
            print("PASS" if ok else "FAIL", name, e, flush=True)
            bad += 0 if ok else 1
        except Exception as ex:
            bad += 1
            print("ERROR", name, type(ex).__name__, ex, flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("sticky:", e2)
                break
    print("failures:", bad)
    sys.exit(1 if bad else 0)
