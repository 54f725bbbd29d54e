"""Parity cases for the tcgen05 GEMM engine (mdm_gemm_raw), checked against torch fp32 on the same
fp16 inputs.  Used by tests/test_gemm_gpu.py and runnable as a script (prints every case)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ml-mdm_b200"))
from mdm_b200 import _lib  # noqa: E402

DEV = "cuda"


def _stream():
    return torch.cuda.current_stream().cuda_stream


def run_plain(M, N, K, a_mn, b_mn, block_n, nz1=1, nz2=1, bias=False, residual=False, act=False,
              f16_out=False, nsplit=1, alpha=1.0, seed=0, ggrad=False, f32_out=True):
    """C[z2,z1] = alpha * A @ B^T (+bias) (+residual); operands stored K-major or MN-major."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    nb = nz1 * nz2
    A = (torch.randn(nb, M, K, generator=g) * 0.5).to(torch.float16).to(DEV)
    B = (torch.randn(nb, N, K, generator=g) * 0.5).to(torch.float16).to(DEV)
    ref = alpha * torch.matmul(A.float(), B.float().transpose(1, 2))  # (nb, M, N)
    A_st = A.transpose(1, 2).contiguous() if a_mn else A  # MN-major: stored [K][M]
    B_st = B.transpose(1, 2).contiguous() if b_mn else B
    if a_mn:
        sa = _lib.tmap(A_st.data_ptr(), (M, K, nz1, nz2), (1, M, K * M, K * M * nz1), (64, 64, 1, 1))
    else:
        sa = _lib.tmap(A_st.data_ptr(), (K, M, nz1, nz2), (1, K, K * M, K * M * nz1), (64, 128, 1, 1))
    if b_mn:
        sb = _lib.tmap(B_st.data_ptr(), (N, K, nz1, nz2), (1, N, K * N, K * N * nz1), (64, 64, 1, 1))
    else:
        sb = _lib.tmap(B_st.data_ptr(), (K, N, nz1, nz2), (1, K, K * N, K * N * nz1),
                       (64, block_n, 1, 1))
    p = _lib.GemmParams()
    p.kind = 0
    p.M, p.N, p.K = M, N, K
    p.block_n = block_n
    p.nz1, p.nz2, p.nsplit = nz1, nz2, nsplit
    p.a_use_z = p.b_use_z = 1
    p.num_kblocks = (K + 63) // 64
    p.alpha = alpha
    p.ldc = N
    p.c_z1_stride = M * N
    p.c_z2_stride = M * N * nz1
    bias_t = res_t = None
    if bias:
        bias_t = torch.randn(N, generator=g).to(DEV)
        p.bias = bias_t.data_ptr()
        ref = ref + bias_t
    if residual:
        res_t = torch.randn(nb, M, N, generator=g).to(DEV)
        p.residual = res_t.data_ptr()
        ref = ref + res_t
    if ggrad:  # FFN backward epilogue: result *= gelu'(src), src indexed like the output
        src = (torch.randn(nb, M, N, generator=g) * 1.5).to(torch.float16).to(DEV)
        p.gelu_grad_src = src.data_ptr()
        x = src.float()
        ref = ref * (0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5)
    out32 = torch.zeros(nb, M, N, device=DEV)
    if f32_out:
        p.out_f32 = out32.data_ptr()
    out16 = outact = None
    if f16_out:
        out16 = torch.zeros(nb, M, N, device=DEV, dtype=torch.float16)
        p.out_f16 = out16.data_ptr()
    if act:
        outact = torch.zeros(nb, M, N, device=DEV, dtype=torch.float16)
        p.out_act_f16 = outact.data_ptr()
        p.act = 1
    if nsplit > 1:
        p.atomic = 1
    _lib.gemm_raw(sa, sb, a_mn, b_mn, p, _stream())
    torch.cuda.synchronize()
    errs = {"f32": rel_err(out32, ref)} if f32_out else {}
    if f16_out:
        errs["f16"] = rel_err(out16.float(), ref)
    if act:
        errs["act"] = rel_err(outact.float(), torch.nn.functional.gelu(ref))
    return errs


def rel_err(x, ref):
    return float((x - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def pack_conv_weight(w):
    """OIHW fp32 -> [Cout][tap][Cin] fp16 (the layout the conv kernels read)."""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).contiguous().to(torch.float16)


def conv_geometry(H, W):
    PW = 16 if W >= 16 else 8
    return PW


def run_conv_fwd(nimg, H, W, Cin, Cout, block_n, bias=True, seed=0, residual=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(nimg, H, W, Cin, generator=g) * 0.5).to(torch.float16).to(DEV)  # NHWC
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1)
    wp = pack_conv_weight(w).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV) if bias else None
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wp.float().reshape(Cout, 3, 3, Cin)
                                     .permute(0, 3, 1, 2), b, padding=1).permute(0, 2, 3, 1)
    PW = conv_geometry(H, W)
    PH = 128 // PW
    sa = _lib.tmap(x.data_ptr(), (Cin, W, H, nimg), (1, Cin, W * Cin, H * W * Cin), (64, PW, PH, 1))
    sb = _lib.tmap(wp.data_ptr(), (Cin, Cout, 9, 1), (1, 9 * Cin, Cin, 9 * Cin * Cout),
                   (64, block_n, 1, 1))
    p = _lib.GemmParams()
    p.kind = 1
    p.N, p.K = Cout, Cin
    p.block_n = block_n
    p.H, p.W, p.PW, p.PH = H, W, PW, PH
    p.tiles_w, p.tiles_h, p.nimg = (W + PW - 1) // PW, (H + PH - 1) // PH, nimg
    p.taps = 9
    p.kblocks_c = (Cin + 63) // 64
    p.num_kblocks = 9 * p.kblocks_c
    p.alpha = 1.0
    p.ldc = Cout
    if b is not None:
        p.bias = b.data_ptr()
    out = torch.zeros(nimg, H, W, Cout, device=DEV)
    p.out_f32 = out.data_ptr()
    if residual:
        res = torch.randn(nimg, H, W, Cout, generator=g).to(DEV)
        p.residual = res.data_ptr()
        ref = ref + res
    _lib.gemm_raw(sa, sb, 0, 0, p, _stream())
    torch.cuda.synchronize()
    return {"f32": rel_err(out, ref)}


def run_conv_dgrad(nimg, H, W, Cin, Cout, block_n, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    dy = (torch.randn(nimg, H, W, Cout, generator=g) * 0.5).to(torch.float16).to(DEV)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1)
    wp = pack_conv_weight(w).to(DEV)
    w_oihw = wp.float().reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    ref = torch.nn.grad.conv2d_input((nimg, Cin, H, W), w_oihw, dy.float().permute(0, 3, 1, 2),
                                     padding=1).permute(0, 2, 3, 1)
    PW = conv_geometry(H, W)
    PH = 128 // PW
    sa = _lib.tmap(dy.data_ptr(), (Cout, W, H, nimg), (1, Cout, W * Cout, H * W * Cout), (64, PW, PH, 1))
    # B = W viewed MN-major: N' = ci (contiguous), K' = co (rows), z1 = tap
    sb = _lib.tmap(wp.data_ptr(), (Cin, Cout, 9, 1), (1, 9 * Cin, Cin, 9 * Cin * Cout), (64, 64, 1, 1))
    p = _lib.GemmParams()
    p.kind = 1
    p.N, p.K = Cin, Cout
    p.block_n = block_n
    p.H, p.W, p.PW, p.PH = H, W, PW, PH
    p.tiles_w, p.tiles_h, p.nimg = (W + PW - 1) // PW, (H + PH - 1) // PH, nimg
    p.taps = 9
    p.flip = 1
    p.kblocks_c = (Cout + 63) // 64
    p.num_kblocks = 9 * p.kblocks_c
    p.alpha = 1.0
    p.ldc = Cin
    out = torch.zeros(nimg, H, W, Cin, device=DEV)
    p.out_f32 = out.data_ptr()
    _lib.gemm_raw(sa, sb, 0, 1, p, _stream())
    torch.cuda.synchronize()
    return {"f32": rel_err(out, ref)}


def run_conv_wgrad(nimg, H, W, Cin, Cout, block_n, nsplit=1, seed=0, kfactor=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(nimg, H, W, Cin, generator=g) * 0.5).to(torch.float16).to(DEV)
    dy = (torch.randn(nimg, H, W, Cout, generator=g) * 0.5).to(torch.float16).to(DEV)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, 3, 3),
                                      dy.float().permute(0, 3, 1, 2), padding=1)  # OIHW
    ref = ref.permute(0, 2, 3, 1).reshape(Cout, 9, Cin)
    PW = conv_geometry(H, W)
    PH = 64 * kfactor // PW
    sa = _lib.tmap(dy.data_ptr(), (Cout, W, H, nimg), (1, Cout, W * Cout, H * W * Cout), (64, PW, PH, 1))
    sb = _lib.tmap(x.data_ptr(), (Cin, W, H, nimg), (1, Cin, W * Cin, H * W * Cin), (64, PW, PH, 1))
    p = _lib.GemmParams()
    p.kind = 2
    p.kfactor = kfactor
    p.M, p.N = Cout, Cin
    p.block_n = block_n
    p.H, p.W, p.PW, p.PH = H, W, PW, PH
    p.tiles_w, p.tiles_h, p.nimg = (W + PW - 1) // PW, (H + PH - 1) // PH, nimg
    p.taps = 9
    p.nz1 = 9
    p.nsplit = nsplit
    p.num_kblocks = nimg * p.tiles_w * p.tiles_h
    p.alpha = 1.0
    p.ldc = 9 * Cin
    p.c_z1_stride = Cin
    p.atomic = 1 if nsplit > 1 else 0
    out = torch.zeros(Cout, 9, Cin, device=DEV)
    p.out_f32 = out.data_ptr()
    _lib.gemm_raw(sa, sb, 1, 1, p, _stream())
    torch.cuda.synchronize()
    return {"f32": rel_err(out, ref)}


CASES = [
    ("kk_256", lambda: run_plain(256, 256, 256, 0, 0, 256)),
    ("kk_bias_128", lambda: run_plain(128, 128, 64, 0, 0, 128, bias=True)),
    ("kk_bigK", lambda: run_plain(256, 512, 4608, 0, 0, 256, bias=True)),
    ("kk_n64", lambda: run_plain(384, 64, 192, 0, 0, 64)),
    ("kk_n16", lambda: run_plain(256, 16, 128, 0, 0, 16)),
    ("kk_epilogue", lambda: run_plain(256, 256, 128, 0, 0, 128, bias=True, residual=True, act=True,
                                      f16_out=True, alpha=0.5)),
    ("kk_batched", lambda: run_plain(128, 128, 128, 0, 0, 128, nz1=3, nz2=2)),
    ("kk_ragged", lambda: run_plain(200, 80, 136, 0, 0, 80, bias=True)),
    ("kk_smallM", lambda: run_plain(5, 96, 64, 0, 0, 96, bias=True)),
    ("kmn_256", lambda: run_plain(256, 256, 256, 0, 1, 256)),
    ("kmn_n96", lambda: run_plain(256, 96, 256, 0, 1, 96)),
    ("mnmn_256", lambda: run_plain(256, 256, 256, 1, 1, 256)),
    ("mnmn_split", lambda: run_plain(256, 128, 1024, 1, 1, 128, nsplit=4)),
    ("mnk_128", lambda: run_plain(256, 128, 256, 1, 0, 128)),
    ("mnmn_ragged", lambda: run_plain(200, 96, 328, 1, 1, 96)),
    ("kk_multitile_epi", lambda: run_plain(300, 640, 192, 0, 0, 256, bias=True, residual=True, act=True, f16_out=True)),
    ("kk_batched_epi", lambda: run_plain(130, 96, 64, 0, 0, 96, nz1=4, nz2=2, f16_out=True, residual=True)),
    ("conv_fwd_16", lambda: run_conv_fwd(2, 16, 16, 128, 128, 128)),
    ("conv_fwd_ragged_hw", lambda: run_conv_fwd(3, 24, 24, 64, 192, 64)),
    ("conv_fwd_32", lambda: run_conv_fwd(1, 32, 32, 64, 256, 256)),
    ("conv_fwd_8", lambda: run_conv_fwd(2, 8, 8, 64, 64, 64)),
    ("conv_fwd_c96", lambda: run_conv_fwd(1, 16, 16, 96, 64, 64)),
    ("conv_dgrad_16", lambda: run_conv_dgrad(2, 16, 16, 128, 128, 128)),
    ("conv_dgrad_mix", lambda: run_conv_dgrad(1, 32, 32, 64, 192, 64)),
    ("conv_wgrad_16", lambda: run_conv_wgrad(2, 16, 16, 128, 128, 128)),
    ("conv_wgrad_split", lambda: run_conv_wgrad(4, 32, 32, 64, 256, 64, nsplit=4)),
    # >= 4 m tiles and >= 128 columns: clusters of 2 CTAs share the B tile by TMA multicast (MDM_GEMM_CLUSTER)
    ("mc_kk_1024", lambda: run_plain(1024, 256, 512, 0, 0, 256, bias=True)),
    ("mc_kk_ragged_m", lambda: run_plain(640 + 37, 384, 256, 0, 0, 128, bias=True, residual=True, f16_out=True)),
    ("mc_kmn_1024", lambda: run_plain(1024, 256, 512, 0, 1, 256)),
    ("mc_kmn_bn128", lambda: run_plain(768, 128, 320, 0, 1, 128)),
    ("mc_mnmn_split", lambda: run_plain(1024, 256, 2048, 1, 1, 256, nsplit=4)),
    ("mc_kk_batched", lambda: run_plain(512, 256, 128, 0, 0, 256, nz1=3, nz2=2)),
    ("mc_conv_fwd", lambda: run_conv_fwd(4, 32, 32, 128, 256, 256)),
    ("mc_conv_fwd_odd_tiles", lambda: run_conv_fwd(3, 24, 40, 64, 128, 128)),
    ("mc_conv_dgrad", lambda: run_conv_dgrad(4, 32, 32, 256, 128, 256)),
    ("mc_conv_dgrad_bn128", lambda: run_conv_dgrad(2, 32, 32, 128, 192, 128)),
    ("mc_conv_wgrad", lambda: run_conv_wgrad(4, 32, 32, 256, 512, 256, nsplit=4)),
    # long contractions on wide tiles: the one-tile kernel's CTA pairs (cta_group::2, M = 256 instructions; MDM_GEMM_PAIR)
    ("pair_kk_bigK", lambda: run_plain(512, 512, 4608, 0, 0, 256, bias=True)),
    ("pair_kk_bn192_odd_m", lambda: run_plain(128 * 3 + 5, 384, 3200, 0, 0, 192, bias=True, residual=True, f16_out=True)),
    ("pair_kk_bn128", lambda: run_plain(512, 256, 3200, 0, 0, 128, act=True, f16_out=True)),
    ("pair_kmn_bigK", lambda: run_plain(512, 512, 3200, 0, 1, 256)),
    ("pair_kmn_bn192", lambda: run_plain(512, 384, 3200, 0, 1, 192)),
    ("pair_mnk_bigK", lambda: run_plain(512, 256, 3200, 1, 0, 256)),
    ("pair_kk_batched_bigK", lambda: run_plain(256, 256, 3200, 0, 0, 256, nz1=2, nz2=2)),
    ("pair_conv_fwd_768", lambda: run_conv_fwd(4, 16, 16, 768, 768, 256)),
    ("pair_conv_fwd_768_bn192_odd", lambda: run_conv_fwd(3, 8, 24, 768, 768, 192)),
    ("pair_conv_dgrad_768", lambda: run_conv_dgrad(4, 16, 16, 768, 768, 256)),
    ("pair_conv_wgrad_768", lambda: run_conv_wgrad(8, 16, 16, 768, 768, 256, nsplit=4)),
    ("pair_conv_wgrad_384_bn128", lambda: run_conv_wgrad(4, 16, 16, 128, 384, 128, nsplit=2)),
    ("pair_conv_dgrad_768_bn192", lambda: run_conv_dgrad(3, 8, 24, 768, 768, 192)),
    ("pair_conv_wgrad_cin192", lambda: run_conv_wgrad(4, 16, 16, 192, 256, 192, nsplit=2)),
    ("pair_conv_wgrad_long", lambda: run_conv_wgrad(16, 16, 16, 256, 256, 256, nsplit=1)),
    # persistent form with the epilogue operand (fp32 residual / fp16 GELU' source) TMA-loaded into the staging tile
    ("op_res_many_tiles", lambda: run_plain(128 * 80 + 3, 768, 256, 0, 0, 192, bias=True, residual=True)),
    ("op_res_partial_n", lambda: run_plain(128 * 9, 320 + 40, 128, 0, 0, 256, bias=True, residual=True, f16_out=True)),
    ("op_res_bn96_batched", lambda: run_plain(260, 96, 192, 0, 0, 96, nz1=3, nz2=2, residual=True, act=True, f16_out=True)),
    ("op_res_conv", lambda: run_conv_fwd(6, 32, 32, 128, 256, 256, residual=True)),
    ("op_res_conv_ragged", lambda: run_conv_fwd(3, 24, 40, 64, 192, 192, residual=True)),
    ("op_ggrad_f16", lambda: run_plain(128 * 80, 768, 256, 0, 1, 192, f16_out=True, ggrad=True, f32_out=False)),
    ("op_ggrad_f16_f32_ragged", lambda: run_plain(128 * 5 + 77, 640 + 24, 192, 0, 1, 256, f16_out=True, ggrad=True)),
    ("ggrad_one_tile", lambda: run_plain(512, 256, 3200, 0, 1, 256, f16_out=True, ggrad=True)),
    # narrow layers (the 32/64-channel levels of the 256/1024-px nests): 256-pixel stages, one A slab when M <= 64
    ("conv_wgrad_tall_c32", lambda: run_conv_wgrad(2, 64, 64, 32, 32, 32, nsplit=4, kfactor=4)),
    ("conv_wgrad_tall_c64", lambda: run_conv_wgrad(3, 32, 48, 64, 64, 64, nsplit=3, kfactor=4)),
    ("conv_wgrad_tall_c32_64", lambda: run_conv_wgrad(1, 48, 32, 32, 64, 32, nsplit=1, kfactor=4)),
    ("conv_wgrad_tall_c96_32", lambda: run_conv_wgrad(2, 32, 32, 32, 96, 32, nsplit=2, kfactor=2)),
]

TOL = {"f32": 2e-5, "f16": 1.5e-3, "act": 1.5e-3}


def bench_one(M, N, K, iters=20, bn=256):
    A = torch.randn(M, K, device=DEV).to(torch.float16)
    B = torch.randn(N, K, device=DEV).to(torch.float16)
    out = torch.zeros(M, N, device=DEV, dtype=torch.float16)
    sa = _lib.tmap(A.data_ptr(), (K, M, 1, 1), (1, K, K * M, K * M), (64, 128, 1, 1))
    sb = _lib.tmap(B.data_ptr(), (K, N, 1, 1), (1, K, K * N, K * N), (64, bn, 1, 1))
    p = _lib.GemmParams()
    p.kind = 0
    p.M, p.N, p.K = M, N, K
    p.block_n = bn
    p.num_kblocks = K // 64
    p.alpha = 1.0
    p.ldc = N
    p.out_f16 = out.data_ptr()
    for _ in range(3):
        _lib.gemm_raw(sa, sb, 0, 0, p, _stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _lib.gemm_raw(sa, sb, 0, 0, p, _stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        torch.matmul(A, B.t())
    t0.record()
    for _ in range(iters):
        torch.matmul(A, B.t())
    t1.record()
    torch.cuda.synchronize()
    ms_ref = t0.elapsed_time(t1) / iters
    fl = 2.0 * M * N * K
    return ms, fl / ms / 1e9, ms_ref, fl / ms_ref / 1e9


def bench_epi(M, N, K, bn, mode, iters=20, b_mn=False):
    """Epilogue variants of one plain GEMM: f16 | f32 | gelu (bias, pre-activation + GELU copies) | ggrad."""
    A = torch.randn(M, K, device=DEV).to(torch.float16)
    sa = _lib.tmap(A.data_ptr(), (K, M, 1, 1), (1, K, K * M, K * M), (64, 128, 1, 1))
    if b_mn:
        B = torch.randn(K, N, device=DEV).to(torch.float16)
        sb = _lib.tmap(B.data_ptr(), (N, K, 1, 1), (1, N, N * K, N * K), (64, 64, 1, 1))
    else:
        B = torch.randn(N, K, device=DEV).to(torch.float16)
        sb = _lib.tmap(B.data_ptr(), (K, N, 1, 1), (1, K, K * N, K * N), (64, bn, 1, 1))
    out = torch.zeros(M, N, device=DEV, dtype=torch.float16)
    out2 = torch.zeros(M, N, device=DEV, dtype=torch.float16)
    out32 = torch.zeros(M, N, device=DEV, dtype=torch.float32) if mode == "f32" else None
    bias = torch.randn(N, device=DEV)
    src = torch.randn(M, N, device=DEV).to(torch.float16)
    p = _lib.GemmParams()
    p.kind = 0
    p.M, p.N, p.K = M, N, K
    p.block_n = bn
    p.num_kblocks = K // 64
    p.alpha = 1.0
    p.ldc = N
    if mode == "f32":
        p.out_f32 = out32.data_ptr()
    else:
        p.out_f16 = out.data_ptr()
    if mode == "gelu":
        p.bias = bias.data_ptr()
        p.out_act_f16 = out2.data_ptr()
        p.act = 1
    if mode == "ggrad":
        p.gelu_grad_src = src.data_ptr()
    for _ in range(3):
        _lib.gemm_raw(sa, sb, 0, 1 if b_mn else 0, p, _stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _lib.gemm_raw(sa, sb, 0, 1 if b_mn else 0, p, _stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K / ms / 1e9


def ncu_target():
    """Three launches of the largest 64x64-level conv of the cc12m_64x64 U-Net at batch 64
    (3x3, 256->256 @ 64x64: M = 262144 pixels, N = 256, K = 2304) for an `ncu --set full` capture."""
    for _ in range(3):
        run_conv_fwd(64, 64, 64, 256, 256, 256, bias=True)


if __name__ == "__main__":
    if "--ncu-conv" in sys.argv:
        ncu_target()
        sys.exit(0)
    if "--bench-epi" in sys.argv:
        for (M, N, K) in [(16384, 3072, 768), (16384, 768, 3072), (16384, 768, 768), (65536, 2048, 512)]:
            for mode, b_mn in (("f16", False), ("f32", False), ("gelu", False), ("ggrad", True), ("f16", True)):
                for bn in (256, 192, 128):
                    ms, tf = bench_epi(M, N, K, bn, mode, b_mn=b_mn)
                    print(f"EPI {M}x{N}x{K} {mode:5s} b_mn={int(b_mn)} bn={bn}: {ms*1e3:7.1f} us {tf:6.0f} TFLOP/s", flush=True)
        sys.exit(0)
    bad = 0
    for name, fn in CASES:
        try:
            errs = fn()
            ok = all(v <= TOL[k] for k, v in errs.items())
            print(f"{'PASS' if ok else 'FAIL'} {name}: {errs}", flush=True)
            bad += 0 if ok else 1
        except Exception as e:  # keep going: one call should report every case
            bad += 1
            print(f"ERROR {name}: {type(e).__name__}: {e}", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("  device error is sticky:", e2, flush=True)
                break
    if bad == 0 or "--bench" in sys.argv:
        bn = int(os.environ.get("MDM_BENCH_BN", "256"))
        for (M, N, K) in [(8192, 768, 6912), (16384, 256, 2304), (8192, 8192, 8192), (16384, 3072, 768)]:
            try:
                ms, tf, msr, tfr = bench_one(M, N, K, bn=bn)
                print(f"BENCH {M}x{N}x{K}: ours {ms:.3f} ms {tf:.0f} TFLOP/s | cuBLAS {msr:.3f} ms {tfr:.0f} TFLOP/s",
                      flush=True)
            except Exception as e:
                print("BENCH error", e, flush=True)
    print("failures:", bad)
    sys.exit(1 if bad else 0)
