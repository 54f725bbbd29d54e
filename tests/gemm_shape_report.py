"""Per-shape efficiency report of the tcgen05 engine over one training step (development aid).
usage: python tests/gemm_shape_report.py [cfg] [batch]  -> gpurun_out/gemm_shapes.csv + summary"""
import collections, csv, ctypes as C, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import torch
import bench
from mdm_b200 import _lib

def flops(r):
    kind, M, N, K, nz, kb = int(r["kind"]), int(r["M"]), int(r["N"]), int(r["K"]), int(r["nz"]), int(r["kblocks"])
    if kind == 0:
        return 2.0 * M * N * K * nz
    if kind == 1:   # conv fwd/dgrad: pixels x N x 9K
        return 2.0 * int(r["nimg"]) * int(r["H"]) * int(r["W"]) * N * K * 9
    return 2.0 * M * N * 9 * int(r["nimg"]) * int(r["H"]) * int(r["W"])   # wgrad (all taps)

def main(cfg="cc12m_64x64", B=64):
    dev = torch.device("cuda", 0)
    pipe, nested = bench.build_pipeline(cfg, dev)
    pipe.train()
    host = bench.synthetic_host_batch(cfg, B, 1234)
    sample = {k: v.to(dev) for k, v in host.items()}
    vm = pipe.get_model().vision_model
    def step():
        loss, *_ = pipe.get_loss(sample); loss.mean().backward(); vm.zero_grad(set_to_none=True)
    for _ in range(2): step()
    torch.cuda.synchronize()
    lib = _lib.lib()
    lib.mdm_profile_gemm(1); step(); torch.cuda.synchronize(); lib.mdm_profile_gemm(0)
    os.makedirs(os.path.join(HERE, "..", "gpurun_out"), exist_ok=True)
    path = os.path.join(HERE, "..", "gpurun_out", "gemm_shapes.csv")
    lib.mdm_profile_dump(path.encode())
    tot = C.c_double(); n = C.c_longlong(); lib.mdm_profile_read(C.byref(tot), C.byref(n))
    rows = list(csv.DictReader(open(path)))
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
    for r in rows:
        key = (r["kind"], r["majors"], r["M"], r["N"], r["K"], r["block_n"], r["nz"], r["nsplit"], r["H"])
        agg[key][0] += float(r["ms"]); agg[key][1] += flops(r); agg[key][2] += 1
    T = sum(v[0] for v in agg.values()); F = sum(v[1] for v in agg.values())
    print(f"total {T:.2f} ms, {F/1e12:.1f} TFLOP issued, {F/T/1e9:.0f} TFLOP/s avg over {len(rows)} launches")
    bucket = collections.defaultdict(lambda: [0.0, 0])
    for r in rows:
        n = int(r["N"])
        b = "N<=32" if n <= 32 else ("N<=64" if n <= 64 else ("N<=128" if n <= 128 else "N>128"))
        kind = {"0": "plain", "1": "conv", "2": "wgrad"}[r["kind"]]
        bucket[(kind, b, "persistent" if int(r["majors"]) & 4 else "plain-launch")][0] += float(r["ms"])
        bucket[(kind, b, "persistent" if int(r["majors"]) & 4 else "plain-launch")][1] += 1
    for k, (ms, cnt) in sorted(bucket.items()):
        print(f"   bucket {k}: {ms:8.3f} ms over {cnt} launches")
    print("  ms     n   TF/s  lost_ms(@1300)  kind maj M N K bn nz split H")
    for k, (ms, fl, cnt) in sorted(agg.items(), key=lambda kv: -(kv[1][0] - kv[1][1] / 1.3e12))[:int(os.environ.get('MDM_REPORT_TOP', '25'))]:
        print(f"{ms:7.2f} {cnt:4d} {fl/ms/1e9:6.0f} {ms - fl/1.3e12:8.2f}   {' '.join(k)}")

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cc12m_64x64", int(sys.argv[2]) if len(sys.argv) > 2 else 64)
