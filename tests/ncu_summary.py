"""Summarise `ncu --set full` captures (gpurun_out/*.ncu-rep) into a table for profiles/: per kernel launch the
duration, DRAM bytes read + written (-> achieved HBM GB/s), tensor-pipe activity, occupancy, registers.

    python tests/ncu_summary.py gpurun_out/prof_a.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_summary.txt
    python tests/ncu_summary.py --traffic-json KERNEL_REGEX gpurun_out/prof_conv.ncu-rep  -> profiles/traffic_top_kernel.json
(the JSON is what bench.py reports as roofline.traffic: measured, never a literal in the source)"""
import csv
import io
import json
import re
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "ns",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "launch__registers_per_thread": "regs",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
        "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"name": re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("mdm::<unnamed>::", "")}
        for m, short in WANT.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    d[short] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1)
                except ValueError:
                    pass
        out.append(d)
    return out


def main():
    args = sys.argv[1:]
    traffic = None
    if args and args[0] == "--traffic-json":
        traffic, args = args[1], args[2:]
    rows = []
    for rep in args:
        rows += [dict(r, rep=rep.split("/")[-1]) for r in load(rep)]
    if traffic is not None:
        m = [r for r in rows if re.search(traffic, r["name"]) and "rd" in r]
        best = max(m, key=lambda r: r["ns"])
        print(json.dumps({"kernel": best["name"], "dram_bytes_per_launch": best["rd"] + best["wr"],
                          "dram_read": best["rd"], "dram_write": best["wr"], "duration_us": best["ns"] / 1e3,
                          "achieved_gbs": round((best["rd"] + best["wr"]) / best["ns"], 1),
                          "note": f"ncu --set full, {best['rep']}: longest launch matching /{traffic}/; "
                                  "dram__bytes_read.sum + dram__bytes_write.sum of that launch"}))
        return
    print(f"{'kernel':46s} {'grid':>8s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>7s} {'dram%':>6s} {'tensor%':>8s} {'occ%':>6s} {'regs':>5s}")
    for r in rows:
        mb = (r.get("rd", 0) + r.get("wr", 0)) / 1e6
        gbs = (r.get("rd", 0) + r.get("wr", 0)) / max(r.get("ns", 1), 1)
        print(f"{r['name'][:46]:46s} {int(r.get('grid', 0)):8d} {r.get('ns', 0) / 1e3:9.1f} {mb:9.1f} {gbs:7.0f} "
              f"{r.get('dram%', 0):6.1f} {r.get('tensor%', 0):8.1f} {r.get('occ%', 0):6.1f} {int(r.get('regs', 0)):5d}")


if __name__ == "__main__":
    main()
