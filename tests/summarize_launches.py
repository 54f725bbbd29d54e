"""Summarise an ncu --csv gpu__time_duration launch list by kernel name (shares of the step)."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for row in r:
    try:
        v = float(row[vi].replace(",", ""))
    except ValueError:
        continue
    u = row[ui]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
    name = re.sub(r"\(.*", "", row[ki])
    name = re.sub(r"^void ", "", name)
    agg[name][0] += ns
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"total {tot/1e6:.2f} ms over {sum(v[1] for v in agg.values())} launches")
for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{ns/1e6:9.3f} ms {100*ns/tot:5.1f}% {n:6d}  {k[:110]}")
