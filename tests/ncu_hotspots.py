"""Aggregate an `ncu --page source --csv` dump: top SASS instructions by stall samples and stall-reason totals."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
i_src, i_samp = hdr.index("Source"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {hdr[i]: 0 for i in stall}
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[i_samp].isdigit(): continue
    n = int(r[i_samp] or 0)
    for i in stall: tot[hdr[i]] += int(r[i] or 0)
    data.append((n, r[i_src].strip(), {hdr[i]: int(r[i] or 0) for i in stall if int(r[i] or 0)}))
total = sum(d[0] for d in data)
print("total samples", total)
print({k: round(100 * v / max(total, 1), 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for idx in sorted(range(len(data)), key=lambda k: -data[k][0])[:top]:
    n, s, st = data[idx]
    print(f"{100*n/total:5.1f}%  #{idx:5d} {s[:70]:70s} {st}")
