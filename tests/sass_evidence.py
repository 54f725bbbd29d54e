"""Counts the Blackwell-specific SASS instructions per kernel in the built objects (development aid; the
output is kept in profiles/). UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store /
reduce, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU"]
rows = []
for obj in ["gemm_tc", "gemm_persistent", "attention", "kernels", "diffusion", "optim"]:
    path = os.path.join(ROOT, "build", obj + ".o")
    if not os.path.exists(path):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = fn.replace("mdm::(anonymous namespace)::", "").replace("void ", "")
            fn = re.sub(r"\(.*", "", fn)
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            op = m.group(1)
            if op in OPS:
                counts[fn][op] += 1
    for fn, c in counts.items():
        if any(c[o] for o in OPS[:-1]):
            rows.append((obj, fn, c))
print(f"{'object':16s} {'kernel':44s} " + " ".join(f"{o:>8s}" for o in OPS))
for obj, fn, c in rows:
    print(f"{obj:16s} {fn[:44]:44s} " + " ".join(f"{c[o]:8d}" for o in OPS))
