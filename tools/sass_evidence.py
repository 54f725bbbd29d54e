"""Per-kernel counts of the Blackwell-only SASS mnemonics in the built library (cuobjdump -sass), the evidence that the
contraction kernels are tcgen05 / TMA / TMEM code: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG /
UTMAREDG = TMA tensor load / store / reduce-add, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = 1-D bulk
copy.  usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ml-mdm_b200", "mdm_b200", "libmdm_b200.so")
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "UTCBAR",
        "UTCBAR.2CTA", "SYNCS", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    counts, order, cur, i = {}, [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\((?!bool).*", "", names[i].replace("(bool)", "").replace("(int)", "")).replace("void ", "").replace("mdm::(anonymous namespace)::", "")
            i += 1
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base in ("UTCHMMA", "UTMALDG", "UTCBAR"):
            counts[cur][base + (".2CTA" if ".2CTA" in op else "")] += 1
        elif base in ("UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "SYNCS", "MUFU"):
            counts[cur][base] += 1
    print("cuobjdump -sass ml-mdm_b200/mdm_b200/libmdm_b200.so (sm_100a), instruction counts per kernel; kernels without any of"
          " these mnemonics omitted")
    print(f"{'kernel':62s}" + "".join(f"{c:>13s}" for c in COLS))
    for k in order:
        c = counts[k]
        if not any(c[x] for x in COLS[:11]):
            continue
        print(f"{k[:61]:62s}" + "".join(f"{c[x]:13d}" for x in COLS))


if __name__ == "__main__":
    sys.exit(main())
