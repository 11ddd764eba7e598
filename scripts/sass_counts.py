"""SASS instruction counts per kernel of the built library (profiles/*_sass_counts.md):
    cuobjdump -sass avici_b200/csrc/libavici_b200.so | python scripts/sass_counts.py > profiles/rXX_sass_counts.md"""
import collections
import re
import subprocess
import sys

keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "MUFU.EX2", "SYNCS", "FFMA2", "FADD2", "F2FP",
        "LDG", "STG", "LDS", "STS", "LD.E", "USETMAXREG"]
cur, counts, tot = None, collections.OrderedDict(), collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); tot[cur] = 0
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        tot[cur] += 1
        body = line.split("*/", 1)[1]
        for k in keys:
            if re.search(r"(^|[\s@!P0-9])" + re.escape(k), body):
                counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.split("\n")
print("# SASS instruction counts per kernel (libavici_b200.so, sm_100a)\n")
print("`cuobjdump -sass avici_b200/csrc/libavici_b200.so | python scripts/sass_counts.py`; UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, "
      "LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = 1-D bulk copy, SYNCS = mbarrier ops, "
      "FFMA2 / FADD2 = packed f32x2, LD.E = generic loads.  Kernels under 60 instructions omitted.\n")
print("| kernel | instrs | " + " | ".join(keys) + " |")
print("|---|---:|" + "---:|" * len(keys))
for (f, c), name in zip(counts.items(), names):
    if tot[f] < 60:
        continue
    name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("avb::", "")
    name = re.sub(r"\(.*", "", name)
    print(f"| `{name}` | {tot[f]} | " + " | ".join(str(c.get(k, 0)) for k in keys) + " |")
