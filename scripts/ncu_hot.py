"""Top stall sites of a kernel from an ncu report (SASS level): python scripts/ncu_hot.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(rows[hdr])}
data = rows[hdr + 1:]
stalls = [h for h in rows[hdr] if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print(rows[0][1][:120] if len(rows[0]) > 1 else "", "| total samples", tot, "| instructions", len(data))
agg = {}
for r in data:
    for h in stalls:
        agg[h[6:]] = agg.get(h[6:], 0) + int(r[ix[h]] or 0)
print("stall mix:", ", ".join(f"{k} {v * 100 / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:topn]
for i in sorted(order):
    r = data[i]
    n = int(r[ix["# Samples"]] or 0)
    top = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {r[ix['Source']][:72]:72s} {n:6d} {n * 100 / tot:5.1f}%  x{r[ix['Instructions Executed']]:>9s} {top}")
