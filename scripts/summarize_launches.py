"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections
import csv
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        ms = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0}.get(row["Metric Unit"], 1e-6) * v
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    lines = ["| kernel | launches | total ms | share | avg ms |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {v[1] / tot * 100:.1f}% | {v[1] / v[0]:.3f} |")
    lines.append(f"| total | {sum(v[0] for v in agg.values())} | {tot:.3f} | 100% | |")
    text = "\n".join(lines)
    print(text)
    if out:
        with open(out, "a") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:3])
