#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launch count, total device time and share of the run.  Usage: summarize_launches.py in.csv > out.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        tot[name][0] += 1
        tot[name][1] += v
    total = sum(v[1] for v in tot.values())
    print(f"source: {path}  (ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / total:.1f}% |")
    print(f"| **all** | {sum(v[0] for v in tot.values())} | {total:.3f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1])
