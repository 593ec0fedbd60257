#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        agg.setdefault(x["Kernel Name"].split("(")[0][:48], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':50s} {'n':>4s} {'mean_us':>9s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:50s} {len(v):4d} {sum(v)/len(v):9.1f} {100*sum(v)/tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
