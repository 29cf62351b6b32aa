#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean time and share.
usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/r01_launches.txt"""
import collections
import csv
import sys


def main():
    lines = [ln for ln in open(sys.argv[1]) if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        agg.setdefault(row["Kernel Name"].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {sys.argv[1]}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)")
    print(f"{'kernel':72s} {'launches':>8s} {'mean us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:72s} {len(v):8d} {sum(v) / len(v):10.1f} {sum(v) / tot:7.3f}")


if __name__ == "__main__":
    main()
