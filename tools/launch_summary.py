"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share, plus per-launch
durations grouped by grid (to tell the GEMM shapes apart).  Usage: python tools/launch_summary.py launches.csv [steps]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    grids = collections.defaultdict(lambda: [0, 0.0])
    for x in rows:
        name = re.sub(r"\(.*", "", x["Kernel Name"])[:60]
        v = float(x["Metric Value"].replace(",", ""))
        v = v / 1e3 if x["Metric Unit"] == "ns" else v * 1e3 if x["Metric Unit"] == "ms" else v
        agg[name][0] += 1; agg[name][1] += v
        g = (name, x["Grid Size"].replace(" ", ""))
        grids[g][0] += 1; grids[g][1] += v
    tot = sum(v[1] for v in agg.values())
    div = steps if steps else 1
    print("%-60s %6s %12s %7s" % ("kernel", "count", "us" + ("/step" if steps else ""), "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %6d %12.1f %6.1f%%" % (n, c, t / div, 100 * t / tot))
    print("\nper grid:")
    for (n, g), (c, t) in sorted(grids.items(), key=lambda kv: -kv[1][1])[:24]:
        print("%-44s %-16s %5d x %9.1f us = %10.1f us" % (n[:44], g, c, t / c, t / div))


main()
