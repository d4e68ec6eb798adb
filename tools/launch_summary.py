#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-56s n=%4d total=%9.1f us  avg=%8.1f us  share=%5.1f%%" % (k[:56], n, t, t / n, 100 * t / tot))
    print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
