#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples) of an `ncu --page source --csv --print-source cuda,sass` dump."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    ie, ss, te = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Thread Instructions Executed")
    agg = collections.OrderedDict()
    cur = None
    fname = ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if len(r) < ie + 1 or r[0] == "Line No":
            continue
        if r[0]:
            cur = (fname, r[0], r[1].strip())
        if cur is None or not r[ie].isdigit():
            continue
        a = agg.setdefault(cur, [0, 0, 0])
        a[0] += int(r[ie])
        a[1] += int(r[ss]) if r[ss].isdigit() else 0
        a[2] += int(r[te]) if r[te].isdigit() else 0
    tot = sum(a[0] for a in agg.values())
    tots = sum(a[1] for a in agg.values())
    tt = sum(a[2] for a in agg.values())
    print("warp instructions %d, thread instructions %d (%.1f active lanes), stall samples %d" % (tot, tt, tt / max(tot, 1), tots))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print("%5.1f%% inst %5.1f%% stall lanes %4.1f | %s:%s %s" % (100 * a[0] / tot, 100 * a[1] / max(tots, 1), a[2] / max(a[0], 1), k[0], k[1], k[2][:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
