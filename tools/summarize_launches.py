#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one line per launch
(short kernel name, grid, duration) and totals per kernel family."""
import collections
import csv
import re
import sys


def main(path, per_launch=True):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    cols = rows[hdr]
    ki, vi, gi = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Grid Size")
    tot = collections.OrderedDict()
    total = 0.0
    for n, r in enumerate(rows[hdr + 1:]):
        if len(r) <= vi or not r[vi]:
            continue
        try:
            us = float(r[vi].replace(",", "")) / 1000.0
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("vp3d::", "")
        name = re.sub(r"at::native::.*?<", "at::", name)[:48]
        total += us
        tot.setdefault(name, [0, 0.0])
        tot[name][0] += 1
        tot[name][1] += us
        if per_launch:
            print(f"{n:4d} {name:48s} grid {r[gi]:>14s} {us:9.2f} us")
    print("---- totals")
    for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:48s} x{c:<4d} {us:10.2f} us  {100 * us / total:5.1f}%")
    print(f"{'TOTAL':48s}       {total:10.2f} us")


if __name__ == "__main__":
    main(sys.argv[1], per_launch="--totals" not in sys.argv)
