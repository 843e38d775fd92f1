#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][:80]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised: compare shares)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e6:10.4f} ms {100 * t / tot:6.2f}%  x{n:<4} {k}")


if __name__ == "__main__":
    main(sys.argv[1])
