#!/usr/bin/env python
"""Developer tool: per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/ncu_launches.py launches.csv ["header comment" ...]"""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iV].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1e-3)
        a = agg.setdefault(r[iK], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for c in sys.argv[2:]:
        print("# " + c)
    print("%-72s %6s %12s %7s" % ("kernel", "count", "total_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %12.1f %6.1f%%" % (k[:72], n, t, 100.0 * t / tot))


if __name__ == "__main__":
    main()
