#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and mean ms."""
import collections
import csv
import sys


def summarise(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        agg.setdefault(r[ki], []).append(v)
    return agg


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(p)
        agg = summarise(p)
        tot = sum(sum(v) for v in agg.values())
        for k, v in agg.items():
            print("  %-60s n=%4d total=%9.3f ms  mean=%8.3f ms  share=%5.1f%%" % (k[:60], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
