#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share, and
optionally the last N launches in order.  usage: launch_summary.py FILE.csv [--tail N] [--last-frac F]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, gs, bs = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    out = []
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            t = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        out.append((r[kn], t / 1e3, r[gs], r[bs]))
    return out


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("ed2::<unnamed>::", "").replace("void ", "").replace("ed2::", "")
    name = re.sub(r"at::native::|at::", "", name)
    return name[:100]


def main():
    path = sys.argv[1]
    tail = int(sys.argv[sys.argv.index("--tail") + 1]) if "--tail" in sys.argv else 0
    frac = float(sys.argv[sys.argv.index("--last-frac") + 1]) if "--last-frac" in sys.argv else 1.0
    seq = load(path)
    seq = seq[int(len(seq) * (1 - frac)):]
    agg = collections.OrderedDict()
    for k, t, g, b in seq:
        a = agg.setdefault(short(k), [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    print(f"{len(seq)} launches, {total:.1f} us total")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} us {100*t/total:5.1f}%  x{n:<4d} {k}")
    if tail:
        print("---- last launches in order")
        for k, t, g, b in seq[-tail:]:
            print(f"{t:9.1f} us  grid {g:14s} block {b:12s} {short(k)}")


if __name__ == "__main__":
    main()
