#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total, average, max, share.
usage: python profiles/summarize_launches.py launches.csv [title] > summary.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "us")
        if unit in ("ns", "nsecond"):
            v /= 1000.0
        elif unit in ("ms", "msecond"):
            v *= 1000.0
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void ", "").strip()
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print("# %s\n" % title)
    print("%d kernel launches, %.1f ms of kernel time (serialised, cold cache, `--clock-control none`: compare shares, not absolutes).\n" % (n, tot / 1000.0))
    print("| kernel | launches | total ms | avg us | max us | share |")
    print("|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.2f | %.1f | %.1f | %.1f %% |" % (k, a[0], a[1] / 1000.0, a[1] / a[0], a[2], 100.0 * a[1] / tot))


if __name__ == "__main__":
    main()
