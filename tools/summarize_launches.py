#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name -> markdown table."""
import collections
import csv
import re
import sys


def main(path, title, steps=1):
    lines = open(path).read().splitlines()
    start = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        name = re.sub(r'<.*', '', name)[:80]
        agg[name][0] += 1
        agg[name][1] += float(r['Metric Value']) / 1e6
    total = sum(v[1] for v in agg.values())
    print("# %s\n" % title)
    print("%d launches over %d identical eager steps = %d launches and %.2f ms of summed device time per step (cold-cache, "
          "serialised under ncu: compare SHARES; the graph-replayed step is faster).  Launch counts and ms below are totals "
          "over the %d steps.\n" % (len(rows), steps, len(rows) // steps, total / steps, steps))
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("| `%s` | %d | %.3f | %.1f%% |" % (name, c, t, 100 * t / total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
