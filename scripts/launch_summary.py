#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    n = r[ki].split('(')[0][:80]
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':82s} {'n':>4s} {'total_us':>10s} {'share':>6s} {'avg_us':>8s}")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n:82s} {a[0]:4d} {a[1]:10.1f} {a[1] / tot * 100:5.1f}% {a[1] / a[0]:8.1f}")
print(f"{'TOTAL':82s} {sum(a[0] for a in agg.values()):4d} {tot:10.1f}")
