#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel.
usage: python tools/launch_summary.py gpurun_out/X_launches.csv > profiles/X_launches_summary.txt"""
import collections
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
tot = collections.OrderedDict()
n = 0
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', '')) * scale[r[ui]]
    t = tot.setdefault(r[ki], [0.0, 0])
    t[0] += v
    t[1] += 1
    n += 1
T = sum(v for v, _ in tot.values())
print('# launch list summary of %s (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)' % path)
print('# total %.1f us over %d launches' % (T, n))
for k, (v, c) in sorted(tot.items(), key=lambda x: -x[1][0]):
    print('%9.1f us  %5.1f%%  x%-4d %s' % (v, 100 * v / T, c, k))
