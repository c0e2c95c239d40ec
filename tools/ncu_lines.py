#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
cur = None; data = []; tot = 0; itot = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) >= 8 and r[0] not in ('', 'Line No') and r[2] == '-':
        try:
            n = int(r[6]); ins = int(r[7])
        except ValueError:
            continue
        data.append((n, ins, cur, r[0], r[1].strip()[:100])); tot += n; itot += ins
data.sort(reverse=True)
print('total samples', tot, 'warp instructions', itot)
for n, ins, f, l, s in data[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print('%6d %5.1f%%  ins %5.1f%%  %s:%s  %s' % (n, 100 * n / tot, 100 * ins / itot, f, l, s))
