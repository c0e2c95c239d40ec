#!/usr/bin/env python
"""Read tools/ab_direct.py's JSON: totals per library, per-class deltas against the first library, and the per-class best."""
import json, sys
d = json.load(open(sys.argv[1]))
base = sys.argv[2] if len(sys.argv) > 2 else list(d)[0]
libs = [k for k in d if 'class_ms' in d[k]]
print('%-28s %9s %9s %10s' % ('library', 'best ms', 'class sum', 'max dev'))
for k in libs:
    print('%-28s %9.3f %9.3f %10.1e' % (k, d[k]['best_ms'], d[k]['class_ms_sum'], d[k]['max_abs_dev_vs_first']))
for k in d:
    if 'error' in d[k]:
        print('%-28s FAILED %s' % (k, d[k]['error'][:80]))
classes = sorted(d[base]['class_ms'], key=lambda c: -d[base]['class_ms'][c])
print('\nper class: base ms, then delta (ms) of every other library; * marks the best')
print('%-7s %7s ' % ('class', 'base') + ' '.join('%9s' % k.replace('libb200jk_', '').replace('.so', '')[:9] for k in libs if k != base))
best_sum = 0.0
for c in classes:
    b = d[base]['class_ms'][c]
    vals = {k: d[k]['class_ms'].get(c, float('nan')) for k in libs}
    kbest = min(vals, key=lambda k: vals[k])
    best_sum += vals[kbest]
    print('%-7s %7.3f ' % (c, b) + ' '.join('%8.3f%s' % (vals[k] - b, '*' if k == kbest else ' ') for k in libs if k != base))
print('sum of per-class best: %.3f ms (base %.3f)' % (best_sum, d[base]['class_ms_sum']))
