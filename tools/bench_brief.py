#!/usr/bin/env python
"""Short view of a bench.py JSON line (headline + df records): tools/bench_brief.py <file.json>"""
import json
import sys


def brief(tag, r):
    if not isinstance(r, dict) or 'value' not in r:
        print(tag, r)
        return
    roof = r.get('roofline', {})
    print('%s: %.3f ms/step (e2e %.3f ms), N=%d, setup %.1f s, roofline %s frac %.3f, parity %s' % (
        tag, r['ms_per_step'], r['e2e']['value'] * 1e3, r['n_gpus'], r.get('setup_s', 0), roof.get('bound'), roof.get('frac') or 0,
        json.dumps(r.get('parity'))))
    for k, v in (roof.get('stages') or {}).items():
        print('    %-8s %8.2f ms/step  frac %s' % (k, v['ms_per_step'], ('%.3f' % v['frac']) if 'frac' in v else '-'))
    if 'per_rank_kernel_ms' in r:
        print('    per-rank kernel ms', r['per_rank_kernel_ms'])
    if 'cpu_baseline' in r:
        c = dict(r['cpu_baseline'])
        c.pop('sample', None), c.pop('split', None)
        print('    cpu_baseline', c)


for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith('{'):
        continue
    r = json.loads(line)
    brief(r.get('config', {}).get('workload', r.get('impl', '?')), r)
    for k, v in (r.get('df') or {}).items():
        brief('  df/' + k, v)
