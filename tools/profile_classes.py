#!/usr/bin/env python
"""Per-class kernel times of one direct J/K build (CUDA events around each class launch)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pyscf_b200 import gto
from pyscf_b200.gto.mole import geometry
from pyscf_b200.jk import VHFOpt
from pyscf_b200 import lib as _lib

geom, basis = (sys.argv[1:3] + ['benzene', 'cc-pvtz'])[:2] if len(sys.argv) > 2 else ('benzene', 'cc-pvtz')
mol = gto.M(atom=geometry(geom), basis=basis)
nao = mol.nao
rng = np.random.RandomState(1)
c, _ = np.linalg.qr(rng.standard_normal((nao, 21)))
dm = 2 * c.dot(c.T)
opt = VHFOpt(mol)
h = opt.handle
h.lib.b200jk_set_profile(h._h, 1)
for _ in range(3):
    opt.get_jk(dm)
ms = np.zeros(100)
h.lib.b200jk_get_class_times(h._h, _lib.dptr(ms), 100)
names = ['ss', 'ps', 'pp', 'ds', 'dp', 'dd', 'fs', 'fp', 'fd', 'ff']
rows = []
for cb in range(10):
    for ck in range(cb + 1):
        if ms[cb * 10 + ck] > 0:
            rows.append((ms[cb * 10 + ck], names[cb] + '|' + names[ck]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print('total class ms %.3f   stats %s' % (tot, opt.stats()))
for t, n in rows:
    print('%-8s %8.3f ms  %5.1f%%' % (n, t, 100 * t / tot))
json.dump({n: t for t, n in rows}, open(os.path.join(ROOT, 'gpurun_out', 'class_times_%s_%s.json' % (geom, basis)), 'w'), indent=1)
