#!/usr/bin/env python
"""A/B timing of direct-J/K library variants (tools/build_variant.sh) on one workload: for every library given, the
device time of a full build (library CUDA events, concurrent class streams, best and mean of N) and the max deviation of
J,K from the first library's result.  usage: python tools/ab_direct.py [--geom benzene --basis cc-pvtz] lib1.so lib2.so ..."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pyscf_b200 import gto
from pyscf_b200.gto.mole import geometry
from pyscf_b200.jk import VHFOpt
from pyscf_b200 import lib as _lib

ap = argparse.ArgumentParser()
ap.add_argument('--geom', default='benzene')
ap.add_argument('--basis', default='cc-pvtz')
ap.add_argument('--nocc', type=int, default=21)
ap.add_argument('--n', type=int, default=10)
ap.add_argument('--ref', default=None, help='npz of J,K to compare with (written by the first library)')
ap.add_argument('libs', nargs='+')
a = ap.parse_args()
if len(a.libs) > 1:
    # one process per library: two builds of the library in one process share the function-local `configured` flags of the
    # launch helpers (GNU unique symbols), so the second one would skip its cudaFuncSetAttribute calls
    import subprocess
    merged = {}
    refp = os.path.join(ROOT, 'gpurun_out', 'ab_ref_%d.npz' % os.getpid())
    for path in a.libs:
        subprocess.call([sys.executable, os.path.abspath(__file__), '--geom', a.geom, '--basis', a.basis, '--nocc', str(a.nocc),
                         '--n', str(a.n), '--ref', refp, path])
        try:
            merged.update(json.load(open(os.path.join(ROOT, 'gpurun_out', 'ab_direct_%s_%s.json' % (a.geom, a.basis)))))
        except Exception:
            pass
    json.dump(merged, open(os.path.join(ROOT, 'gpurun_out', 'ab_direct_%s_%s.json' % (a.geom, a.basis)), 'w'), indent=1)
    if os.path.exists(refp):
        os.remove(refp)
    sys.exit(0)
mol = gto.M(atom=geometry(a.geom), basis=a.basis)
c, _ = np.linalg.qr(np.random.RandomState(1).standard_normal((mol.nao, a.nocc)))
dm = 2 * c.dot(c.T)
ref = None
if a.ref and os.path.exists(a.ref):
    z = np.load(a.ref)
    ref = (z['vj'], z['vk'])
out = {}
for path in a.libs:
    name = os.path.basename(path)
    try:
        opt = VHFOpt(mol, libpath=os.path.abspath(path))
        for _ in range(3):
            vj, vk = opt.get_jk(dm)
        ms = []
        for _ in range(a.n):
            vj, vk = opt.get_jk(dm)
            ms.append(opt.stats()['ms_kernels'])
        if ref is None:
            ref = (vj, vk)
            if a.ref:
                os.makedirs(os.path.dirname(a.ref), exist_ok=True)
                np.savez(a.ref, vj=vj, vk=vk)
        err = max(abs(vj - ref[0]).max(), abs(vk - ref[1]).max())
        out[name] = {'best_ms': min(ms), 'mean_ms': float(np.mean(ms)), 'max_abs_dev_vs_first': float(err)}
        # per-class times, classes serialised (CUDA events around each class launch)
        h = opt.handle
        h.lib.b200jk_set_profile(h._h, 1)
        acc = np.zeros(100)
        for _ in range(3):
            opt.get_jk(dm)
            cm = np.zeros(100)
            h.lib.b200jk_get_class_times(h._h, _lib.dptr(cm), 100)
            acc = cm if acc.sum() == 0 else np.minimum(acc, cm)
        names = ['ss', 'ps', 'pp', 'ds', 'dp', 'dd', 'fs', 'fp', 'fd', 'ff']
        out[name]['class_ms'] = {names[cb] + '|' + names[ck]: float(acc[cb * 10 + ck]) for cb in range(10) for ck in range(cb + 1)
                                 if acc[cb * 10 + ck] > 0}
        out[name]['class_ms_sum'] = float(acc.sum())
        print('%-28s best %8.3f ms  mean %8.3f ms  |dJK| vs first %.1e' % (name, min(ms), np.mean(ms), err), flush=True)
        opt.close()
    except Exception as e:   # a variant that fails to launch must not hide the others
        print('%-28s FAILED: %s' % (name, e), flush=True)
        out[name] = {'error': str(e)}
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'ab_direct_%s_%s.json' % (a.geom, a.basis)), 'w'), indent=1)
