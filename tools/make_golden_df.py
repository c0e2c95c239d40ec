#!/usr/bin/env python
"""Golden DF J/K fixtures from the CPU oracle (run in the build container; committed under tests/golden).
The oracle's int3c2e / int2c2e / cholesky_eri / df_get_jk are pinned to the reference's fingerprints by
tests/test_oracle_golden.py.  Usage: python tools/make_golden_df.py [name ...]   names: gly4_dz, bz_tz_df"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pyscf_b200 import gto
from pyscf_b200.gto.mole import geometry, make_auxmol
from oracle import oracle as O

CASES = {
    'gly4_dz': ('gly4', 'cc-pvdz', None),              # config-5 chemistry at oracle size: cc-pvdz-jkfit, aux up to f
    'bz_tz_df': ('benzene', 'cc-pvtz', None),          # f orbital shells, cc-pvtz-jkfit aux up to g
}

for name in (sys.argv[1:] or CASES):
    geom, basis, aux = CASES[name]
    mol = gto.M(atom=geometry(geom), basis=basis)
    auxmol = make_auxmol(mol, aux)
    t = time.time()
    cderi, nao = O.cholesky_eri(mol, auxmol)
    print(name, 'nao', nao, 'naux', cderi.shape[0], 'aux lmax', int(auxmol._bas[:, 1].max()), 'oracle seconds', time.time() - t, flush=True)
    np.random.seed(1)
    dms = np.random.random((2, nao, nao))
    vj, vk = O.df_get_jk(cderi, nao, dms)
    rows = np.linspace(0, cderi.shape[0] - 1, 7).astype(int)      # a few tensor rows, the rest through fingerprints
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'df_%s.npz' % name), vj=vj, vk=vk, rows=rows, cderi_rows=cderi[rows],
                        fp_cderi=O.fp(cderi), naux=cderi.shape[0])
