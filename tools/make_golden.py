#!/usr/bin/env python
"""Golden J/K fixtures from the CPU oracle (run in the build container; committed under tests/golden).

The oracle itself is pinned to the reference's known-answer fingerprints by
tests/test_oracle_golden.py, so these vectors are "reference-pinned oracle outputs".
Usage: python tools/make_golden.py [name ...]   names: bz_tz, bz_dz, h2o_tz
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pyscf_b200 import gto
from pyscf_b200.gto.mole import geometry
from oracle import oracle as O

CASES = {
    'bz_tz': ('benzene', 'cc-pvtz'),
    'bz_dz': ('benzene', 'cc-pvdz'),
    'h2o_tz': ('h2o', 'cc-pvtz'),
}

def parity_dm(nao):
    np.random.seed(1)  # the reference's own test idiom (pyscf/scf/test/test_rhf.py:897-899)
    dm = np.random.random((nao, nao))
    return dm + dm.T

for name in (sys.argv[1:] or CASES):
    geom, basis = CASES[name]
    mol = gto.M(atom=geometry(geom), basis=basis)
    dm = parity_dm(mol.nao)
    t = time.time()
    vj, vk, n = O.get_jk(mol, dm, return_count=True)
    print(name, mol.nao, 'quartets', n, 'oracle seconds', time.time() - t, flush=True)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'jk_%s.npz' % name), vj=vj, vk=vk,
                        fp_j=O.fp(vj), fp_k=O.fp(vk), nquartets=n)
