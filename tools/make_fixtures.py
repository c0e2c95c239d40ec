#!/usr/bin/env python
"""Extract the basis-set *data* and benchmark geometries this repo needs from the
read-only reference tree and freeze them as small JSON/XYZ fixtures.

Run in the build container only (needs /root/reference); the outputs are committed
so that nothing on the GPU box reads /root/reference.

Sources (reference file:line):
  * basis text: pyscf/gto/basis/*.dat (NWChem format; alias table pyscf/gto/basis/__init__.py:49-208)
  * benzene geometry: examples/2-benchmark/bz.py:9-22
  * C60 geometry: pyscf/tools/c60struct.py:39 make60(1.46, 1.38) (examples/2-benchmark/c60.py:10)
  * H2O geometry: pyscf/scf/test/test_rhf.py:36-41
The NWChem text is parsed with THIS repo's parser (pyscf_b200.gto.basis.parse_nwchem),
not the reference's.
"""
import json, os, re, sys, importlib.util
import numpy as np

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyscf_b200.gto.basis import parse_nwchem, extract_element_block

FILES = {
    'sto-3g': 'sto-3g.dat',
    '6-31g': 'pople-basis/6-31G.dat',
    'cc-pvdz': 'cc-pvdz.dat',
    'cc-pvtz': 'cc-pvtz.dat',
    'def2-svp': 'def2-svp.dat',
    'def2-tzvp': 'def2-tzvp.dat',
    'def2-universal-jkfit': 'def2-universal-jkfit.dat',
    'def2-universal-jfit': 'def2-universal-jfit.dat',
    'cc-pvdz-jkfit': 'cc-pvdz-jkfit.dat',
    'cc-pvtz-jkfit': 'cc-pvtz-jkfit.dat',
}
ELEMENTS = ['H', 'He', 'C', 'N', 'O', 'Ne']

def main():
    outdir = os.path.join(ROOT, 'pyscf_b200', 'data', 'basis')
    for name, fn in FILES.items():
        text = open(os.path.join(REF, 'pyscf/gto/basis', fn)).read()
        out = {}
        for el in ELEMENTS:
            blk = extract_element_block(text, el)
            if blk is None:
                continue
            out[el] = parse_nwchem(blk)
        with open(os.path.join(outdir, name + '.json'), 'w') as f:
            json.dump(out, f, separators=(',', ':'))
        print(name, {k: len(v) for k, v in out.items()})

    gdir = os.path.join(ROOT, 'pyscf_b200', 'data', 'geom')
    # C60 from the reference's generator (pure numpy module, importable stand-alone)
    spec = importlib.util.spec_from_file_location('c60struct', os.path.join(REF, 'pyscf/tools/c60struct.py'))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    xyz = m.make60(1.46, 1.38)
    with open(os.path.join(gdir, 'c60.xyz'), 'w') as f:
        f.write('60\nC60 make60(1.46,1.38) Angstrom\n')
        for r in xyz:
            f.write('C %.15f %.15f %.15f\n' % tuple(r))
    print('c60', xyz.shape)

if __name__ == '__main__':
    main()
