#!/usr/bin/env python
"""Deterministic extended polyglycine H-(NH-CH2-CO)n-OH (all backbone dihedrals 180 deg, ideal bond lengths/angles).
(Gly)30 = C60H92N30O31, 213 atoms — BASELINE config 5 (the geometry is not in the reference).
Usage: python tools/make_glyn.py 30 > pyscf_b200/data/geom/gly30.xyz"""
import sys
import numpy as np


def place(a, b, c, bond, angle, dihedral):
    """Position of D with |CD|=bond, angle BCD, dihedral ABCD (degrees)."""
    ang, dih = np.radians(angle), np.radians(dihedral)
    bc = c - b
    bc /= np.linalg.norm(bc)
    n = np.cross(b - a, bc)
    n /= np.linalg.norm(n)
    m = np.cross(n, bc)
    d2 = np.array([-bond * np.cos(ang), bond * np.sin(ang) * np.cos(dih), bond * np.sin(ang) * np.sin(dih)])
    return c + d2[0] * bc + d2[1] * m + d2[2] * n


def build(n):
    atoms = []
    N = np.array([0.0, 0.0, 0.0])
    CA = np.array([1.458, 0.0, 0.0])
    C = place(np.array([0.0, 1.0, 0.0]), N, CA, 1.525, 111.0, 180.0)
    prevC = None
    for i in range(n):
        if i > 0:
            N = place(prevN, prevCA, prevC, 1.329, 116.2, 180.0)
            CA = place(prevCA, prevC, N, 1.458, 121.7, 180.0)
            C = place(prevC, N, CA, 1.525, 111.0, 180.0)
        atoms.append(('N', N)); atoms.append(('C', CA)); atoms.append(('C', C))
        # carbonyl O (trans to N of the next residue), amide H, two alpha H
        O = place(N, CA, C, 1.231, 120.5, 0.0)
        atoms.append(('O', O))
        if i == 0:
            atoms.append(('H', place(C, CA, N, 1.01, 109.5, 60.0)))
            atoms.append(('H', place(C, CA, N, 1.01, 109.5, -60.0)))
        else:
            atoms.append(('H', place(CA, N, prevC, 1.01, 119.0, 180.0) if False else place(prevCA, prevC, N, 1.01, 119.5, 0.0)))
        atoms.append(('H', place(N, C, CA, 1.09, 109.5, 120.0)))
        atoms.append(('H', place(N, C, CA, 1.09, 109.5, -120.0)))
        prevN, prevCA, prevC = N, CA, C
    OH = place(prevN, prevCA, prevC, 1.34, 113.0, 180.0)
    atoms.append(('O', OH))
    atoms.append(('H', place(prevCA, prevC, OH, 0.96, 108.0, 180.0)))
    return atoms


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    atoms = build(n)
    print(len(atoms))
    print('(Gly)%d extended chain, tools/make_glyn.py (Angstrom)' % n)
    for s, r in atoms:
        print('%s %.10f %.10f %.10f' % (s, r[0], r[1], r[2]))
