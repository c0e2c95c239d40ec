#!/usr/bin/env python
"""Deterministic 3-D geometry of paclitaxel (Taxol), C47H51NO14, 113 atoms — BASELINE config 4.

The reference ships no Taxol geometry ("any chemically sane 113-atom C47H51NO14 geometry ... generated
once with a fixed seed and committed").  This script builds one from the molecular graph alone:

  1. the heavy-atom graph is parsed from the paclitaxel SMILES below (own 40-line parser: C/N/O, branches, ring bonds,
     '=' bonds, aromatic rings written in Kekule form);
  2. hydrogens are added to fill the valences (C 4, N 3, O 2);
  3. coordinates are embedded by minimising a small force field from a seeded random start (distance-geometry style:
     4-D start, the 4th coordinate is squeezed out): harmonic bonds (lengths by element pair and bond order), harmonic
     1-3 distances (109.5 / 120 degree angles by hybridisation), flat 6-rings / carbonyls / amide through 1-4
     distances and improper terms, soft repulsion between atoms more than three bonds apart.

Stereo-centres come out as the embedding finds them (the result is a stereo-isomer/conformer of paclitaxel, not the
crystal structure); what the benchmark needs — the atom list, hence 886 shells / 2228 AOs with def2-TZVP, and a compact
3-D shape with realistic interatomic distances for the Schwarz screening — does not depend on that.

Usage: python tools/make_taxol.py > pyscf_b200/data/geom/taxol.xyz
"""
import sys
import numpy as np
from scipy.optimize import minimize

SMILES = ('CC1=C2C(C(=O)C3(C(CC4C(C3C(C(C2(C)C)(CC1OC(=O)C(C(C5=CC=CC=C5)NC(=O)C6=CC=CC=C6)O)O)'
          'OC(=O)C7=CC=CC=C7)(CO4)OC(=O)C)O)C)OC(=O)C')
VALENCE = {'C': 4, 'N': 3, 'O': 2}


def parse_smiles(s):
    """Heavy atoms and bonds {(i,j): order} of a SMILES restricted to C/N/O, (), =, single-digit ring bonds."""
    atoms, bonds, stack, rings = [], {}, [], {}
    prev, order = None, 1
    for ch in s:
        if ch in 'CNO':
            atoms.append(ch)
            i = len(atoms) - 1
            if prev is not None:
                bonds[(prev, i)] = order
            prev, order = i, 1
        elif ch == '=':
            order = 2
        elif ch == '(':
            stack.append(prev)
        elif ch == ')':
            prev = stack.pop()
        elif ch.isdigit():
            if ch in rings:
                j, o = rings.pop(ch)
                bonds[(j, prev)] = max(o, order)
            else:
                rings[ch] = (prev, order)
            order = 1
        else:
            raise ValueError(ch)
    assert not rings and not stack
    return atoms, bonds


def add_hydrogens(atoms, bonds):
    atoms = list(atoms)
    bonds = dict(bonds)
    used = [0] * len(atoms)
    for (i, j), o in bonds.items():
        used[i] += o
        used[j] += o
    for i in range(len(used)):
        for _ in range(VALENCE[atoms[i]] - used[i]):
            atoms.append('H')
            bonds[(i, len(atoms) - 1)] = 1
    return atoms, bonds


def bond_length(a, b, order, arom):
    key = ''.join(sorted(a + b))
    if 'H' in key:
        return {'CH': 1.09, 'HN': 1.01, 'HO': 0.96}[key]
    if key == 'CC':
        return 1.395 if arom else (1.34 if order == 2 else 1.53)
    if key == 'CO':
        return 1.21 if order == 2 else 1.43
    if key == 'CN':
        return 1.46
    raise KeyError(key)


def build(seed=7):
    heavy, hb = parse_smiles(SMILES)
    atoms, bonds = add_hydrogens(heavy, hb)
    n = len(atoms)
    nbr = [[] for _ in range(n)]
    for (i, j) in bonds:
        nbr[i].append(j)
        nbr[j].append(i)
    dbl = [any(bonds.get((min(i, j), max(i, j)), 1) == 2 for j in nbr[i]) for i in range(n)]
    # aromatic carbons: sp2 carbons of the three C6 rings written as C5=CC=CC=C5 etc. (all ring atoms carry a double bond)
    arom = [False] * n
    for i in range(n):
        if atoms[i] == 'C' and dbl[i] and sum(1 for j in nbr[i] if atoms[j] == 'C' and dbl[j]) >= 2:
            arom[i] = True
    # conjugated single bonds: ester C(=O)-O and amide C(=O)-N are shorter; the N and the ester O are planar
    sp2 = list(dbl)
    for i in range(n):
        if atoms[i] == 'N' and any(dbl[j] for j in nbr[i]):
            sp2[i] = True
    # --- topological distances
    big = 99
    dist = np.full((n, n), big, dtype=int)
    for i in range(n):
        dist[i, i] = 0
        frontier, d = [i], 0
        while frontier and d < 4:
            d += 1
            nxt = []
            for u in frontier:
                for v in nbr[u]:
                    if dist[i, v] > d:
                        dist[i, v] = d
                        nxt.append(v)
            frontier = nxt
    r0 = {}
    for (i, j), o in bonds.items():
        L = bond_length(atoms[i], atoms[j], o, arom[i] and arom[j])
        if o == 1 and {atoms[i], atoms[j]} == {'C', 'O'} and (dbl[i] or dbl[j]):
            L = 1.35            # ester C(=O)-O
        if o == 1 and {atoms[i], atoms[j]} == {'C', 'N'} and (dbl[i] or dbl[j]):
            L = 1.34            # amide C(=O)-N
        if o == 1 and atoms[i] == atoms[j] == 'C' and (dbl[i] != dbl[j] or (dbl[i] and dbl[j] and not (arom[i] and arom[j]))):
            L = 1.50            # sp2-sp3 / sp2-sp2 single
        r0[(i, j)] = L
    def blen(i, j):
        return r0[(min(i, j), max(i, j))]
    pairs13 = []
    for c in range(n):
        ang = np.radians(120.0 if sp2[c] else (109.5 if atoms[c] != 'O' else 112.0))
        for a in range(len(nbr[c])):
            for b in range(a + 1, len(nbr[c])):
                i, j = nbr[c][a], nbr[c][b]
                d = np.sqrt(blen(i, c) ** 2 + blen(j, c) ** 2 - 2 * blen(i, c) * blen(j, c) * np.cos(ang))
                if dist[i, j] == 2:          # not in a 3-ring (none here) — 4-ring (oxetane) angles are ~90 deg
                    pairs13.append((i, j, d))
    # the oxetane (4-ring: C-C-C-O): override its 1-3 distances with the ring diagonal
    four = set()
    for (i, j) in bonds:
        for k in nbr[j]:
            if k == i:
                continue
            for l in nbr[k]:
                if l != j and l != i and (min(l, i), max(l, i)) in bonds:
                    four.add(tuple(sorted((i, j, k, l))))
    ring4 = set(a for r in four for a in r)
    p13 = []
    for (i, j, d) in pairs13:
        if i in ring4 and j in ring4 and any(i in r and j in r for r in four):
            d = 2.10
        p13.append((i, j, d))
    # planar groups: for every sp2 centre with three neighbours keep the centre in their plane (improper), and keep
    # aromatic rings flat through para (1-4) distances
    impropers = [(c, nbr[c][0], nbr[c][1], nbr[c][2]) for c in range(n) if sp2[c] and len(nbr[c]) == 3]
    p14 = []
    for i in range(n):
        for j in range(i + 1, n):
            if arom[i] and arom[j] and dist[i, j] == 3:
                # para carbons of one ring: two 3-bond paths
                paths = sum(1 for u in nbr[i] for v in nbr[j] if (min(u, v), max(u, v)) in bonds and arom[u] and arom[v])
                if paths == 2:
                    p14.append((i, j, 2.79))
    rep = [(i, j) for i in range(n) for j in range(i + 1, n) if dist[i, j] > 3]
    rep3 = [(i, j) for i in range(n) for j in range(i + 1, n) if dist[i, j] == 3]
    vdw = {'C': 1.7, 'N': 1.6, 'O': 1.5, 'H': 1.1}
    bi = np.array([(i, j) for (i, j) in r0]); bl = np.array([r0[k] for k in r0])
    ai = np.array([(i, j) for (i, j, d) in p13]); al = np.array([d for (i, j, d) in p13])
    fi = np.array([(i, j) for (i, j, d) in p14]); fl = np.array([d for (i, j, d) in p14])
    ri = np.array(rep); rl = np.array([0.95 * (vdw[atoms[i]] + vdw[atoms[j]]) for (i, j) in rep])
    r3 = np.array(rep3); r3l = np.array([0.68 * (vdw[atoms[i]] + vdw[atoms[j]]) for (i, j) in rep3])
    imp = np.array(impropers)

    def energy(x, dim, w4, wrep):
        X = x.reshape(n, dim)
        G = np.zeros_like(X)
        E = 0.0

        def harmonic(idx, L, k):
            nonlocal E
            d = X[idx[:, 0]] - X[idx[:, 1]]
            r = np.sqrt((d * d).sum(1)) + 1e-12
            E += k * ((r - L) ** 2).sum()
            g = (2 * k * (r - L) / r)[:, None] * d
            np.add.at(G, idx[:, 0], g)
            np.add.at(G, idx[:, 1], -g)

        def repulse(idx, L, k):
            nonlocal E
            d = X[idx[:, 0]] - X[idx[:, 1]]
            r = np.sqrt((d * d).sum(1)) + 1e-12
            m = r < L
            E += k * ((L[m] - r[m]) ** 2).sum()
            g = np.zeros_like(d)
            g[m] = (-2 * k * (L[m] - r[m]) / r[m])[:, None] * d[m]
            np.add.at(G, idx[:, 0], g)
            np.add.at(G, idx[:, 1], -g)

        harmonic(bi, bl, 100.0)
        harmonic(ai, al, 40.0)
        if len(fi):
            harmonic(fi, fl, 40.0)
        repulse(ri, rl, wrep)
        repulse(r3, r3l, wrep)
        if dim == 3 and len(imp):
            # improper: signed volume of (n1-c, n2-c, n3-c) -> 0
            a = X[imp[:, 1]] - X[imp[:, 0]]
            b = X[imp[:, 2]] - X[imp[:, 0]]
            c = X[imp[:, 3]] - X[imp[:, 0]]
            bxc, cxa, axb = np.cross(b, c), np.cross(c, a), np.cross(a, b)
            v = (a * bxc).sum(1)
            k = 10.0
            E += k * (v * v).sum()
            ga, gb, gc = (2 * k * v)[:, None] * bxc, (2 * k * v)[:, None] * cxa, (2 * k * v)[:, None] * axb
            np.add.at(G, imp[:, 1], ga)
            np.add.at(G, imp[:, 2], gb)
            np.add.at(G, imp[:, 3], gc)
            np.add.at(G, imp[:, 0], -(ga + gb + gc))
        if dim == 4:
            E += w4 * (X[:, 3] ** 2).sum()
            G[:, 3] += 2 * w4 * X[:, 3]
        return E, G.ravel()

    rng = np.random.RandomState(seed)
    X = rng.standard_normal((n, 4)) * 6.0
    for w4, wrep in [(0.0, 1.0), (0.05, 5.0), (1.0, 20.0), (20.0, 40.0)]:
        X = minimize(energy, X.ravel(), args=(4, w4, wrep), jac=True, method='L-BFGS-B',
                     options={'maxiter': 4000, 'maxfun': 8000}).x.reshape(n, 4)
    X = X[:, :3].copy()
    for wrep in (40.0, 80.0):
        res = minimize(energy, X.ravel(), args=(3, 0.0, wrep), jac=True, method='L-BFGS-B',
                       options={'maxiter': 20000, 'maxfun': 40000, 'ftol': 1e-14, 'gtol': 1e-8})
        X = res.x.reshape(n, 3)
    X -= X.mean(0)
    # principal axes, fixed handedness -> deterministic orientation
    w, v = np.linalg.eigh(X.T.dot(X))
    X = X.dot(v[:, ::-1])
    return atoms, X, bonds, res.fun


def report(atoms, X, bonds):
    """Closest contact between atoms that are not bonded to each other (1-3 pairs included)."""
    n = len(atoms)
    d = np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1)) + 10 * np.eye(n)
    bonded = np.zeros((n, n), bool)
    for (i, j) in bonds:
        bonded[i, j] = bonded[j, i] = True
    return d[~bonded].min()


if __name__ == '__main__':
    atoms, X, bonds, e = build()
    from collections import Counter
    c = Counter(atoms)
    assert (c['C'], c['H'], c['N'], c['O']) == (47, 51, 1, 14), c
    closest = report(atoms, X, bonds)
    sys.stderr.write('formula %s  residual %.4f  closest non-bonded contact %.3f A  extent %s\n'
                     % (dict(c), e, closest, np.ptp(X, axis=0).round(2)))
    print(len(atoms))
    print('paclitaxel C47H51NO14, embedded from its SMILES graph by tools/make_taxol.py (seed 7), Angstrom')
    for s, r in zip(atoms, X):
        print('%s %.8f %.8f %.8f' % (s, r[0], r[1], r[2]))
