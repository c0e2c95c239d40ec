"""Host-side mirror of the part of pyscf.gto.Mole the J/K path needs: the libcint-layout
tables `_atm`, `_bas`, `_env` and the AO offsets.

Layout contract (reference file:line):
  * slot constants                     pyscf/gto/mole.py:58-88
  * atm/env rows per atom              pyscf/gto/mole.py:963-983  (make_atm_env)
  * bas/env rows per shell             pyscf/gto/mole.py:986-1018 (make_bas_env): exponents sorted
    descending, coefficients stored [nctr][nprim], pre-multiplied by gto_norm(l, e) and
    renormalised per contraction (_nomalize_contracted_ao, :1020-1029)
  * gto_norm / gaussian_int            pyscf/gto/mole.py:122-157
  * shells ordered atom by atom, per atom by angular momentum (format_basis, :420-470)
  * ao_loc                             pyscf/gto/moleintor.py:805 (make_loc)
  * Angstrom -> Bohr                   pyscf/data/nist.py:24 (BOHR = 0.52917721092)
The implementation is this repo's own; the reference cannot be imported here.
"""
import contextlib
import math
import re

import numpy as np

from . import basis as _basis_mod

BOHR = 0.52917721092

CHARGE_OF, PTR_COORD, NUC_MOD_OF, PTR_ZETA, PTR_FRAC_CHARGE, PTR_RADIUS, ATM_SLOTS = 0, 1, 2, 3, 4, 5, 6
ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF, KAPPA_OF, PTR_EXP, PTR_COEFF, BAS_SLOTS = 0, 1, 2, 3, 4, 5, 6, 8
PTR_EXPCUTOFF, PTR_RANGE_OMEGA, PTR_ENV_START = 0, 8, 20
NUC_POINT = 1

_ELEMENTS = ['X', 'H', 'He', 'Li', 'Be', 'B', 'C', 'N', 'O', 'F', 'Ne', 'Na', 'Mg', 'Al', 'Si', 'P',
             'S', 'Cl', 'Ar']
_CHARGE = {s.upper(): z for z, s in enumerate(_ELEMENTS)}


def _pure_symbol(label):
    m = re.match(r'[A-Za-z]+', label)
    s = m.group(0)
    return s[0].upper() + s[1:].lower()


def gaussian_int(n, alpha):
    n1 = (n + 1) * 0.5
    return math.gamma(n1) / (2.0 * np.asarray(alpha, dtype=float) ** n1)


def gto_norm(l, expnt):
    return 1.0 / np.sqrt(gaussian_int(l * 2 + 2, 2.0 * np.asarray(expnt, dtype=float)))


def _parse_atoms(atom, unit):
    scale = 1.0 if unit.lower().startswith(('b', 'au')) else 1.0 / BOHR
    out = []
    if isinstance(atom, str):
        for line in re.split(r'[;\n]', atom):
            tok = line.replace(',', ' ').split()
            if not tok:
                continue
            out.append((tok[0], [float(x) * scale for x in tok[1:4]]))
    else:
        for a in atom:
            lab = a[0]
            xyz = a[1] if len(a) == 2 else a[1:4]
            out.append((str(lab), [float(x) * scale for x in xyz]))
    return out


class Mole:
    """Minimal Mole: `atom`, `basis`, `unit`, `cart`, `charge`; `build()` fills _atm/_bas/_env."""

    def __init__(self, atom=None, basis='sto-3g', unit='Angstrom', cart=False, charge=0, verbose=0):
        self.atom = atom
        self.basis = basis
        self.unit = unit
        self.cart = cart
        self.charge = charge
        self.verbose = verbose
        self.omega = None
        self._built = False

    # ---- construction -------------------------------------------------------------------
    def build(self):
        self._atom = _parse_atoms(self.atom, self.unit)
        labels = []
        for lab, _ in self._atom:
            if lab not in labels:
                labels.append(lab)
        # resolve basis per label
        bdict = {}
        if isinstance(self.basis, dict):
            for lab in labels:
                if lab in self.basis:
                    spec = self.basis[lab]
                elif _pure_symbol(lab) in self.basis:
                    spec = self.basis[_pure_symbol(lab)]
                elif 'default' in self.basis:
                    spec = self.basis['default']
                else:
                    raise KeyError('no basis for atom %s' % lab)
                bdict[lab] = self._resolve(spec, lab)
        else:
            for lab in labels:
                bdict[lab] = self._resolve(self.basis, lab)
        self._basis = bdict

        env = [np.zeros(PTR_ENV_START)]
        ptr = PTR_ENV_START
        atm = []
        for lab, xyz in self._atom:
            z = _CHARGE[_pure_symbol(lab).upper()]
            atm.append([z, ptr, NUC_POINT, ptr + 3, 0, 0])
            env.append(np.array(list(xyz) + [0.0]))
            ptr += 4
        basrows = {}
        for lab, shells in bdict.items():
            rows = []
            for b in shells:
                l = b[0]
                ec = np.array(sorted(b[1:], reverse=True))
                es, cs = ec[:, 0], ec[:, 1:]
                nprim, nctr = cs.shape
                cs = cs * gto_norm(l, es)[:, None]
                ee = gaussian_int(l * 2 + 2, es[:, None] + es[None, :])
                s1 = 1.0 / np.sqrt(np.einsum('pi,pq,qi->i', cs, ee, cs))
                cs = cs * s1[None, :]
                env.append(es)
                env.append(cs.T.reshape(-1))
                rows.append([0, l, nprim, nctr, 0, ptr, ptr + nprim, 0])
                ptr += nprim + nprim * nctr
            basrows[lab] = rows
        bas = []
        for ia, (lab, _) in enumerate(self._atom):
            for r in basrows[lab]:
                rr = list(r)
                rr[ATOM_OF] = ia
                bas.append(rr)
        self._atm = np.array(atm, dtype=np.int32).reshape(-1, ATM_SLOTS)
        self._bas = np.array(bas, dtype=np.int32).reshape(-1, BAS_SLOTS)
        self._env = np.concatenate(env).astype(np.float64)
        self.natm = len(atm)
        self.nbas = len(bas)
        self.nelectron = int(self._atm[:, CHARGE_OF].sum()) - self.charge
        self._built = True
        return self

    @staticmethod
    def _resolve(spec, lab):
        if isinstance(spec, str):
            if '\n' in spec:
                return _basis_mod.parse_nwchem(spec)
            return _basis_mod.load(spec, _pure_symbol(lab))
        return [list(b) for b in spec]

    # ---- queries ------------------------------------------------------------------------
    def atom_coords(self):
        return np.array([self._env[p:p + 3] for p in self._atm[:, PTR_COORD]])

    def atom_charges(self):
        return self._atm[:, CHARGE_OF].astype(int)

    def atom_symbol(self, i):
        return self._atom[i][0]

    def ao_loc_nr(self, cart=None):
        cart = self.cart if cart is None else cart
        l = self._bas[:, ANG_OF].astype(np.int64)
        nf = (l + 1) * (l + 2) // 2 if cart else 2 * l + 1
        dims = nf * self._bas[:, NCTR_OF]
        loc = np.zeros(self.nbas + 1, dtype=np.int32)
        loc[1:] = np.cumsum(dims)
        return loc

    def nao_nr(self, cart=None):
        return int(self.ao_loc_nr(cart)[-1])

    @property
    def nao(self):
        return self.nao_nr()

    def energy_nuc(self):
        z = self.atom_charges().astype(float)
        r = self.atom_coords()
        d = np.linalg.norm(r[:, None, :] - r[None, :, :], axis=-1)
        iu = np.triu_indices(len(z), 1)
        return float((z[:, None] * z[None, :])[iu].dot(1.0 / d[iu]))

    @contextlib.contextmanager
    def with_range_coulomb(self, omega):
        """pyscf/gto/mole.py:2940-2951, 3049-3063: omega>0 erf(w r)/r, omega<0 erfc, 0 full Coulomb; None leaves the
        operator the molecule already carries (mol.omega / an enclosing context) untouched."""
        old = self._env[PTR_RANGE_OMEGA]
        if omega is not None:
            self._env[PTR_RANGE_OMEGA] = omega
        try:
            yield self
        finally:
            self._env[PTR_RANGE_OMEGA] = old

    def copy(self):
        import copy
        m = copy.copy(self)
        m._atm, m._bas, m._env = self._atm.copy(), self._bas.copy(), self._env.copy()
        return m


def M(**kw):
    return Mole(**kw).build()


def make_auxmol(mol, auxbasis=None):
    """pyscf/df/addons.py:230-270: same atoms, auxiliary basis."""
    if auxbasis is None:
        name = mol.basis if isinstance(mol.basis, str) else 'weigend'
        auxbasis = _basis_mod.predefined_auxbasis(name)
    aux = Mole(atom=[(lab, xyz) for lab, xyz in mol._atom], basis=auxbasis, unit='Bohr', cart=mol.cart)
    return aux.build()


def load_xyz(path):
    with open(path) as f:
        lines = f.read().splitlines()
    n = int(lines[0].split()[0])
    return [(t.split()[0], [float(x) for x in t.split()[1:4]]) for t in lines[2:2 + n]]


def geometry(name):
    """Atoms of a committed benchmark geometry (pyscf_b200/data/geom/<name>.xyz, Angstrom)."""
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return load_xyz(os.path.join(here, 'data', 'geom', name + '.xyz'))
