"""Basis-set data access for the host-side Mole mirror.

Mirrors the *behaviour* of the reference's NWChem-format reader
(pyscf/gto/basis/parse_nwchem.py:105-153 `_parse` with optimize=False: shells grouped by
angular momentum in file order, SP shells split into an s and a p shell, primitives whose
contraction coefficients are all zero dropped, pyscf/gto/basis/parse_nwchem.py:298) and of the
alias table (pyscf/gto/basis/__init__.py:49-208).  The committed JSON fixtures under
pyscf_b200/data/basis were produced by tools/make_fixtures.py from the reference's .dat files.

Internal shell format (same as Mole._basis in the reference):
    [l, [exp, c_1, c_2, ...], [exp, c_1, ...], ...]
"""
import json
import os
import re

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'data', 'basis')
_L = {'S': 0, 'P': 1, 'D': 2, 'F': 3, 'G': 4, 'H': 5, 'I': 6}

# normalised alias -> fixture name (pyscf/gto/basis/__init__.py:49-208, subset)
_ALIAS = {
    'sto3g': 'sto-3g', '631g': '6-31g', 'ccpvdz': 'cc-pvdz', 'ccpvtz': 'cc-pvtz',
    'def2svp': 'def2-svp', 'def2tzvp': 'def2-tzvp',
    'def2svpjkfit': 'def2-universal-jkfit', 'def2tzvpjkfit': 'def2-universal-jkfit',
    'def2tzvppjkfit': 'def2-universal-jkfit', 'def2qzvpjkfit': 'def2-universal-jkfit',
    'def2universaljkfit': 'def2-universal-jkfit', 'weigendjkfit': 'def2-universal-jkfit',
    'weigend': 'def2-universal-jfit', 'def2universaljfit': 'def2-universal-jfit',
    'ccpvdzjkfit': 'cc-pvdz-jkfit', 'ccpvtzjkfit': 'cc-pvtz-jkfit',
}
_cache = {}


def _norm_name(name):
    return re.sub(r'[-_ ]', '', name.lower())


def extract_element_block(text, symb):
    """Return the lines of `text` (a whole NWChem .dat file) that define element `symb`."""
    out = []
    for line in text.splitlines():
        s = line.split('#')[0].strip()
        if not s:
            continue
        tok = s.split()
        if tok[0][0].isalpha():
            if tok[0].upper() in ('BASIS', 'END'):
                cur = None
                continue
            cur = tok[0]
            if cur.lower() == symb.lower() and len(tok) >= 2:
                out.append(s)
            continue
        if cur is not None and cur.lower() == symb.lower():
            out.append(s)
    return '\n'.join(out) if out else None


def parse_nwchem(text):
    """Parse NWChem-format basis text for ONE element into the internal shell list."""
    by_l = [[] for _ in range(8)]
    key = None
    cur = None
    for line in text.splitlines():
        s = line.split('#')[0].strip()
        if not s:
            continue
        up = s.upper()
        if up.startswith('END') or up.startswith('BASIS'):
            continue
        if s[0].isalpha():
            tok = s.split()
            key = (tok[0] if len(tok) == 1 else tok[1]).upper()
            if key == 'SP':
                cur = ([0], [1])
                by_l[0].append(cur[0])
                by_l[1].append(cur[1])
            elif key in _L:
                cur = [_L[key]]
                by_l[_L[key]].append(cur)
            else:
                raise ValueError('not basis data: %r' % s)
        else:
            dat = [float(x) for x in s.replace('D', 'e').replace('d', 'e').split()]
            if key is None:
                raise ValueError('not basis data: %r' % s)
            if key == 'SP':
                cur[0].append([dat[0], dat[1]])
                cur[1].append([dat[0], dat[2]])
            else:
                cur.append(dat)
    shells = [b for bs in by_l for b in bs]
    # drop primitives with all-zero coefficients, and empty shells
    out = []
    for b in shells:
        prims = [p for p in b[1:] if any(c != 0.0 for c in p[1:])]
        if prims:
            out.append([b[0]] + prims)
    if not out:
        raise ValueError('basis data not found')
    return out


def load(name, symb):
    """Internal-format shells of basis `name` for element `symb` (cf. gto.basis.load)."""
    key = _norm_name(name)
    if key not in _ALIAS:
        raise KeyError('basis %r is not among the fixtures shipped with pyscf_b200 '
                       '(pass parsed shells or NWChem text instead)' % name)
    fn = _ALIAS[key]
    if fn not in _cache:
        with open(os.path.join(_DATA, fn + '.json')) as f:
            _cache[fn] = json.load(f)
    tab = _cache[fn]
    el = symb[0].upper() + symb[1:].lower()
    if el not in tab:
        raise KeyError('basis %s has no fixture for element %s' % (name, symb))
    return [list(map(lambda x: x if isinstance(x, int) else list(x), b)) for b in tab[el]]


# JK-fit auxiliary basis chosen for an orbital basis (pyscf/df/addons.py:42-72 DEFAULT_AUXBASIS)
DEFAULT_JKFIT = {
    'ccpvdz': 'cc-pvdz-jkfit', 'ccpvtz': 'cc-pvtz-jkfit',
    'def2svp': 'def2-svp-jkfit', 'def2tzvp': 'def2-tzvp-jkfit',
    'sto3g': 'def2-svp-jkfit', '631g': 'cc-pvdz-jkfit',
}


def predefined_auxbasis(basis_name):
    """pyscf/df/addons.py:335-361 (JK-fit column); falls back to 'weigend' like DFBASIS."""
    return DEFAULT_JKFIT.get(_norm_name(basis_name), 'weigend')
