"""Checks of a density-fitting engine against the oracle fixtures made AT CONFIGURATION SIZE by tools/make_golden_df_size.py
(tests/golden/df_size_<name>.npz): sampled AO-pair columns of the tensor (all auxiliary rows) and J/K of a density supported
on a few shells, whose oracle value only needs the slab (P|s nu).  Shared by tests/test_df_size.py and bench.py's parity leg.
Nothing here touches oracle/: the fixtures are plain arrays."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def fp(a):
    """pyscf/lib/misc.py:1359-1363"""
    a = np.asarray(a)
    return float(np.dot(np.cos(np.arange(a.size)), a.ravel()))


def load(name):
    path = os.path.join(GOLDEN, 'df_size_%s.npz' % name)
    return np.load(path) if os.path.exists(path) else None


def slab_coeff(z):
    """C_S of the fixture: random normal on the rows z['sao'], zero elsewhere (tools/make_golden_df_size.py:slab_coeff)."""
    nao, nocc, sao = int(z['nao']), int(z['nocc']), z['sao']
    rng = np.random.RandomState(int(z['seed']))
    c = np.zeros((nao, nocc))
    c[sao] = rng.standard_normal((len(sao), nocc)) / np.sqrt(nocc)
    return c


def check_columns(dfobj, z):
    """max |cderi[:, cols] - oracle| over the rows this rank holds; None when the fixture's tensor is not unique (metric not
    positive definite: eigen-decomposition fallback, rows defined up to rotations in near-degenerate eigenspaces)."""
    import ctypes
    if 'chol' in z and int(z['chol']) == 0:
        return None
    h = dfobj._handle
    row0, nrow = ctypes.c_int(0), ctypes.c_int(0)
    dfobj.get_naoaux()
    h = dfobj._handle
    h.check(h.lib.b200jk_df_local_rows(h._h, ctypes.byref(row0), ctypes.byref(nrow)), 'b200jk_df_local_rows')
    got = dfobj.cderi_columns(z['cols'])
    ref = z['cderi_cols'][row0.value:row0.value + nrow.value]
    return float(abs(got - ref).max()) if nrow.value else 0.0


def compare_jk(z, vj, vk):
    """Deviations of a J/K pair computed for the slab density 2 C_S C_S^T from the oracle values."""
    out = {}
    if vj is not None:
        out['max_abs_dJ'] = float(abs(vj[z['sao']] - z['vj_rows']).max())
    if vk is not None:
        i, l = z['vk_idx'][:, 0], z['vk_idx'][:, 1]
        out['max_abs_dK'] = float(abs(vk[i, l] - z['vk_val']).max())
        out['max_abs_dK_diag'] = float(abs(np.diag(vk) - z['vk_diag']).max())
        out['d_fp_K'] = float(abs(fp(vk) - float(z['vk_fp'])))
    return out
