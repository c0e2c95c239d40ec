"""CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (pyscf_b200.*) never does.

Python face of oracle/liboracle.so (McMurchie-Davidson integrals + a restatement of the reference's
direct-SCF driver) plus numpy restatements of the reference's Python-level algebra:
  * lib.fp                         pyscf/lib/misc.py:1359-1363
  * hf.dot_eri_dm / get_jk         pyscf/scf/hf.py:902-961, 963-1034
  * incore.cholesky_eri            pyscf/df/incore.py:129-220, _eig_decompose :263-270
  * df_jk.get_jk                   pyscf/df/df_jk.py:280-413
"""
import ctypes
import os
import subprocess

import numpy as np
import scipy.linalg

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build():
    """Compile liboracle.so (gcc); also oracle/_ref when the reference tree is present."""
    subprocess.check_call(['make', '-s', '-C', _HERE, 'liboracle.so'])
    if os.path.isdir('/root/reference/pyscf/lib/vhf') and os.path.exists(os.path.join(_HERE, 'Makefile.ref')):
        subprocess.call(['make', '-s', '-C', _HERE, 'ref'])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_direct_jk.restype = ctypes.c_long
    return _lib


def _p(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _tables(mol):
    atm = np.ascontiguousarray(mol._atm, dtype=np.int32)
    bas = np.ascontiguousarray(mol._bas, dtype=np.int32)
    env = np.ascontiguousarray(mol._env, dtype=np.float64)
    return atm, bas, env


def fp(a):
    """Reference fingerprint, pyscf/lib/misc.py:1359-1363."""
    a = np.asarray(a)
    return float(np.dot(np.cos(np.arange(a.size)), a.ravel()))


def int1e(mol, kind):
    """kind in {'ovlp','kin','nuc'}; spherical [nao,nao]."""
    atm, bas, env = _tables(mol)
    loc = mol.ao_loc_nr(cart=False)
    nao = int(loc[-1])
    out = np.zeros((nao, nao))
    lib().oracle_int1e(ctypes.c_int({'ovlp': 0, 'kin': 1, 'nuc': 2}[kind]), _p(out), ctypes.c_int(nao), _ip(loc),
                       _ip(atm), ctypes.c_int(mol.natm), _ip(bas), ctypes.c_int(mol.nbas), _p(env))
    return out


def int2e(mol, cart=False):
    """Full (ij|kl) tensor [nao]*4 (aosym s1)."""
    atm, bas, env = _tables(mol)
    loc = mol.ao_loc_nr(cart=cart)
    nao = int(loc[-1])
    eri = np.zeros((nao,) * 4)
    lib().oracle_fill_int2e(_p(eri), ctypes.c_int(nao), _ip(loc), ctypes.c_int(int(cart)), _ip(atm),
                            ctypes.c_int(mol.natm), _ip(bas), ctypes.c_int(mol.nbas), _p(env))
    return eri


def s8_pack(eri):
    """aosym='s8' packing used by mol.intor('int2e', aosym='s8') (pyscf/lib/vhf/fill_nr_s8.c:110)."""
    n = eri.shape[0]
    idx = np.tril_indices(n)
    e4 = eri[idx[0], idx[1]][:, idx[0], idx[1]]
    return e4[np.tril_indices(e4.shape[0])]


def conc_mol(mol, auxmol):
    """pyscf/gto/mole.py:805-838 conc_env: concatenated tables for 3-centre integrals."""
    off = len(mol._env)
    natm_off = mol.natm
    atm2 = auxmol._atm.copy()
    atm2[:, 1] += off
    atm2[:, 3] += off
    bas2 = auxmol._bas.copy()
    bas2[:, 0] += natm_off
    bas2[:, 5] += off
    bas2[:, 6] += off
    atm = np.ascontiguousarray(np.vstack([mol._atm, atm2]), dtype=np.int32)
    bas = np.ascontiguousarray(np.vstack([mol._bas, bas2]), dtype=np.int32)
    env = np.concatenate([mol._env, auxmol._env])
    return atm, bas, env


def int3c2e(mol, auxmol):
    """(ij|P), [nao,nao,naux] (aosym s1), operator from mol._env[8] (omega)."""
    atm, bas, env = conc_mol(mol, auxmol)
    env[8] = mol._env[8]
    loc, aloc = mol.ao_loc_nr(cart=False), auxmol.ao_loc_nr(cart=False)
    nao, naux = int(loc[-1]), int(aloc[-1])
    out = np.zeros((nao, nao, naux))
    lib().oracle_fill_int3c2e(_p(out), ctypes.c_int(nao), ctypes.c_int(naux), _ip(loc), _ip(aloc),
                              ctypes.c_int(mol.nbas), ctypes.c_int(auxmol.nbas), _ip(atm), ctypes.c_int(len(atm)),
                              _ip(bas), ctypes.c_int(len(bas)), _p(env))
    return out


def int3c2e_pairs(mol, auxmol, pairs):
    """(ij|P) for a list of AO shell pairs [(ish, jsh), ...]: returns (out[naux, ncol], col0[npair]); pair p owns the
    columns col0[p] + a*dj + b (a, b = functions of shells ish, jsh).  Operator from mol._env[8]."""
    atm, bas, env = conc_mol(mol, auxmol)
    env[8] = mol._env[8]
    loc, aloc = mol.ao_loc_nr(cart=False), auxmol.ao_loc_nr(cart=False)
    loc = np.ascontiguousarray(loc, dtype=np.int32)
    aloc = np.ascontiguousarray(aloc, dtype=np.int32)
    pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    dims = (loc[pairs[:, 0] + 1] - loc[pairs[:, 0]]).astype(np.int64) * (loc[pairs[:, 1] + 1] - loc[pairs[:, 1]])
    col0 = np.ascontiguousarray(np.concatenate([[0], np.cumsum(dims)[:-1]]), dtype=np.int64)
    ncol, naux = int(dims.sum()), int(aloc[-1])
    out = np.zeros((naux, ncol))
    lib().oracle_fill_int3c2e_pairs(_p(out), ctypes.c_long(ncol), ctypes.c_int(len(pairs)), _ip(pairs),
                                    col0.ctypes.data_as(ctypes.POINTER(ctypes.c_long)), ctypes.c_int(naux), _ip(loc), _ip(aloc),
                                    ctypes.c_int(mol.nbas), ctypes.c_int(auxmol.nbas), _ip(atm), ctypes.c_int(len(atm)),
                                    _ip(bas), ctypes.c_int(len(bas)), _p(env))
    return out, col0


def int2c2e(auxmol, omega=None):
    atm, bas, env = _tables(auxmol)
    if omega is not None:
        env = env.copy()
        env[8] = omega
    loc = auxmol.ao_loc_nr(cart=False)
    n = int(loc[-1])
    out = np.zeros((n, n))
    lib().oracle_fill_int2c2e(_p(out), ctypes.c_int(n), _ip(loc), ctypes.c_int(0), ctypes.c_int(auxmol.nbas),
                              _ip(atm), ctypes.c_int(auxmol.natm), _ip(bas), ctypes.c_int(auxmol.nbas), _p(env))
    return out


def q_cond(mol, omega=None):
    atm, bas, env = _tables(mol)
    if omega is not None:
        env = env.copy()
        env[8] = omega
    loc = mol.ao_loc_nr(cart=False)
    q = np.zeros((mol.nbas, mol.nbas))
    lib().oracle_q_cond(_p(q), _ip(loc), _ip(atm), ctypes.c_int(mol.natm), _ip(bas), ctypes.c_int(mol.nbas), _p(env))
    return q


def get_jk(mol, dm, omega=None, direct_scf_tol=1e-13, screen=True, return_count=False):
    """Direct-SCF J/K for arbitrary real dm[..., nao, nao] (pyscf/scf/hf.py:963 semantics)."""
    atm, bas, env = _tables(mol)
    env = env.copy()
    env[8] = 0.0 if omega is None else omega
    loc = mol.ao_loc_nr(cart=False)
    nao = int(loc[-1])
    dm = np.asarray(dm, dtype=np.float64)
    shape = dm.shape
    dms = np.ascontiguousarray(dm.reshape(-1, nao, nao))
    n_dm = len(dms)
    vj = np.zeros_like(dms)
    vk = np.zeros_like(dms)
    if screen:
        q = np.zeros((mol.nbas, mol.nbas))
        lib().oracle_q_cond(_p(q), _ip(loc), _ip(atm), ctypes.c_int(mol.natm), _ip(bas), ctypes.c_int(mol.nbas), _p(env))
        dc = np.zeros((mol.nbas, mol.nbas))
        lib().oracle_dm_cond(_p(dc), _p(dms), ctypes.c_int(n_dm), ctypes.c_int(nao), _ip(loc), ctypes.c_int(mol.nbas))
        qp, dp = _p(q), _p(dc)
    else:
        qp = dp = None
    n = lib().oracle_direct_jk(_p(vj), _p(vk), _p(dms), ctypes.c_int(n_dm), ctypes.c_int(nao), _ip(loc), _ip(atm),
                               ctypes.c_int(mol.natm), _ip(bas), ctypes.c_int(mol.nbas), _p(env), qp, dp,
                               ctypes.c_double(direct_scf_tol))
    vj, vk = vj.reshape(shape), vk.reshape(shape)
    return (vj, vk, n) if return_count else (vj, vk)


def jk_from_eri(eri, dm):
    """pyscf/scf/hf.py:906-907: J_kl = sum_ij (ij|kl) D_ji ; K_il = sum_jk (ij|kl) D_jk."""
    vj = np.einsum('ijkl,ji->kl', eri, dm)
    vk = np.einsum('ijkl,jk->il', eri, dm)
    return vj, vk


# ---------------------------------------------------------------- density fitting (numpy restatement)
def pack_tril(a):
    n = a.shape[-1]
    i, j = np.tril_indices(n)
    return a[..., i, j]


def unpack_tril(t, n):
    out = np.zeros(t.shape[:-1] + (n, n))
    i, j = np.tril_indices(n)
    out[..., i, j] = t
    out[..., j, i] = t
    return out


def cholesky_eri(mol, auxmol, lindep=1e-7, omega=None):
    """cderi[naux', nao(nao+1)/2]; pyscf/df/incore.py:129-220 (CD first, eig fallback :150-158)."""
    ctx = mol.with_range_coulomb(omega) if omega is not None else None
    if ctx is not None:
        ctx.__enter__()
    try:
        j3c = int3c2e(mol, auxmol)
        j2c = int2c2e(auxmol, omega=mol._env[8])
    finally:
        if ctx is not None:
            ctx.__exit__(None, None, None)
    nao = j3c.shape[0]
    j3c = pack_tril(j3c.transpose(2, 0, 1))  # [naux, npair]
    try:
        low = scipy.linalg.cholesky(j2c, lower=True)
        cderi = scipy.linalg.solve_triangular(low, j3c, lower=True)
    except scipy.linalg.LinAlgError:
        w, v = scipy.linalg.eigh(j2c)
        mask = w > lindep
        low = (v[:, mask] / np.sqrt(w[mask])).T
        cderi = low.dot(j3c)
    return np.ascontiguousarray(cderi), nao


def df_get_jk(cderi, nao, dm):
    """pyscf/df/df_jk.py:329-410 general-DM algebra: J = cderi^T (cderi . dmtril), K = sum_P (P|.i)(P|.i)^T."""
    dm = np.asarray(dm)
    shape = dm.shape
    dms = dm.reshape(-1, nao, nao)
    eri = unpack_tril(cderi, nao)  # [naux, nao, nao]
    vj = np.empty_like(dms)
    vk = np.empty_like(dms)
    for s, d in enumerate(dms):
        rho = np.einsum('pij,ji->p', eri, d)
        vj[s] = np.einsum('p,pij->ij', rho, eri)
        tmp = np.einsum('pij,jk->pik', eri, d)
        vk[s] = np.einsum('pik,pkl->il', tmp, eri)
    return vj.reshape(shape), vk.reshape(shape)
