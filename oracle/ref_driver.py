"""The REFERENCE's own direct-SCF driver, screening and J/K digestion (compiled from the reference's C sources
where they lie, oracle/Makefile.ref -> oracle/_ref/libcvhf_ref.so), fed with this oracle's integral function.

TEST INFRASTRUCTURE ONLY.  Mirrors the call sequence of pyscf/scf/_vhf.py:151-206 (`_VHFOpt.init_cvhf_direct`),
:224-244 (`set_dm`), :370-429 (`direct`) and :505-604 (`nr_direct_drv`): scripts 'ji->s2kl' + 'li->s2kj' (hermi=1)
or 'li->s1kj' (hermi=0), `CVHFdot_nrs8`, `CVHFnrs8_prescreen`, then lib.hermi_triu.  libcint itself is absent, so the
`intor` function pointer is oracle_cint.c's `int2e_sph`; everything else executed here is reference code.
"""
import ctypes
import os
import time

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, '_ref', 'libcvhf_ref.so')
_lib = None


class CVHFOpt(ctypes.Structure):   # pyscf/lib/vhf/optimizer.h:23-35, pyscf/scf/_vhf.py:267-275
    _fields_ = [('nbas', ctypes.c_int), ('ngrids', ctypes.c_int), ('direct_scf_cutoff', ctypes.c_double),
                ('q_cond', ctypes.c_void_p), ('dm_cond', ctypes.c_void_p), ('fprescreen', ctypes.c_void_p),
                ('r_vkscreen', ctypes.c_void_p)]


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
    return _lib


def _fptr(cdll, name):
    return ctypes.c_void_p(ctypes.cast(getattr(cdll, name), ctypes.c_void_p).value)


def set_threads(n):
    """Make the OpenMP regions of the reference C (and of liboracle) use n threads from now on, whatever OMP_NUM_THREADS said
    when libgomp was loaded (torchrun exports OMP_NUM_THREADS=1).  Returns omp_get_max_threads() afterwards."""
    orc = O.lib()
    orc.oracle_omp_set_threads(ctypes.c_int(int(n)))
    return int(orc.oracle_omp_max_threads())


def q_cond(mol, omega=None):
    """The reference's own CVHFnr_int2e_q_cond (pyscf/lib/vhf/optimizer.c:408-454, compiled in place) on the oracle's int2e_sph."""
    ref, orc = lib(), O.lib()
    atm = np.ascontiguousarray(mol._atm, dtype=np.int32)
    bas = np.ascontiguousarray(mol._bas, dtype=np.int32)
    env = np.array(mol._env, dtype=np.float64)
    env[8] = 0.0 if omega is None else omega
    ao_loc = np.ascontiguousarray(mol.ao_loc_nr(cart=False), dtype=np.int32)
    q = np.empty((len(bas), len(bas)))
    ref.CVHFnr_int2e_q_cond(_fptr(orc, 'int2e_sph'), None, q.ctypes.data_as(ctypes.c_void_p), ao_loc.ctypes.data_as(ctypes.c_void_p),
                            atm.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(atm)), bas.ctypes.data_as(ctypes.c_void_p),
                            ctypes.c_int(len(bas)), env.ctypes.data_as(ctypes.c_void_p))
    return q


def get_jk(mol, dm, hermi=1, direct_scf_tol=1e-13, omega=None, screen=True, sample_stride=0, info=None):
    """J, K through CVHFnr_direct_drv (reference C) for real dm [..., nao, nao].

    sample_stride >= 1 routes the integrals through oracle_jk.c's int2e_sph_sampled: stride 1 only times the integral
    function (info['intor_thread_s'], info['calls']), stride m > 1 evaluates every m-th surviving shell quartet per thread
    (a bounded sample for bench.py's CPU arm; the returned J/K are then NOT the full matrices)."""
    ref, orc = lib(), O.lib()
    atm = np.ascontiguousarray(mol._atm, dtype=np.int32)
    bas = np.ascontiguousarray(mol._bas, dtype=np.int32)
    env = np.array(mol._env, dtype=np.float64)
    env[8] = 0.0 if omega is None else omega
    natm, nbas = ctypes.c_int(len(atm)), ctypes.c_int(len(bas))
    ao_loc = np.ascontiguousarray(mol.ao_loc_nr(cart=False), dtype=np.int32)
    nao = int(ao_loc[-1])
    dm = np.asarray(dm, dtype=np.float64)
    shape = dm.shape
    dms = np.ascontiguousarray(dm.reshape(-1, nao, nao))
    n_dm = len(dms)
    intor = _fptr(orc, 'int2e_sph')
    intor_drv = intor
    if sample_stride >= 1:
        orc.oracle_sample_reset(ctypes.c_int(int(sample_stride)))
        intor_drv = _fptr(orc, 'int2e_sph_sampled')

    opt = None
    if screen:
        q_cond = np.empty((len(bas), len(bas)))
        ref.CVHFnr_int2e_q_cond(intor, None, q_cond.ctypes.data_as(ctypes.c_void_p), ao_loc.ctypes.data_as(ctypes.c_void_p),
                                atm.ctypes.data_as(ctypes.c_void_p), natm, bas.ctypes.data_as(ctypes.c_void_p), nbas,
                                env.ctypes.data_as(ctypes.c_void_p))
        t_dmc0 = time.perf_counter()
        dm_cond = np.empty((len(bas), len(bas)))
        ref.CVHFnr_dm_cond(dm_cond.ctypes.data_as(ctypes.c_void_p), dms.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n_dm),
                           ao_loc.ctypes.data_as(ctypes.c_void_p), atm.ctypes.data_as(ctypes.c_void_p), natm,
                           bas.ctypes.data_as(ctypes.c_void_p), nbas, env.ctypes.data_as(ctypes.c_void_p))
        opt = CVHFOpt(len(bas), 0, direct_scf_tol, q_cond.ctypes.data, dm_cond.ctypes.data,
                      _fptr(ref, 'CVHFnrs8_prescreen').value, None)
        if info is not None:
            info['dm_cond_s'] = time.perf_counter() - t_dmc0

    t_iter0 = time.perf_counter()   # per-iteration part starts after q_cond (once per geometry, _vhf.py:151-206); dm_cond is per call
    kname = 'CVHFnrs8_li_s2kj' if hermi == 1 else 'CVHFnrs8_li_s1kj'
    njk = 2 * n_dm
    fjk = (ctypes.c_void_p * njk)()
    dmptr = (ctypes.c_void_p * njk)()
    vptr = (ctypes.c_void_p * njk)()
    out = np.zeros((njk, nao, nao))
    for i in range(n_dm):
        fjk[i] = _fptr(ref, 'CVHFnrs8_ji_s2kl').value
        fjk[n_dm + i] = _fptr(ref, kname).value
        dmptr[i] = dmptr[n_dm + i] = dms[i].ctypes.data
        vptr[i] = out[i].ctypes.data
        vptr[n_dm + i] = out[n_dm + i].ctypes.data
    shls_slice = (ctypes.c_int * 8)(*([0, len(bas)] * 4))
    ref.CVHFnr_direct_drv(intor_drv, _fptr(ref, 'CVHFdot_nrs8'), fjk, dmptr, vptr, ctypes.c_int(njk), ctypes.c_int(1), shls_slice,
                          ao_loc.ctypes.data_as(ctypes.c_void_p), None, ctypes.byref(opt) if opt is not None else None,
                          atm.ctypes.data_as(ctypes.c_void_p), natm, bas.ctypes.data_as(ctypes.c_void_p), nbas,
                          env.ctypes.data_as(ctypes.c_void_p))
    if info is not None:
        info['driver_s'] = time.perf_counter() - t_iter0     # CVHFnr_direct_drv alone (q_cond / dm_cond excluded)
    if sample_stride >= 1 and info is not None:
        st = (ctypes.c_double * 3)()
        orc.oracle_sample_stats(st)
        info.update(intor_thread_s=float(st[0]), calls=int(st[1]), evaluated=int(st[2]))
    vj, vk = out[:n_dm], out[n_dm:]
    # lib.hermi_triu (pyscf/lib/numpy_helper.py:499): fill the upper triangle from the lower one
    il = np.tril_indices(nao, -1)
    for v in vj:
        v[il[1], il[0]] = v[il]
    if hermi == 1:
        for v in vk:
            v[il[1], il[0]] = v[il]
    return vj.reshape(shape), vk.reshape(shape)
