"""Direct-SCF J/K on B200 behind the reference's plugin surface.

Mirrors (same names, argument meaning, shapes and error behaviour):
  * scf.hf.get_jk(mol, dm, hermi, vhfopt, with_j, with_k, omega)      pyscf/scf/hf.py:963-1034
  * scf.hf.SCF.get_jk / get_j / get_k                                  pyscf/scf/hf.py:2136-2170
  * scf._vhf._VHFOpt (cached screening state, keyed by omega)          pyscf/scf/_vhf.py:151-275;
    SCF._opt dict                                                      pyscf/scf/hf.py:1803,2141-2146
Install on a PySCF mean-field object with `patch(mf)` (instance override of get_jk, the hook
documented in examples/scf/43-custom_get_jk.py:36-45).
"""
import numpy as np

from . import lib as _lib


class VHFOpt:
    """Device-resident shell-pair data + Schwarz bounds for one (mol, omega); cf. _vhf._VHFOpt."""

    def __init__(self, mol, direct_scf_tol=1e-13, omega=None, device=0, libpath=None):
        self.mol = mol
        self.direct_scf_tol = direct_scf_tol
        self.omega = 0.0 if omega is None else float(omega)
        env = np.array(mol._env, dtype=np.float64, copy=True)
        self.handle = _lib.Handle(mol._atm, mol._bas, env, device=device, libpath=libpath)
        self.nao = int(mol.ao_loc_nr(cart=False)[-1]) if hasattr(mol, 'ao_loc_nr') else mol.nao
        if getattr(mol, 'cart', False):
            raise NotImplementedError('cart=True molecules are not supported')
        h = self.handle
        h.check(h.lib.b200jk_set_screening(h._h, direct_scf_tol, self.omega), 'b200jk_set_screening')

    @property
    def q_cond(self):
        h = self.handle
        nbas = len(h.bas)
        q = np.empty((nbas, nbas))
        h.check(h.lib.b200jk_get_q_cond(h._h, _lib.dptr(q), nbas), 'b200jk_get_q_cond')
        return q

    def get_jk(self, dm, hermi=1, with_j=True, with_k=True):
        dm = np.asarray(dm)
        if np.iscomplexobj(dm):
            # pyscf/scf/hf.py:1017-1031: real and imaginary parts are contracted separately
            vjr, vkr = self.get_jk(dm.real, 0 if hermi else 0, with_j, with_k)
            vji, vki = self.get_jk(dm.imag, 0, with_j, with_k)
            vj = None if vjr is None else vjr + 1j * vji
            vk = None if vkr is None else vkr + 1j * vki
            return vj, vk
        nao = self.nao
        if dm.shape[-1] != nao or dm.shape[-2] != nao:
            raise RuntimeError('dm shape %s does not match nao=%d' % (dm.shape, nao))
        shape = dm.shape
        dms = np.ascontiguousarray(dm.reshape(-1, nao, nao), dtype=np.float64)
        n_dm = len(dms)
        vj = np.empty_like(dms) if with_j else None
        vk = np.empty_like(dms) if with_k else None
        h = self.handle
        h.check(h.lib.b200jk_direct_jk(h._h, _lib.dptr(dms), n_dm, nao, int(hermi), _lib.dptr(vj), _lib.dptr(vk)),
                'b200jk_direct_jk')
        if vj is not None:
            vj = vj.reshape(shape)
        if vk is not None:
            vk = vk.reshape(shape)
        return vj, vk

    def stats(self):
        return self.handle.stats()

    def close(self):
        self.handle.close()


_opt_cache = {}


def get_jk(mol, dm, hermi=1, vhfopt=None, with_j=True, with_k=True, omega=None):
    """Drop-in for pyscf.scf.hf.get_jk (pyscf/scf/hf.py:963): returns (vj, vk) shaped like dm."""
    if vhfopt is None:
        key = (id(mol), omega)
        vhfopt = _opt_cache.get(key)
        if vhfopt is None or vhfopt.mol is not mol:
            vhfopt = _opt_cache[key] = VHFOpt(mol, omega=omega)
    return vhfopt.get_jk(dm, hermi, with_j, with_k)


def patch(mf, device=0, libpath=None):
    """Install the B200 builder as `mf.get_jk` on a PySCF SCF object (instance override).

    Keeps the reference semantics of SCF.get_jk (pyscf/scf/hf.py:2136-2160): one cached optimizer per
    omega in mf._opt, rebuilt after mf.reset()."""
    opts = {}

    def _get_jk(mol=None, dm=None, hermi=1, with_j=True, with_k=True, omega=None):
        if mol is None:
            mol = mf.mol
        if dm is None:
            dm = mf.make_rdm1()
        key = (id(mol), omega)
        if key not in opts:
            opts[key] = VHFOpt(mol, direct_scf_tol=getattr(mf, 'direct_scf_tol', 1e-13), omega=omega, device=device,
                               libpath=libpath)
        return opts[key].get_jk(dm, hermi, with_j, with_k)

    mf.get_jk = _get_jk
    mf._b200_opts = opts
    return mf
