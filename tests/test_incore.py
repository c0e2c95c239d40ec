"""In-core path (BASELINE configs[0] H2O/STO-3G): J/K from stored two-electron integrals, the role of
_vhf.incore / CVHFnrs8_incore_drv (pyscf/scf/_vhf.py:283-366, pyscf/lib/vhf/nr_incore.c:624) behind RHF.get_jk's `mf._eri`
branch (pyscf/scf/hf.py:2499-2508).  The integrals come from the oracle here (in a PySCF script: mol.intor('int2e', aosym='s8'))."""
import numpy as np
import pytest

from pyscf_b200 import gto, jk
from oracle import oracle as O

H2O = 'O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587'


def _s4(eri):
    n = eri.shape[0]
    i, j = np.tril_indices(n)
    return np.ascontiguousarray(eri[i, j][:, i, j])


def _check(libpath, basis):
    mol = gto.M(atom=H2O, basis=basis)
    nao = mol.nao
    eri = O.int2e(mol)
    np.random.seed(3)
    dms = np.random.random((2, nao, nao))            # no symmetry: hermi = 0 semantics
    rj = np.einsum('ijkl,sji->skl', eri, dms)        # pyscf/scf/hf.py:906-907
    rk = np.einsum('ijkl,sjk->sil', eri, dms)
    for packed in (O.s8_pack(eri), _s4(eri), eri):
        vj, vk = jk.incore(mol, packed, dms, hermi=0, libpath=libpath)
        assert abs(vj - rj).max() < 1e-11 and abs(vk - rk).max() < 1e-11
    # the reference's in-core test idiom: incore == direct (pyscf/scf/test/test_vhf.py:53-59)
    dm = dms[0] + dms[0].T
    vj, vk = jk.incore(mol, O.s8_pack(eri), dm, hermi=1, with_k=False, libpath=libpath)
    assert vk is None and abs(vj - O.get_jk(mol, dm)[0]).max() < 1e-10
    with pytest.raises(RuntimeError):
        jk.incore(mol, np.zeros(17), dm, libpath=libpath)
    return mol, O.s8_pack(eri)


def test_incore_emulated(emu_lib):
    mol, eri8 = _check(emu_lib, 'sto-3g')            # BASELINE configs[0]: 7 AOs, 406 stored integrals
    assert mol.nao == 7 and eri8.size == 406
    _check(emu_lib, '6-31g')


def test_patch_uses_stored_integrals(emu_lib):
    """patch(mf): an object that carries mf._eri is served from it (hf.py:2499-2508), others by the direct path."""
    mol = gto.M(atom=H2O, basis='sto-3g')

    class MF:
        def __init__(self, mol):
            self.mol, self._eri, self.direct_scf_tol = mol, None, 1e-13

        def reset(self, mol=None):
            return self

    mf = jk.patch(MF(mol), libpath=emu_lib)
    dm = np.eye(mol.nao)
    vj0, vk0 = mf.get_jk(mol, dm)
    mf._eri = O.s8_pack(O.int2e(mol))
    vj1, vk1 = mf.get_jk(mol, dm)
    assert abs(vj0 - vj1).max() < 1e-10 and abs(vk0 - vk1).max() < 1e-10
    mf._eri = mf._eri * 2.0                          # a different array: picked up, not the cached copy
    assert abs(mf.get_jk(mol, dm)[0] - 2 * vj0).max() < 1e-9
    mf.reset()


@pytest.mark.gpu
def test_incore_gpu():
    _check(None, 'sto-3g')
    _check(None, 'cc-pvdz')
