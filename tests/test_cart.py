"""Cartesian AOs (mol.cart = True): J/K over libcint's Cartesian functions (int2e_cart), the reference's `cart=True` molecules
(pyscf/scf/test/test_jk.py:30-38, pyscf/scf/test/test_rhf.py:418-422).  Checked against the oracle's int2e_cart tensor with the
definitions of pyscf/scf/hf.py:906-907, incl. d and f shells, a non-symmetric density and the erf-attenuated operator."""
import numpy as np
import pytest

from pyscf_b200 import gto
from pyscf_b200.jk import VHFOpt
from oracle import oracle as O


def _check(libpath):
    for atom, basis, omega in [('O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587', 'cc-pvdz', None),
                               ('He 0 0 0; Ne 1.2 0.3 0', 'cc-pvtz', None),
                               ('O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587', '6-31g', 0.2)]:
        mol = gto.M(atom=atom, basis=basis, cart=True)
        if omega is not None:
            mol._env[8] = omega
        nao = mol.nao_nr(cart=True)
        eri = O.int2e(mol, cart=True)
        assert eri.shape[0] == nao
        np.random.seed(5)
        dm = np.random.random((nao, nao))
        opt = VHFOpt(mol, libpath=libpath)
        assert opt.nao == nao
        for d, hermi in ((dm + dm.T, 1), (dm, 0)):
            vj, vk = opt.get_jk(d, hermi=hermi)
            assert abs(vj - np.einsum('ijkl,ji->kl', eri, d)).max() < 1e-10
            assert abs(vk - np.einsum('ijkl,jk->il', eri, d)).max() < 1e-10
        opt.close()


def test_cart_emulated(emu_lib):
    _check(emu_lib)


@pytest.mark.gpu
def test_cart_gpu():
    _check(None)
