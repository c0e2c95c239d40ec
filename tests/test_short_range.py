"""Short-range operator erfc(|omega| r12)/r12 (omega < 0; pyscf/gto/mole.py:2940-2951, used by HSE-type functionals
through get_jk(..., omega=-w), pyscf/scf/hf.py:1021): the kernels visit every primitive quartet twice, once with the
Coulomb roots and once with the erf-attenuated roots and negated weights.  Emulated on the CPU and on the GPU."""
import numpy as np
import pytest

from pyscf_b200 import gto
from pyscf_b200.df import DF
from pyscf_b200.gto.mole import make_auxmol
from pyscf_b200.jk import VHFOpt
from oracle import oracle as O

H2O = 'O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587'


def _check_direct(libpath):
    # reference fingerprint pyscf/scf/test/test_vhf.py:183-198 (6-31G, omega = 0.15, erfc)
    mol = gto.M(atom=H2O, basis='6-31g')
    np.random.seed(1)
    dm = np.random.random((mol.nao, mol.nao))
    opt = VHFOpt(mol, omega=-0.15, libpath=libpath)
    vj, vk = opt.get_jk(dm, hermi=0)
    assert abs(O.fp(np.array([vj, vk])) - 25.317344717490613) < 1e-9
    # d shells against the oracle; short + long range == full, each from its own screening state
    mol = gto.M(atom=H2O, basis='cc-pvdz')
    np.random.seed(4)
    dm = np.random.random((mol.nao, mol.nao))
    dm = dm + dm.T
    vj, vk = VHFOpt(mol, omega=-0.4, libpath=libpath).get_jk(dm)
    rj, rk = O.get_jk(mol, dm, omega=-0.4)
    assert abs(vj - rj).max() < 1e-10 and abs(vk - rk).max() < 1e-10
    vj0, vk0 = VHFOpt(mol, libpath=libpath).get_jk(dm)
    vj1, vk1 = VHFOpt(mol, omega=0.4, libpath=libpath).get_jk(dm)
    assert abs(vj + vj1 - vj0).max() < 1e-10 and abs(vk + vk1 - vk0).max() < 1e-10


def _check_df(libpath):
    # DF.range_coulomb(omega < 0) (pyscf/df/df.py:298-333): 3c/2c integrals of the erfc operator
    mol = gto.M(atom=H2O, basis='6-31g')
    d = DF(mol, 'weigend', libpath=libpath)
    d.omega = -0.3
    d.build()
    ref, nao = O.cholesky_eri(mol, make_auxmol(mol, 'weigend'), omega=-0.3)
    assert abs(d._cderi - ref).max() < 1e-9
    np.random.seed(0)
    dm = np.random.random((nao, nao))
    dm = dm + dm.T
    vj, vk = d.get_jk(dm)
    rj, rk = O.df_get_jk(ref, nao, dm)
    assert abs(vj - rj).max() < 1e-9 and abs(vk - rk).max() < 1e-9


def test_short_range_direct_emulated(emu_lib):
    _check_direct(emu_lib)


def test_short_range_df_emulated(emu_lib):
    _check_df(emu_lib)


@pytest.mark.gpu
def test_short_range_direct_gpu():
    _check_direct(None)


@pytest.mark.gpu
def test_short_range_df_gpu():
    _check_df(None)
