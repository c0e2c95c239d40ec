"""Pins the CPU oracle to the reference's own known-answer fingerprints.

Every expected number below is copied from a test in /root/reference (file:line cited); none was
produced by this repository.  CPU only."""
import numpy as np
import pytest

from pyscf_b200 import gto
from pyscf_b200.gto.mole import make_auxmol
from oracle import oracle as O

H2O = 'O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587'


@pytest.fixture(scope='module')
def h2o():
    return gto.M(atom=H2O, basis='cc-pvdz')


@pytest.fixture(scope='module')
def h2o_eri(h2o):
    return O.int2e(h2o)


def test_int2e_s8_he_ne():
    # pyscf/gto/test/test_moleintor.py:317-320
    mol = gto.M(atom='He 0 0 0; Ne 3 0 0', basis='ccpvdz')
    assert abs(O.fp(O.s8_pack(O.int2e(mol))) - (-10.685918926843847)) < 1e-9


def test_int2c2e_benzene_like():
    # pyscf/gto/test/test_moleintor.py:20-67,331-333 (the ECP on C1 does not enter int2c2e)
    atoms = [["C", (-0.65830719, 0.61123287, -0.00800148)], ["C1", (0.73685281, 0.61123287, -0.00800148)],
             ["C2", (1.43439081, 1.81898387, -0.00800148)], ["C3", (0.73673681, 3.02749287, -0.00920048)],
             ["C4", (-0.65808819, 3.02741487, -0.00967948)], ["C5", (-1.35568919, 1.81920887, -0.00868348)],
             ["H", (-1.20806619, -0.34108413, -0.00755148)], ["H", (1.28636081, -0.34128013, -0.00668648)],
             ["H", (2.53407081, 1.81906387, -0.00736748)], ["H", (1.28693681, 3.97963587, -0.00925948)],
             ["H", (-1.20821019, 3.97969587, -0.01063248)], ["H", (-2.45529319, 1.81939187, -0.00886348)]]
    mol = gto.M(atom=atoms, basis='cc-pvdz')
    assert abs(O.fp(O.int2c2e(mol)) - (-460.83033192375615)) < 1e-9


def test_int3c2e_h2o_weigend(h2o):
    # pyscf/df/test/test_incore.py:46-72
    aux = make_auxmol(h2o, 'weigend')
    j3c = O.int3c2e(h2o, aux)
    assert abs(O.fp(j3c) - 45.27912877994409) < 1e-9
    idx = np.tril_indices(h2o.nao)
    assert abs(O.fp(j3c[idx]) - 12.407403711205063) < 1e-9


def test_get_vj_norm_and_identity(h2o, h2o_eri):
    # pyscf/scf/test/test_rhf.py:896-922
    nao = h2o.nao
    np.random.seed(1)
    dm = np.random.random((nao, nao))
    vj, _ = O.jk_from_eri(h2o_eri, dm)
    assert abs(np.linalg.norm(vj) - 77.035779188661465) < 1e-9
    vj, vk = O.jk_from_eri(h2o_eri, np.eye(nao))
    assert abs(O.fp(vj) - 1.6593323222866125) < 1e-9
    assert abs(O.fp(vk) - (-1.4662135224053987)) < 1e-9


def test_get_vk_hermi0(h2o, h2o_eri):
    # pyscf/scf/test/test_rhf.py:924-934
    np.random.seed(1)
    dm = np.random.random((h2o.nao, h2o.nao))
    _, vk = O.jk_from_eri(h2o_eri, dm)
    assert abs(O.fp(vk) - (-12.365527167710301)) < 1e-10
    # the direct driver restatement must agree with the dense contraction
    vj2, vk2 = O.get_jk(h2o, dm)
    vj1, vk1 = O.jk_from_eri(h2o_eri, dm)
    assert abs(vj2 - vj1).max() < 1e-11 and abs(vk2 - vk1).max() < 1e-11


def test_long_range_jk(h2o):
    # pyscf/scf/test/test_rhf.py:936-958 (omega = 1.5)
    np.random.seed(1)
    dm = np.random.random((h2o.nao, h2o.nao))
    with h2o.with_range_coulomb(1.5):
        eri = O.int2e(h2o)
    vj, vk = O.jk_from_eri(eri, dm)
    assert abs(O.fp(vj) - (-10.015956161068031)) < 1e-10
    assert abs(O.fp(vk) - (-11.399103957754445)) < 1e-10
    vj2, vk2 = O.get_jk(h2o, dm, omega=1.5)
    assert abs(vj2 - vj).max() < 1e-11 and abs(vk2 - vk).max() < 1e-11


def test_nr_get_jk_two_dms(h2o, h2o_eri):
    # pyscf/df/test/test_df_jk.py:157-165 (the non-DF branch of test_nr_get_jk)
    np.random.seed(1)
    dms = np.random.random((2, h2o.nao, h2o.nao))
    vj = np.array([O.jk_from_eri(h2o_eri, d)[0] for d in dms])
    vk = np.array([O.jk_from_eri(h2o_eri, d)[1] for d in dms])
    assert abs(O.fp(vj) - (-194.08878302990749)) < 1e-9
    assert abs(O.fp(vk) - (-46.530782983591152)) < 1e-9


def test_df_get_jk_weigend(h2o):
    # pyscf/df/test/test_df_jk.py:144-156: DF J/K fingerprints, H2O cc-pVDZ / weigend, 2 random DMs
    aux = make_auxmol(h2o, 'weigend')
    cderi, nao = O.cholesky_eri(h2o, aux)
    np.random.seed(1)
    dms = np.random.random((2, nao, nao))
    vj, vk = O.df_get_jk(cderi, nao, dms)
    assert abs(O.fp(vj) - (-194.15910890730066)) < 1e-8
    assert abs(O.fp(vk) - (-46.365071587653517)) < 1e-8


def test_short_range_jk_631g():
    # pyscf/scf/test/test_vhf.py:183-198 (omega = 0.15, erfc operator), fp of stacked [vj, vk]
    mol = gto.M(atom=H2O, basis='6-31g')
    np.random.seed(1)
    dm = np.random.random((mol.nao, mol.nao))
    vj, vk = O.get_jk(mol, dm, omega=-0.15)
    assert abs(O.fp(np.array([vj, vk])) - 25.317344717490613) < 1e-9


def test_direct_jk_631g_and_veff_norms(h2o):
    # pyscf/scf/test/test_vhf.py:158-173: K ('jk->s1il') and J ('ji->s1kl') of a symmetrised random density, H2O/6-31G
    mol = gto.M(atom=H2O, basis='6-31g')
    np.random.seed(1)
    dm = np.random.random((mol.nao, mol.nao))
    dm = dm + dm.T
    vj, vk = O.get_jk(mol, dm)
    assert abs(O.fp(vk) - 5.0067176755619975) < 1e-9 and abs(O.fp(vj) - 48.61070262547175) < 1e-9
    # pyscf/scf/test/test_rhf.py:453-460: || J - K/2 || of two densities, H2O/cc-pVDZ
    nao = h2o.nao
    np.random.seed(1)
    d1 = np.random.random((nao, nao))
    d2 = np.random.random((nao, nao))
    d = np.array((d1 + d1.T, d2 + d2.T))
    vj, vk = O.get_jk(h2o, d)
    assert abs(np.linalg.norm(vj - .5 * vk) - 199.66041114502335) < 1e-9


def test_overlap_is_normalised(h2o):
    s = O.int1e(h2o, 'ovlp')
    assert abs(np.diag(s) - 1).max() < 1e-12
