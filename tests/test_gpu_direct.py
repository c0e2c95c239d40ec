"""GPU parity tests of the 4-center direct J/K path: CUDA kernels (through the C ABI) vs the CPU oracle,
the committed golden vectors and the reference's fingerprints.  Tolerance: 1e-9 Eh max-abs (north_star)."""
import os

import numpy as np
import pytest

from pyscf_b200 import gto
from pyscf_b200.gto.mole import geometry
from pyscf_b200.jk import VHFOpt, get_jk
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9
H2O = 'O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587'
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def parity_dm(nao, seed=1):
    np.random.seed(seed)
    dm = np.random.random((nao, nao))
    return dm + dm.T


@pytest.mark.parametrize('basis', ['sto-3g', '6-31g', 'cc-pvdz', 'cc-pvtz', 'def2-svp', 'def2-tzvp'])
def test_h2o_parity(basis):
    mol = gto.M(atom=H2O, basis=basis)
    dm = parity_dm(mol.nao)
    opt = VHFOpt(mol)
    vj, vk = opt.get_jk(dm, hermi=1)
    rj, rk = O.get_jk(mol, dm)
    assert abs(vj - rj).max() < TOL and abs(vk - rk).max() < TOL
    assert abs(vj - vj.T).max() < 1e-12 and abs(vk - vk.T).max() < 1e-12


def test_reference_fingerprints_on_gpu():
    # pyscf/scf/test/test_rhf.py:896-958
    mol = gto.M(atom=H2O, basis='cc-pvdz')
    nao = mol.nao
    opt = VHFOpt(mol)
    np.random.seed(1)
    dm = np.random.random((nao, nao))
    vj, vk = opt.get_jk(dm, hermi=0)
    assert abs(np.linalg.norm(vj) - 77.035779188661465) < TOL
    assert abs(O.fp(vk) - (-12.365527167710301)) < TOL
    vj, vk = opt.get_jk(np.eye(nao), hermi=1)
    assert abs(O.fp(vj) - 1.6593323222866125) < TOL and abs(O.fp(vk) - (-1.4662135224053987)) < TOL
    vj, vk = VHFOpt(mol, omega=1.5).get_jk(dm, hermi=0)
    assert abs(O.fp(vj) - (-10.015956161068031)) < TOL and abs(O.fp(vk) - (-11.399103957754445)) < TOL
    # pyscf/df/test/test_df_jk.py:157-165 (non-DF branch), two DMs
    np.random.seed(1)
    dms = np.random.random((2, nao, nao))
    vj, vk = get_jk(mol, dms, hermi=0)
    assert abs(O.fp(vj) - (-194.08878302990749)) < TOL and abs(O.fp(vk) - (-46.530782983591152)) < TOL


@pytest.mark.parametrize('name,geom,basis', [('bz_dz', 'benzene', 'cc-pvdz'), ('bz_tz', 'benzene', 'cc-pvtz'),
                                             ('h2o_tz', 'h2o', 'cc-pvtz')])
def test_golden_vectors(name, geom, basis):
    g = np.load(os.path.join(GOLD, 'jk_%s.npz' % name))
    mol = gto.M(atom=geometry(geom), basis=basis)
    dm = parity_dm(mol.nao)
    vj, vk = VHFOpt(mol).get_jk(dm, hermi=1)
    assert abs(vj - g['vj']).max() < TOL and abs(vk - g['vk']).max() < TOL
    assert abs(O.fp(vj) - float(g['fp_j'])) < 1e-7 and abs(O.fp(vk) - float(g['fp_k'])) < 1e-7


def test_full_size_properties_benzene_tz():
    """BASELINE config 2 at full size: size-independent properties (linearity, symmetry, hermi split,
    J/K-only calls, batching) plus agreement of repeated builds to accumulation-order noise."""
    mol = gto.M(atom=geometry('benzene'), basis='cc-pvtz')
    nao = mol.nao
    opt = VHFOpt(mol)
    rng = np.random.RandomState(5)
    a = rng.random_sample((nao, nao)); a = a + a.T
    b = rng.random_sample((nao, nao)); b = b + b.T
    (ja, jb), (ka, kb) = [x for x in opt.get_jk(np.array([a, b]), hermi=1)]
    jab, kab = opt.get_jk(0.3 * a - 1.7 * b, hermi=1)
    assert abs(jab - (0.3 * ja - 1.7 * jb)).max() < 1e-9 * 50
    assert abs(kab - (0.3 * ka - 1.7 * kb)).max() < 1e-9 * 50
    assert abs(ja - ja.T).max() < 1e-11 and abs(ka - ka.T).max() < 1e-11
    # general matrix = sym + antisym; J ignores the antisymmetric part, K of it is antisymmetric
    g = rng.random_sample((nao, nao))
    jg, kg = opt.get_jk(g, hermi=0)
    js, ks = opt.get_jk(0.5 * (g + g.T), hermi=1)
    j2, k2 = opt.get_jk(0.5 * (g - g.T), hermi=2)
    assert abs(jg - js).max() < 1e-9 and abs(kg - (ks + k2)).max() < 1e-9
    assert abs(k2 + k2.T).max() < 1e-11
    # energy-like checksum: tr(D J) symmetric bilinear form
    assert abs(np.einsum('ij,ji', a, jb) - np.einsum('ij,ji', b, ja)) < 1e-7
    assert abs(np.einsum('ij,ji', a, kb) - np.einsum('ij,ji', b, ka)) < 1e-7
    j_only, none = opt.get_jk(a, with_k=False)
    assert none is None and abs(j_only - ja).max() < 1e-10
    none, k_only = opt.get_jk(a, with_j=False)
    assert none is None and abs(k_only - ka).max() < 1e-10


def test_edge_cases():
    # single atom, single s shell; zero density; ragged batch dims; error behaviour
    mol = gto.M(atom='He 0 0 0', basis='sto-3g')
    vj, vk = VHFOpt(mol).get_jk(np.ones((1, 1)))
    rj, rk = O.get_jk(mol, np.ones((1, 1)))
    assert abs(vj - rj).max() < TOL and abs(vk - rk).max() < TOL
    mol = gto.M(atom=H2O, basis='cc-pvdz')
    nao = mol.nao
    opt = VHFOpt(mol)
    vj, vk = opt.get_jk(np.zeros((nao, nao)))
    assert abs(vj).max() == 0 and abs(vk).max() == 0
    dms = np.random.RandomState(0).random_sample((2, 3, nao, nao))
    vj, vk = opt.get_jk(dms, hermi=0)
    assert vj.shape == (2, 3, nao, nao)
    rj, rk = O.get_jk(mol, dms)
    assert abs(vj - rj).max() < TOL and abs(vk - rk).max() < TOL
    with pytest.raises(RuntimeError):
        opt.get_jk(np.zeros((nao + 1, nao + 1)))
    # complex density: real and imaginary parts contracted separately (pyscf/scf/hf.py:1017-1031)
    dmc = dms[0, 0] + 1j * dms[0, 1]
    vj, vk = opt.get_jk(dmc, hermi=0)
    assert abs(vj - (rj[0, 0] + 1j * rj[0, 1])).max() < TOL and abs(vk - (rk[0, 0] + 1j * rk[0, 1])).max() < TOL


def test_screening_on_device():
    mol = gto.M(atom='O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587; O 0 0 12; H 0 -0.757 12.587; H 0 0.757 12.587',
                basis='cc-pvdz')
    dm = parity_dm(mol.nao) * 1e-2
    opt = VHFOpt(mol)
    vj, vk = opt.get_jk(dm)
    rj, rk = O.get_jk(mol, dm, screen=False)
    assert abs(vj - rj).max() < TOL and abs(vk - rk).max() < TOL
