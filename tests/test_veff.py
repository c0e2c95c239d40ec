"""get_veff on top of the J/K builders (pyscf_b200/veff.py): incremental (difference-density) Fock builds of the direct-SCF
loop (pyscf/scf/hf.py:2172-2201, uhf.py:1066-1095) and the hybrid / range-separated exchange mixes of RKS
(pyscf/dft/rks.py:98-127), checked against the CPU oracle."""
import numpy as np
import pytest

from pyscf_b200 import gto, veff
from pyscf_b200.df import DF
from pyscf_b200.jk import VHFOpt
from oracle import oracle as O

H2O = 'O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587'


def _sym(n, seed):
    a = np.random.RandomState(seed).random_sample((n, n))
    return a + a.T


def _check_incremental(libpath):
    mol = gto.M(atom=H2O, basis='cc-pvdz')
    nao = mol.nao
    get_jk = veff.make_get_jk(mol, libpath=libpath)
    dm0 = _sym(nao, 1)
    # difference density 1e-3 (D' - D) exercises the density half of CVHFnrs8_prescreen
    dm1 = dm0 + 1e-3 * (_sym(nao, 2) - dm0)
    v0 = veff.get_veff_rhf(get_jk, mol, dm0)
    rj, rk = O.get_jk(mol, dm0)
    assert abs(v0 - (rj - .5 * rk)).max() < 1e-9
    assert abs(v0.ecoul - .5 * np.einsum('ij,ji', dm0, rj)) < 1e-9
    opt = get_jk.opts[None]
    n_full = opt.stats()['quartets_computed']
    v1 = veff.get_veff_rhf(get_jk, mol, dm1, dm_last=dm0, vhf_last=v0)
    n_incr = opt.stats()['quartets_computed']
    rj, rk = O.get_jk(mol, dm1)
    assert abs(v1 - (rj - .5 * rk)).max() < 1e-9
    assert abs(v1.ecoul - .5 * np.einsum('ij,ji', dm1, rj)) < 1e-8
    assert n_incr <= n_full                       # a smaller density can only screen more
    # converged SCF: a 1e-9 change of the density leaves (almost) nothing to compute
    veff.get_veff_rhf(get_jk, mol, dm1 + 1e-9 * (dm1 - dm0), dm_last=dm1, vhf_last=v1)
    assert opt.stats()['quartets_computed'] < n_full
    # not direct_scf: full build
    v1b = veff.get_veff_rhf(get_jk, mol, dm1, dm_last=dm0, vhf_last=v0, direct_scf=False)
    assert abs(v1b - v1).max() < 1e-9
    # UHF: (alpha, beta) pair, incremental
    dma = np.array([_sym(nao, 3), _sym(nao, 4)])
    dmb = dma + 1e-2 * np.array([_sym(nao, 5), _sym(nao, 6)])
    u0 = veff.get_veff_uhf(get_jk, mol, dma)
    u1 = veff.get_veff_uhf(get_jk, mol, dmb, dm_last=dma, vhf_last=u0)
    rj, rk = O.get_jk(mol, dmb)
    assert abs(u1 - (rj[0] + rj[1] - rk)).max() < 1e-9
    assert abs(u1.ecoul - .5 * np.einsum('nij,ji->', dmb, rj[0] + rj[1])) < 1e-8
    r2 = veff.get_veff_uhf(get_jk, mol, dm0)       # 2-D dm treated as a closed-shell density
    assert abs(r2[0] - v0).max() < 1e-9 and abs(r2[1] - v0).max() < 1e-9


def _check_rks_mixes(libpath):
    mol = gto.M(atom=H2O, basis='6-31g')
    nao = mol.nao
    get_jk = veff.make_get_jk(mol, libpath=libpath)
    dm = _sym(nao, 7)
    rj, rk = O.get_jk(mol, dm)
    for xc in ('pbe', 'b3lyp', 'hse06', 'lc-wpbe', 'wb97x', 'camb3lyp'):
        omega, alpha, hyb = veff.RSH_AND_HYBRID_COEFF[xc.replace('-', '')]
        v = veff.get_veff_rks(get_jk, mol, dm, xc=xc)
        if omega:
            rklr = O.get_jk(mol, dm, omega=omega)[1]
            ref_k = hyb * (rk - rklr) + alpha * rklr        # hyb x short range + alpha x long range
        else:
            ref_k = hyb * rk
        assert abs(v.vj - rj).max() < 1e-9, xc
        if hyb == 0 and alpha == 0:
            assert v.vk is None and abs(v - rj).max() < 1e-9
        else:
            assert abs(v.vk - ref_k).max() < 1e-9, xc
            assert abs(v - (rj - .5 * ref_k)).max() < 1e-9, xc
            assert abs(v.exc + .25 * np.einsum('ij,ji', dm, ref_k)) < 1e-9
    # incremental RKS build reuses the tagged vj / vk of the previous cycle (rks.py:98-103,128-131)
    dm2 = dm + 1e-3 * _sym(nao, 8)
    v0 = veff.get_veff_rks(get_jk, mol, dm, xc='wb97x')
    v1 = veff.get_veff_rks(get_jk, mol, dm2, xc='wb97x', dm_last=dm, vhf_last=v0)
    v1f = veff.get_veff_rks(get_jk, mol, dm2, xc='wb97x')
    assert abs(v1 - v1f).max() < 1e-9 and abs(v1.ecoul - v1f.ecoul) < 1e-8
    # a semilocal term supplied by the caller is added untouched
    vx = _sym(nao, 9)
    v2 = veff.get_veff_rks(get_jk, mol, dm, xc='b3lyp', nr_rks=lambda d: (10.0, -1.25, vx))
    v2r = veff.get_veff_rks(get_jk, mol, dm, xc='b3lyp')
    assert abs(v2 - v2r - vx).max() < 1e-12 and abs(v2.exc - v2r.exc + 1.25) < 1e-12


def _check_rks_df(libpath):
    # the same mixes through a DF object: range_coulomb tensors are built on demand (pyscf/df/df.py:298-333)
    mol = gto.M(atom=H2O, basis='6-31g')
    d = DF(mol, 'weigend', libpath=libpath).build()
    get_jk = veff.make_get_jk(mol, with_df=d)
    dm = _sym(mol.nao, 11)
    v = veff.get_veff_rks(get_jk, mol, dm, xc='wb97x')
    omega, alpha, hyb = veff.RSH_AND_HYBRID_COEFF['wb97x']
    aux = d.auxmol
    c0, nao = O.cholesky_eri(mol, aux)
    c1, _ = O.cholesky_eri(mol, aux, omega=omega)
    rj, rk = O.df_get_jk(c0, nao, dm)
    rklr = O.df_get_jk(c1, nao, dm)[1]
    assert abs(v.vj - rj).max() < 1e-9
    assert abs(v.vk - (hyb * rk + (alpha - hyb) * rklr)).max() < 1e-9


def _check_reference_fingerprints(libpath):
    # pyscf/scf/test/test_rhf.py:453-460: scf.hf.get_veff of two densities, H2O/cc-pVDZ
    mol = gto.M(atom=H2O, basis='cc-pvdz')
    nao = mol.nao
    np.random.seed(1)
    d1 = np.random.random((nao, nao))
    d2 = np.random.random((nao, nao))
    d = np.array((d1 + d1.T, d2 + d2.T))
    get_jk = veff.make_get_jk(mol, libpath=libpath)
    v = veff.get_veff_rhf(get_jk, mol, d)
    assert abs(np.linalg.norm(v) - 199.66041114502335) < 1e-9
    # pyscf/df/test/test_df_jk.py:127-133: UHF get_veff through density fitting, dm of shape (2, 4, nao, nao), hermi=0
    dfobj = DF(mol, 'weigend', libpath=libpath).build()
    np.random.seed(1)
    dm = np.random.random((2, 4, nao, nao))
    vhf = veff.get_veff_uhf(veff.make_get_jk(mol, with_df=dfobj), mol, dm, hermi=0)
    assert vhf.shape == (2, 4, nao, nao) and abs(np.linalg.norm(vhf) - 413.82341595365853) < 1e-9
    # pyscf/scf/test/test_vhf.py:158-173: K and J of a symmetrised random density, H2O/6-31G
    mol = gto.M(atom=H2O, basis='6-31g')
    np.random.seed(1)
    dm = np.random.random((mol.nao, mol.nao))
    dm = dm + dm.T
    vj, vk = veff.make_get_jk(mol, libpath=libpath)(mol, dm)
    assert abs(O.fp(vk) - 5.0067176755619975) < 1e-9 and abs(O.fp(vj) - 48.61070262547175) < 1e-9


def test_reference_fingerprints_emulated(emu_lib):
    _check_reference_fingerprints(emu_lib)


@pytest.mark.gpu
def test_reference_fingerprints_gpu():
    _check_reference_fingerprints(None)


def test_patch_instance_override(emu_lib):
    """jk.patch(mf) installs the builder as mf.get_jk, the instance-override hook PySCF documents
    (examples/scf/43-custom_get_jk.py:36-45), with SCF.get_jk's defaults (mol=None -> mf.mol, dm=None -> mf.make_rdm1(),
    one cached optimizer per omega; pyscf/scf/hf.py:2136-2160).  A stand-in object plays the SCF instance."""
    from pyscf_b200 import jk

    class FakeSCF:
        def __init__(self, mol, dm):
            self.mol, self._dm, self.direct_scf_tol = mol, dm, 1e-13

        def make_rdm1(self):
            return self._dm

        def get_veff(self, mol=None, dm=None):          # what the reference's get_veff does with the hook
            vj, vk = self.get_jk(mol, dm, 1)
            return vj - vk * .5

    mol = gto.M(atom=H2O, basis='6-31g')
    dm = _sym(mol.nao, 2)
    mf = jk.patch(FakeSCF(mol, dm), libpath=emu_lib)
    rj, rk = O.get_jk(mol, dm)
    vj, vk = mf.get_jk()
    assert abs(vj - rj).max() < 1e-10 and abs(vk - rk).max() < 1e-10
    assert abs(mf.get_veff() - (rj - .5 * rk)).max() < 1e-10
    vj, vk = mf.get_jk(mol, dm, hermi=1, with_k=False)
    assert vk is None and abs(vj - rj).max() < 1e-10
    vklr = mf.get_jk(mol, dm, with_j=False, omega=0.4)[1]
    assert abs(vklr - O.get_jk(mol, dm, omega=0.4)[1]).max() < 1e-10
    assert len(mf._b200_opts) == 2                       # one optimizer per omega, reused on the next call
    mf.get_jk(omega=0.4)
    assert len(mf._b200_opts) == 2


def test_incremental_veff_emulated(emu_lib):
    _check_incremental(emu_lib)


def test_rks_exchange_mixes_emulated(emu_lib):
    _check_rks_mixes(emu_lib)


def test_rks_df_emulated(emu_lib):
    _check_rks_df(emu_lib)


@pytest.mark.gpu
def test_incremental_veff_gpu():
    _check_incremental(None)


@pytest.mark.gpu
def test_rks_exchange_mixes_gpu():
    _check_rks_mixes(None)


@pytest.mark.gpu
def test_rks_df_gpu():
    _check_rks_df(None)
