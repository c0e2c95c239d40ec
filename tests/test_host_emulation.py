"""CPU tests of the host logic (Mole tables, C-ABI loading, error behaviour) and of the kernel
arithmetic through the CPU SIMT emulation of the CUDA templates (tests/emu/libb200jk_emu.so, built
from the SAME sources with -DB200JK_EMULATE).  The emulation is test infrastructure only."""
import ctypes
import os

import numpy as np
import pytest

from pyscf_b200 import gto, lib as b2lib
from pyscf_b200.gto.mole import geometry, make_auxmol
from pyscf_b200.jk import VHFOpt
from oracle import oracle as O

H2O = 'O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_table_sizes_match_survey():
    # nbas / nao / naux of the benchmark configs
    m = gto.M(atom=H2O, basis='sto-3g')
    assert (m.nbas, m.nao, m.nelectron) == (5, 7, 10)
    m = gto.M(atom=geometry('benzene'), basis='cc-pvtz')
    assert (m.nbas, m.nao, m.nelectron) == (90, 264, 42)
    m = gto.M(atom=geometry('c60'), basis='def2-svp')
    assert (m.nbas, m.nao) == (360, 840)
    aux = make_auxmol(m)
    assert (aux.nbas, aux.nao, int(aux._bas[:, 1].max())) == (1500, 4500, 4)
    m = gto.M(atom=geometry('gly30'), basis='cc-pvdz')       # config 5: C60H92N30O31
    assert (m.natm, m.nbas, m.nao, m.nelectron // 2) == (213, 881, 2154, 455)
    aux = make_auxmol(m)
    assert (aux.nbas, aux.nao) == (3732, 10586)
    m = gto.M(atom=geometry('taxol'), basis='def2-tzvp')     # config 4: C47H51NO14 (tools/make_taxol.py)
    assert (m.natm, m.nbas, m.nao, m.nelectron // 2, int(m._bas[:, 1].max())) == (113, 886, 2228, 226, 3)
    aux = make_auxmol(m)
    assert (aux.nbas, aux.nao, int(aux._bas[:, 1].max())) == (1856, 5598, 4)
    # chemically sane: no two atoms closer than a bond, no non-hydrogen pair closer than 1.19 A (C=O)
    r = m.atom_coords() * 0.52917721092
    d = np.sqrt(((r[:, None] - r[None]) ** 2).sum(-1)) + 10 * np.eye(m.natm)
    heavy = m._atm[:, 0] > 1
    assert d.min() > 0.94 and d[np.ix_(heavy, heavy)].min() > 1.19


def test_env_layout():
    m = gto.M(atom=H2O, basis='sto-3g')
    # pyscf/gto/mole.py:58-88: PTR_ENV_START = 20, coordinates in Bohr
    assert m._atm[0, 1] == 20 and m._atm.shape == (3, 6) and m._bas.shape == (5, 8)
    assert abs(m._env[m._atm[1, 1] + 1] - (-0.757 / 0.52917721092)) < 1e-14
    # gto_norm(0, 1) documented value, pyscf/gto/mole.py:146-147
    from pyscf_b200.gto.mole import gto_norm
    assert abs(gto_norm(0, 1.0) - 2.5264751109842591) < 1e-14


def test_cabi_library_exports_every_symbol():
    path = b2lib.DEFAULT_LIB
    assert os.path.exists(path), 'build the library first (python -c "import __graft_entry__ as g; g.build()")'
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, 'include', 'b200jk.h')).read()
    import re
    declared = sorted(set(re.findall(r'\b(b200jk_[a-z0-9_]+)\s*\(', hdr)))
    assert len(declared) >= 14
    for s in declared:
        assert hasattr(lib, s), s
    lib.b200jk_version.restype = ctypes.c_char_p
    assert b'sm_100a' in lib.b200jk_version()


def test_no_silent_cpu_fallback():
    # without a GPU the product library must fail loudly, never compute on the CPU
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    m = gto.M(atom=H2O, basis='sto-3g')
    with pytest.raises(RuntimeError, match='no CUDA device'):
        VHFOpt(m)


@pytest.mark.parametrize('basis', ['sto-3g', '6-31g', 'cc-pvdz'])
def test_emulated_kernels_match_oracle(emu_lib, basis):
    mol = gto.M(atom=H2O, basis=basis)
    nao = mol.nao
    np.random.seed(1)
    dm = np.random.random((nao, nao))
    dm = dm + dm.T
    opt = VHFOpt(mol, libpath=emu_lib)
    vj, vk = opt.get_jk(dm, hermi=1)
    rj, rk = O.get_jk(mol, dm)
    assert abs(vj - rj).max() < 1e-11 and abs(vk - rk).max() < 1e-11
    # hermi=0, two density matrices, leading dims preserved
    dms = np.random.random((2, nao, nao))
    vj, vk = opt.get_jk(dms, hermi=0)
    rj, rk = O.get_jk(mol, dms)
    assert vj.shape == dms.shape and vk.shape == dms.shape
    assert abs(vj - rj).max() < 1e-11 and abs(vk - rk).max() < 1e-11
    # with_j / with_k switches return None for the one not requested (pyscf/scf/hf.py:963)
    vj1, vk1 = opt.get_jk(dm, hermi=1, with_k=False)
    assert vk1 is None and abs(vj1 - O.get_jk(mol, dm)[0]).max() < 1e-11
    vj1, vk1 = opt.get_jk(dm, hermi=1, with_j=False)
    assert vj1 is None


def test_emulated_f_functions_and_fingerprints(emu_lib):
    mol = gto.M(atom='He 0 0 0; Ne 1.2 0.3 0', basis='cc-pvtz')  # f shell on Ne
    nao = mol.nao
    np.random.seed(3)
    dm = np.random.random((nao, nao))
    dm = dm + dm.T
    opt = VHFOpt(mol, libpath=emu_lib)
    vj, vk = opt.get_jk(dm)
    rj, rk = O.get_jk(mol, dm)
    assert abs(vj - rj).max() < 1e-10 and abs(vk - rk).max() < 1e-10


def test_emulated_reference_fingerprints(emu_lib):
    # pyscf/scf/test/test_rhf.py:896-934 through the device code path (emulated)
    mol = gto.M(atom=H2O, basis='cc-pvdz')
    nao = mol.nao
    opt = VHFOpt(mol, libpath=emu_lib)
    np.random.seed(1)
    dm = np.random.random((nao, nao))
    vj, vk = opt.get_jk(dm, hermi=0)
    assert abs(np.linalg.norm(vj) - 77.035779188661465) < 1e-9
    assert abs(O.fp(vk) - (-12.365527167710301)) < 1e-9
    vj, vk = opt.get_jk(np.eye(nao), hermi=1)
    assert abs(O.fp(vj) - 1.6593323222866125) < 1e-9 and abs(O.fp(vk) - (-1.4662135224053987)) < 1e-9
    opt_lr = VHFOpt(mol, omega=1.5, libpath=emu_lib)
    vj, vk = opt_lr.get_jk(dm, hermi=0)
    assert abs(O.fp(vj) - (-10.015956161068031)) < 1e-9 and abs(O.fp(vk) - (-11.399103957754445)) < 1e-9


def test_emulated_screening_and_errors(emu_lib):
    # a stretched molecule: screening must drop quartets without changing J/K beyond the tolerance
    mol = gto.M(atom='O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587; O 0 0 12; H 0 -0.757 12.587; H 0 0.757 12.587',
                basis='6-31g')
    nao = mol.nao
    np.random.seed(2)
    dm = np.random.random((nao, nao)) * 1e-2
    dm = dm + dm.T
    opt = VHFOpt(mol, libpath=emu_lib)
    vj, vk = opt.get_jk(dm)
    rj, rk = O.get_jk(mol, dm, screen=False)
    assert abs(vj - rj).max() < 1e-10 and abs(vk - rk).max() < 1e-10
    # a loose threshold must reject quartets on device and stay within the implied error bound
    n_all = opt.stats()['quartets_computed']
    opt2 = VHFOpt(mol, direct_scf_tol=1e-9, libpath=emu_lib)
    vj2, vk2 = opt2.get_jk(dm * 1e-6)
    st = opt2.stats()
    assert st['quartets_screened'] > 0 and st['quartets_computed'] < n_all
    assert abs(vj2 - rj * 1e-6).max() < 1e-7 and abs(vk2 - rk * 1e-6).max() < 1e-7
    with pytest.raises(RuntimeError):
        opt.get_jk(np.zeros((nao + 1, nao + 1)))
    q = opt.q_cond
    qo = O.q_cond(mol)
    big = qo > 1e-12  # negligible pairs are dropped on the device side (reported as the 1e-100 floor)
    assert q.shape == qo.shape and abs(np.log(q[big] / qo[big])).max() < 1e-9  # s,p: identical to CVHFnr_int2e_q_cond


def test_emulated_experimental_layouts(emu_lib, emu_lib_experimental):
    """Primitive batching (QClass::PB) and the part-per-warp lane layout (GroupCfg::PPW) are compile-time options that are
    off in the shipped library until measured; they must give the same J/K as the default layout (same primitive order,
    so only the order of the reductions differs) for d/f shells, hermi 0/1 and the erf / erfc operators."""
    for atom, basis, omega in [(H2O, 'cc-pvdz', None), ('He 0 0 0; Ne 1.2 0.3 0', 'cc-pvtz', None), (H2O, 'cc-pvdz', -0.4),
                               ('O 0 0 0; O 0 0 1.2', 'cc-pvdz', 0.7)]:
        mol = gto.M(atom=atom, basis=basis)
        nao = mol.nao
        np.random.seed(3)
        dms = np.random.random((2, nao, nao))
        a = VHFOpt(mol, omega=omega, libpath=emu_lib_experimental).get_jk(dms, hermi=0)
        b = VHFOpt(mol, omega=omega, libpath=emu_lib).get_jk(dms, hermi=0)
        assert abs(a[0] - b[0]).max() < 1e-11 and abs(a[1] - b[1]).max() < 1e-11
        d = dms[0] + dms[0].T
        a = VHFOpt(mol, omega=omega, libpath=emu_lib_experimental).get_jk(d, hermi=1)
        r = O.get_jk(mol, d, omega=omega)
        assert abs(a[0] - r[0]).max() < 1e-10 and abs(a[1] - r[1]).max() < 1e-10
