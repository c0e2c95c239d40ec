"""Nuclear-gradient J/K (pyscf/grad/rhf.py:191-235: vj[x,i,j] = -sum_kl (nabla_x i j|kl) D_lk,
vk[x,i,l] = -sum_jk (nabla_x i j|kl) D_jk from int2e_ip1).  The oracle has no derivative integrals, so the checks are finite
differences of the oracle's ORDINARY J/K under a displacement of one atom (independent of how pyscf_b200.grad builds them):
  * element-wise: for a density supported off atom A, d J_ij / d A_x = vj[x,i,j] and d K_il / d A_x = vk[x,i,l] for i on A, j / l off A;
  * contracted: d E_J / d A_x = 2 sum_{i on A, j} D_ij vj[x,i,j], d E_K / d A_x = 2 sum_{i on A, l} D_il vk[x,i,l] for any symmetric D."""
import numpy as np
import pytest

from pyscf_b200 import gto, grad
from oracle import oracle as O

ATOMS = [('O', (0.0, 0.0, 0.0)), ('H', (0.0, -0.757, 0.587)), ('H', (0.3, 0.757, 0.587))]


def _mol(basis, shift=None):
    atoms = [(s, list(r)) for s, r in ATOMS]
    if shift is not None:
        a, x, h = shift
        atoms[a][1][x] += h * 0.52917721092          # h in Bohr, geometry in Angstrom
    return gto.M(atom=[(s, tuple(r)) for s, r in atoms], basis=basis)


def _fd(f, h):
    """fourth-order central difference of a tuple-valued function of the displacement"""
    p1, m1, p2, m2 = f(h), f(-h), f(2 * h), f(-2 * h)
    return [(8.0 * (a - b) - (c - e)) / (12.0 * h) for a, b, c, e in zip(p1, m1, p2, m2)]


def _check(libpath, basis):
    mol = _mol(basis)
    nao = mol.nao
    loc = mol.ao_loc_nr()
    on = [np.concatenate([np.arange(loc[b], loc[b + 1]) for b in range(mol.nbas) if mol._bas[b, 0] == a]) for a in range(3)]
    rng = np.random.RandomState(4)
    h = 4e-3
    for A in (0, 1):
        off = np.setdiff1d(np.arange(nao), on[A])
        # ---- element-wise, density supported off atom A
        d = np.zeros((nao, nao))
        r = rng.standard_normal((len(off), len(off)))
        d[np.ix_(off, off)] = r + r.T
        vj, vk = grad.get_jk(mol, d, libpath=libpath)
        assert vj.shape == (3, nao, nao) and vk.shape == (3, nao, nao)
        for x in range(3):
            fj, fk = _fd(lambda s_: O.get_jk(_mol(basis, (A, x, s_)), d), h)
            sel = np.ix_(on[A], off)
            assert abs(vj[x][sel] - fj[sel]).max() < 1e-8, (basis, A, x, abs(vj[x][sel] - fj[sel]).max())
            assert abs(vk[x][sel] - fk[sel]).max() < 1e-8, (basis, A, x, abs(vk[x][sel] - fk[sel]).max())
    # ---- contracted, any symmetric density
    r = rng.standard_normal((nao, nao)) * 0.3
    d = r + r.T
    vj, vk = grad.get_jk(mol, d, libpath=libpath)
    assert abs(grad.get_veff(mol, d, libpath=libpath) - (vj - 0.5 * vk)).max() < 1e-12
    for A in range(3):
        for x in range(3):
            fe = _fd(lambda s_: [np.array(0.5 * np.einsum('ij,ji', v, d)) for v in O.get_jk(_mol(basis, (A, x, s_)), d)], h)
            for v, num in ((vj, fe[0]), (vk, fe[1])):
                ana = 2.0 * np.einsum('ij,ij', v[x][on[A]], d[on[A]])
                assert abs(ana - num) < 1e-7 * max(1.0, abs(ana)), (basis, A, x, ana, float(num))
    with pytest.raises(NotImplementedError):
        grad.get_jk(gto.M(atom='Ne 0 0 0', basis='cc-pvtz'), np.eye(30), libpath=libpath)


def test_grad_jk_emulated(emu_lib):
    _check(emu_lib, '6-31g')           # s, p shells (general-contracted-free), companions up to d
    _check(emu_lib, 'cc-pvdz')         # general contractions and d shells: companions up to f


@pytest.mark.gpu
def test_grad_jk_gpu():
    _check(None, 'cc-pvdz')
