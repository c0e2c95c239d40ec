"""A minimal restricted SCF driver used to exercise the J/K builders end to end.

This is NOT a replacement for pyscf.scf (out of scope): it mirrors the control flow of
scf.hf.kernel (pyscf/scf/hf.py:49-241: core-Hamiltonian guess, Fock = h + J - K/2, Pulay DIIS, generalized
eigenproblem, aufbau occupation, energy_elec :287-293) just enough to check that the converged total
energy obtained through `get_jk` equals the reference's published value.  One-electron matrices are inputs.
"""
import numpy as np
import scipy.linalg


class RHF:
    def __init__(self, mol, get_jk, hcore, ovlp, conv_tol=1e-10, max_cycle=50):
        self.mol, self._get_jk, self.hcore, self.ovlp = mol, get_jk, hcore, ovlp
        self.conv_tol, self.max_cycle = conv_tol, max_cycle
        self.e_tot = None
        self.converged = False

    def _eig(self, f):
        e, c = scipy.linalg.eigh(f, self.ovlp)
        return e, c

    def make_rdm1(self, c, nocc):
        co = c[:, :nocc]
        return 2.0 * co.dot(co.T), co

    def kernel(self):
        nocc = self.mol.nelectron // 2
        e, c = self._eig(self.hcore)
        dm, co = self.make_rdm1(c, nocc)
        enuc = self.mol.energy_nuc()
        fs, es = [], []
        e_last = 0.0
        for it in range(self.max_cycle):
            vj, vk = self._get_jk(dm, co)
            f = self.hcore + vj - 0.5 * vk
            e_tot = 0.5 * np.einsum('ij,ji', dm, self.hcore + f) + enuc
            err = f.dot(dm).dot(self.ovlp)
            err = err - err.T
            fs.append(f); es.append(err)
            fs, es = fs[-8:], es[-8:]
            if len(fs) > 1:   # Pulay DIIS
                n = len(fs)
                b = np.zeros((n + 1, n + 1))
                for i in range(n):
                    for j in range(n):
                        b[i, j] = np.vdot(es[i], es[j])
                b[n, :n] = b[:n, n] = -1.0
                rhs = np.zeros(n + 1); rhs[n] = -1.0
                try:
                    w = np.linalg.solve(b, rhs)[:n]
                    f = sum(wi * fi for wi, fi in zip(w, fs))
                except np.linalg.LinAlgError:
                    pass
            if abs(e_tot - e_last) < self.conv_tol and np.abs(err).max() < 1e-6:
                self.converged = True
                self.e_tot = e_tot
                break
            e_last = e_tot
            e, c = self._eig(f)
            dm, co = self.make_rdm1(c, nocc)
        self.e_tot = e_tot
        self.mo_energy, self.mo_coeff = e, c
        return e_tot
