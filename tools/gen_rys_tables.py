#!/usr/bin/env python
"""Generate the Rys-quadrature root/weight tables used by the CUDA kernels.

For nroots n = 1..NMAX and x in [0, XMAX): piecewise Chebyshev interpolants (degree DEG on
intervals of width H) of the nodes u_r(x) = t_r^2 and weights w_r(x) of the n-point Gauss rule for
    int_0^1 f(t^2) exp(-x t^2) dt ,
whose moments in u = t^2 are the Boys functions F_k(x).  The exact nodes/weights at the Chebyshev
points come from a Golub-Welsch construction in 80-digit arithmetic (mpmath), i.e. from first
principles — no table from libcint or any other package is used.  For x >= XMAX the rule is the
positive half of the 2n-point Gauss-Hermite rule: u_r = s_r^2 / x, w_r = W_r / sqrt(x).

Output: pyscf_b200/csrc/rys_tables.bin  (float64, little endian)
    header  : [NMAX, DEG, NINT, H, XMAX]              (5 doubles)
    hermite : for n in 1..NMAX: u_r*x (n), w_r*sqrt(x) (n)
    cheb    : for n in 1..NMAX: [NINT][n][2][DEG+1]    coefficients c_0..c_DEG (c_0 already halved)
and a self-check against fresh mpmath evaluations at random x.
"""
import os
import struct
import sys
from multiprocessing import Pool

import mpmath as mp
import numpy as np

NMAX = 9
DEG = 7
H = 0.3125
NINT = 320
XMAX = H * NINT
mp.mp.dps = 80


def boys_moments(x, nm):
    x = mp.mpf(x)
    if x == 0:
        return [mp.mpf(1) / (2 * k + 1) for k in range(nm)]
    return [mp.gammainc(k + mp.mpf(1) / 2, 0, x) / (2 * x ** (k + mp.mpf(1) / 2)) for k in range(nm)]


def gauss_from_moments(mu, n):
    """Golub-Welsch from moments via Cholesky of the Hankel matrix."""
    Hm = mp.matrix(n + 1, n + 1)
    for i in range(n + 1):
        for j in range(n + 1):
            Hm[i, j] = mu[i + j]
    L = mp.cholesky(Hm)  # H = L L^T ; R = L^T upper
    R = L.T
    alpha = []
    beta = []
    for j in range(n):
        a = R[j, j + 1] / R[j, j]
        if j > 0:
            a -= R[j - 1, j] / R[j - 1, j - 1]
        alpha.append(a)
    for j in range(1, n):
        beta.append(R[j, j] / R[j - 1, j - 1])
    J = mp.matrix(n, n)
    for j in range(n):
        J[j, j] = alpha[j]
    for j in range(n - 1):
        J[j, j + 1] = beta[j]
        J[j + 1, j] = beta[j]
    E, Q = mp.eigsy(J)
    idx = sorted(range(n), key=lambda i: E[i])
    u = [E[i] for i in idx]
    w = [mu[0] * Q[0, i] ** 2 for i in idx]
    return u, w


def rys_exact(n, x):
    mu = boys_moments(x, 2 * n + 1)
    return gauss_from_moments(mu, n)


def hermite_half(n):
    mu = [mp.gamma(k + mp.mpf(1) / 2) / 2 for k in range(2 * n + 1)]
    return gauss_from_moments(mu, n)


def fit_interval(args):
    n, iv = args
    mp.mp.dps = 80
    x0 = iv * H
    # Chebyshev nodes of the first kind
    K = DEG + 1
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / K) for k in range(K)]
    vals_u = [[None] * K for _ in range(n)]
    vals_w = [[None] * K for _ in range(n)]
    for k, tk in enumerate(nodes):
        x = x0 + H * (tk + 1) / 2
        u, w = rys_exact(n, x)
        for r in range(n):
            vals_u[r][k] = u[r]
            vals_w[r][k] = w[r]
    out = np.zeros((n, 2, K))
    for r in range(n):
        for which, vals in enumerate((vals_u[r], vals_w[r])):
            for j in range(K):
                c = mp.mpf(2) / K * sum(vals[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / K) for k in range(K))
                if j == 0:
                    c /= 2
                out[r, which, j] = float(c)
    return n, iv, out


def cheb_eval(c, t):
    # c_0 already halved: f = sum_j c_j T_j(t)
    b1 = b2 = 0.0
    for j in range(len(c) - 1, 0, -1):
        b1, b2 = 2 * t * b1 - b2 + c[j], b1
    return t * b1 - b2 + c[0]


def main():
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'pyscf_b200', 'csrc',
                            'rys_tables.bin')
    tasks = [(n, iv) for n in range(1, NMAX + 1) for iv in range(NINT)]
    tabs = {n: np.zeros((NINT, n, 2, DEG + 1)) for n in range(1, NMAX + 1)}
    with Pool(8) as pool:
        for n, iv, arr in pool.imap_unordered(fit_interval, tasks, chunksize=4):
            tabs[n][iv] = arr
    herm = []
    for n in range(1, NMAX + 1):
        u, w = hermite_half(n)
        herm.append(np.array([float(v) for v in u] + [float(v) for v in w]))
    with open(out_path, 'wb') as f:
        f.write(struct.pack('<5d', NMAX, DEG, NINT, H, XMAX))
        for h in herm:
            f.write(h.astype('<f8').tobytes())
        for n in range(1, NMAX + 1):
            f.write(tabs[n].astype('<f8').tobytes())
    print('wrote', out_path, os.path.getsize(out_path), 'bytes')

    # self check
    rng = np.random.RandomState(7)
    worst_u = worst_w = 0.0
    for n in range(1, NMAX + 1):
        xs = list(rng.uniform(0, XMAX, 12)) + [0.0, 1e-9, H - 1e-12, H + 1e-12, XMAX - 1e-9]
        for x in xs:
            iv = min(int(x / H), NINT - 1)
            t = (x - iv * H) * 2 / H - 1
            u, w = rys_exact(n, x)
            for r in range(n):
                eu = abs(cheb_eval(tabs[n][iv, r, 0], t) - float(u[r]))
                ew = abs(cheb_eval(tabs[n][iv, r, 1], t) - float(w[r]))
                worst_u = max(worst_u, eu)
                worst_w = max(worst_w, ew)
        # asymptotic branch at XMAX
        u, w = rys_exact(n, XMAX)
        hu, hw = herm[n - 1][:n], herm[n - 1][n:]
        eu = max(abs(hu[r] / XMAX - float(u[r])) for r in range(n))
        ew = max(abs(hw[r] / np.sqrt(XMAX) - float(w[r])) for r in range(n))
        print('n=%d  cheb max abs err so far: u %.2e w %.2e ; asymptotic@XMAX: u %.2e w %.2e' % (n, worst_u, worst_w, eu, ew))


if __name__ == '__main__':
    main()
