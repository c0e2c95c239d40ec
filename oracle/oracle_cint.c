/*
 * oracle_cint.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * A plain-C restatement of the integrals the reference obtains from libcint v6.1.3
 * (third-party, pinned at pyscf/lib/CMakeLists.txt:178, NOT vendored in /root/reference):
 *     int2e_sph / int2e_cart      (ij|kl)     called at pyscf/lib/vhf/nr_direct.c:73
 *     int3c2e_sph                 (ij|P)      called via pyscf/lib/gto/fill_nr_3c.c:68,176
 *     int2c2e_sph                 (P|Q)       called via pyscf/lib/gto/fill_int2c.c:36
 *     int1e_ovlp/kin/nuc_sph      (host-side SCF checks only)
 * with libcint's calling convention (out Fortran-ordered, contraction index slowest inside a
 * shell: pyscf/gto/moleintor.py:729-736,772; pyscf/lib/vhf/nr_direct_dot.c:232) and libcint's AO
 * conventions (pyscf/gto/mole.py:122-181,986-1029; pyscf/lib/parameters.py:69-77).
 *
 * The arithmetic is deliberately NOT the Rys quadrature used by libcint and by this repo's CUDA
 * kernels: it is the McMurchie–Davidson scheme (Hermite Gaussians + Boys function, J. Comput.
 * Phys. 26, 218 (1978)), so that agreement between the GPU path and this oracle is an
 * independent check.  Parity is pinned against the reference's own known-answer fingerprints in
 * tests/test_oracle_golden.py.
 *
 * Range separation (env[PTR_RANGE_OMEGA], pyscf/gto/mole.py:80,2940-2951): omega>0 -> erf(w r)/r,
 * omega<0 -> erfc(|w| r)/r, 0 -> 1/r.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ATOM_OF 0
#define ANG_OF 1
#define NPRIM_OF 2
#define NCTR_OF 3
#define PTR_EXP 5
#define PTR_COEFF 6
#define BAS_SLOTS 8
#define PTR_COORD 1
#define ATM_SLOTS 6
#define CHARGE_OF 0
#define PTR_RANGE_OMEGA 8

#define LSH_MAX 6           /* max angular momentum of one shell */
#define LTOT_MAX (4 * LSH_MAX)
#define NCART(l) (((l) + 1) * ((l) + 2) / 2)

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ Boys function */
static void boys(int nmax, double x, double *F)
{
    if (x < 1e-15) {
        for (int n = 0; n <= nmax; n++) F[n] = 1.0 / (2 * n + 1);
        return;
    }
    if (x < 40.0) {
        /* series for the top order, downward recursion (stable) */
        long double ex = expl(-(long double)x);
        long double term = 1.0L / (2 * nmax + 1), sum = term;
        for (int k = 1; k < 400; k++) {
            term *= 2.0L * x / (2 * nmax + 2 * k + 1);
            sum += term;
            if (term < 1e-22L * sum) break;
        }
        long double f = ex * sum;
        F[nmax] = (double)f;
        for (int n = nmax - 1; n >= 0; n--) {
            f = (2.0L * x * f + ex) / (2 * n + 1);
            F[n] = (double)f;
        }
    } else {
        /* erf(sqrt(x)) == 1 to < 1e-17; upward recursion is stable for large x */
        long double ex = expl(-(long double)x);
        long double f = 0.5L * sqrtl((long double)M_PI / x) * erfl(sqrtl((long double)x));
        F[0] = (double)f;
        for (int n = 0; n < nmax; n++) {
            f = ((2 * n + 1) * f - ex) / (2.0L * x);
            F[n + 1] = (double)f;
        }
    }
}

/* ------------------------------------------------------------------ Hermite expansion */
/* E[(i*(lb+1)+j)*(la+lb+1)+t]; one Cartesian direction; includes exp(-mu XAB^2). */
static void hermite_E(int la, int lb, double a, double b, double XAB, double *E)
{
    int nt = la + lb + 1;
    double p = a + b;
    double mu = a * b / p;
    double XPA = -b / p * XAB, XPB = a / p * XAB;
    double hp = 0.5 / p;
    memset(E, 0, sizeof(double) * (la + 1) * (lb + 1) * nt);
#define EE(i, j, t) E[((i) * (lb + 1) + (j)) * nt + (t)]
    EE(0, 0, 0) = exp(-mu * XAB * XAB);
    for (int i = 0; i < la; i++)
        for (int t = 0; t <= i + 1; t++) {
            double v = XPA * EE(i, 0, t);
            if (t > 0) v += hp * EE(i, 0, t - 1);
            if (t + 1 <= i) v += (t + 1) * EE(i, 0, t + 1);
            EE(i + 1, 0, t) = v;
        }
    for (int i = 0; i <= la; i++)
        for (int j = 0; j < lb; j++)
            for (int t = 0; t <= i + j + 1; t++) {
                double v = XPB * EE(i, j, t);
                if (t > 0) v += hp * EE(i, j, t - 1);
                if (t + 1 <= i + j) v += (t + 1) * EE(i, j, t + 1);
                EE(i, j + 1, t) = v;
            }
#undef EE
}

/* Hermite Coulomb integrals R_{tuv} = R^0_{tuv}, t+u+v <= L; out indexed [(t*(L+1)+u)*(L+1)+v] */
static void hermite_R(int L, double alpha, const double *PQ, double prefac, double *R, double *work)
{
    double F[LTOT_MAX + 2];
    double T = alpha * (PQ[0] * PQ[0] + PQ[1] * PQ[1] + PQ[2] * PQ[2]);
    boys(L, T, F);
    int n1 = L + 1;
    /* work[n][t][u][v] */
#define W(n, t, u, v) work[(((n) * n1 + (t)) * n1 + (u)) * n1 + (v)]
    double m2a = 1.0;
    for (int n = 0; n <= L; n++) {
        W(n, 0, 0, 0) = prefac * m2a * F[n];
        m2a *= -2.0 * alpha;
    }
    for (int N = 1; N <= L; N++)          /* total order t+u+v = N */
        for (int t = 0; t <= N; t++)
            for (int u = 0; u <= N - t; u++) {
                int v = N - t - u;
                for (int n = 0; n <= L - N; n++) {
                    double val;
                    if (t > 0) {
                        val = PQ[0] * W(n + 1, t - 1, u, v);
                        if (t > 1) val += (t - 1) * W(n + 1, t - 2, u, v);
                    } else if (u > 0) {
                        val = PQ[1] * W(n + 1, t, u - 1, v);
                        if (u > 1) val += (u - 1) * W(n + 1, t, u - 2, v);
                    } else {
                        val = PQ[2] * W(n + 1, t, u, v - 1);
                        if (v > 1) val += (v - 1) * W(n + 1, t, u, v - 2);
                    }
                    W(n, t, u, v) = val;
                }
            }
    for (int t = 0; t <= L; t++)
        for (int u = 0; u <= L - t; u++)
            for (int v = 0; v <= L - t - u; v++)
                R[(t * n1 + u) * n1 + v] = W(0, t, u, v);
#undef W
}

/* ------------------------------------------------------------------ cart <-> real spherical */
static double binom(int n, int k)
{
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}
static double fact(int n)
{
    double r = 1.0;
    for (int i = 2; i <= n; i++) r *= i;
    return r;
}
static int cart_index(int l, int lx, int ly)
{ /* libcint order: lx descending, then ly descending */
    int idx = 0;
    for (int x = l; x > lx; x--) idx += l - x + 1;
    return idx + (l - lx - ly);
}

/* c2s[m_index * ncart + cart]; rows in libcint's spherical order (p: x,y,z; l>=2: m=-l..l).
 * Real solid harmonics, Helgaker/Jorgensen/Olsen eq. 6.4.47, times sqrt((2l+1)/4pi). */
static void make_c2s(int l, double *c2s)
{
    int nc = NCART(l), ns = 2 * l + 1;
    memset(c2s, 0, sizeof(double) * nc * ns);
    if (l == 0) { c2s[0] = 0.282094791773878143; return; }
    if (l == 1) {
        for (int i = 0; i < 3; i++) c2s[i * 3 + i] = 0.488602511902919921;
        return;
    }
    double ang = sqrt((2 * l + 1) / (4.0 * M_PI));
    for (int m = -l; m <= l; m++) {
        int am = abs(m);
        double N = 1.0 / (pow(2.0, am) * fact(l)) * sqrt(2.0 * fact(l + am) * fact(l - am) / (m == 0 ? 2.0 : 1.0));
        int two_vm = (m < 0) ? 1 : 0; /* v_m = two_vm/2 */
        double *row = c2s + (m + l) * nc;
        for (int t = 0; t <= (l - am) / 2; t++)
            for (int u = 0; u <= t; u++) {
                /* v = v_m, v_m+1, ... , floor(|m|/2 - v_m) + v_m ; work with 2v = two_v */
                int vmax2 = 2 * (int)floor(am / 2.0 - two_vm / 2.0) + two_vm;
                for (int two_v = two_vm; two_v <= vmax2; two_v += 2) {
                    int sgn_pow = t + (two_v - two_vm) / 2;
                    double C = ((sgn_pow & 1) ? -1.0 : 1.0) * pow(0.25, t) * binom(l, t) * binom(l - t, am + t) *
                               binom(t, u) * binom(am, two_v);
                    int lx = 2 * t + am - 2 * u - two_v;
                    int ly = 2 * u + two_v;
                    int lz = l - 2 * t - am;
                    if (lx < 0 || ly < 0 || lz < 0) continue;
                    row[cart_index(l, lx, ly)] += ang * N * C;
                }
            }
    }
}

static double *C2S[LSH_MAX + 1];
static void init_c2s(void)
{
    if (C2S[0]) return;
#pragma omp critical(oracle_c2s)
    {
        if (!C2S[0])
            for (int l = LSH_MAX; l >= 0; l--) {
                double *m = malloc(sizeof(double) * NCART(l) * (2 * l + 1));
                make_c2s(l, m);
                C2S[l] = m;
            }
    }
}
void oracle_c2s(int l, double *out) /* for tests */
{
    init_c2s();
    memcpy(out, C2S[l], sizeof(double) * NCART(l) * (2 * l + 1));
}

/* ------------------------------------------------------------------ shells */
typedef struct {
    int l, nprim, nctr;
    const double *exps;  /* [nprim] */
    const double *coef;  /* [nctr][nprim] */
    const double *r;     /* [3] */
} Shell;

static const double ZERO3[3] = {0, 0, 0};
static const double ONE1[1] = {1.0};
static const double ZEROEXP[1] = {0.0};

static Shell get_shell(int ish, const int *atm, const int *bas, const double *env)
{
    Shell s;
    const int *b = bas + ish * BAS_SLOTS;
    s.l = b[ANG_OF];
    s.nprim = b[NPRIM_OF];
    s.nctr = b[NCTR_OF];
    s.exps = env + b[PTR_EXP];
    s.coef = env + b[PTR_COEFF];
    s.r = env + atm[b[ATOM_OF] * ATM_SLOTS + PTR_COORD];
    return s;
}
static Shell unit_shell(const double *r)
{ /* the function "1" sitting on r: s-type, exponent 0, coefficient 1, no angular factor */
    Shell s = {0, 1, 1, ZEROEXP, ONE1, r};
    return s;
}

static void cart_powers(int l, int *lx, int *ly, int *lz)
{
    int n = 0;
    for (int x = l; x >= 0; x--)
        for (int y = l - x; y >= 0; y--) {
            lx[n] = x; ly[n] = y; lz[n] = l - x - y; n++;
        }
}

/* Primitive Cartesian (ab|cd) block over bare monomials; out[a + na*(b + nb*(c + nc*d))] */
static void prim_eri_cart(int la, int lb, int lc, int ld, double a, double b, double c, double d,
                          const double *A, const double *B, const double *C, const double *D, double omega,
                          double *out, double *scratch)
{
    int na = NCART(la), nb = NCART(lb), nc = NCART(lc), nd = NCART(ld);
    int Lb = la + lb, Lk = lc + ld, L = Lb + Lk;
    double p = a + b, q = c + d;
    double P[3], Q[3], PQ[3];
    for (int x = 0; x < 3; x++) {
        P[x] = (a * A[x] + b * B[x]) / p;
        Q[x] = (c * C[x] + d * D[x]) / q;
        PQ[x] = P[x] - Q[x];
    }
    double alpha = p * q / (p + q);
    double pref = 2.0 * pow(M_PI, 2.5) / (p * q * sqrt(p + q));
    if (omega != 0.0) { /* erf-attenuated: alpha -> alpha*theta, prefactor * sqrt(theta) */
        double theta = omega * omega / (omega * omega + alpha);
        pref *= sqrt(theta);
        alpha *= theta;
    }
    /* scratch carve-up */
    double *Eab[3], *Ecd[3];
    int sab = (la + 1) * (lb + 1) * (Lb + 1), scd = (lc + 1) * (ld + 1) * (Lk + 1);
    double *w = scratch;
    for (int x = 0; x < 3; x++) { Eab[x] = w; w += sab; }
    for (int x = 0; x < 3; x++) { Ecd[x] = w; w += scd; }
    int n1 = L + 1;
    double *R = w; w += n1 * n1 * n1;
    double *Rwork = w; w += n1 * n1 * n1 * n1;
    int nb1 = Lb + 1;
    double *Y = w; w += nb1 * nb1 * nb1; /* per ket component pair */

    for (int x = 0; x < 3; x++) {
        hermite_E(la, lb, a, b, A[x] - B[x], Eab[x]);
        hermite_E(lc, ld, c, d, C[x] - D[x], Ecd[x]);
    }
    hermite_R(L, alpha, PQ, pref, R, Rwork);

    int ax[28], ay[28], az[28], bx[28], by[28], bz[28], cx[28], cy[28], cz[28], dx[28], dy[28], dz[28];
    cart_powers(la, ax, ay, az); cart_powers(lb, bx, by, bz);
    cart_powers(lc, cx, cy, cz); cart_powers(ld, dx, dy, dz);
#define EAB(x, i, j, t) Eab[x][((i) * (lb + 1) + (j)) * (Lb + 1) + (t)]
#define ECD(x, i, j, t) Ecd[x][((i) * (ld + 1) + (j)) * (Lk + 1) + (t)]
    for (int id = 0; id < nd; id++)
        for (int ic = 0; ic < nc; ic++) {
            int mx = cx[ic] + dx[id], my = cy[ic] + dy[id], mz = cz[ic] + dz[id];
            /* Y[T][U][V] = sum_{tau,nu,phi} (-1)^(tau+nu+phi) Ecd R[T+tau][U+nu][V+phi] */
            for (int T = 0; T <= Lb; T++)
                for (int U = 0; U <= Lb - T; U++)
                    for (int V = 0; V <= Lb - T - U; V++) {
                        double s = 0.0;
                        for (int tau = 0; tau <= mx; tau++) {
                            double ex = ECD(0, cx[ic], dx[id], tau);
                            for (int nu = 0; nu <= my; nu++) {
                                double exy = ex * ECD(1, cy[ic], dy[id], nu);
                                for (int phi = 0; phi <= mz; phi++) {
                                    double e = exy * ECD(2, cz[ic], dz[id], phi);
                                    if ((tau + nu + phi) & 1) e = -e;
                                    s += e * R[((T + tau) * n1 + (U + nu)) * n1 + (V + phi)];
                                }
                            }
                        }
                        Y[(T * nb1 + U) * nb1 + V] = s;
                    }
            for (int ib = 0; ib < nb; ib++)
                for (int ia = 0; ia < na; ia++) {
                    int kx = ax[ia] + bx[ib], ky = ay[ia] + by[ib], kz = az[ia] + bz[ib];
                    double s = 0.0;
                    for (int t = 0; t <= kx; t++) {
                        double ex = EAB(0, ax[ia], bx[ib], t);
                        for (int u = 0; u <= ky; u++) {
                            double exy = ex * EAB(1, ay[ia], by[ib], u);
                            for (int v = 0; v <= kz; v++)
                                s += exy * EAB(2, az[ia], bz[ib], v) * Y[(t * nb1 + u) * nb1 + v];
                        }
                    }
                    out[ia + na * (ib + nb * (ic + nc * id))] = s;
                }
        }
#undef EAB
#undef ECD
}

static size_t prim_scratch_size(int la, int lb, int lc, int ld)
{
    int Lb = la + lb, Lk = lc + ld, L = Lb + Lk, n1 = L + 1;
    return 3 * (la + 1) * (lb + 1) * (Lb + 1) + 3 * (lc + 1) * (ld + 1) * (Lk + 1) + n1 * n1 * n1 + n1 * n1 * n1 * n1 +
           (Lb + 1) * (Lb + 1) * (Lb + 1) + 64;
}

/* transform one index of a 4-index Fortran-ordered block: in[n0,nc,n1] -> out[n0,ns,n1] */
static void c2s_index(const double *in, double *out, int n0, int l, int n1, int sph, int is_unit)
{
    int nc = NCART(l);
    if (is_unit) { memcpy(out, in, sizeof(double) * n0 * n1); return; }
    if (!sph) { /* Cartesian output: libcint normalises s and p only */
        double f = (l == 0) ? 0.282094791773878143 : (l == 1 ? 0.488602511902919921 : 1.0);
        for (long i = 0; i < (long)n0 * nc * n1; i++) out[i] = in[i] * f;
        return;
    }
    int ns = 2 * l + 1;
    const double *T = C2S[l];
    for (int k = 0; k < n1; k++)
        for (int m = 0; m < ns; m++) {
            double *o = out + ((long)k * ns + m) * n0;
            for (int i = 0; i < n0; i++) o[i] = 0.0;
            for (int c = 0; c < nc; c++) {
                double t = T[m * nc + c];
                if (t == 0.0) continue;
                const double *s = in + ((long)k * nc + c) * n0;
                for (int i = 0; i < n0; i++) o[i] += t * s[i];
            }
        }
}

/* Contracted shell-quartet block.  unit[x] != 0 marks a "1" shell (3c/2c integrals).
 * out: Fortran order [di,dj,dk,dl], d = nctr * nfunc, contraction slowest inside the shell. */
static int eri_block(const Shell *sh, const int *unit, double omega, int sph, double *out)
{
    init_c2s();
    int l[4], ncart[4], nf[4], nctr[4], dim[4];
    for (int x = 0; x < 4; x++) {
        l[x] = sh[x].l;
        ncart[x] = NCART(l[x]);
        nf[x] = unit[x] ? 1 : (sph ? 2 * l[x] + 1 : ncart[x]);
        nctr[x] = sh[x].nctr;
        dim[x] = nf[x] * nctr[x];
    }
    long nprimblk = (long)ncart[0] * ncart[1] * ncart[2] * ncart[3];
    long nctrblk = (long)nctr[0] * nctr[1] * nctr[2] * nctr[3];
    double *prim = malloc(sizeof(double) * nprimblk * 3);
    double *prim2 = prim + nprimblk;
    double *tmp = prim2 + nprimblk;
    double *ctr = calloc(nprimblk * nctrblk, sizeof(double)); /* [cblk][cart block] */
    double *scratch = malloc(sizeof(double) * prim_scratch_size(l[0], l[1], l[2], l[3]));
    double lr_sign = 1.0;
    double om = omega;
    for (int ip = 0; ip < sh[0].nprim; ip++)
        for (int jp = 0; jp < sh[1].nprim; jp++)
            for (int kp = 0; kp < sh[2].nprim; kp++)
                for (int lp = 0; lp < sh[3].nprim; lp++) {
                    if (omega >= 0.0) {
                        prim_eri_cart(l[0], l[1], l[2], l[3], sh[0].exps[ip], sh[1].exps[jp], sh[2].exps[kp],
                                      sh[3].exps[lp], sh[0].r, sh[1].r, sh[2].r, sh[3].r, om, prim, scratch);
                    } else { /* erfc = full - erf */
                        prim_eri_cart(l[0], l[1], l[2], l[3], sh[0].exps[ip], sh[1].exps[jp], sh[2].exps[kp],
                                      sh[3].exps[lp], sh[0].r, sh[1].r, sh[2].r, sh[3].r, 0.0, prim, scratch);
                        prim_eri_cart(l[0], l[1], l[2], l[3], sh[0].exps[ip], sh[1].exps[jp], sh[2].exps[kp],
                                      sh[3].exps[lp], sh[0].r, sh[1].r, sh[2].r, sh[3].r, -omega, prim2, scratch);
                        for (long n = 0; n < nprimblk; n++) prim[n] -= prim2[n];
                    }
                    for (int cl = 0; cl < nctr[3]; cl++)
                        for (int ck = 0; ck < nctr[2]; ck++)
                            for (int cj = 0; cj < nctr[1]; cj++)
                                for (int ci = 0; ci < nctr[0]; ci++) {
                                    double cc = lr_sign * sh[0].coef[ci * sh[0].nprim + ip] *
                                                sh[1].coef[cj * sh[1].nprim + jp] * sh[2].coef[ck * sh[2].nprim + kp] *
                                                sh[3].coef[cl * sh[3].nprim + lp];
                                    if (cc == 0.0) continue;
                                    double *dst = ctr + (((long)(cl * nctr[2] + ck) * nctr[1] + cj) * nctr[0] + ci) * nprimblk;
                                    for (long n = 0; n < nprimblk; n++) dst[n] += cc * prim[n];
                                }
                }
    /* transform each contraction block and scatter into out */
    int nonzero = 0;
    double *b0 = malloc(sizeof(double) * nprimblk * 2);
    double *b1 = b0 + nprimblk;
    for (int cl = 0; cl < nctr[3]; cl++)
        for (int ck = 0; ck < nctr[2]; ck++)
            for (int cj = 0; cj < nctr[1]; cj++)
                for (int ci = 0; ci < nctr[0]; ci++) {
                    const double *src = ctr + (((long)(cl * nctr[2] + ck) * nctr[1] + cj) * nctr[0] + ci) * nprimblk;
                    /* index 0 */
                    c2s_index(src, b0, 1, l[0], ncart[1] * ncart[2] * ncart[3], sph, unit[0]);
                    c2s_index(b0, b1, nf[0], l[1], ncart[2] * ncart[3], sph, unit[1]);
                    c2s_index(b1, b0, nf[0] * nf[1], l[2], ncart[3], sph, unit[2]);
                    c2s_index(b0, b1, nf[0] * nf[1] * nf[2], l[3], 1, sph, unit[3]);
                    for (int d = 0; d < nf[3]; d++)
                        for (int c = 0; c < nf[2]; c++)
                            for (int b = 0; b < nf[1]; b++)
                                for (int a = 0; a < nf[0]; a++) {
                                    double v = b1[a + nf[0] * (b + nf[1] * (c + nf[2] * d))];
                                    if (v != 0.0) nonzero = 1;
                                    out[(ci * nf[0] + a) +
                                        (long)dim[0] * ((cj * nf[1] + b) +
                                                        (long)dim[1] * ((ck * nf[2] + c) + (long)dim[2] * (cl * nf[3] + d)))] = v;
                                }
                }
    (void)tmp;
    free(b0); free(scratch); free(ctr); free(prim);
    return nonzero;
}

/* ------------------------------------------------------------------ libcint-signature entry points */
/* dims (if non-NULL) gives the leading dimensions of `out` (libcint convention). */
static int run_block(double *out, const int *dims, const Shell *sh, const int *unit, double omega, int sph, int ncenter)
{
    int d[4];
    for (int x = 0; x < 4; x++) {
        int nf = unit[x] ? 1 : (sph ? 2 * sh[x].l + 1 : NCART(sh[x].l));
        d[x] = nf * sh[x].nctr;
    }
    if (out == NULL) return 0;
    if (dims == NULL) return eri_block(sh, unit, omega, sph, out);
    /* strided output: compute compactly then copy. real indices = those with unit==0 */
    long n = (long)d[0] * d[1] * d[2] * d[3];
    double *buf = malloc(sizeof(double) * n);
    int nz = eri_block(sh, unit, omega, sph, buf);
    int real_idx[4], nr = 0;
    for (int x = 0; x < 4; x++) if (!unit[x]) real_idx[nr++] = x;
    (void)ncenter;
    int e[4] = {1, 1, 1, 1}, ld[4] = {1, 1, 1, 1};
    for (int x = 0; x < nr; x++) { e[x] = d[real_idx[x]]; ld[x] = dims[x]; }
    for (int i3 = 0; i3 < e[3]; i3++)
        for (int i2 = 0; i2 < e[2]; i2++)
            for (int i1 = 0; i1 < e[1]; i1++)
                for (int i0 = 0; i0 < e[0]; i0++)
                    out[i0 + (long)ld[0] * (i1 + (long)ld[1] * (i2 + (long)ld[2] * i3))] =
                        buf[i0 + (long)e[0] * (i1 + (long)e[1] * (i2 + (long)e[2] * i3))];
    free(buf);
    return nz;
}

#define DEF_INT2E(name, sph)                                                                                          \
    int name(double *out, int *dims, int *shls, int *atm, int natm, int *bas, int nbas, double *env, void *opt,        \
             double *cache)                                                                                            \
    {                                                                                                                  \
        Shell sh[4];                                                                                                   \
        int unit[4] = {0, 0, 0, 0};                                                                                    \
        for (int x = 0; x < 4; x++) sh[x] = get_shell(shls[x], atm, bas, env);                                         \
        if (out == NULL) return 0; /* cache-size query (pyscf/lib/gto/fill_int2e.c:42-64) */                           \
        return run_block(out, dims, sh, unit, env[PTR_RANGE_OMEGA], sph, 4);                                           \
    }
DEF_INT2E(int2e_sph, 1)
DEF_INT2E(int2e_cart, 0)

#define DEF_INT3C2E(name, sph)                                                                                        \
    int name(double *out, int *dims, int *shls, int *atm, int natm, int *bas, int nbas, double *env, void *opt,        \
             double *cache)                                                                                            \
    {                                                                                                                  \
        Shell sh[4];                                                                                                   \
        int unit[4] = {0, 0, 0, 1};                                                                                    \
        for (int x = 0; x < 3; x++) sh[x] = get_shell(shls[x], atm, bas, env);                                         \
        sh[3] = unit_shell(sh[2].r);                                                                                   \
        if (out == NULL) return 0;                                                                                     \
        return run_block(out, dims, sh, unit, env[PTR_RANGE_OMEGA], sph, 3);                                           \
    }
DEF_INT3C2E(int3c2e_sph, 1)
DEF_INT3C2E(int3c2e_cart, 0)

#define DEF_INT2C2E(name, sph)                                                                                        \
    int name(double *out, int *dims, int *shls, int *atm, int natm, int *bas, int nbas, double *env, void *opt,        \
             double *cache)                                                                                            \
    {                                                                                                                  \
        Shell sh[4];                                                                                                   \
        int unit[4] = {0, 1, 0, 1};                                                                                    \
        sh[0] = get_shell(shls[0], atm, bas, env);                                                                     \
        sh[2] = get_shell(shls[1], atm, bas, env);                                                                     \
        sh[1] = unit_shell(sh[0].r);                                                                                   \
        sh[3] = unit_shell(sh[2].r);                                                                                   \
        if (out == NULL) return 0;                                                                                     \
        return run_block(out, dims, sh, unit, env[PTR_RANGE_OMEGA], sph, 2);                                           \
    }
DEF_INT2C2E(int2c2e_sph, 1)
DEF_INT2C2E(int2c2e_cart, 0)

/* ------------------------------------------------------------------ one-electron integrals (SCF checks) */
/* kind: 0 overlap, 1 kinetic, 2 nuclear attraction; spherical or cart; out [di,dj] Fortran order */
static void prim_1e_cart(int kind, int la, int lb, double a, double b, const double *A, const double *B,
                         const int *atm, int natm, const double *env, double *out, double *scratch)
{
    int na = NCART(la), nb = NCART(lb);
    double p = a + b;
    int lb2 = lb + 2;
    double *E[3];
    int se = (la + 1) * (lb2 + 1) * (la + lb2 + 1);
    for (int x = 0; x < 3; x++) { E[x] = scratch + x * se; hermite_E(la, lb2, a, b, A[x] - B[x], E[x]); }
#define E1(x, i, j, t) E[x][((i) * (lb2 + 1) + (j)) * (la + lb2 + 1) + (t)]
    int ax[28], ay[28], az[28], bx[28], by[28], bz[28];
    cart_powers(la, ax, ay, az); cart_powers(lb, bx, by, bz);
    double s3 = pow(M_PI / p, 1.5);
    if (kind == 0 || kind == 1) {
        for (int ib = 0; ib < nb; ib++)
            for (int ia = 0; ia < na; ia++) {
                int i[3] = {ax[ia], ay[ia], az[ia]}, j[3] = {bx[ib], by[ib], bz[ib]};
                double S[3], T[3];
                for (int x = 0; x < 3; x++) {
                    S[x] = E1(x, i[x], j[x], 0);
                    double t = 4 * b * b * E1(x, i[x], j[x] + 2, 0) - 2 * b * (2 * j[x] + 1) * S[x];
                    if (j[x] >= 2) t += j[x] * (j[x] - 1) * E1(x, i[x], j[x] - 2, 0);
                    T[x] = -0.5 * t;
                }
                double v = (kind == 0) ? S[0] * S[1] * S[2]
                                       : (T[0] * S[1] * S[2] + S[0] * T[1] * S[2] + S[0] * S[1] * T[2]);
                out[ia + na * ib] = v * s3;
            }
    } else {
        int L = la + lb, n1 = L + 1;
        double *R = scratch + 3 * se, *Rw = R + n1 * n1 * n1;
        double P[3];
        for (int x = 0; x < 3; x++) P[x] = (a * A[x] + b * B[x]) / p;
        for (int n = 0; n < na * nb; n++) out[n] = 0.0;
        for (int ic = 0; ic < natm; ic++) {
            const double *C = env + atm[ic * ATM_SLOTS + PTR_COORD];
            double Z = atm[ic * ATM_SLOTS + CHARGE_OF];
            double PC[3] = {P[0] - C[0], P[1] - C[1], P[2] - C[2]};
            hermite_R(L, p, PC, -Z * 2.0 * M_PI / p, R, Rw);
            for (int ib = 0; ib < nb; ib++)
                for (int ia = 0; ia < na; ia++) {
                    double s = 0.0;
                    for (int t = 0; t <= ax[ia] + bx[ib]; t++)
                        for (int u = 0; u <= ay[ia] + by[ib]; u++)
                            for (int v = 0; v <= az[ia] + bz[ib]; v++)
                                s += E1(0, ax[ia], bx[ib], t) * E1(1, ay[ia], by[ib], u) * E1(2, az[ia], bz[ib], v) *
                                     R[(t * n1 + u) * n1 + v];
                    out[ia + na * ib] += s;
                }
        }
    }
#undef E1
}

/* full matrix of a 1e operator, spherical, row-major [nao,nao] */
void oracle_int1e(int kind, double *mat, int nao, const int *ao_loc, int *atm, int natm, int *bas, int nbas, double *env)
{
    init_c2s();
    for (int ish = 0; ish < nbas; ish++)
        for (int jsh = 0; jsh < nbas; jsh++) {
            Shell si = get_shell(ish, atm, bas, env), sj = get_shell(jsh, atm, bas, env);
            int na = NCART(si.l), nb = NCART(sj.l), nsa = 2 * si.l + 1, nsb = 2 * sj.l + 1;
            int L = si.l + sj.l + 2 + 1;
            double *scratch = malloc(sizeof(double) * (3 * (si.l + 1) * (sj.l + 3) * (si.l + sj.l + 3) + L * L * L + L * L * L * L + 64));
            double *prim = malloc(sizeof(double) * na * nb * 3);
            double *b0 = prim + na * nb, *b1 = b0 + na * nb;
            for (int ci = 0; ci < si.nctr; ci++)
                for (int cj = 0; cj < sj.nctr; cj++) {
                    for (int n = 0; n < na * nb; n++) b0[n] = 0.0;
                    for (int ip = 0; ip < si.nprim; ip++)
                        for (int jp = 0; jp < sj.nprim; jp++) {
                            prim_1e_cart(kind, si.l, sj.l, si.exps[ip], sj.exps[jp], si.r, sj.r, atm, natm, env, prim, scratch);
                            double cc = si.coef[ci * si.nprim + ip] * sj.coef[cj * sj.nprim + jp];
                            for (int n = 0; n < na * nb; n++) b0[n] += cc * prim[n];
                        }
                    c2s_index(b0, b1, 1, si.l, nb, 1, 0);
                    c2s_index(b1, b0, nsa, sj.l, 1, 1, 0);
                    for (int b = 0; b < nsb; b++)
                        for (int a = 0; a < nsa; a++)
                            mat[(long)(ao_loc[ish] + ci * nsa + a) * nao + ao_loc[jsh] + cj * nsb + b] = b0[a + nsa * b];
                }
            free(prim); free(scratch);
        }
}
