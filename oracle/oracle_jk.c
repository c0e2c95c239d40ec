/*
 * oracle_jk.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * Restatement of the reference's J/K drivers on top of oracle_cint.c:
 *   oracle_q_cond      <- CVHFnr_int2e_q_cond      pyscf/lib/vhf/optimizer.c:408-454
 *   oracle_dm_cond     <- CVHFnr_dm_cond           pyscf/lib/vhf/optimizer.c:494-518
 *   prescreen          <- CVHFnrs8_prescreen       pyscf/lib/vhf/optimizer.c:90-117
 *   oracle_direct_jk   <- CVHFnr_direct_drv + CVHFdot_nrs8 + nrs8_ji_s2kl / nrs8_li_s1kj
 *                         pyscf/lib/vhf/nr_direct.c:183-231,361-489; nr_direct_dot.c:1293-1433
 *   oracle_fill_*      <- GTOnr2e_fill_drv / GTOnr3c_drv / GTOint2c
 *                         pyscf/lib/gto/fill_int2e.c:538, fill_nr_3c.c:196, fill_int2c.c:36
 * Definitions: J_kl = sum_ij (ij|kl) D_ji, K_il = sum_jk (ij|kl) D_jk  (pyscf/scf/hf.py:906-907).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int int2e_sph(double *, int *, int *, int *, int, int *, int, double *, void *, double *);
int int2e_cart(double *, int *, int *, int *, int, int *, int, double *, void *, double *);
int int3c2e_sph(double *, int *, int *, int *, int, int *, int, double *, void *, double *);
int int2c2e_sph(double *, int *, int *, int *, int, int *, int, double *, void *, double *);

#define MAXBLK 4096

/* q_cond[i*nbas+j] = sqrt(max |(ij|ij)|) over the shell block, floor 1e-100 */
void oracle_q_cond(double *q, const int *ao_loc, int *atm, int natm, int *bas, int nbas, double *env)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (int ij = 0; ij < nbas * (nbas + 1) / 2; ij++) {
        int i = (int)(sqrt(2 * ij + 0.25) - 0.5 + 1e-7);
        int j = ij - i * (i + 1) / 2;
        int di = ao_loc[i + 1] - ao_loc[i], dj = ao_loc[j + 1] - ao_loc[j];
        double *buf = malloc(sizeof(double) * di * dj * di * dj);
        int shls[4] = {i, j, i, j};
        double qmax = 1e-100;
        if (int2e_sph(buf, NULL, shls, atm, natm, bas, nbas, env, NULL, NULL)) {
            for (int b = 0; b < dj; b++)
                for (int a = 0; a < di; a++) {
                    double v = fabs(buf[a + di * b + (long)di * dj * a + (long)di * dj * di * b]);
                    if (v > qmax * qmax) qmax = sqrt(v);
                }
        }
        q[i * nbas + j] = q[j * nbas + i] = qmax;
        free(buf);
    }
}

void oracle_dm_cond(double *dmc, const double *dm, int n_dm, int nao, const int *ao_loc, int nbas)
{
    for (int i = 0; i < nbas; i++)
        for (int j = 0; j <= i; j++) {
            double m = 0;
            for (int s = 0; s < n_dm; s++)
                for (int a = ao_loc[i]; a < ao_loc[i + 1]; a++)
                    for (int b = ao_loc[j]; b < ao_loc[j + 1]; b++) {
                        double t = .5 * (fabs(dm[(long)s * nao * nao + (long)a * nao + b]) +
                                         fabs(dm[(long)s * nao * nao + (long)b * nao + a]));
                        if (t > m) m = t;
                    }
            dmc[i * nbas + j] = dmc[j * nbas + i] = m;
        }
}

static inline int prescreen(int i, int j, int k, int l, const double *q, const double *d, int n, double tol)
{
    double qijkl = q[i * n + j] * q[k * n + l];
    if (!(qijkl > tol)) return 0;
    double dmin = tol / qijkl;
    return (4 * d[j * n + i] > dmin) || (4 * d[l * n + k] > dmin) || (d[j * n + k] > dmin) || (d[j * n + l] > dmin) ||
           (d[i * n + k] > dmin) || (d[i * n + l] > dmin);
}

/* vj, vk: [n_dm, nao, nao] C-order, fully filled (no hermi_triu step needed). dm arbitrary real.
 * q_cond/dm_cond may be NULL => no screening.  Returns number of shell quartets computed. */
long oracle_direct_jk(double *vj, double *vk, const double *dm, int n_dm, int nao, const int *ao_loc, int *atm,
                      int natm, int *bas, int nbas, double *env, const double *q_cond, const double *dm_cond,
                      double tol)
{
    long nn = (long)nao * nao;
    memset(vj, 0, sizeof(double) * n_dm * nn);
    memset(vk, 0, sizeof(double) * n_dm * nn);
    long ncomputed = 0;
#pragma omp parallel reduction(+ : ncomputed)
    {
        double *pj = calloc(n_dm * nn, sizeof(double));
        double *pk = calloc(n_dm * nn, sizeof(double));
        double *buf = malloc(sizeof(double) * MAXBLK * 16);
#pragma omp for schedule(dynamic, 1)
        for (int ij = nbas * (nbas + 1) / 2 - 1; ij >= 0; ij--) {
            int i = (int)(sqrt(2 * ij + 0.25) - 0.5 + 1e-7);
            int j = ij - i * (i + 1) / 2;
            for (int kl = 0; kl <= ij; kl++) {
                int k = (int)(sqrt(2 * kl + 0.25) - 0.5 + 1e-7);
                int l = kl - k * (k + 1) / 2;
                if (q_cond && !prescreen(i, j, k, l, q_cond, dm_cond, nbas, tol)) continue;
                int shls[4] = {i, j, k, l};
                int i0 = ao_loc[i], j0 = ao_loc[j], k0 = ao_loc[k], l0 = ao_loc[l];
                int di = ao_loc[i + 1] - i0, dj = ao_loc[j + 1] - j0, dk = ao_loc[k + 1] - k0, dl = ao_loc[l + 1] - l0;
                double *b = buf;
                if ((long)di * dj * dk * dl > MAXBLK * 16) b = malloc(sizeof(double) * di * dj * dk * dl);
                ncomputed++;
                if (int2e_sph(b, NULL, shls, atm, natm, bas, nbas, env, NULL, NULL)) {
                    int sw_ij = (i != j), sw_kl = (k != l), sw_bk = (ij != kl);
                    for (int s = 0; s < n_dm; s++) {
                        const double *D = dm + s * nn;
                        double *J = pj + s * nn, *K = pk + s * nn;
                        for (int dd = 0; dd < dl; dd++)
                            for (int c = 0; c < dk; c++)
                                for (int bb = 0; bb < dj; bb++)
                                    for (int a = 0; a < di; a++) {
                                        double v = b[a + di * (bb + dj * (c + (long)dk * dd))];
                                        long p = i0 + a, q = j0 + bb, r = k0 + c, t = l0 + dd;
                                        /* tuple (x0,x1,x2,x3): J[x2,x3] += v D[x1,x0]; K[x0,x3] += v D[x1,x2] */
#define DIGEST(x0, x1, x2, x3)                                  \
    J[(x2) * nao + (x3)] += v * D[(x1) * nao + (x0)];           \
    K[(x0) * nao + (x3)] += v * D[(x1) * nao + (x2)];
                                        DIGEST(p, q, r, t)
                                        if (sw_ij) { DIGEST(q, p, r, t) }
                                        if (sw_kl) { DIGEST(p, q, t, r) }
                                        if (sw_ij && sw_kl) { DIGEST(q, p, t, r) }
                                        if (sw_bk) {
                                            DIGEST(r, t, p, q)
                                            if (sw_kl) { DIGEST(t, r, p, q) }
                                            if (sw_ij) { DIGEST(r, t, q, p) }
                                            if (sw_ij && sw_kl) { DIGEST(t, r, q, p) }
                                        }
#undef DIGEST
                                    }
                    }
                }
                if (b != buf) free(b);
            }
        }
#pragma omp critical
        {
            for (long n = 0; n < n_dm * nn; n++) { vj[n] += pj[n]; vk[n] += pk[n]; }
        }
        free(pj); free(pk); free(buf);
    }
    return ncomputed;
}

/* full (ij|kl) tensor, C-order [nao,nao,nao,nao] (s1), sph or cart */
void oracle_fill_int2e(double *eri, int nao, const int *ao_loc, int cart, int *atm, int natm, int *bas, int nbas,
                       double *env)
{
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int i = 0; i < nbas; i++)
        for (int j = 0; j < nbas; j++) {
            if (j > i) continue;
            for (int k = 0; k < nbas; k++)
                for (int l = 0; l <= k; l++) {
                    if (k * (k + 1) / 2 + l > i * (i + 1) / 2 + j) continue;
                    int shls[4] = {i, j, k, l};
                    int i0 = ao_loc[i], j0 = ao_loc[j], k0 = ao_loc[k], l0 = ao_loc[l];
                    int di = ao_loc[i + 1] - i0, dj = ao_loc[j + 1] - j0, dk = ao_loc[k + 1] - k0, dl = ao_loc[l + 1] - l0;
                    double *b = malloc(sizeof(double) * di * dj * dk * dl);
                    if (cart) int2e_cart(b, NULL, shls, atm, natm, bas, nbas, env, NULL, NULL);
                    else int2e_sph(b, NULL, shls, atm, natm, bas, nbas, env, NULL, NULL);
                    for (int dd = 0; dd < dl; dd++)
                        for (int c = 0; c < dk; c++)
                            for (int bb = 0; bb < dj; bb++)
                                for (int a = 0; a < di; a++) {
                                    double v = b[a + di * (bb + dj * (c + (long)dk * dd))];
                                    long p = i0 + a, q = j0 + bb, r = k0 + c, t = l0 + dd, n = nao;
                                    eri[((p * n + q) * n + r) * n + t] = v;
                                    eri[((q * n + p) * n + r) * n + t] = v;
                                    eri[((p * n + q) * n + t) * n + r] = v;
                                    eri[((q * n + p) * n + t) * n + r] = v;
                                    eri[((r * n + t) * n + p) * n + q] = v;
                                    eri[((t * n + r) * n + p) * n + q] = v;
                                    eri[((r * n + t) * n + q) * n + p] = v;
                                    eri[((t * n + r) * n + q) * n + p] = v;
                                }
                    free(b);
                }
        }
}

/* (ij|P): out C-order [nao, nao, naux]; atm/bas/env is the concatenation mol+auxmol
 * (pyscf/gto/mole.py:805 conc_env); AO shells [0,nbas_ao), aux shells [nbas_ao, nbas_ao+nbas_aux). */
void oracle_fill_int3c2e(double *out, int nao, int naux, const int *ao_loc, const int *aux_loc, int nbas_ao,
                         int nbas_aux, int *atm, int natm, int *bas, int nbas, double *env)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < nbas_ao; i++)
        for (int j = 0; j <= i; j++)
            for (int k = 0; k < nbas_aux; k++) {
                int shls[3] = {i, j, nbas_ao + k};
                int i0 = ao_loc[i], j0 = ao_loc[j], k0 = aux_loc[k];
                int di = ao_loc[i + 1] - i0, dj = ao_loc[j + 1] - j0, dk = aux_loc[k + 1] - k0;
                double *b = malloc(sizeof(double) * di * dj * dk);
                int3c2e_sph(b, NULL, shls, atm, natm, bas, nbas, env, NULL, NULL);
                for (int c = 0; c < dk; c++)
                    for (int bb = 0; bb < dj; bb++)
                        for (int a = 0; a < di; a++) {
                            double v = b[a + di * (bb + (long)dj * c)];
                            out[((long)(i0 + a) * nao + (j0 + bb)) * naux + k0 + c] = v;
                            out[((long)(j0 + bb) * nao + (i0 + a)) * naux + k0 + c] = v;
                        }
                free(b);
            }
}

/* (ij|P) for a LIST of AO shell pairs (tools/make_golden_df_size.py: sampled columns and shell slabs of the 3-center
 * tensor at configuration size): out C-order [naux][ncol]; pair p = (pairs[2p], pairs[2p+1]) owns the columns
 * col0[p] + a * dj + b, a < di, b < dj.  Same table conventions as oracle_fill_int3c2e. */
void oracle_fill_int3c2e_pairs(double *out, long ncol, int npair, const int *pairs, const long *col0, int naux,
                               const int *ao_loc, const int *aux_loc, int nbas_ao, int nbas_aux, int *atm, int natm, int *bas,
                               int nbas, double *env)
{
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int p = 0; p < npair; p++)
        for (int k = 0; k < nbas_aux; k++) {
            int i = pairs[2 * p], j = pairs[2 * p + 1];
            int shls[3] = {i, j, nbas_ao + k};
            int di = ao_loc[i + 1] - ao_loc[i], dj = ao_loc[j + 1] - ao_loc[j];
            int k0 = aux_loc[k], dk = aux_loc[k + 1] - k0;
            double *b = malloc(sizeof(double) * di * dj * dk);
            int3c2e_sph(b, NULL, shls, atm, natm, bas, nbas, env, NULL, NULL);
            for (int c = 0; c < dk; c++)
                for (int bb = 0; bb < dj; bb++)
                    for (int a = 0; a < di; a++)
                        out[(long)(k0 + c) * ncol + col0[p] + (long)a * dj + bb] = b[a + di * (bb + (long)dj * c)];
            free(b);
        }
}

/* (P|Q): out C-order [n,n] over shells [sh0, sh1) of the given tables */
void oracle_fill_int2c2e(double *out, int n, const int *loc, int sh0, int sh1, int *atm, int natm, int *bas, int nbas,
                         double *env)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = sh0; i < sh1; i++)
        for (int j = sh0; j <= i; j++) {
            int shls[2] = {i, j};
            int i0 = loc[i - sh0], j0 = loc[j - sh0];
            int di = loc[i - sh0 + 1] - i0, dj = loc[j - sh0 + 1] - j0;
            double *b = malloc(sizeof(double) * di * dj);
            int2c2e_sph(b, NULL, shls, atm, natm, bas, nbas, env, NULL, NULL);
            for (int bb = 0; bb < dj; bb++)
                for (int a = 0; a < di; a++) {
                    out[(long)(i0 + a) * n + j0 + bb] = b[a + di * bb];
                    out[(long)(j0 + bb) * n + i0 + a] = b[a + di * bb];
                }
            free(b);
        }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Timing / sampling wrapper around int2e_sph for bench.py's CPU arm (TEST INFRASTRUCTURE ONLY).
 * The reference driver (CVHFnr_direct_drv, pyscf/lib/vhf/nr_direct.c:361-489) receives the integral function as a
 * pointer (nr_direct.c:73); handing it this wrapper instead of int2e_sph
 *   - accumulates, per OpenMP thread, the seconds spent inside the integral function (so the arm can report how much
 *     of its time is the oracle's McMurchie-Davidson integrals and how much is reference driver/digestion code), and
 *   - with stride m > 1 evaluates only every m-th surviving shell quartet of each thread (the others return 0 = "block
 *     vanishes", which the driver skips like a libcint zero, nr_direct.c:73-76): a bounded sample of the same workload
 *     whose time, multiplied by m, estimates the full build.
 */
#define ORACLE_MAXTHREADS 1024
static double g_t_intor[ORACLE_MAXTHREADS * 8];
static long g_n_calls[ORACLE_MAXTHREADS * 8], g_n_eval[ORACLE_MAXTHREADS * 8];
static int g_stride = 1;

void oracle_sample_reset(int stride)
{
    g_stride = stride < 1 ? 1 : stride;
    memset(g_t_intor, 0, sizeof g_t_intor);
    memset(g_n_calls, 0, sizeof g_n_calls);
    memset(g_n_eval, 0, sizeof g_n_eval);
}

/* out[0] = summed thread-seconds inside the integral function, out[1] = calls, out[2] = calls evaluated */
void oracle_sample_stats(double *out)
{
    double t = 0, c = 0, e = 0;
    for (int i = 0; i < ORACLE_MAXTHREADS; i++) { t += g_t_intor[8 * i]; c += g_n_calls[8 * i]; e += g_n_eval[8 * i]; }
    out[0] = t; out[1] = c; out[2] = e;
}

int oracle_omp_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_omp_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
}

int int2e_sph_sampled(double *out, int *dims, int *shls, int *atm, int natm, int *bas, int nbas, double *env, void *opt,
                      double *cache)
{
    if (out == NULL) return int2e_sph(out, dims, shls, atm, natm, bas, nbas, env, opt, cache);
#ifdef _OPENMP
    int t = omp_get_thread_num() % ORACLE_MAXTHREADS;
    double t0 = omp_get_wtime();
#else
    int t = 0;
    double t0 = 0;
#endif
    long n = g_n_calls[8 * t]++;
    if (g_stride > 1 && (n % g_stride) != 0) return 0;
    int r = int2e_sph(out, dims, shls, atm, natm, bas, nbas, env, opt, cache);
    g_n_eval[8 * t]++;
#ifdef _OPENMP
    g_t_intor[8 * t] += omp_get_wtime() - t0;
#endif
    return r;
}
