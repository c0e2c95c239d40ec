// i8gemm.cu — host side of the tcgen05 int8-slice GEMM (see i8gemm.cuh) + a C-ABI self-test entry.
#include "host_common.hpp"
#ifndef B200JK_EMULATE
#include <cudaTypedefs.h>
#include "i8gemm.cuh"
#include "i8gemm_host.hpp"

namespace b200jk {
namespace i8g {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult qres;
        void* p = nullptr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled not available");
        fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

static void make_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t kp, uint32_t box_rows)
{
    cuuint64_t dims[2] = {kp, rows};
    cuuint64_t strides[1] = {kp};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
}

void SliceStack::alloc(int rows, int k, int ns_)
{
    if (rows != R || k != K || ns_ != ns) zeroed_for = 0;
    R = rows; K = k; ns = ns_;
    Rp = ((rows + 255) / 256) * 256;   // multiple of both BM and BN
    Kp = ((k + BK - 1) / BK) * BK;
    size_t need = (size_t)ns * Rp * Kp;
    if (need > cap) { dev_free(q); q = (int8_t*)dev_alloc(need); cap = need; }
    if (Rp > ecap) { dev_free(E); E = (int*)dev_alloc((size_t)Rp * 4); ecap = Rp; }
    if (!maxbits_cap || (size_t)Rp > maxbits_cap) { dev_free(maxbits); maxbits = (unsigned long long*)dev_alloc((size_t)Rp * 8); maxbits_cap = Rp; }
}
void SliceStack::release() { dev_free(q); dev_free(E); dev_free(maxbits); q = nullptr; E = nullptr; maxbits = nullptr; cap = 0; ecap = 0; maxbits_cap = 0; }

// rows [row0, row0+rows) of an already allocated stack (S.alloc(total_rows, k, ns) + zero fill done by the caller)
void split_rows_into(SliceStack& S, int row0, const double* X, long ldx, int rows, cudaStream_t st)
{
    split_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(X, ldx, rows, S.K, S.Rp, S.Kp, S.ns, row0, S.q, S.E);
    CK(cudaGetLastError());
}

// first half of split_rows for a matrix that is still being produced: allocation + cleared row maxima, which the producer
// (stage 1 of DF-K, GemmParams::rowmax) fills; split_rows_premax then cuts the slices without a row-maximum pass
void split_rows_prepare(SliceStack& S, int rows, int k, int ns, cudaStream_t st)
{
    S.alloc(rows, k, ns);
    S.zeroed_for = 0;
    CK(cudaMemsetAsync(S.maxbits, 0, (size_t)rows * 8, st));
}
void split_rows_premax(SliceStack& S, const double* X, long ldx, int rows, int k, int ns, cudaStream_t st)
{
    if (S.R != rows || S.K != k || S.ns != ns) throw std::runtime_error("split_rows_premax: call split_rows_prepare first");
    if (S.Rp > rows) {
        for (int s = 0; s < ns; s++)
            CK(cudaMemsetAsync(S.q + ((size_t)s * S.Rp + rows) * S.Kp, 0, (size_t)(S.Rp - rows) * S.Kp, st));
        CK(cudaMemsetAsync(S.E + rows, 0, (size_t)(S.Rp - rows) * 4, st));
    }
    const long seglen = 8192;
    unsigned nseg = (unsigned)((S.Kp + seglen - 1) / seglen);
    split_long_kernel<<<dim3(nseg, rows), 256, 0, st>>>(X, ldx, k, S.Rp, S.Kp, ns, seglen, S.maxbits, S.q, S.E);
    CK(cudaGetLastError());
}

void split_rows(SliceStack& S, const double* X, long ldx, int rows, int k, int ns, cudaStream_t st)
{
    S.alloc(rows, k, ns);
    S.zeroed_for = 0;
    if (S.Rp > rows) {   // zero the pad rows of every slice
        for (int s = 0; s < ns; s++)
            CK(cudaMemsetAsync(S.q + ((size_t)s * S.Rp + rows) * S.Kp, 0, (size_t)(S.Rp - rows) * S.Kp, st));
        CK(cudaMemsetAsync(S.E + rows, 0, (size_t)(S.Rp - rows) * 4, st));
    }
    if ((long)k >= 8192 && rows < 4096) {
        const long seglen = 8192;
        unsigned nseg = (unsigned)((S.Kp + seglen - 1) / seglen);
        CK(cudaMemsetAsync(S.maxbits, 0, (size_t)rows * 8, st));
        rowmax_kernel<<<dim3(nseg, rows), 256, 0, st>>>(X, ldx, k, seglen, S.maxbits);
        split_long_kernel<<<dim3(nseg, rows), 256, 0, st>>>(X, ldx, k, S.Rp, S.Kp, ns, seglen, S.maxbits, S.q, S.E);
    } else {
        split_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(X, ldx, rows, k, S.Rp, S.Kp, ns, 0, S.q, S.E);
    }
    CK(cudaGetLastError());
}

// ---- the DF tensor from its packed rows (see i8gemm.cuh (3)) ----
// rowexp[nr][nao]: exponent of every row (P, a) of the unpacked tensor; cderi points at the first of the nr packed rows
void packed_rowexp(const double* cderi, long npair, int nao, int nr, int* rowexp, float* rownorm2, cudaStream_t st)
{
    CK(cudaMemsetAsync(rowexp, 0x80, (size_t)nr * nao * 4, st));     // EXP_NONE
    if (rownorm2) CK(cudaMemsetAsync(rownorm2, 0, (size_t)nr * nao * 4, st));
    static bool configured = false;
    if (!configured) { CK(cudaFuncSetAttribute(packed_rowexp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); configured = true; }
    if ((size_t)nao * 8 > 160 * 1024) throw std::runtime_error("packed_rowexp: nao too large for the shared-memory tables");
    for (int p0 = 0; p0 < nr; p0 += 32768) {
        int n = std::min(32768, nr - p0);
        packed_rowexp_kernel<<<dim3((nao + 63) / 64, n), 256, (size_t)nao * 8, st>>>(cderi + (size_t)p0 * npair, npair, nao, rowexp + (size_t)p0 * nao,
                                                                                      rownorm2 ? rownorm2 + (size_t)p0 * nao : nullptr);
    }
    CK(cudaGetLastError());
}
// ---- stage 1 cutting the slices of Y itself (GemmParams::yq) ----
// cmax2[0] = max_i ||row i of X||^2 (X = the right factor of stage 1 as rows [ncol][k])
void colnorm_max(const double* X, long ldx, int nrows, int k, double* cmax2, cudaStream_t st)
{
    CK(cudaMemsetAsync(cmax2, 0, 8, st));
    colnorm_max_kernel<<<(nrows + 7) / 8, 256, 0, st>>>(X, ldx, nrows, k, reinterpret_cast<unsigned long long*>(cmax2));
    CK(cudaGetLastError());
}
// the Y stack of one block: rows = nao, columns (P, i) with i padded to a multiple of 16; pads are zeroed when the shape changes
// (the epilogue of stage 1 overwrites every real element of every block), exponents = the Cauchy-Schwarz bound of each row
void y_prepare(SliceStack& S, int nao, int nr, int ncolp, int ns, const float* rownorm2_block, const double* cmax2, cudaStream_t st)
{
    const int k = nr * ncolp;
    S.alloc(nao, k, ns);
    const bool same = (S.zeroed_for == (long)nao * 1000003L + k);
    if (!same) {
        CK(cudaMemsetAsync(S.q, 0, (size_t)ns * S.Rp * S.Kp, st));
        CK(cudaMemsetAsync(S.E, 0, (size_t)S.Rp * 4, st));
        S.zeroed_for = (long)nao * 1000003L + k;
    }
    yexp_bound_kernel<<<(nao + 255) / 256, 256, 0, st>>>(rownorm2_block, nr, nao, cmax2, S.E);
    CK(cudaGetLastError());
}
// slices of the unpacked rows (P, a), P in [0, nr), into the rows out_row0 + P nao + a of an allocated stack
void split_packed_into(SliceStack& S, int out_row0, const double* cderi, long npair, int nao, int nr, const int* rowexp, cudaStream_t st)
{
    if (S.K != nao) throw std::runtime_error("split_packed: stack width does not match nao");
    const unsigned nt = (unsigned)(S.Kp / PT);
    for (int p0 = 0; p0 < nr; p0 += 32768) {
        int n = std::min(32768, nr - p0);
        if (S.ns == 7)
            split_packed_kernel<true><<<dim3(nt, nt, n), 256, 0, st>>>(cderi + (size_t)p0 * npair, npair, nao, rowexp + (size_t)p0 * nao, S.ns, S.Rp, S.Kp,
                                                                       out_row0 + p0 * nao, S.q, S.E);
        else
            split_packed_kernel<false><<<dim3(nt, nt, n), 256, 0, st>>>(cderi + (size_t)p0 * npair, npair, nao, rowexp + (size_t)p0 * nao, S.ns, S.Rp, S.Kp,
                                                                        out_row0 + p0 * nao, S.q, S.E);
    }
    CK(cudaGetLastError());
}
// a stack holding exactly these nr packed rows (allocated here, pad rows zeroed)
void split_packed(SliceStack& S, const double* cderi, long npair, int nao, int nr, const int* rowexp, int ns, cudaStream_t st)
{
    const int rows = nr * nao;
    S.alloc(rows, nao, ns);
    if (S.Rp > rows) {
        for (int s = 0; s < ns; s++)
            CK(cudaMemsetAsync(S.q + ((size_t)s * S.Rp + rows) * S.Kp, 0, (size_t)(S.Rp - rows) * S.Kp, st));
        CK(cudaMemsetAsync(S.E + rows, 0, (size_t)(S.Rp - rows) * 4, st));
    }
    split_packed_into(S, 0, cderi, npair, nao, nr, rowexp, st);
}

// stage-1 GEMM of DF-K with all slice-pair groups resident in TMEM (i8gemm_ar_kernel): rows [a_row0, a_row0+M) of A
void gemm_ar(const SliceStack& A, int a_row0, int M, const SliceStack& B, double* C, long ldc, int inner, cudaStream_t st,
             unsigned long long* rowmax, const SliceStack* Yout, int y_ncolp)
{
    if (A.Kp != B.Kp || A.ns != B.ns) throw std::runtime_error("i8gemm_ar: operand stacks disagree");
    if (A.ns * AR_BN > 512) throw std::runtime_error("i8gemm_ar: too many slices for TMEM");
    static bool configured = false;
    static int nsm = 148;
    if (!configured) {
        CK(cudaFuncSetAttribute(i8gemm_ar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AR_SMEM_MAX));
        int dev = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    static const int nsa_env = getenv("B200JK_AR_NSA") ? atoi(getenv("B200JK_AR_NSA")) : 0;   // tuning: A ring depth
    int nsa = AR_MAXA;
    while (nsa > 2 && ar_smem_bytes(nsa, A.ns) > AR_SMEM_MAX) nsa--;
    if (nsa_env >= 2 && nsa_env < nsa) nsa = nsa_env;
    CUtensorMap ta, tb;
    make_tmap(&ta, A.q, (uint64_t)A.ns * A.Rp, A.Kp, BM);
    make_tmap(&tb, B.q, (uint64_t)B.ns * B.Rp, B.Kp, AR_BN);
    GemmParams P{};
    P.M = M; P.N = B.R; P.Kp = A.Kp; P.Mp = A.Rp; P.Np = B.Rp; P.ns = A.ns; P.symmetric = 0;
    P.Ea = A.E; P.Eb = B.E; P.C = C; P.ldc = ldc; P.inner = inner; P.a_row0 = a_row0; P.ksplit = 1; P.dbg = nullptr;
    static const int stack = getenv("B200JK_AR_STACK") ? atoi(getenv("B200JK_AR_STACK")) : 1;   // 0: one slice pair per MMA (tuning yardstick)
    P.stack = stack;
    P.nsa = nsa;
    const int ntiles = ((B.R + AR_BN - 1) / AR_BN) * ((M + BM - 1) / BM);
    P.ar_ntiles = ntiles; P.ar_ksplit = 1; P.ar_kb_per = A.Kp / BK; P.accumulate = 0; P.rowmax = rowmax;
    if (Yout) {
        if (inner <= 0 || (y_ncolp & 15) || y_ncolp < B.R || (long)((M + inner - 1) / inner) * y_ncolp > Yout->Kp || Yout->R != inner)
            throw std::runtime_error("i8gemm_ar: inconsistent Y stack");
        P.yq = Yout->q; P.Ey = Yout->E; P.y_Rp = Yout->Rp; P.y_Kp = Yout->Kp; P.y_ncolp = y_ncolp;
    }
    static const int persist = getenv("B200JK_AR_PERSIST") ? atoi(getenv("B200JK_AR_PERSIST")) : 1;   // 0: one tile per CTA (yardstick)
    dim3 grid(persist ? std::min(ntiles, nsm) : ntiles);
    static const bool dbg = getenv("B200JK_I8_DEBUG") != nullptr;   // cycle stamps of CTA 0 (tuning)
    static long long* d_dbg = nullptr;
    static int dbg_left = 2;
    if (dbg && dbg_left > 0) {
        if (!d_dbg) d_dbg = (long long*)dev_alloc(64 * 8);
        dev_zero(d_dbg, 64 * 8, st);
        P.dbg = d_dbg;
    }
    i8gemm_ar_kernel<<<grid, NTHREADS, ar_smem_bytes(nsa, A.ns), st>>>(ta, tb, P);
    if (P.dbg) {
        long long hd[64];
        d2h(hd, d_dbg, 64 * 8, st);
        CK(cudaStreamSynchronize(st));
        dbg_left--;
        for (int it = 0; it < 6; it++)
            fprintf(stderr, "[i8gemm_ar CTA0 tile %d] mma: wait_tmem %lld issue %lld | epi: wait_acc %lld drain %lld store %lld | tile period %lld cycles\n", it,
                    hd[it * 8 + 1] - hd[it * 8 + 0], hd[it * 8 + 2] - hd[it * 8 + 1], hd[it * 8 + 4] - hd[it * 8 + 3],
                    hd[it * 8 + 5] - hd[it * 8 + 4], hd[it * 8 + 6] - hd[it * 8 + 5], it ? hd[it * 8 + 6] - hd[(it - 1) * 8 + 6] : 0LL);
    }
    CK(cudaGetLastError());
}

// C += A B^T (upper triangle only when symmetric) on the A-stationary kernel with all slice-pair groups resident in TMEM
// (i8gemm_ar_kernel): stage 2 of DF-K.  Against i8gemm_kernel's one (A_k, B_l) tile pair per pipeline stage this loads every
// A slice tile once per K block and multiplies it with all its B slices: half the operand traffic per output column.
// Work items = tiles x K ranges, sized to fill whole waves of the persistent grid.
void gemm_ar_acc(const SliceStack& A, const SliceStack& B, double* C, long ldc, bool symmetric, cudaStream_t st)
{
    if (A.Kp != B.Kp || A.ns != B.ns) throw std::runtime_error("i8gemm_ar_acc: operand stacks disagree");
    if (A.ns * AR_BN > 512) throw std::runtime_error("i8gemm_ar_acc: too many slices for TMEM");
    static bool configured = false;
    static int nsm = 148;
    if (!configured) {
        CK(cudaFuncSetAttribute(i8gemm_ar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AR_SMEM_MAX));
        int dev = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    int nsa = AR_MAXA;
    while (nsa > 2 && ar_smem_bytes(nsa, A.ns) > AR_SMEM_MAX) nsa--;
    CUtensorMap ta, tb;
    make_tmap(&ta, A.q, (uint64_t)A.ns * A.Rp, A.Kp, BM);
    make_tmap(&tb, B.q, (uint64_t)B.ns * B.Rp, B.Kp, AR_BN);
    GemmParams P{};
    P.M = A.R; P.N = B.R; P.Kp = A.Kp; P.Mp = A.Rp; P.Np = B.Rp; P.ns = A.ns; P.symmetric = symmetric ? 1 : 0;
    P.Ea = A.E; P.Eb = B.E; P.C = C; P.ldc = ldc; P.inner = 0; P.a_row0 = 0; P.ksplit = 1; P.dbg = nullptr;
    P.stack = 1; P.nsa = nsa; P.accumulate = 1;
    const int ntm = (A.R + BM - 1) / BM, ntn = (B.R + AR_BN - 1) / AR_BN;
    int tiles = 0;
    for (int mt = 0; mt < ntm; mt++) tiles += symmetric ? std::max(0, ntn - (BM / AR_BN) * mt) : ntn;
    if (symmetric && ntn < (BM / AR_BN) * (ntm - 1) + 1) throw std::runtime_error("i8gemm_ar_acc: symmetric product needs a square output");
    const int nkb = A.Kp / BK;
    // K ranges: >= 8 K blocks each; among those the count that wastes the least of the last wave of the persistent grid
    int best = 1; double best_eff = -1.0;
    for (int ks = 1; ks <= std::max(1, nkb / 8); ks++) {
        const int per = (nkb + ks - 1) / ks, kse = (nkb + per - 1) / per;
        const long items = (long)tiles * kse;
        const double eff = (double)items / ((double)nsm * ((items + nsm - 1) / nsm));
        if (eff > best_eff + 1e-9) { best_eff = eff; best = kse; }
        if (items > 12L * nsm) break;
    }
    static const int ks_env = getenv("B200JK_G2_KS") ? atoi(getenv("B200JK_G2_KS")) : 0;   // tuning experiment: force the K-range count
    if (ks_env > 0) best = std::min(ks_env, std::max(1, nkb));
    P.ar_kb_per = (nkb + best - 1) / best;
    P.ar_ksplit = (nkb + P.ar_kb_per - 1) / P.ar_kb_per;
    P.ar_ntiles = tiles;
    const long items = (long)tiles * P.ar_ksplit;
    dim3 grid((unsigned)std::min<long>(items, nsm));
    i8gemm_ar_kernel<<<grid, NTHREADS, ar_smem_bytes(nsa, A.ns), st>>>(ta, tb, P);
    CK(cudaGetLastError());
}

void gemm(const SliceStack& A, const SliceStack& B, double* C, long ldc, int inner, bool symmetric, cudaStream_t st, long long* dbg)
{
    if (A.Kp != B.Kp || A.ns != B.ns) throw std::runtime_error("i8gemm: operand stacks disagree");
    static bool configured = false;
    if (!configured) {
        CK(cudaFuncSetAttribute(i8gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    CUtensorMap ta, tb;
    make_tmap(&ta, A.q, (uint64_t)A.ns * A.Rp, A.Kp, BM);
    make_tmap(&tb, B.q, (uint64_t)B.ns * B.Rp, B.Kp, BN);
    GemmParams P{};
    P.M = A.R; P.N = B.R; P.Kp = A.Kp; P.Mp = A.Rp; P.Np = B.Rp; P.ns = A.ns; P.symmetric = symmetric ? 1 : 0;
    P.Ea = A.E; P.Eb = B.E; P.C = C; P.ldc = ldc; P.inner = inner; P.dbg = dbg; P.a_row0 = 0;
    const int mtiles = (A.R + BM - 1) / BM, ntiles = (B.R + BN - 1) / BN;
    int tiles = 0;
    for (int mt = 0; mt < mtiles; mt++)
        for (int nt = 0; nt < ntiles; nt++)
            if (!(symmetric && (nt + 1) * BN <= mt * BM)) tiles++;
    int nkb = A.Kp / BK;
    // split K so that the CTAs fill one (or two) waves of 148 SMs as exactly as possible, >= 8 K blocks each
    int ksplit = 1, best_waste = 1 << 30;
    for (int ks = 1; ks <= 64 && ks * 8 <= std::max(nkb, 8); ks++) {
        int ctas = tiles * ks, waves = (ctas + 147) / 148;
        if (waves > 2) break;
        int waste = (waves * 148 - ctas) * 1000 / (waves * 148);
        if (waste < best_waste || (waste == best_waste && ks > ksplit)) { best_waste = waste; ksplit = ks; }
    }
    P.ksplit = ksplit;
    dim3 grid((B.R + BN - 1) / BN, (A.R + BM - 1) / BM, ksplit);
    i8gemm_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(ta, tb, P);
    CK(cudaGetLastError());
}

}  // namespace i8g
}  // namespace b200jk
#endif

// C = A B^T through the tcgen05 int8-slice path; A [M,K], B [N,K], C [M,N] host fp64 (self-test / tests).
extern "C" int b200jk_i8gemm_test(b200jk_handle h, int M, int N, int K, const double* A, const double* B, double* C, int ns,
                                  int symmetric)
{
    if (!h) return 1;
#ifndef B200JK_EMULATE
    try {
        using namespace b200jk::i8g;
        if (ns < 1 || ns > MAXS) throw std::runtime_error("ns out of range");
        CK(cudaSetDevice(h->device));
        cudaStream_t st = h->stream;
        double* dA = (double*)dev_alloc((size_t)M * K * 8);
        double* dB = (double*)dev_alloc((size_t)N * K * 8);
        double* dC = (double*)dev_alloc((size_t)M * N * 8);
        h2d(dA, A, (size_t)M * K * 8, st);
        h2d(dB, B, (size_t)N * K * 8, st);
        dev_zero(dC, (size_t)M * N * 8, st);
        SliceStack SA, SB;
        split_rows(SA, dA, K, M, K, ns, st);
        split_rows(SB, dB, K, N, K, ns, st);
        CK(cudaEventRecord(h->ev0, st));
        long long* ddbg = (long long*)dev_alloc(64 * 8);
        dev_zero(ddbg, 64 * 8, st);
        gemm(SA, SB, dC, N, 0, symmetric != 0, st, ddbg);
        CK(cudaEventRecord(h->ev1, st));
        d2h(C, dC, (size_t)M * N * 8, st);
        long long hd[64];
        d2h(hd, ddbg, 64 * 8, st);
        CK(cudaStreamSynchronize(st));
        if (getenv("B200JK_I8_DEBUG")) {
            fprintf(stderr, "i8gemm cycles (CTA 0,0) rel. to start:");
            for (int i = 8; i < 8 + 4 * ns; i++) fprintf(stderr, "%s%lld", (i % 4 == 0) ? " | " : " ", hd[i] ? hd[i] - hd[0] : -1);
            fprintf(stderr, "\n");
        }
        dev_free(ddbg);
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->stats.ms_kernels = ms;
        SA.release(); SB.release();
        dev_free(dA); dev_free(dB); dev_free(dC);
    } catch (std::exception& e) { set_err(h, e.what()); return 2; }
    return 0;
#else
    (void)M; (void)N; (void)K; (void)A; (void)B; (void)C; (void)ns; (void)symmetric;
    set_err(h, "tcgen05 path is not emulated on the CPU");
    return 3;
#endif
}
