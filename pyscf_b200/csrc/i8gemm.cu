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
    R = rows; K = k; ns = ns_;
    Rp = ((rows + 255) / 256) * 256;   // multiple of both BM and BN
    Kp = ((k + BK - 1) / BK) * BK;
    size_t need = (size_t)ns * Rp * Kp;
    if (need > cap) { dev_free(q); q = (int8_t*)dev_alloc(need); cap = need; }
    if (Rp > ecap) { dev_free(E); E = (int*)dev_alloc((size_t)Rp * 4); ecap = Rp; }
}
void SliceStack::release() { dev_free(q); dev_free(E); q = nullptr; E = nullptr; cap = 0; ecap = 0; }

void split_rows(SliceStack& S, const double* X, long ldx, int rows, int k, int ns, cudaStream_t st)
{
    S.alloc(rows, k, ns);
    split_rows_kernel<<<(S.Rp + 7) / 8, 256, 0, st>>>(X, ldx, rows, k, S.Rp, S.Kp, ns, S.q, S.E);
    CK(cudaGetLastError());
}

void gemm(const SliceStack& A, const SliceStack& B, double* C, long ldc, int inner, bool symmetric, cudaStream_t st)
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
    P.Ea = A.E; P.Eb = B.E; P.C = C; P.ldc = ldc; P.inner = inner;
    dim3 grid((B.R + BN - 1) / BN, (A.R + BM - 1) / BM);
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
        gemm(SA, SB, dC, N, 0, symmetric != 0, st);
        CK(cudaEventRecord(h->ev1, st));
        d2h(C, dC, (size_t)M * N * 8, st);
        CK(cudaStreamSynchronize(st));
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
