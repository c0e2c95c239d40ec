// i8gemm.cuh — FP64-accurate GEMM on the 5th-generation tensor cores (tcgen05, kind::i8) by error-free
// slicing (Ozaki scheme): the DF-K contractions of pyscf/df/df_jk.py:373-380 (AO2MOnr_e2_drv's dsymm and
// lib.dot/NPdgemm) executed as exact int8 x int8 -> int32 slice products.
//
//   C[m,n] (+)= sum_k A[m,k] B[n,k]            A: [M,K], B: [N,K] fp64, both K-major (row-major, K contiguous)
//
// 1. split_rows_kernel: every row is scaled by 2^-E (E = ceil(log2 max|row|)) and cut into NS signed 7-bit slices
//        a = 2^E ( q0 2^-6 + q1 2^-13 + ... + q_{NS-1} 2^-(6+7(NS-1)) ) + tail,  |q| <= 64.
// 2. i8gemm_kernel: for every slice-pair group g = k+l (same power of two) the products A_k B_l^T are
//    accumulated EXACTLY in a TMEM int32 accumulator by tcgen05.mma.kind::i8 (operands TMA-loaded into
//    128B-swizzled shared memory through a 4-stage mbarrier ring); the epilogue warps read the accumulator
//    back with tcgen05.ld, convert to fp64, apply 2^(Ea[m]+Eb[n]-12-7g) and add into C.  Two TMEM
//    accumulators (2 x 256 columns) let the epilogue of group g overlap the MMAs of group g+1.
// Only pairs with k+l < NS are formed (the rest is below the slice truncation): NS(NS+1)/2 slice GEMMs.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200jk {
namespace i8g {

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BN = 256;        // UMMA N
constexpr int BK = 128;        // bytes (= int8 elements) of K per pipeline stage: one 128B swizzle row
constexpr int UK = 32;         // K per tcgen05.mma.kind::i8
constexpr int NSTAGE = 4;
constexpr int MAXS = 8;        // max slices
constexpr int A_STAGE_BYTES = BM * BK;
constexpr int B_STAGE_BYTES = BN * BK;
constexpr int SMEM_BYTES = NSTAGE * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NTHREADS = 192;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 epilogue

struct GemmParams {
    int M, N, Kp;          // Kp: padded K (multiple of BK)
    int Mp, Np;            // padded rows of the slice stacks (multiples of BM / BN)
    int ns;                // slices
    int symmetric;         // 1: B == A, only tiles with n-tile >= m-tile*(BM/BN...) are computed (upper part)
    const int* Ea; const int* Eb;   // per-row exponents
    double* C; long ldc;   // fp64 output, row-major [M, ldc]
    // optional transposed-scatter epilogue (stage 1 of DF-K): C element (m, n) is stored at
    //   C[(m % inner) * ldc + (m / inner) * N + n]   when inner > 0
    int inner;
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ double pow2i(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major), canonical value 1
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: 8 rows x 128 B between row groups
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// instruction descriptor for kind::i8: c=S32, a=b=INT8, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------- GEMM kernel
// tmapA / tmapB: 2-D uint8 tensors [ns*Mp (resp. ns*Np) rows][Kp bytes], box {BK, BM} / {BK, BN}, SWIZZLE_128B.
__global__ void __launch_bounds__(NTHREADS, 1)
i8gemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const GemmParams P)
{
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + NSTAGE * A_STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(smem + NSTAGE * (A_STAGE_BYTES + B_STAGE_BYTES));
    uint64_t* full = bars;                 // [NSTAGE]
    uint64_t* empty = bars + NSTAGE;       // [NSTAGE]
    uint64_t* tfull = bars + 2 * NSTAGE;   // [2]
    uint64_t* tempty = bars + 2 * NSTAGE + 2;  // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * NSTAGE + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.y, nt = blockIdx.x;
    if (P.symmetric && (nt + 1) * BN <= mt * BM) return;   // tile entirely below the diagonal
    const int nkb = P.Kp / BK;
    const int ns = P.ns;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmapA);
        prefetch_tmap(&tmapB);
        for (int i = 0; i < NSTAGE; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int g = ns - 1; g >= 0; g--)
                for (int k = 0; k <= g; k++) {
                    const int l = g - k;
                    for (int kb = 0; kb < nkb; kb++) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], A_STAGE_BYTES + B_STAGE_BYTES);
                        tma_load_2d(sA + stage * A_STAGE_BYTES, &tmapA, &full[stage], kb * BK, k * P.Mp + mt * BM);
                        tma_load_2d(sB + stage * B_STAGE_BYTES, &tmapB, &full[stage], kb * BK, l * P.Np + nt * BN);
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_i8(BM, BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int g = ns - 1; g >= 0; g--, it++) {
                const int buf = it & 1;
                const uint32_t tphase = (it >> 1) & 1;
                mbar_wait(&tempty[buf], tphase ^ 1);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * BN;
                uint32_t acc = 0;
                for (int k = 0; k <= g; k++)
                    for (int kb = 0; kb < nkb; kb++) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(sA + stage * A_STAGE_BYTES), b0 = smem_u32(sB + stage * B_STAGE_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / UK; kk++) {
                            mma_i8(tacc, make_desc_k_sw128(a0 + kk * UK), make_desc_k_sw128(b0 + kk * UK), idesc, acc);
                            acc = 1;
                        }
                        mma_commit(&empty[stage]);            // frees the smem stage when these MMAs retire
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                mma_commit(&tfull[buf]);                      // accumulator of group g complete
            }
        }
    } else {
        // ===== epilogue warps: TMEM -> registers -> fp64 scale -> C =====
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int m = mt * BM + row;
        const bool mrow_ok = m < P.M;
        const double sa = mrow_ok ? pow2i(P.Ea[m]) : 0.0;
        int it = 0;
        for (int g = ns - 1; g >= 0; g--, it++) {
            const int buf = it & 1;
            const uint32_t tphase = (it >> 1) & 1;
            mbar_wait(&tfull[buf], tphase);
            tc_fence_after();
            const double sg = sa * pow2i(-12 - 7 * g);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + c0, r);
                const int n0 = nt * BN + c0;
                if (mrow_ok && n0 < P.N) {
                    double* crow;
                    if (P.inner > 0) crow = P.C + (long)(m % P.inner) * P.ldc + (long)(m / P.inner) * P.N;
                    else crow = P.C + (long)m * P.ldc;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int n = n0 + j;
                        if (n < P.N && (!P.symmetric || n >= m)) {
                            const double v = (double)(int)r[j] * sg * pow2i(P.Eb[n]);
                            crow[n] += v;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------- slicing kernel
// X: [R, K] fp64 row-major (row stride ldx).  out: [ns][Rp][Kp] int8 (rows >= R and cols >= K zero).  E[r] exponents.
// One warp per row.
__global__ void __launch_bounds__(256) split_rows_kernel(const double* __restrict__ X, long ldx, int R, int K, int Rp, int Kp, int ns,
                                                         int8_t* __restrict__ out, int* __restrict__ E)
{
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= Rp) return;
    if (r >= R) {
        for (int s = 0; s < ns; s++)
            for (int k = lane; k < Kp; k += 32) out[((long)s * Rp + r) * Kp + k] = 0;
        if (lane == 0) E[r] = 0;
        return;
    }
    const double* x = X + (long)r * ldx;
    double mx = 0.0;
    for (int k = lane; k < K; k += 32) mx = fmax(mx, fabs(x[k]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e = 0;
    if (mx > 0.0) { frexp(mx, &e); }         // mx = f * 2^e, f in [0.5,1)  =>  |x| / 2^e < 1
    if (lane == 0) E[r] = e;
    const double sc = ldexp(1.0, 6 - e);
    for (int k = lane; k < Kp; k += 32) {
        double rr = (k < K) ? x[k] * sc : 0.0;
        for (int s = 0; s < ns; s++) {
            double qv = rint(rr);
            out[((long)s * Rp + r) * Kp + k] = (int8_t)(int)qv;
            rr = (rr - qv) * 128.0;
        }
    }
}

}  // namespace i8g
}  // namespace b200jk
