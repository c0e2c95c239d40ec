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
constexpr int EPI_STAGE_INTS = 32 * 33;   // per epilogue warp: 32x32 int32 transpose buffer, padded
constexpr int SMEM_BYTES = NSTAGE * (A_STAGE_BYTES + B_STAGE_BYTES) + 4 * EPI_STAGE_INTS * 4 + 1024 /*align*/ + 256 /*barriers*/;
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
    int a_row0;            // first row of the A stack used by this GEMM (row blocks of a persistent stack)
    int ksplit;            // K blocks are divided among gridDim.z CTAs (fp64 reductions make this safe)
    long long* dbg;        // optional cycle stamps of CTA (0,0) (tests/tuning)
    int stack;             // i8gemm_ar_kernel: multiply A_k with up to 4 stacked B slices per MMA (N = 256)
    int nsa;               // i8gemm_ar_kernel: depth of the A ring (host: as many 16 KB stages as fit in 227 KB)
    // i8gemm_ar_kernel as stage 2 of DF-K (K += Y Y^T / Y G^T): work item = (tile, K range); ar_ksplit K ranges per tile,
    // results meet in fp64 reductions (accumulate), only tiles touching the upper triangle when symmetric
    int ar_ksplit, ar_kb_per, ar_ntiles, accumulate;
    // stage 1: optional per-output-row maxima (bit pattern of max |C| per row m % inner, 64-bit atomicMax), so that the slicing
    // of Y needs no row-maximum pass of its own
    unsigned long long* rowmax;
    // stage 1 writing the int8 slices of Y itself (no fp64 Y): yq = base of the Y stack [ns][y_Rp][y_Kp], row = m % inner,
    // column = (m / inner) * y_ncolp + n (y_ncolp = N padded to 16), Ey[row] = exponent BOUND of the row (known before the GEMM:
    // Cauchy-Schwarz on the row norms of the tensor block and the column norms of the right factor)
    int8_t* yq; const int* Ey; int y_Rp, y_Kp, y_ncolp;
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

// the same load without the wait: several loads in flight, then one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld_32x32b_x32_nowait(uint32_t taddr, uint32_t (&r)[32])
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// exact int32 -> fp64 on the FP64 add pipe (one LOP + one DADD instead of a conversion instruction):
// the double with high word 0x43300000 and low word (x ^ 2^31) is 2^52 + 2^31 + x
__device__ __forceinline__ double i2d_exact(uint32_t x)
{
    return __hiloint2double(0x43300000, (int)(x ^ 0x80000000u)) - 4503601774854144.0;
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
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still known to be shared memory
    uint8_t* sA = smem;
    uint8_t* sB = smem + NSTAGE * A_STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(smem + NSTAGE * (A_STAGE_BYTES + B_STAGE_BYTES));
    uint64_t* full = bars;                 // [NSTAGE]
    uint64_t* empty = bars + NSTAGE;       // [NSTAGE]
    uint64_t* tfull = bars + 2 * NSTAGE;   // [2]
    uint64_t* tempty = bars + 2 * NSTAGE + 2;  // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * NSTAGE + 4);
    int* epi_stage = (int*)(smem + NSTAGE * (A_STAGE_BYTES + B_STAGE_BYTES) + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.y, nt = blockIdx.x;
    if (P.symmetric && (nt + 1) * BN <= mt * BM) return;   // tile entirely below the diagonal
    const int nkb_all = P.Kp / BK;
    const int kb_per = (nkb_all + P.ksplit - 1) / P.ksplit;
    const int kb0 = blockIdx.z * kb_per;
    const int kb1 = (kb0 + kb_per < nkb_all) ? kb0 + kb_per : nkb_all;
    const int nkb = kb1 - kb0;
    if (nkb <= 0) return;
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
    const bool dbg_on = P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    if (dbg_on && threadIdx.x == 0) P.dbg[0] = clock64();

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
                        tma_load_2d(sA + stage * A_STAGE_BYTES, &tmapA, &full[stage], (kb0 + kb) * BK, k * P.Mp + P.a_row0 + mt * BM);
                        tma_load_2d(sB + stage * B_STAGE_BYTES, &tmapB, &full[stage], (kb0 + kb) * BK, l * P.Np + nt * BN);
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
                if (dbg_on) P.dbg[8 + 4 * it] = clock64();
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
                if (dbg_on) P.dbg[9 + 4 * it] = clock64();
            }
        }
    } else {
        // ===== epilogue warps: TMEM -> registers -> smem transpose -> coalesced fp64 read-modify-write of C =====
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        int* stg = epi_stage + q * EPI_STAGE_INTS;
        const int mrow0 = mt * BM + q * 32;      // first row of this warp's 32-row band
        int it = 0;
        for (int g = ns - 1; g >= 0; g--, it++) {
            const int buf = it & 1;
            const uint32_t tphase = (it >> 1) & 1;
            mbar_wait(&tfull[buf], tphase);
            tc_fence_after();
            if (dbg_on && q == 0 && lane == 0) P.dbg[10 + 4 * it] = clock64();
            const int eg = -12 - 7 * g;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + c0, r);
                const int n = nt * BN + c0 + lane;       // this lane's column after the transpose
                if (nt * BN + c0 >= P.N) continue;
#pragma unroll
                for (int j = 0; j < 32; j++) stg[lane * 33 + j] = (int)r[j];
                __syncwarp();
                const bool ncol_ok = n < P.N;
                const int ebn = ncol_ok ? __ldg(P.Eb + n) : 0;
#pragma unroll 4
                for (int rr = 0; rr < 32; rr++) {
                    const int m = mrow0 + rr;
                    if (m >= P.M) break;
                    if (ncol_ok && (!P.symmetric || n >= m)) {
                        // __ldg: a plain load could not be hoisted above the preceding reduction (possible alias) and the
                        // 32 rows would serialise on the load latency
                        const double v = (double)stg[rr * 33 + lane] * pow2i(__ldg(P.Ea + P.a_row0 + m) + ebn + eg);
                        double* dst;
                        if (P.inner > 0) dst = P.C + (long)(m % P.inner) * P.ldc + (long)(m / P.inner) * P.N + n;
                        else dst = P.C + (long)m * P.ldc + n;
                        atomicAdd(dst, v);   // RED.ADD.F64: no read latency, safe under split-K
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (dbg_on && q == 0 && lane == 0) P.dbg[11 + 4 * it] = clock64();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------- GEMM kernel, "all groups resident"
// Variant for a SHORT contraction (K = nao) and a narrow N (stage 1 of DF-K: N = nocc): tile 128 x 64 with ALL slice-pair
// groups resident in TMEM (group g at columns g*64, ns*64 <= 512).  The A_k tile is loaded ONCE per (K block, k) and
// reused for every B_l (A-stationary: no operand re-reads), and there is ONE epilogue per tile that combines the groups in
// fp64 registers and stores each output once (no read-modify-write, no zero fill).
// Persistent: one CTA per SM walks over the tiles (N tile fastest, so the CTAs that share an A tile run side by side
// and meet in L2).  The TMA producer keeps prefetching the next tile's operands while the epilogue drains TMEM; the
// accumulators are handed back to the MMA warp as soon as the last tcgen05.ld has landed, before the stores.
constexpr int AR_BN = 64;
constexpr int AR_NSB = 2;                        // B ring: ALL slices of a K block per stage
constexpr int AR_MAXA = 8;                       // A ring: one slice tile per stage, depth chosen by the host (P.nsa)
constexpr int AR_A_BYTES = BM * BK, AR_B1_BYTES = AR_BN * BK;
constexpr int AR_BAR_BYTES = 512, AR_EPI_BYTES = 0;   // the epilogue needs no staging: every lane stores its own row
constexpr int AR_SMEM_MAX = 232448;              // 227 KB opt-in limit per CTA
__host__ __device__ constexpr int ar_smem_bytes(int nsa, int ns) { return nsa * AR_A_BYTES + AR_NSB * ns * AR_B1_BYTES + AR_BAR_BYTES + AR_EPI_BYTES + 1024; }

// work item -> (m tile, n tile, K-block range).  Items are numbered K range slowest, tile fastest (CTAs running side by side
// share the A tile of their m tile in L2); symmetric: only the tiles with (nt + 1) * AR_BN > mt * BM, i.e. nt >= 2 mt.
__device__ __forceinline__ void ar_decode_item(const GemmParams& P, int item, int ntn, int nkb, int& mt, int& nt, int& kb0, int& kb1)
{
    const int ks = item / P.ar_ntiles;
    int tile = item - ks * P.ar_ntiles;
    if (P.symmetric) {
        mt = 0;
        for (;;) {
            const int cnt = ntn - (BM / AR_BN) * mt;
            if (tile < cnt) break;
            tile -= cnt; mt++;
        }
        nt = (BM / AR_BN) * mt + tile;
    } else {
        mt = tile / ntn; nt = tile - mt * ntn;
    }
    kb0 = ks * P.ar_kb_per;
    kb1 = (kb0 + P.ar_kb_per < nkb) ? kb0 + P.ar_kb_per : nkb;
}

__global__ void __launch_bounds__(NTHREADS, 1)
i8gemm_ar_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const GemmParams P)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still known to be shared memory
    const int ns = P.ns, nsa = P.nsa;
    const int bstage = ns * AR_B1_BYTES;
    uint8_t* sA = smem;
    uint8_t* sB = smem + nsa * AR_A_BYTES;
    uint64_t* bars = (uint64_t*)(sB + AR_NSB * bstage);
    uint64_t* afull = bars;                      // [AR_MAXA]
    uint64_t* aempty = afull + AR_MAXA;
    uint64_t* bfull = aempty + AR_MAXA;          // [AR_NSB]
    uint64_t* bempty = bfull + AR_NSB;
    uint64_t* tfull = bempty + AR_NSB;           // [1] accumulators complete (MMA -> epilogue)
    uint64_t* tempty = tfull + 1;                // [MAXS] accumulator of slice-pair group g drained (epilogue -> MMA), 4 arrivals each
    uint32_t* tmem_slot = (uint32_t*)(tempty + MAXS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = P.Kp / BK;
    const int ntn = (P.N + AR_BN - 1) / AR_BN, ntm = (P.M + BM - 1) / BM;
    const int ntiles = P.ar_ntiles * P.ar_ksplit;     // work items
    (void)ntm;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmapA);
        prefetch_tmap(&tmapB);
        for (int i = 0; i < nsa; i++) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < AR_NSB; i++) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
        mbar_init(tfull, 1);
        for (int g = 0; g < MAXS; g++) mbar_init(&tempty[g], 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // The A slices of a K block are taken in DESCENDING order k = ns-1 .. 0: slice k multiplies B slices 0 .. ns-1-k into the
    // groups k .. ns-1, so group g is first touched by A_g, and the next tile's MMAs can start as soon as the epilogue has
    // drained group ns-1 (it drains in the same order) instead of waiting for the whole accumulator.
    if (warp == 0) {
        if (lane == 0) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int mt, nt, kb0, kb1;
                ar_decode_item(P, tile, ntn, nkb, mt, nt, kb0, kb1);
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait(&bempty[sb], pb ^ 1);
                    mbar_expect_tx(&bfull[sb], ns * AR_B1_BYTES);
                    for (int l = 0; l < ns; l++)
                        tma_load_2d(sB + sb * bstage + l * AR_B1_BYTES, &tmapB, &bfull[sb], kb * BK, l * P.Np + nt * AR_BN);
                    if (++sb == AR_NSB) { sb = 0; pb ^= 1; }
                    for (int k = ns - 1; k >= 0; k--) {
                        mbar_wait(&aempty[sa], pa ^ 1);
                        mbar_expect_tx(&afull[sa], AR_A_BYTES);
                        tma_load_2d(sA + sa * AR_A_BYTES, &tmapA, &afull[sa], kb * BK, k * P.Mp + P.a_row0 + mt * BM);
                        if (++sa == nsa) { sa = 0; pa ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
                if (P.dbg && blockIdx.x == 0 && it < 6) P.dbg[it * 8 + 0] = clock64();
                int mt_, nt_, kb0, kb1;
                ar_decode_item(P, tile, ntn, nkb, mt_, nt_, kb0, kb1);
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait(&bfull[sb], pb);
                    tc_fence_after();
                    const uint32_t bbase = smem_u32(sB + sb * bstage);
                    for (int k = ns - 1; k >= 0; k--) {
                        mbar_wait(&afull[sa], pa);
                        if (kb == kb0) {
                            mbar_wait(&tempty[k], (uint32_t)((it & 1) ^ 1));   // the epilogue has read group k of the previous tile
                            if (k == ns - 1 && P.dbg && blockIdx.x == 0 && it < 6) P.dbg[it * 8 + 1] = clock64();
                        }
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(sA + sa * AR_A_BYTES);
                        // The B slices l = 0..ns-1-k of this K block lie back to back in shared memory (64 rows x 128 B each),
                        // i.e. they ARE one K-major tile of (ns-k)*64 rows, and their groups k+l are adjacent TMEM column
                        // blocks: one MMA with N = 64*cnt multiplies A_k with cnt slices at once.  A 128-row MMA costs the
                        // same ~128 cycles for any N <= 256, so stacking cuts the 28 slice-pair MMAs per K block to 10.
                        // The first touch of group g is (kb = 0, k = g, l = 0) and has to overwrite: in the first K block that
                        // pair gets an MMA of its own (N = 64), the stacked MMAs start at l = 1.
                        const int lstep = P.stack ? 4 : 1;
                        int l = 0;
                        if (kb == kb0) {
                            const uint32_t idesc_1 = make_idesc_i8(BM, AR_BN);
                            const uint32_t tacc = tmem_base + k * AR_BN;
                            uint32_t acc = 0;
#pragma unroll
                            for (int kk = 0; kk < BK / UK; kk++) {
                                mma_i8(tacc, make_desc_k_sw128(a0 + kk * UK), make_desc_k_sw128(bbase + kk * UK), idesc_1, acc);
                                acc = 1;
                            }
                            l = 1;
                        }
                        for (; l < ns - k; l += lstep) {
                            const int cnt = (ns - k - l < lstep) ? ns - k - l : lstep;
                            const uint32_t idesc_n = make_idesc_i8(BM, cnt * AR_BN);
                            const uint32_t b0 = bbase + l * AR_B1_BYTES;
                            const uint32_t tacc = tmem_base + (k + l) * AR_BN;
#pragma unroll
                            for (int kk = 0; kk < BK / UK; kk++)
                                mma_i8(tacc, make_desc_k_sw128(a0 + kk * UK), make_desc_k_sw128(b0 + kk * UK), idesc_n, 1u);
                        }
                        mma_commit(&aempty[sa]);
                        if (++sa == nsa) { sa = 0; pa ^= 1; }
                    }
                    mma_commit(&bempty[sb]);
                    if (++sb == AR_NSB) { sb = 0; pb ^= 1; }
                }
                mma_commit(tfull);
                if (P.dbg && blockIdx.x == 0 && it < 6) P.dbg[it * 8 + 2] = clock64();   // all MMAs of the tile issued
            }
        }
    } else {
        const int q = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            int mt, nt, kb0_, kb1_;
            ar_decode_item(P, tile, ntn, nkb, mt, nt, kb0_, kb1_);
            const int mrow0 = mt * BM + q * 32;
            // exponent and output offset of this warp's 32 rows, one row per lane, fetched once per tile (see i8gemm_kernel)
            const int mlane = mrow0 + lane;
            const int ea_lane = (mlane < P.M) ? __ldg(P.Ea + P.a_row0 + mlane) : 0;
            const long off_lane = (P.inner > 0) ? (long)(mlane % P.inner) * P.ldc + (long)(mlane / P.inner) * P.N : (long)mlane * P.ldc;
            const bool stamp = P.dbg && blockIdx.x == 0 && it < 6 && q == 0 && lane == 0;
            if (stamp) P.dbg[it * 8 + 3] = clock64();
            mbar_wait(tfull, (uint32_t)(it & 1));
            tc_fence_after();
            if (stamp) P.dbg[it * 8 + 4] = clock64();   // accumulators complete
            // combine the groups in fp64, smallest weight first, row = this lane, all 64 columns; every group is handed back to
            // the MMA warp as soon as its loads have landed
            double accv[AR_BN];
#pragma unroll
            for (int j = 0; j < AR_BN; j++) accv[j] = 0.0;
            const bool on0 = nt * AR_BN < P.N, on1 = nt * AR_BN + 32 < P.N;
#pragma unroll 1
            for (int g = ns - 1; g >= 0; g--) {
                uint32_t r0[32], r1[32];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + g * AR_BN;
                if (on0) tmem_ld_32x32b_x32_nowait(ta, r0);
                if (on1) tmem_ld_32x32b_x32_nowait(ta + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[g]);
                const double w = pow2i(-12 - 7 * g);
                if (on0) {
#pragma unroll
                    for (int j = 0; j < 32; j++) accv[j] += i2d_exact(r0[j]) * w;
                }
                if (on1) {
#pragma unroll
                    for (int j = 0; j < 32; j++) accv[32 + j] += i2d_exact(r1[j]) * w;
                }
            }
            if (stamp) P.dbg[it * 8 + 5] = clock64();   // TMEM handed back
            // lane = output row (the TMEM lane): its 64 columns are contiguous in C, so every lane streams its own 512 B;
            // the partial sectors of neighbouring stores merge in L2.  No transpose, no shuffles; the column exponents come
            // through the read-only path, so they are not ordered behind the stores of the previous row.
            if (P.yq) {
                // Y never exists in fp64: scale the row by its exponent bound and cut the balanced 7-bit digits here
                // (round-to-nearest by the 1.5*2^52 trick: two adds per digit instead of a rounding and a conversion instruction)
                if (mlane < P.M) {
                    const int nb = nt * AR_BN;
                    const int yrow = mlane % P.inner;
                    const long ycol0 = (long)(mlane / P.inner) * P.y_ncolp + nb;
                    const int esc = ea_lane + 6 - __ldg(P.Ey + yrow);
                    int8_t* dst0 = P.yq + (long)yrow * P.y_Kp + ycol0;
                    const long sstride = (long)P.y_Rp * P.y_Kp;
#pragma unroll
                    for (int j0 = 0; j0 < AR_BN; j0 += 16) {
                        if (nb + j0 >= P.y_ncolp) break;
                        double rr[16];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const int4 eb = __ldg(reinterpret_cast<const int4*>(P.Eb + nb + j0 + j));
                            rr[j] = accv[j0 + j] * pow2i(esc + eb.x); rr[j + 1] = accv[j0 + j + 1] * pow2i(esc + eb.y);
                            rr[j + 2] = accv[j0 + j + 2] * pow2i(esc + eb.z); rr[j + 3] = accv[j0 + j + 3] * pow2i(esc + eb.w);
                        }
                        for (int s_ = 0; s_ < ns; s_++) {
                            unsigned w[4];
#pragma unroll
                            for (int q4 = 0; q4 < 4; q4++) {
                                unsigned pack = 0;
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const double t_ = rr[4 * q4 + j] + 6755399441055744.0;
                                    const double qv = t_ - 6755399441055744.0;
                                    pack |= ((unsigned)__double2loint(t_) & 255u) << (8 * j);
                                    rr[4 * q4 + j] = (rr[4 * q4 + j] - qv) * 128.0;
                                }
                                w[q4] = pack;
                            }
                            *reinterpret_cast<uint4*>(dst0 + s_ * sstride + j0) = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                }
            } else if (P.accumulate) {
                // partial result of this K range: fp64 reductions; every lane walks its own row, so the 32 reductions of one
                // instruction go to 32 rows (scattered, but this epilogue runs once per 128 x 64 x K-range item)
                if (mlane < P.M) {
                    const int nb = nt * AR_BN;
                    double* dst = P.C + off_lane + nb;
#pragma unroll
                    for (int j = 0; j < AR_BN; j++) {
                        const int n = nb + j;
                        if (n < P.N && (!P.symmetric || n >= mlane)) atomicAdd(dst + j, accv[j] * pow2i(ea_lane + __ldg(P.Eb + n)));
                    }
                }
            } else if (mlane < P.M) {
                const int nb = nt * AR_BN;
                double* dst = P.C + off_lane + nb;
                double vmax = 0.0;
                if (((P.N | P.ldc) & 3) == 0 && nb + AR_BN <= P.N) {
                    // every row segment starts on a 32-byte boundary: 256-bit stores, one full sector per lane and instruction
                    // (scalar stores leave 32 eight-byte fragments per instruction for L2 to merge)
#pragma unroll
                    for (int j = 0; j < AR_BN; j += 4) {
                        const int4 eb = __ldg(reinterpret_cast<const int4*>(P.Eb + nb + j));
                        const double v0 = accv[j] * pow2i(ea_lane + eb.x), v1 = accv[j + 1] * pow2i(ea_lane + eb.y);
                        const double v2 = accv[j + 2] * pow2i(ea_lane + eb.z), v3 = accv[j + 3] * pow2i(ea_lane + eb.w);
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "d"(v0), "d"(v1), "d"(v2), "d"(v3) : "memory");
                        vmax = fmax(fmax(vmax, fmax(fabs(v0), fabs(v1))), fmax(fabs(v2), fabs(v3)));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < AR_BN; j++)
                        if (nb + j < P.N) { const double v = accv[j] * pow2i(ea_lane + __ldg(P.Eb + nb + j)); dst[j] = v; vmax = fmax(vmax, fabs(v)); }
                }
                if (P.rowmax && vmax > 0.0)
                    atomicMax(P.rowmax + (P.inner > 0 ? mlane % P.inner : mlane), (unsigned long long)__double_as_longlong(vmax));
            }
            if (stamp) P.dbg[it * 8 + 6] = clock64();   // tile stored
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------- slicing kernels
// X: [R, K] fp64 row-major (row stride ldx).  out: [ns][Rp][Kp] int8, rows written at out_row0 + r; E[out_row0 + r].
// Pad rows / pad columns of the stack must be zero (stack_alloc memsets once; pad columns are rewritten here).
// (1) one warp per row: many short rows (the unpacked tensor: K = nao).
__global__ void __launch_bounds__(256) split_rows_kernel(const double* __restrict__ X, long ldx, int R, int K, int Rp, int Kp, int ns,
                                                         int out_row0, int8_t* __restrict__ out, int* __restrict__ E)
{
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const double* x = X + (long)r * ldx;
    double mx = 0.0;
    for (int k = lane; k < K; k += 32) mx = fmax(mx, fabs(x[k]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e = 0;
    if (mx > 0.0) { frexp(mx, &e); }         // mx = f * 2^e, f in [0.5,1)  =>  |x| / 2^e < 1
    if (lane == 0) E[out_row0 + r] = e;
    const double sc = ldexp(1.0, 6 - e);
    for (int k = lane; k < Kp; k += 32) {
        double rr = (k < K) ? x[k] * sc : 0.0;
        for (int s = 0; s < ns; s++) {
            double qv = rint(rr);
            out[((long)s * Rp + out_row0 + r) * Kp + k] = (int8_t)(int)qv;
            rr = (rr - qv) * 128.0;
        }
    }
}
// (2) few long rows (Y: K = naux_block * nocc): grid (segments, rows); row maxima through 64-bit atomicMax on |x| bits
__global__ void __launch_bounds__(256) rowmax_kernel(const double* __restrict__ X, long ldx, int K, long seglen,
                                                     unsigned long long* __restrict__ maxbits)
{
    const int r = blockIdx.y;
    const long k0 = blockIdx.x * seglen, k1 = (k0 + seglen < K) ? k0 + seglen : K;
    const double* x = X + (long)r * ldx;
    double mx = 0.0;
    for (long k = k0 + threadIdx.x; k < k1; k += 256) mx = fmax(mx, fabs(x[k]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(&maxbits[r], (unsigned long long)__double_as_longlong(mx));
}
__global__ void __launch_bounds__(256) split_long_kernel(const double* __restrict__ X, long ldx, int K, int Rp, int Kp, int ns, long seglen,
                                                         const unsigned long long* __restrict__ maxbits, int8_t* __restrict__ out,
                                                         int* __restrict__ E)
{
    const int r = blockIdx.y;
    const long k0 = blockIdx.x * seglen, k1 = (k0 + seglen < Kp) ? k0 + seglen : Kp;
    const double* x = X + (long)r * ldx;
    const double mx = __longlong_as_double((long long)maxbits[r]);
    int e = 0;
    if (mx > 0.0) { frexp(mx, &e); }
    if (blockIdx.x == 0 && threadIdx.x == 0) E[r] = e;
    const double sc = ldexp(1.0, 6 - e);
    // 8 consecutive elements per thread: one 64-bit store per slice (a warp writes 256 contiguous bytes per instruction);
    // k0, k1 and Kp are multiples of 8 (segments of 8192, Kp multiple of 128)
    for (long k = k0 + threadIdx.x * 8L; k < k1; k += 256 * 8L) {
        double rr[8];
#pragma unroll
        for (int j = 0; j < 8; j++) rr[j] = (k + j < K) ? x[k + j] * sc : 0.0;
        for (int s = 0; s < ns; s++) {
            unsigned long long pack = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double qv = rint(rr[j]);
                pack |= (unsigned long long)(unsigned char)(int8_t)(int)qv << (8 * j);
                rr[j] = (rr[j] - qv) * 128.0;
            }
            *reinterpret_cast<unsigned long long*>(out + ((long)s * Rp + r) * Kp + k) = pack;
        }
    }
}


// (3) the DF tensor, straight from its PACKED rows (reference layout cderi[P][a(a+1)/2 + b], a >= b, pyscf/df/incore.py:134-136)
//     to the int8 slices of the UNPACKED matrices A_P[a][b] = A_P[b][a] that stage 1 of DF-K multiplies: no fp64 unpacked
//     copy is ever written.  Row (P, a) of the stack is scaled by its own exponent (max over the whole unpacked row).
constexpr int EXP_NONE = (int)0x80808080;    // "no non-zero element seen yet" (byte pattern of a memset with 0x80)
__device__ __forceinline__ int frexp_exp(double v)   // e with |v| = f 2^e, f in [0.5, 1); v != 0
{
    return (int)((__double_as_longlong(v) >> 52) & 0x7ff) - 1022;
}
// rowexp[P][a] = exponent of max_b |A_P[a][b]|, accumulated with integer atomicMax; the caller fills rowexp with EXP_NONE.
// grid (ceil(nao / 64), nr): a CTA reads the packed rows a0 .. a0+63 of auxiliary row P once, coalesced; an element (a, b)
// counts for row a (warp reduction) and for row b (shared-memory atomicMax, flushed once per CTA).
// rownorm2[P][a] (optional, zeroed by the caller) = sum_b A_P[a][b]^2 in single precision: the row norms behind the exponent
// bound of Y when stage 1 cuts the slices of Y itself.
__global__ void __launch_bounds__(256) packed_rowexp_kernel(const double* __restrict__ cderi, long npair, int nao, int* __restrict__ rowexp,
                                                            float* __restrict__ rownorm2)
{
    extern __shared__ int emax_s[];          // [nao] exponents, then [nao] partial squared norms
    float* ss = reinterpret_cast<float*>(emax_s + nao);
    const int P = blockIdx.y;
    const double* row = cderi + (long)P * npair;
    const int a0 = blockIdx.x * 64, a1 = (a0 + 64 < nao) ? a0 + 64 : nao;
    for (int b = threadIdx.x; b < a1; b += 256) { emax_s[b] = EXP_NONE; ss[b] = 0.0f; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int a = a0 + warp; a < a1; a += 8) {
        const double* x = row + (long)a * (a + 1) / 2;
        int em = EXP_NONE;
        float sq = 0.0f;
        for (int b = lane; b <= a; b += 32) {
            const double v = fabs(x[b]);
            if (v > 0.0) {
                const int e = frexp_exp(v);
                em = e > em ? e : em;
                if (e > emax_s[b]) atomicMax(&emax_s[b], e);
                const float v2 = (float)(v * v);
                sq += v2;
                if (rownorm2 && b < a) atomicAdd(&ss[b], v2);      // the element also belongs to row b of the symmetric matrix
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const int t = __shfl_xor_sync(0xffffffffu, em, o); em = t > em ? t : em;
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
        if (lane == 0 && em != EXP_NONE) {
            atomicMax(&rowexp[(long)P * nao + a], em);
            if (rownorm2) atomicAdd(&rownorm2[(long)P * nao + a], sq);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < a1; b += 256)
        if (emax_s[b] != EXP_NONE) {
            atomicMax(&rowexp[(long)P * nao + b], emax_s[b]);
            if (rownorm2 && ss[b] > 0.0f) atomicAdd(&rownorm2[(long)P * nao + b], ss[b]);
        }
}

// Exponent bound of the rows of Y = A_P C~ over a block of nr auxiliary rows, before the GEMM (Cauchy-Schwarz):
//   |Y[nu,(P,i)]| <= ||A_P[nu,:]||_2 * max_i ||C~[:,i]||_2      Ey[nu] = exponent of 1.001 * max_P ... (|y| < 2^Ey)
// cmax2: device scalar, max_i sum_mu C~[mu,i]^2 (colnorm_max_kernel).  One thread per nu.
__global__ void yexp_bound_kernel(const float* __restrict__ rownorm2, int nr, int nao, const double* __restrict__ cmax2, int* __restrict__ Ey)
{
    const int nu = blockIdx.x * blockDim.x + threadIdx.x;
    if (nu >= nao) return;
    float m = 0.0f;
    for (int P = 0; P < nr; P++) m = fmaxf(m, rownorm2[(long)P * nao + nu]);
    const double bound = 1.001 * sqrt((double)m * 1.0001 * cmax2[0]);
    Ey[nu] = bound > 0.0 ? frexp_exp(bound) : 0;
}
// cmax2[0] = max over the rows i of X[nrows][k] (row stride ldx) of sum_k X[i][k]^2; one warp per row; cmax2 zeroed by the caller
__global__ void __launch_bounds__(256) colnorm_max_kernel(const double* __restrict__ X, long ldx, int nrows, int k, unsigned long long* __restrict__ cmax2)
{
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= nrows) return;
    const double* x = X + (long)r * ldx;
    double sq = 0.0;
    for (int c = lane; c < k; c += 32) sq += x[c] * x[c];
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0 && sq > 0.0) atomicMax(cmax2, (unsigned long long)__double_as_longlong(sq));
}

constexpr int PT = 64;   // tile edge of split_packed_kernel
// grid (Kp/PT, Kp/PT, nr), CTAs with tile column > tile row leave at once.  CTA (ta, tb, P): loads the 64 x 64 tile
// A_P[ta*64 .., tb*64 ..] from the packed row (coalesced: 64 consecutive doubles per a), writes its slices to the rows
// (P, a) at columns b and — for off-diagonal tiles — the slices of the transposed tile to the rows (P, b) at columns a,
// four int8 per 32-bit store.  Columns nao..Kp-1 are written as zeros; pad ROWS of the stack are the caller's (memset).
// NS7: the slice count is the compile-time 7 (constant shift amounts, digits 3..6 from the low word, 0..1 from the high word)
template <bool NS7>
__global__ void __launch_bounds__(256) split_packed_kernel(const double* __restrict__ cderi, long npair, int nao,
                                                           const int* __restrict__ rowexp, int ns, int Rp, int Kp, int out_row0,
                                                           int8_t* __restrict__ out, int* __restrict__ E)
{
    const int ta = blockIdx.x, tb = blockIdx.y, P = blockIdx.z;
    if (tb > ta) return;
    __shared__ double S[PT][PT + 1];
    const double* row = cderi + (long)P * npair;
    const int* ex = rowexp + (long)P * nao;
    const int t = threadIdx.x;
#pragma unroll 4
    for (int i = 0; i < PT * PT / 256; i++) {
        const int idx = t + 256 * i, al = idx >> 6, bl = idx & 63;
        const int a = ta * PT + al, b = tb * PT + bl;
        double v = 0.0;
        if (a < nao && b < nao) { const int hi = a > b ? a : b, lo = a > b ? b : a; v = row[(long)hi * (hi + 1) / 2 + lo]; }
        S[al][bl] = v;
    }
    __syncthreads();
    const int c4 = (t & 15) * 4;     // four consecutive output columns per thread
    const long orow0 = (long)out_row0 + (long)P * nao;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1 && ta == tb) break;
        const int trow = pass ? tb : ta, tcol = pass ? ta : tb;    // output rows / columns of this pass
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            const int rl = (t >> 4) + 16 * i;
            const int r = trow * PT + rl;
            if (r >= nao) continue;
            int e = ex[r];
            if (e == EXP_NONE) e = 0;
            if (tcol == 0 && c4 == 0) E[orow0 + r] = e;
            int8_t* dst = out + (orow0 + r) * Kp + tcol * PT + c4;
            if constexpr (NS7) {
                // N = rint(x 2^(48-e)), |N| <= 2^48, as (hi, lo) words of the mantissa of x*scale + 1.5*2^52 (offset 2^51 removed):
                // digit s sits at bit 7(6-s): s = 3..6 in lo[0,28), s = 2 across the words, s = 1 at hi[3,10), s = 0 = hi >> 10 (signed)
                const double scN = pow2i(48 - e);
                unsigned lo[4]; int hi[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double t_ = fma(pass ? S[c4 + j][rl] : S[rl][c4 + j], scN, 6755399441055744.0);
                    lo[j] = (unsigned)__double2loint(t_);
                    hi[j] = (__double2hiint(t_) & 0x000FFFFF) - 0x00080000;     // remove exponent bits and the 2^51 offset
                }
                const long sstride = (long)Rp * Kp;
                unsigned pk[7];
#pragma unroll
                for (int s_ = 0; s_ < 7; s_++) pk[s_] = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const unsigned l_ = lo[j];
                    const int h_ = hi[j];
                    pk[6] |= (l_ & 127u) << (8 * j);
                    pk[5] |= ((l_ >> 7) & 127u) << (8 * j);
                    pk[4] |= ((l_ >> 14) & 127u) << (8 * j);
                    pk[3] |= ((l_ >> 21) & 127u) << (8 * j);
                    pk[2] |= (__funnelshift_r(l_, (unsigned)h_, 28) & 127u) << (8 * j);
                    pk[1] |= (((unsigned)h_ >> 3) & 127u) << (8 * j);
                    pk[0] |= ((unsigned)(h_ >> 10) & 255u) << (8 * j);
                }
#pragma unroll
                for (int s_ = 0; s_ < 7; s_++) *reinterpret_cast<unsigned*>(dst + s_ * sstride) = pk[s_];
            } else if (ns <= 7) {
                // One FP64 operation per element instead of four per digit: N = rint(x 2^(6-e+7(ns-1))) (|N| < 2^48) sits in the
                // mantissa of x*scale + 1.5*2^52; its digits come out with integer shifts: the top one signed (-64..64), the others
                // 0..127 (two's-complement style, no carries).  Still an exact representation; the digit products of stage 1 are
                // bounded by 127*64 instead of 64*64, which the int32 accumulation bound covers up to K = nao < 37 000.
                const double scN = pow2i(6 - e + 7 * (ns - 1));
                long long N[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double t_ = fma(pass ? S[c4 + j][rl] : S[rl][c4 + j], scN, 6755399441055744.0);
                    N[j] = (long long)(__double_as_longlong(t_) & 0x000FFFFFFFFFFFFFLL) - (1LL << 51);
                }
                for (int s_ = 0; s_ < ns; s_++) {
                    const int sh = 7 * (ns - 1 - s_);
                    unsigned pack = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int d_ = s_ == 0 ? (int)(N[j] >> sh) : (int)((N[j] >> sh) & 127);
                        pack |= (unsigned)(d_ & 0xff) << (8 * j);
                    }
                    *reinterpret_cast<unsigned*>(dst + (long)s_ * Rp * Kp) = pack;
                }
            } else {
                const double sc = pow2i(6 - e);
                double rr[4];
#pragma unroll
                for (int j = 0; j < 4; j++) rr[j] = (pass ? S[c4 + j][rl] : S[rl][c4 + j]) * sc;
                for (int s_ = 0; s_ < ns; s_++) {
                    unsigned pack = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const double qv = rint(rr[j]);
                        pack |= (unsigned)(unsigned char)(int8_t)(int)qv << (8 * j);
                        rr[j] = (rr[j] - qv) * 128.0;
                    }
                    *reinterpret_cast<unsigned*>(dst + (long)s_ * Rp * Kp) = pack;
                }
            }
        }
    }
}

}  // namespace i8g
}  // namespace b200jk
