// df_classes.cuh — instantiation/dispatch of the (LI LJ | LK 1) three-center kernels.
// Bra classes: the 10 AO pair classes (s..f) plus (g,1) for the two-center metric; LK = 0..4 (s..g aux).
#pragma once
#include <stdexcept>
#include <string>
#include "df_block.cuh"
#include "jk_classes.cuh"

namespace b200jk {

template <int LI, int LJ, int LK>
struct J3cCfg {
    static constexpr int NP = choose_np(ncart(LI), ncart(LJ), ncart(LK));
    using C = QClass<LI, LJ, LK, 0, NP>;
    using GC = GroupCfg<C>;
};

#ifndef B200JK_EMULATE
template <class C, bool SR>
__global__ void __launch_bounds__(GroupCfg<C>::NT) j3c_kernel(const J3cParams P)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    BlockSmem<C>& sm = *reinterpret_cast<BlockSmem<C>*>(smraw);
    j3c_block<C, SR>(P, blockIdx.x, blockIdx.y, sm);
}
#endif

#ifndef B200JK_EMULATE
template <class C, bool SR>
void launch_j3c_kernel(const J3cParams& P, dim3 grid, int nt, size_t smem, b2_stream_t st)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(j3c_kernel<C, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(j3c): ") + cudaGetErrorString(e));
        configured = true;
    }
    j3c_kernel<C, SR><<<grid, nt, smem, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string("j3c_kernel launch: ") + cudaGetErrorString(e));
}
#endif

template <int LI, int LJ, int LK>
void launch_j3c_one(J3cParams P, b2_stream_t st)
{
    using C = typename J3cCfg<LI, LJ, LK>::C;
    using GC = GroupCfg<C>;
    P.kchunk = pick_kchunk(P.nbra, P.nket, GC::NSLOT, 4096);
    int ny = (P.nket + P.kchunk - 1) / P.kchunk;
#ifndef B200JK_EMULATE
    size_t smem = sizeof(BlockSmem<C>);
    dim3 grid(P.nbra, ny);
    if (P.omega < 0.0) launch_j3c_kernel<C, true>(P, grid, GC::NT, smem, st);
    else launch_j3c_kernel<C, false>(P, grid, GC::NT, smem, st);
#else
    (void)st;
    BlockSmem<C>* sm = new BlockSmem<C>();
    for (int bx = 0; bx < P.nbra; bx++)
        for (int by = 0; by < ny; by++) {
            if (P.omega < 0.0) j3c_block<C, true>(P, bx, by, *sm);
            else j3c_block<C, false>(P, bx, by, *sm);
        }
    delete sm;
#endif
}

// bra class ids 0..9 as in B2_PAIR_CASES, 10 = (g,1)
#define B2_J3C_BRA_CASES(X) B2_PAIR_CASES(X) X(10, 4, 0)

template <int LK>
void launch_j3c_lk(int cb, const J3cParams& P, b2_stream_t st)
{
    switch (cb) {
#define X(id, li, lj)                              \
    case id:                                       \
        launch_j3c_one<li, lj, LK>(P, st);         \
        return;
        B2_J3C_BRA_CASES(X)
#undef X
    }
    throw std::runtime_error("bad three-center bra class");
}

// one translation unit per auxiliary angular momentum (df_class_tu.cu with -DB2_LK=<l>)
void launch_j3c_lk0(int cb, const J3cParams& P, b2_stream_t st);
void launch_j3c_lk1(int cb, const J3cParams& P, b2_stream_t st);
void launch_j3c_lk2(int cb, const J3cParams& P, b2_stream_t st);
void launch_j3c_lk3(int cb, const J3cParams& P, b2_stream_t st);
void launch_j3c_lk4(int cb, const J3cParams& P, b2_stream_t st);

inline void launch_j3c(int cb, int lk, const J3cParams& P, b2_stream_t st)
{
    switch (lk) {
    case 0: launch_j3c_lk0(cb, P, st); return;
    case 1: launch_j3c_lk1(cb, P, st); return;
    case 2: launch_j3c_lk2(cb, P, st); return;
    case 3: launch_j3c_lk3(cb, P, st); return;
    case 4: launch_j3c_lk4(cb, P, st); return;
    }
    throw std::runtime_error("auxiliary angular momentum > g is not supported");
}

}  // namespace b200jk
