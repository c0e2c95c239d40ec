// jk_block.cuh — the CTA-level procedure of the direct J/K kernels (one template per class).
// Compiles as a CUDA kernel body (threads = CUDA threads, B2_SYNC = __syncthreads) and, with
// B200JK_EMULATE, as a sequential SIMT emulation on the CPU (tests only).
#pragma once
#include "jk_core.cuh"

namespace b200jk {

constexpr int KCH_MAX = 512;  // ket pairs examined per CTA

struct KParams {
    const ShellPair* bra_pairs; int nbra;
    const ShellPair* ket_pairs; int nket;
    int same_class;
    const PrimPair* prims;
    RysTables tb;
    double omega, tol;
    const double* dmc; int nsh;
    const double* dmj; const double* dmk;
    double* vj; double* vk;
    int n, n_dm_j, n_dm_k;
    int kchunk;
    unsigned long long* counters;  // [0] quartets computed, [1] quartets screened out (may be null)
};

template <class C, int NQ>
struct BlockSmem {
    SlotSmem<C> slot[NQ];
    BraInfo bra;
    int klist[KCH_MAX];
    int nk;
    int npmax;
};

template <class C, int NQ>
struct BlockCfg {
    static constexpr int NT = ((NQ * C::G + 31) / 32) * 32;
};

#if defined(__CUDA_ARCH__)
#define B2_FOR_THREADS(tid) { const int tid = threadIdx.x;
#define B2_END_THREADS }
#define B2_SYNC() __syncthreads()
#define B2_CTX(tid) ctx
#define B2_DECL_CTX ThreadCtx<C> ctx;
#else
#define B2_FOR_THREADS(tid) for (int tid = 0; tid < BlockCfg<C, NQ>::NT; tid++) {
#define B2_END_THREADS }
#define B2_SYNC()
#define B2_CTX(tid) ctxs[tid]
#define B2_DECL_CTX ThreadCtx<C>* ctxs = new ThreadCtx<C>[BlockCfg<C, NQ>::NT];
#endif

template <class C, int NQ>
#ifdef __CUDACC__
__device__ __forceinline__
#else
inline
#endif
void jk_block(const KParams& P, int bx, int by, BlockSmem<C, NQ>& sm)
{
    B2_DECL_CTX
    const ShellPair& bpair = P.bra_pairs[bx];
    const int kmax = P.same_class ? (bx + 1) : P.nket;
    const int kbeg = by * P.kchunk;
    const int kend = (kbeg + P.kchunk < kmax) ? kbeg + P.kchunk : kmax;
    if (kbeg >= kend) {
#if !defined(__CUDA_ARCH__)
        delete[] ctxs;
#endif
        return;
    }

    B2_FOR_THREADS(tid)
        thread_decode<C>(B2_CTX(tid), tid);
        if (tid == 0) {
            sm.bra.ABx = bpair.ABx; sm.bra.ABy = bpair.ABy; sm.bra.ABz = bpair.ABz;
            sm.bra.i0 = bpair.i0; sm.bra.j0 = bpair.j0;
            sm.bra.nprim = bpair.nprim; sm.bra.prim_off = bpair.prim_off;
            sm.bra.same = bpair.same; sm.bra.idx = bx;
            sm.nk = 0;
        }
    B2_END_THREADS
    B2_SYNC();

    // ---- on-device screening: compact the surviving kets of this chunk
    B2_FOR_THREADS(tid)
        for (int kk = kbeg + tid; kk < kend; kk += BlockCfg<C, NQ>::NT) {
            const ShellPair& kp = P.ket_pairs[kk];
            bool keep = keep_quartet(bpair.q, kp.q, bpair.ish, bpair.jsh, kp.ish, kp.jsh, P.dmc, P.nsh, P.tol,
                                     P.vj != nullptr, P.vk != nullptr);
            if (keep) {
#if defined(__CUDA_ARCH__)
                int pos = atomicAdd(&sm.nk, 1);
#else
                int pos = sm.nk++;
#endif
                sm.klist[pos] = kk;
            }
        }
    B2_END_THREADS
    B2_SYNC();
    const int nk = sm.nk;
#if defined(__CUDA_ARCH__)
    if (P.counters && threadIdx.x == 0) {
        atomicAdd(&P.counters[0], (unsigned long long)nk);
        atomicAdd(&P.counters[1], (unsigned long long)(kend - kbeg - nk));
    }
#else
    if (P.counters) { P.counters[0] += nk; P.counters[1] += kend - kbeg - nk; }
#endif

    for (int base = 0; base < nk; base += NQ) {
        // ---- per-batch slot setup
        B2_FOR_THREADS(tid)
            ThreadCtx<C>& t = B2_CTX(tid);
            if (t.q < NQ) {
                SlotSmem<C>& s = sm.slot[t.q];
                if (t.g == 0) {
                    int e = base + t.q;
                    s.active = (e < nk);
                    if (s.active) {
                        int kk = sm.klist[e];
                        const ShellPair& kp = P.ket_pairs[kk];
                        s.kl = kk; s.k0 = kp.i0; s.l0 = kp.j0;
                        s.nprim_k = kp.nprim; s.prim_off_k = kp.prim_off;
                        double f = 1.0;
                        if (sm.bra.same) f *= 0.5;
                        if (kp.same) f *= 0.5;
                        if (P.same_class && kk == bx) f *= 0.5;
                        s.fac = f;
                        slot_set_cd<C>(s, kp.ABx, kp.ABy, kp.ABz);
                    } else {
                        s.nprim_k = 0;
                    }
                }
                B2_UNROLL
                for (int e = 0; e < C::NV; e++) t.v[e] = 0.0;
            }
            if (tid == 0) sm.npmax = 0;
        B2_END_THREADS
        B2_SYNC();
        B2_FOR_THREADS(tid)
            if (tid == 0) {
                int m = 0;
                for (int q = 0; q < NQ; q++) m = (sm.slot[q].nprim_k > m) ? sm.slot[q].nprim_k : m;
                sm.npmax = m * sm.bra.nprim;
            }
        B2_END_THREADS
        B2_SYNC();
        const int npmax = sm.npmax;
        const int nbp = sm.bra.nprim;

        for (int ip = 0; ip < npmax; ip++) {
            // ---- phase A: Rys roots
            B2_FOR_THREADS(tid)
                ThreadCtx<C>& t = B2_CTX(tid);
                if (t.q < NQ) {
                    SlotSmem<C>& s = sm.slot[t.q];
                    if (s.active) {
                        int nkp = s.nprim_k;
                        bool on = ip < nbp * nkp;
                        if (on) {
                            int ibp = ip / nkp, ikp = ip - ibp * nkp;
                            phase_roots<C>(s, t.g, P.prims[sm.bra.prim_off + ibp], P.prims[s.prim_off_k + ikp], P.tb,
                                           P.omega);
                        }
                    }
                }
            B2_END_THREADS
            B2_SYNC();
            // ---- phase B: vertical recurrences into shared memory
            B2_FOR_THREADS(tid)
                ThreadCtx<C>& t = B2_CTX(tid);
                if (t.q < NQ) {
                    SlotSmem<C>& s = sm.slot[t.q];
                    if (s.active && ip < nbp * s.nprim_k) phase_vrr<C>(s, t.g);
                }
            B2_END_THREADS
            B2_SYNC();
            // ---- phase D: horizontal recurrences + root sum in registers
            B2_FOR_THREADS(tid)
                ThreadCtx<C>& t = B2_CTX(tid);
                if (t.q < NQ) {
                    SlotSmem<C>& s = sm.slot[t.q];
                    if (s.active && ip < nbp * s.nprim_k) phase_accumulate<C>(s, t, sm.bra.ABx, sm.bra.ABy, sm.bra.ABz);
                }
            B2_END_THREADS
        }
        // ---- phase E: digestion
        B2_FOR_THREADS(tid)
            ThreadCtx<C>& t = B2_CTX(tid);
            if (t.q < NQ) {
                SlotSmem<C>& s = sm.slot[t.q];
                if (s.active) {
                    if (P.vj) phase_digest<C>(s, t, sm.bra.i0, sm.bra.j0, P.n, P.n_dm_j, P.dmj, nullptr, P.vj, nullptr);
                    if (P.vk) phase_digest<C>(s, t, sm.bra.i0, sm.bra.j0, P.n, P.n_dm_k, nullptr, P.dmk, nullptr, P.vk);
                }
            }
        B2_END_THREADS
        B2_SYNC();
    }
#if !defined(__CUDA_ARCH__)
    delete[] ctxs;
#endif
}

}  // namespace b200jk
