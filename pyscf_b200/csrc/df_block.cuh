// df_block.cuh — three-center (ij|P) and two-center (P|Q) Coulomb integrals for density fitting.
//
// Replaces libcint int3c2e_sph / int2c2e_sph behind GTOnr3c_drv / GTOint2c
// (pyscf/lib/gto/fill_nr_3c.c:196, fill_int2c.c:36; called from pyscf/df/incore.py:149,194).
// Same Rys machinery as the 4-center kernels with the fourth centre replaced by the unit function:
// class (LI LJ | LK 0), ket "pair" = (auxiliary shell, 1).  A CTA owns one AO shell pair (ij) and a
// chunk of auxiliary shells; the auxiliary Cartesian component is the thread index, the (a,b) block
// lives in registers (jk_core.cuh phases A/B/D), and the epilogue stores the Cartesian block
//     out[(P0 + c) * row_stride + pair_offset + b*NI + a]
// which the transform kernels in df.cu turn into the reference layout cderi[naux, nao(nao+1)/2].
#pragma once
#include "jk_block.cuh"

namespace b200jk {

struct J3cParams {
    const ShellPair* bra_pairs; int nbra;     // AO shell pairs of one class (or (aux,1) pairs for (P|Q))
    const int64_t* bra_out_off;               // per bra pair: offset of its NI*NJ block inside a row
    const ShellPair* ket_shells; int nket;    // auxiliary shells of one angular momentum as (P,1) pairs
    const PrimPair* bra_prims; const PrimPair* ket_prims;
    RysTables tb;
    double omega;
    double* out; int64_t row_stride;          // out row = auxiliary Cartesian function (ket i0 + c)
    int64_t col0;                              // column offset of this batch of bra pairs inside the full row
    int kchunk;
};

template <class C, bool SR>
#ifdef __CUDACC__
__device__ __forceinline__
#else
inline
#endif
void j3c_block(const J3cParams& P, int bx, int by, BlockSmem<C>& sm)
{
    using GC = GroupCfg<C>;
    static_assert(C::LL == 0, "three-center classes have a unit fourth function");
    const ShellPair& bpair = P.bra_pairs[bx];
    const int kbeg = by * P.kchunk;
    const int kend = (kbeg + P.kchunk < P.nket) ? kbeg + P.kchunk : P.nket;
    if (kbeg >= kend) return;
    const int64_t boff = P.bra_out_off[bx] - P.col0;
#if defined(__CUDA_ARCH__)
    LaneCtx<C> ctx;
#else
    LaneCtx<C>* ctxs = new LaneCtx<C>[GC::NT];
#endif
    B2_ALL_THREADS(tid)
        LaneCtx<C>& L = B2_CTX(tid);
        L.grp = tid / GC::TG;
        L.lt = tid % GC::TG;
        if constexpr (GC::PPW) {
            int sl, g;
            L.valid = GC::decode(L.lt, sl, g);
            L.slot = L.grp * GC::QPG + sl;
            thread_decode<C>(L.t, g);
        } else {
            int sl = L.lt / GC::GP, g = L.lt % GC::GP;
            L.valid = (sl < GC::QPG) && (g < C::G);
            L.slot = L.grp * GC::QPG + (sl < GC::QPG ? sl : 0);
            thread_decode<C>(L.t, g < C::G ? g : 0);
        }
        L.t.q = L.slot;
        if (tid == 0) {
            sm.bra.ABx = bpair.ABx; sm.bra.ABy = bpair.ABy; sm.bra.ABz = bpair.ABz;
            sm.bra.nprim = bpair.nprim; sm.bra.prim_off = bpair.prim_off;
            sm.next = kbeg;
        }
    B2_END
    B2_SYNC();
    const int nbp = bpair.nprim;

#if defined(__CUDA_ARCH__)
    const int grp = threadIdx.x / GC::TG;
    {
#else
    for (int grp = 0; grp < GC::NG; grp++) {
#endif
        const int tid0 = grp * GC::TG;
        (void)tid0;
        for (;;) {
            B2_GROUP_LANES(lt)
                if (lt == 0) {
#if defined(__CUDA_ARCH__)
                    sm.gbase[grp] = atomicAdd(&sm.next, GC::QPG);
#else
                    sm.gbase[grp] = sm.next; sm.next += GC::QPG;
#endif
                }
            B2_END
            group_sync<C>(grp);
            const int base = sm.gbase[grp];
            if (base >= kend) break;
            B2_GROUP_LANES(lt)
                LaneCtx<C>& L = B2_CTX(tid0 + lt);
                if (L.valid) {
                    SlotSmem<C>& s = sm.slot[L.slot];
                    if (L.t.g == 0) {
                        int kk = base + (L.slot - grp * GC::QPG);
                        s.active = (kk < kend);
                        if (s.active) {
                            const ShellPair& kp = P.ket_shells[kk];
                            s.kl = kk; s.k0 = kp.i0;
                            s.nprim_k = kp.nprim; s.prim_off_k = kp.prim_off;
                            slot_set_cd<C>(s, 0.0, 0.0, 0.0);
                        } else {
                            s.nprim_k = 0;
                        }
                    }
                    B2_UNROLL
                    for (int e = 0; e < C::NV; e++) L.t.v[e] = 0.0;
                    L.ibp = 0; L.ikp = 0; L.sr = 0;
                }
            B2_END
            group_sync<C>(grp);
            int npmax = 0;
            for (int q = 0; q < GC::QPG; q++) {
                int nk_ = sm.slot[grp * GC::QPG + q].nprim_k;
                npmax = nk_ > npmax ? nk_ : npmax;
            }
            npmax *= nbp;
            constexpr bool sr_op = SR;   // omega < 0: erfc = Coulomb - erf, two root sets per primitive pair
            if (sr_op) npmax *= 2;
            for (int ip = 0; ip < npmax; ip++) {
                B2_GROUP_LANES(lt)
                    LaneCtx<C>& L = B2_CTX(tid0 + lt);
                    if (L.valid) {
                        SlotSmem<C>& s = sm.slot[L.slot];
                        if (s.active && L.ibp < nbp)
                            phase_roots<C>(s, L.t.g, P.bra_prims[sm.bra.prim_off + L.ibp], P.ket_prims[s.prim_off_k + L.ikp], P.tb,
                                           sr_op ? (L.sr ? -P.omega : 0.0) : P.omega, (sr_op && L.sr) ? -1.0 : 1.0);
                    }
                B2_END
                group_sync<C>(grp);
                B2_GROUP_LANES(lt)
                    LaneCtx<C>& L = B2_CTX(tid0 + lt);
                    if (L.valid) {
                        SlotSmem<C>& s = sm.slot[L.slot];
                        if (s.active && L.ibp < nbp) phase_vrr<C>(s, L.t.g);
                    }
                B2_END
                group_sync<C>(grp);
                B2_GROUP_LANES(lt)
                    LaneCtx<C>& L = B2_CTX(tid0 + lt);
                    if (L.valid) {
                        SlotSmem<C>& s = sm.slot[L.slot];
                        if (s.active && L.ibp < nbp) {
                            phase_accumulate<C>(s, L.t, sm.bra.ABx, sm.bra.ABy, sm.bra.ABz);
                            if (sr_op && !L.sr) L.sr = 1;
                            else { L.sr = 0; if (++L.ikp == s.nprim_k) { L.ikp = 0; L.ibp++; } }
                        }
                    }
                B2_END
            }
            // ---- epilogue: store the Cartesian block
            B2_GROUP_LANES(lt)
                LaneCtx<C>& L = B2_CTX(tid0 + lt);
                if (L.valid) {
                    SlotSmem<C>& s = sm.slot[L.slot];
                    if (s.active) {
                        double* row = P.out + (int64_t)(s.k0 + L.t.c) * P.row_stride + boff;
                        const int b0 = L.t.p * C::NJP;
                        B2_UNROLL
                        for (int bb = 0; bb < C::NJP; bb++) {
                            B2_UNROLL
                            for (int a = 0; a < C::NI; a++) row[(b0 + bb) * C::NI + a] = L.t.v[bb * C::NI + a];
                        }
                    }
                }
            B2_END
            group_sync<C>(grp);
        }
    }
#if !defined(__CUDA_ARCH__)
    delete[] ctxs;
#endif
}

}  // namespace b200jk
