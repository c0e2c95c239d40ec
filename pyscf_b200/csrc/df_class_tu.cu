// df_class_tu.cu — compiled once per auxiliary angular momentum (-DB2_LK=0..4); without B2_LK: all.
#ifndef B200JK_EMULATE
#include <cuda_runtime.h>
#endif
#include "df_classes.cuh"

namespace b200jk {
#if !defined(B2_LK) || B2_LK == 0
void launch_j3c_lk0(int cb, const J3cParams& P, b2_stream_t st) { launch_j3c_lk<0>(cb, P, st); }
#endif
#if !defined(B2_LK) || B2_LK == 1
void launch_j3c_lk1(int cb, const J3cParams& P, b2_stream_t st) { launch_j3c_lk<1>(cb, P, st); }
#endif
#if !defined(B2_LK) || B2_LK == 2
void launch_j3c_lk2(int cb, const J3cParams& P, b2_stream_t st) { launch_j3c_lk<2>(cb, P, st); }
#endif
#if !defined(B2_LK) || B2_LK == 3
void launch_j3c_lk3(int cb, const J3cParams& P, b2_stream_t st) { launch_j3c_lk<3>(cb, P, st); }
#endif
#if !defined(B2_LK) || B2_LK == 4
void launch_j3c_lk4(int cb, const J3cParams& P, b2_stream_t st) { launch_j3c_lk<4>(cb, P, st); }
#endif
}  // namespace b200jk
