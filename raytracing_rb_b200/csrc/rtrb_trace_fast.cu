// rtrb_trace_fast.cu — RTRB_PREC_FAST64 dispatch (FP32 filter + exact FP64 refine).  The kernels are instantiated
// per work-stack capacity in rtrb_trace_fast_{d1,t10,t32,t128}.cu and, for the scene / frame classes of
// rtrb_trace_fast.cuh, once more in the *_l1 / *_l1n / d1lean translation units (all compiled -fmad=false: the exact
// parts must round like STRICT; the filter uses explicit fmaf()).
#include "rtrb_launch.h"

#define RTRB_DECLARE_NS(ns)                                       \
  namespace ns {                                                  \
  cudaError_t pre_d1(const FrameParams& P, cudaStream_t s);       \
  cudaError_t extra_d1(const FrameParams& P, cudaStream_t s);     \
  cudaError_t pre_t10(const FrameParams& P, cudaStream_t s);      \
  cudaError_t extra_t10(const FrameParams& P, cudaStream_t s);    \
  cudaError_t pre_t32(const FrameParams& P, cudaStream_t s);      \
  cudaError_t extra_t32(const FrameParams& P, cudaStream_t s);    \
  cudaError_t pre_t128(const FrameParams& P, cudaStream_t s);     \
  cudaError_t extra_t128(const FrameParams& P, cudaStream_t s);   \
  }
RTRB_DECLARE_NS(rtrb_fast)       // generic: d1, t10, t32, t128
RTRB_DECLARE_NS(rtrb_fast_lean)  // d1 only
RTRB_DECLARE_NS(rtrb_fast_l1)    // t10, t32
RTRB_DECLARE_NS(rtrb_fast_l1n)   // t10, t32

// A persistent-THREAD variant (single lanes refetch a new sample when their stack empties) was measured in round 1
// and rejected: it mixed unrelated rays into one warp (22.9 instead of 29.1 of 32 lanes active on config 3).  Round 2
// measured warp-granular refill as well (profiles/README.md): also rejected.

#define RTRB_FAST_DISPATCH(fn)                                                                              \
  const bool one_light = (P.scene_class & RTRB_SCENE_CLASS_ONE_LIGHT) != 0;                                 \
  if (P.trace_depth <= 1)                                                                                   \
    return (P.scene_class & RTRB_SCENE_CLASS_LEAN) ? rtrb_fast_lean::fn##_d1(P, s) : rtrb_fast::fn##_d1(P, s); \
  if (stack_need <= 10)                                                                                     \
    return !one_light ? rtrb_fast::fn##_t10(P, s) : (P.mc == 0 ? rtrb_fast_l1n::fn##_t10(P, s) : rtrb_fast_l1::fn##_t10(P, s)); \
  if (stack_need <= 32)                                                                                     \
    return !one_light ? rtrb_fast::fn##_t32(P, s) : (P.mc == 0 ? rtrb_fast_l1n::fn##_t32(P, s) : rtrb_fast_l1::fn##_t32(P, s)); \
  return rtrb_fast::fn##_t128(P, s);

cudaError_t rtrb_launch_trace_pre_fast(const FrameParams& P, int stack_need, cudaStream_t s) { RTRB_FAST_DISPATCH(pre) }
cudaError_t rtrb_launch_trace_extra_fast(const FrameParams& P, int stack_need, cudaStream_t s) { RTRB_FAST_DISPATCH(extra) }
