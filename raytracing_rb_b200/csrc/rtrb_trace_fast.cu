// rtrb_trace_fast.cu — RTRB_PREC_FAST64 dispatch (FP32 filter + exact FP64 refine).  The kernels are instantiated
// per work-stack capacity in rtrb_trace_fast_{d1,t10,t32,t128}.cu (all compiled -fmad=false: the exact parts must
// round like STRICT; the filter uses explicit fmaf()).
#include "rtrb_launch.h"

namespace rtrb_fast {
cudaError_t pre_d1(const FrameParams& P, cudaStream_t s);
cudaError_t extra_d1(const FrameParams& P, cudaStream_t s);
cudaError_t pre_t10(const FrameParams& P, cudaStream_t s);
cudaError_t extra_t10(const FrameParams& P, cudaStream_t s);
cudaError_t pre_t32(const FrameParams& P, cudaStream_t s);
cudaError_t extra_t32(const FrameParams& P, cudaStream_t s);
cudaError_t pre_t128(const FrameParams& P, cudaStream_t s);
cudaError_t extra_t128(const FrameParams& P, cudaStream_t s);
}  // namespace rtrb_fast
namespace rtrb_fast_lean {  // rtrb_trace_fast_d1lean.cu
cudaError_t pre_d1(const FrameParams& P, cudaStream_t s);
cudaError_t extra_d1(const FrameParams& P, cudaStream_t s);
}  // namespace rtrb_fast_lean

// A persistent-THREAD variant (single lanes refetch a new sample when their stack empties) was measured in round 1
// and rejected: it mixed unrelated rays into one warp (22.9 instead of 29.1 of 32 lanes active on config 3).  The
// ray-tree kernels refill at WARP granularity instead (rtrb_trace_fast.cuh, trace_pre_warp_body).

cudaError_t rtrb_launch_trace_pre_fast(const FrameParams& P, int stack_need, cudaStream_t s) {
  if (P.trace_depth <= 1) return P.lean_scene ? rtrb_fast_lean::pre_d1(P, s) : rtrb_fast::pre_d1(P, s);
  if (stack_need <= 10) return rtrb_fast::pre_t10(P, s);
  if (stack_need <= 32) return rtrb_fast::pre_t32(P, s);
  return rtrb_fast::pre_t128(P, s);
}
cudaError_t rtrb_launch_trace_extra_fast(const FrameParams& P, int stack_need, cudaStream_t s) {
  if (P.trace_depth <= 1) return P.lean_scene ? rtrb_fast_lean::extra_d1(P, s) : rtrb_fast::extra_d1(P, s);
  if (stack_need <= 10) return rtrb_fast::extra_t10(P, s);
  if (stack_need <= 32) return rtrb_fast::extra_t32(P, s);
  return rtrb_fast::extra_t128(P, s);
}
