// rtrb_trace_fast.cu — RTRB_PREC_FAST64.  Placeholder routing: until the FP32-cull + exact-FP64
// refine kernels land, FAST64 runs the STRICT kernels (identical results by definition of the mode).
#include "rtrb_launch.h"

cudaError_t rtrb_launch_trace_pre_fast(const FrameParams& P, int stack_need, cudaStream_t s) {
  return rtrb_launch_trace_pre_strict(P, stack_need, s);
}
cudaError_t rtrb_launch_trace_extra_fast(const FrameParams& P, int stack_need, cudaStream_t s) {
  return rtrb_launch_trace_extra_strict(P, stack_need, s);
}
