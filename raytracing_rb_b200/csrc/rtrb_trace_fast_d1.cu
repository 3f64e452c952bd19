// FAST64, trace_depth <= 1: the stackless depth-1 kernels (config 2's kernel lives here).  -fmad=false.
#include "rtrb_trace_fast_launch.cuh"
namespace rtrb_fast {
cudaError_t pre_d1(const FrameParams& P, cudaStream_t s) {
  return P.count_detail ? launch_pre_d1<true>(P, s) : launch_pre_d1<false>(P, s);
}
cudaError_t extra_d1(const FrameParams& P, cudaStream_t s) {
  return P.count_detail ? launch_extra<1, true>(P, s) : launch_extra<1, false>(P, s);
}
}  // namespace rtrb_fast
