// rtrb_trace_strict.cu — RTRB_PREC_STRICT instantiation of the trace kernels.
// MUST be compiled with -fmad=false: see the arithmetic contract in rtrb_trace.cuh.
#include "rtrb_launch.h"
#include "rtrb_trace.cuh"

namespace {

constexpr int kBlock = 128;

template <int MAXS, bool DETAIL>
__global__ void __launch_bounds__(kBlock) trace_pre_strict_kernel(const __grid_constant__ FrameParams P) {
  rtrb::trace_pre_body<MAXS, DETAIL>(P);
}
template <int MAXS, bool DETAIL>
__global__ void __launch_bounds__(kBlock) trace_extra_strict_kernel(const __grid_constant__ FrameParams P) {
  rtrb::trace_extra_body<MAXS, DETAIL>(P);
}

template <int MAXS, bool DETAIL>
__global__ void __launch_bounds__(kBlock) trace_mt_strict_kernel(const __grid_constant__ FrameParams P) {
  rtrb::trace_mt_body<MAXS, DETAIL>(P);
}
template <int MAXS, bool DETAIL>
cudaError_t launch_mt(const FrameParams& P, cudaStream_t s) {
  unsigned long long total = (unsigned long long)P.n_tiles * RTRB_SUPER_PIXELS;  // one thread per PIXEL
  if (total == 0) return cudaSuccess;
  unsigned long long blocks = (total + kBlock - 1) / kBlock;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  trace_mt_strict_kernel<MAXS, DETAIL><<<(unsigned)blocks, kBlock, 0, s>>>(P);
  return cudaGetLastError();
}

template <int MAXS, bool DETAIL>
cudaError_t launch_pre(const FrameParams& P, cudaStream_t s) {
  unsigned long long total = (unsigned long long)P.n_tiles * RTRB_SUPER_PIXELS * (unsigned long long)P.pre;
  if (total == 0) return cudaSuccess;
  unsigned long long blocks = (total + kBlock - 1) / kBlock;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  trace_pre_strict_kernel<MAXS, DETAIL><<<(unsigned)blocks, kBlock, 0, s>>>(P);
  return cudaGetLastError();
}
template <int MAXS, bool DETAIL>
cudaError_t launch_extra(const FrameParams& P, cudaStream_t s) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  trace_extra_strict_kernel<MAXS, DETAIL><<<sms * 8, kBlock, 0, s>>>(P);
  return cudaGetLastError();
}

}  // namespace

int rtrb_max_stack_supported(void) { return 128; }

#define RTRB_DISPATCH(fn, P, need, s)                                         \
  do {                                                                        \
    const bool det = (P).count_detail != 0;                                   \
    if ((need) <= 10) return det ? fn<10, true>(P, s) : fn<10, false>(P, s);  \
    if ((need) <= 32) return det ? fn<32, true>(P, s) : fn<32, false>(P, s);  \
    return det ? fn<128, true>(P, s) : fn<128, false>(P, s);                  \
  } while (0)

cudaError_t rtrb_launch_trace_pre_strict(const FrameParams& P, int stack_need, cudaStream_t s) {
  RTRB_DISPATCH(launch_pre, P, stack_need, s);
}
cudaError_t rtrb_launch_trace_extra_strict(const FrameParams& P, int stack_need, cudaStream_t s) {
  RTRB_DISPATCH(launch_extra, P, stack_need, s);
}
cudaError_t rtrb_launch_trace_mt_strict(const FrameParams& P, int stack_need, cudaStream_t s) {
  RTRB_DISPATCH(launch_mt, P, stack_need, s);
}
