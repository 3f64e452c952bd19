// rtrb_trace_fast.cu — RTRB_PREC_FAST64 instantiation (FP32 filter + exact FP64 refine).
// Compiled with -fmad=false: the exact parts must round like STRICT; the filter uses explicit fmaf().
#include <stdlib.h>

#include "rtrb_launch.h"
#include "rtrb_trace_fast.cuh"

namespace {

// Launch shape per kernel family (16 warps per SM at 128 registers either way; measured on B200, profiles/README.md):
//   depth-1 kernels (MAXS == 1): 128 threads x 4 CTAs per SM  (256 x 2 is 5 % slower on config 2)
//   ray-tree kernels (MAXS > 1): lockstep item loop (rtrb_trace.cuh), so the CTA is the unit that shares the
//                                instruction caches: 512 threads x 1 CTA per SM on frames of more than a few waves
//                                (config 4: 4.44 / 3.78 / 3.57 ms and config 5: 5.13 / 4.61 / 4.24 ms with 128 / 256 /
//                                512 threads; config 3: 3.87 / 3.44 / 3.52), 128 threads on small frames.
#ifndef RTRB_FAST_BLOCK
#define RTRB_FAST_BLOCK 128
#endif
#ifndef RTRB_FAST_MIN_BLOCKS
#define RTRB_FAST_MIN_BLOCKS 4
#endif
#ifndef RTRB_TREE_BLOCK
#define RTRB_TREE_BLOCK 512
#endif
#ifndef RTRB_TREE_MIN_BLOCKS
#define RTRB_TREE_MIN_BLOCKS 1
#endif
template <int MAXS> constexpr int block_of() { return MAXS == 1 ? RTRB_FAST_BLOCK : RTRB_TREE_BLOCK; }
template <int MAXS> constexpr int min_blocks_of() { return MAXS == 1 ? RTRB_FAST_MIN_BLOCKS : RTRB_TREE_MIN_BLOCKS; }

template <int MAXS, bool DETAIL, bool BVH>
__global__ void __launch_bounds__(block_of<MAXS>(), min_blocks_of<MAXS>()) trace_pre_fast_kernel(const __grid_constant__ FrameParams P) {
  rtrb::trace_pre_body<MAXS, DETAIL, BVH ? 2 : 1>(P);
}
template <int MAXS, bool DETAIL, bool BVH>
__global__ void __launch_bounds__(block_of<MAXS>(), min_blocks_of<MAXS>()) trace_extra_fast_kernel(const __grid_constant__ FrameParams P) {
  rtrb::trace_extra_body<MAXS, DETAIL, BVH ? 2 : 1>(P);
}

// A persistent-thread variant (lanes refetch a new sample when their stack empties) was measured in
// round 1 and REJECTED: on config 3 the flat grid already runs at 29.1/32 active threads per
// instruction because neighbouring samples have similar ray trees; refetching mixed unrelated rays
// into one warp (22.9/32) and ran 1.8x slower (profiles/README.md).

template <int MAXS, bool DETAIL>
cudaError_t launch_pre(const FrameParams& P, cudaStream_t s) {
  unsigned long long total = (unsigned long long)P.n_tiles * RTRB_SUPER_PIXELS * (unsigned long long)P.pre;
  if (total == 0) return cudaSuccess;
  int kBlock = block_of<MAXS>();
  if (kBlock > 128 && total < 4ull * 148ull * 512ull) kBlock = 128;  // small frames: more, smaller CTAs
  if (MAXS > 1) {  // development override: RTRB_TREE_BLOCK_RT=<threads> (must not exceed the compiled launch bound)
    static const int env_block = getenv("RTRB_TREE_BLOCK_RT") ? atoi(getenv("RTRB_TREE_BLOCK_RT")) : 0;
    if (env_block >= 32 && env_block <= block_of<MAXS>() && env_block % 32 == 0) kBlock = env_block;
  }
  unsigned long long blocks = (total + kBlock - 1) / kBlock;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  if (P.use_bvh) trace_pre_fast_kernel<MAXS, DETAIL, true><<<(unsigned)blocks, kBlock, 0, s>>>(P);
  else trace_pre_fast_kernel<MAXS, DETAIL, false><<<(unsigned)blocks, kBlock, 0, s>>>(P);
  return cudaGetLastError();
}
template <int MAXS, bool DETAIL>
cudaError_t launch_extra(const FrameParams& P, cudaStream_t s) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  constexpr int kBlock = 128;
  if (P.use_bvh) trace_extra_fast_kernel<MAXS, DETAIL, true><<<sms * 8, kBlock, 0, s>>>(P);
  else trace_extra_fast_kernel<MAXS, DETAIL, false><<<sms * 8, kBlock, 0, s>>>(P);
  return cudaGetLastError();
}

}  // namespace

#define RTRB_DISPATCH(fn, P, need, s)                                         \
  do {                                                                        \
    const bool det = (P).count_detail != 0;                                   \
    if ((P).trace_depth <= 1) return det ? fn<1, true>(P, s) : fn<1, false>(P, s); \
    if ((need) <= 10) return det ? fn<10, true>(P, s) : fn<10, false>(P, s);  \
    if ((need) <= 32) return det ? fn<32, true>(P, s) : fn<32, false>(P, s);  \
    return det ? fn<128, true>(P, s) : fn<128, false>(P, s);                  \
  } while (0)

cudaError_t rtrb_launch_trace_pre_fast(const FrameParams& P, int stack_need, cudaStream_t s) {
  RTRB_DISPATCH(launch_pre, P, stack_need, s);
}
cudaError_t rtrb_launch_trace_extra_fast(const FrameParams& P, int stack_need, cudaStream_t s) {
  RTRB_DISPATCH(launch_extra, P, stack_need, s);
}
