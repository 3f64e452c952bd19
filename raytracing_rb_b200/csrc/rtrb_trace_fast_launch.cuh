// rtrb_trace_fast_launch.cuh — kernels and launchers of the FAST64 trace path, instantiated per stack capacity
// in separate translation units (rtrb_trace_fast_d1.cu, _t10.cu, _t32.cu, _t128.cu) so they compile in parallel.
#pragma once
#include "rtrb_launch.h"
#include "rtrb_trace_fast.cuh"

#ifndef RTRB_FAST_NS
#define RTRB_FAST_NS rtrb_fast  // the lean-scene build of the depth-1 kernels lives in rtrb_fast_lean
#endif
namespace RTRB_FAST_NS {

// Launch shape per kernel family (16 warps per SM at 128 registers either way; measured on B200, profiles/README.md):
//   depth-1 kernels (MAXS == 1): 128 threads x 5 CTAs per SM = 96 registers (round 2, config 2: 0.101 ms at 4 CTAs /
//                                128 registers, 0.095 - 0.098 at 5 .. 8 CTAs / 96 .. 64 registers; 256 x 2: 0.103).
//                                The extra-sample kernels keep 4 CTAs (config 1: 0.707 vs 0.735 ms)
//   ray-tree kernels (MAXS > 1): lockstep item loop, so the CTA is the unit that shares the instruction caches:
//                                512 threads x 1 CTA per SM on frames of more than a few waves (config 4: 4.44 / 3.78 /
//                                3.57 ms and config 5: 5.13 / 4.61 / 4.24 ms with 128 / 256 / 512 threads; config 3:
//                                3.87 / 3.44 / 3.52), 128 threads on small frames
#ifndef RTRB_TREE_BLOCK
#define RTRB_TREE_BLOCK 512
#endif
#ifndef RTRB_FAST_MIN_BLOCKS
#ifdef RTRB_SCENE_LEAN
#define RTRB_FAST_MIN_BLOCKS 8  // lean depth-1 kernel (124 registers unconstrained): 0.086 / 0.080 / 0.078 / 0.076 / 0.078 / 0.088 ms
#else                           // on config 2 at 4 / 5 / 7 / 8 / 10 / 12 CTAs per SM (124 / 96 / 72 / 64 / 48 / 40 registers)
#define RTRB_FAST_MIN_BLOCKS 5
#endif
#endif
constexpr int kFastBlock = 128, kFastMinBlocks = RTRB_FAST_MIN_BLOCKS, kExtraMinBlocks = 4, kTreeBlock = RTRB_TREE_BLOCK, kTreeMinBlocks = 512 / RTRB_TREE_BLOCK > 0 ? 512 / RTRB_TREE_BLOCK : 1;

template <int MAXS, bool DETAIL, bool BVH>
__global__ void __launch_bounds__(kFastBlock, kFastMinBlocks) trace_pre_fast_kernel(const __grid_constant__ FrameParams P) {
  rtrb::trace_pre_body<MAXS, DETAIL, BVH ? 2 : 1>(P);
}
template <int MAXS, bool DETAIL, bool BVH>
__global__ void __launch_bounds__(kTreeBlock, kTreeMinBlocks) trace_pre_tree_kernel(const __grid_constant__ FrameParams P) {
  rtrb::trace_pre_tree_body<MAXS, BVH, DETAIL>(P);
}
template <int MAXS, bool DETAIL, bool BVH>
__global__ void __launch_bounds__(kFastBlock, kExtraMinBlocks) trace_extra_fast_kernel(const __grid_constant__ FrameParams P) {
  rtrb::trace_extra_body<MAXS, DETAIL, BVH ? 2 : 1>(P);
}

inline int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// depth-1 frames: one thread per (pixel, sample)
template <bool DETAIL>
cudaError_t launch_pre_d1(const FrameParams& P, cudaStream_t s) {
  const unsigned long long total = (unsigned long long)P.n_tiles * RTRB_SUPER_PIXELS * (unsigned long long)P.pre;
  if (total == 0) return cudaSuccess;
  const unsigned long long blocks = (total + kFastBlock - 1) / kFastBlock;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  if (P.use_bvh) trace_pre_fast_kernel<1, DETAIL, true><<<(unsigned)blocks, kFastBlock, 0, s>>>(P);
  else trace_pre_fast_kernel<1, DETAIL, false><<<(unsigned)blocks, kFastBlock, 0, s>>>(P);
  return cudaGetLastError();
}

// ray-tree frames: one thread per (pixel, sample), lockstep CTAs (rtrb_trace_fast.cuh, trace_pre_tree_body)
template <int MAXS, bool DETAIL>
cudaError_t launch_pre_tree(const FrameParams& P, cudaStream_t s) {
  const unsigned long long total = (unsigned long long)P.n_tiles * RTRB_SUPER_PIXELS * (unsigned long long)P.pre;
  if (total == 0) return cudaSuccess;
  int block = kTreeBlock;
  if (total < 4ull * (unsigned long long)sm_count() * kTreeBlock) block = 128;  // small frames: more, smaller CTAs
  const unsigned long long blocks = (total + block - 1) / block;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  // in-CTA resolve (fuse_resolve == 2): the host only selects it for sample counts that divide both CTA sizes
  if (P.fuse_resolve == 2 && (block % P.pre) != 0) return cudaErrorInvalidConfiguration;
  const size_t smem = P.fuse_resolve == 2 ? (size_t)block * 3u * sizeof(double) : 0u;
  if (P.use_bvh) trace_pre_tree_kernel<MAXS, DETAIL, true><<<(unsigned)blocks, block, smem, s>>>(P);
  else trace_pre_tree_kernel<MAXS, DETAIL, false><<<(unsigned)blocks, block, smem, s>>>(P);
  return cudaGetLastError();
}

template <int MAXS, bool DETAIL>
cudaError_t launch_extra(const FrameParams& P, cudaStream_t s) {
  const int sms = sm_count();
  if (P.use_bvh) trace_extra_fast_kernel<MAXS, DETAIL, true><<<sms * 8, kFastBlock, 0, s>>>(P);
  else trace_extra_fast_kernel<MAXS, DETAIL, false><<<sms * 8, kFastBlock, 0, s>>>(P);
  return cudaGetLastError();
}

// per-capacity entry points, one translation unit each
cudaError_t pre_d1(const FrameParams& P, cudaStream_t s);
cudaError_t extra_d1(const FrameParams& P, cudaStream_t s);
cudaError_t pre_t10(const FrameParams& P, cudaStream_t s);
cudaError_t extra_t10(const FrameParams& P, cudaStream_t s);
cudaError_t pre_t32(const FrameParams& P, cudaStream_t s);
cudaError_t extra_t32(const FrameParams& P, cudaStream_t s);
cudaError_t pre_t128(const FrameParams& P, cudaStream_t s);
cudaError_t extra_t128(const FrameParams& P, cudaStream_t s);

}  // namespace RTRB_FAST_NS

#define RTRB_FAST_TREE_TU(N)                                                                         \
  namespace RTRB_FAST_NS {                                                                              \
  cudaError_t pre_t##N(const FrameParams& P, cudaStream_t s) {                                       \
    return P.count_detail ? launch_pre_tree<N, true>(P, s) : launch_pre_tree<N, false>(P, s);        \
  }                                                                                                  \
  cudaError_t extra_t##N(const FrameParams& P, cudaStream_t s) {                                     \
    return P.count_detail ? launch_extra<N, true>(P, s) : launch_extra<N, false>(P, s);              \
  }                                                                                                  \
  }
