// rtrb_launch.h — host-callable launchers of the trace kernels.  Each arithmetic mode lives in its
// own translation unit because FMA contraction is a per-TU compiler flag:
//   rtrb_trace_strict.cu  (-fmad=false)  RTRB_PREC_STRICT
//   rtrb_trace_fast.cu    (-fmad=true)   RTRB_PREC_FAST64
#pragma once
#include <cuda_runtime.h>

#include "rtrb_types.h"

// stack_need = deepest work stack the frame can reach (trace_depth * (1 + mc) + 2).
cudaError_t rtrb_launch_trace_pre_strict(const FrameParams& P, int stack_need, cudaStream_t s);
cudaError_t rtrb_launch_trace_extra_strict(const FrameParams& P, int stack_need, cudaStream_t s);
cudaError_t rtrb_launch_trace_pre_fast(const FrameParams& P, int stack_need, cudaStream_t s);
cudaError_t rtrb_launch_trace_extra_fast(const FrameParams& P, int stack_need, cudaStream_t s);
cudaError_t rtrb_launch_trace_mt_strict(const FrameParams& P, int stack_need, cudaStream_t s);  // RTRB_RNG_MT
int rtrb_max_stack_supported(void);
