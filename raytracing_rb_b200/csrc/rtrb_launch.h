// rtrb_launch.h — host-callable launchers of the trace kernels.  Each arithmetic mode lives in its
// own translation units, all compiled -fmad=false (the exact FP64 parts must round like the reference's scalar code;
// the FP32 filter of FAST64 uses explicit fmaf()):
//   rtrb_trace_strict.cu                          RTRB_PREC_STRICT
//   rtrb_trace_fast.cu + rtrb_trace_fast_*.cu     RTRB_PREC_FAST64 (dispatch + one TU per work-stack capacity)
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
