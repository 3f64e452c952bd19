// FAST64 ray-tree kernels with a work stack of up to 32 items (trace_depth * (1 + mc) + 1 <= 32).  -fmad=false.
#include "rtrb_trace_fast_launch.cuh"
RTRB_FAST_TREE_TU(32)
