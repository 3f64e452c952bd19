// FAST64, trace_depth <= 1, LEAN scenes (FrameParams::lean_scene: one light of radius 0, no textures, exponent 2): the
// same source as rtrb_trace_fast_d1.cu compiled with RTRB_LEAN_SCENE, in namespace rtrb_fast_lean.  -fmad=false.
#define RTRB_LEAN_SCENE 1
#define RTRB_FAST_NS rtrb_fast_lean
#include "rtrb_trace_fast_d1.cu"
