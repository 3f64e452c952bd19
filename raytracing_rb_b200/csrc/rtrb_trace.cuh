// rtrb_trace.cuh — the ray-tree engine of raytracing_rb as device code (sm_100a).
//
// One thread owns one pixel-sample: it builds the thin-lens primary ray (camera.rb:129-151), then
// runs RayTracer#trace_sync (ray_tracer.rb:16-46) with the work stack as an explicit per-thread
// LIFO array in local memory, popping the LAST pushed item exactly like `@queue.pop` (:32), and
// accumulates the colours in emission order (rt_reduce, :292-298).
//
// STRICT arithmetic contract: this translation unit is compiled with -fmad=false and every
// expression keeps the reference's association order, every division and square root the
// reference performs is performed, so FP64 results equal the CPU oracle's bit for bit except where
// libm's transcendental functions (sin cos acos asin pow) round differently from CUDA's.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/rtrb_b200.h"
#include "rtrb_types.h"

namespace rtrb {

#define RTRB_EPSILON 1e-5                 // src/libs/algebra.rb:2
#define RTRB_PI 3.141592653589793         // Math::PI

struct d3 { double x, y, z; };

// libm's FP64 transcendentals are 100-500 SASS instructions each; inlined at every call site they push the
// ray-tree kernels past 160 KB of code and the warps stall on instruction fetch (ncu: no_instruction is the
// top stall of the depth-8 kernel).  One out-of-line copy each would keep more of the hot loop in the instruction
// cache; results are unchanged (same routines).  MEASURED (round 1): out-of-line copies made configs 3-5
// 4-5 % SLOWER (4.38 -> 4.62 ms on config 3), so the inlined form stays the default;
// -DRTRB_OUTLINE_LIBM selects the out-of-line form.
#ifndef RTRB_OUTLINE_LIBM
#define RTRB_LIBM __device__ __forceinline__
#else
#define RTRB_LIBM static __device__ __noinline__
#endif
RTRB_LIBM double m_sin(double x) { return sin(x); }
RTRB_LIBM double m_cos(double x) { return cos(x); }
RTRB_LIBM double m_asin(double x) { return asin(x); }
RTRB_LIBM double m_acos(double x) { return acos(x); }
RTRB_LIBM double m_pow(double x, double y) { return pow(x, y); }
RTRB_LIBM double m_fmod(double x, double y) { return fmod(x, y); }

__device__ __forceinline__ d3 mk(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator-(d3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ d3 operator*(d3 a, d3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ d3 operator*(d3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ d3 operator/(d3 a, double s) { return mk(a.x / s, a.y / s, a.z / s); }
// Vec3_method_dot (fast_4d_matrix.c:98-108): ((a0*b0) + a1*b1) + a2*b2
__device__ __forceinline__ double dot(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// sum of squares as Vec3_c_create computes it (c:67)
__device__ __forceinline__ double sumsq(d3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
__device__ __forceinline__ double norm(d3 a) { return sqrt(sumsq(a)); }
__device__ __forceinline__ d3 cross(d3 a, d3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

struct ThreadCtx {
  uint32_t status;
  // lean counters (always)
  uint32_t rays, shadow;
  // detailed counters (count_detail only)
  uint32_t c[RTRB_CNT_N];
  bool detail;
  uint32_t max_stack;
};

#define RTRB_COUNT(ctx, which) do { if ((ctx).detail) (ctx).c[which]++; } while (0)

// Vec3#normalize (c:286-293)
__device__ __forceinline__ d3 normalize(d3 a, ThreadCtx& ctx) {
  double r = norm(a);
  if (r == 0) ctx.status |= RTRB_ST_ZERO_VECTOR;
  return mk(a.x / r, a.y / r, a.z / r);
}
// Vec3#cos (c:109-129): |cos| clamped to <= 1
__device__ __forceinline__ double vcos(d3 a, d3 b, ThreadCtx& ctx) {
  double ret = dot(a, b);
  double r1 = sumsq(a), r2 = sumsq(b);
  if (r1 == 0 || r2 == 0) ctx.status |= RTRB_ST_ZERO_VECTOR;
  double v = sqrt(ret * ret / r1 / r2);
  if (v > 1) v = 1;
  return v;
}
__device__ __forceinline__ double rb_sqrt(double x, ThreadCtx& ctx) {
  if (x < 0) ctx.status |= RTRB_ST_MATH_DOMAIN;
  return sqrt(x);
}
__device__ __forceinline__ double rb_acos(double x, ThreadCtx& ctx) {
  if (x < -1 || x > 1) ctx.status |= RTRB_ST_MATH_DOMAIN;
  return m_acos(x);
}
// Float ** exponent (world.rb:76).  pow(x, 2.0) is the correctly rounded x*x.
__device__ __forceinline__ double rb_pow(double x, double y) { return (y == 2.0) ? x * x : m_pow(x, y); }

// ---- counter RNG: Philox4x32-10 keyed by (seed; pixel, sample, ray path, purpose) ---------------
__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                              uint32_t& c3) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
}
__device__ __forceinline__ double res53(uint32_t a, uint32_t b) {
  a >>= 5; b >>= 6;
  return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

// ---- object tests ------------------------------------------------------------------------------
struct HitRec {
  d3 p;        // intersection
  bool dir_in; // :in / :out
  uint8_t face;  // box only: Box#intersect's data[:index] (box.rb:89)
};

// Sphere#intersect (sphere.rb:60-85).  d_r = |d|, dn = d / d_r are ray invariants.
__device__ __forceinline__ bool sphere_intersect(const DevGeom& g, d3 o, d3 d, double d_r, d3 dn, HitRec& h) {
  d3 c = mk(g.px, g.py, g.pz);
  d3 oc = c - o;
  double t = dot(oc, d) / (d_r * d_r);  // r2 = r*r (c:280-284)
  d3 np = o + d * t;
  d3 ncv = np - c;
  double nd = norm(ncv);
  if (!(nd <= g.radius)) return false;  // unless inner?(nearest_point)
  double hh = sqrt(g.radius * g.radius - nd * nd);
  d3 vec = dn * hh;
  // from_inner = |O - C| <= R (|O - C| == |C - O| bit for bit)
#ifdef RTRB_FAST_TU
  // FAST64 translation units: sqrt is monotone and correctly rounded, so away from the surface the comparison of the
  // squares decides (1e-14 >> 2^-52 covers the roundings of R*R and of the sum); inside that band, and for NaNs,
  // negative or overflowing radii, the reference's own expression is evaluated.
  const double oc2 = sumsq(oc), rr2 = g.radius * g.radius;
  bool from_inner;
  if (g.radius >= 0 && rr2 < 1e300 && oc2 < rr2 * (1.0 - 1e-14)) from_inner = true;
  else if (g.radius >= 0 && rr2 < 1e300 && oc2 > rr2 * (1.0 + 1e-14)) from_inner = false;
  else from_inner = sqrt(oc2) <= g.radius;
#else
  const bool from_inner = norm(oc) <= g.radius;
#endif
  h.dir_in = !from_inner;
  h.p = h.dir_in ? (np - vec) : (np + vec);
  if (!from_inner && t < 0) return false;
  return true;
}
// Plane#intersect (plane.rb:38-51)
__device__ __forceinline__ bool plane_intersect(const DevGeom& g, d3 o, d3 d, HitRec& h, double& den_out) {
  d3 n = mk(g.nx, g.ny, g.nz);
  double den = dot(n, d);
  den_out = den;
  if (den == 0) return false;
  double t = dot(mk(g.px, g.py, g.pz) - o, n) / den;
  h.p = o + d * t;
  if (t < 0) return false;
  h.dir_in = den < 0;
  return true;
}

// Box#intersect (box.rb:80-99): Plane#intersect on each of the six faces in order, kept when Plane#get_uv
// of the hit lies in [-0.5, 0.5]^2, nearest by |hit - origin| with a strict `<` (lowest face wins ties).
// Returns the winning face or -1.  Not inlined and register-only interface: boxes are rare, and the
// hot sphere/plane kernels must not pay registers or local memory for this loop.
static __device__ __noinline__ int box_nearest_face(const DevBox* bx, double ox, double oy, double oz, double dx,
                                                    double dy, double dz) {
  const d3 o = mk(ox, oy, oz), d = mk(dx, dy, dz);
  double nearest = CUDART_INF;
  int face = -1;
#pragma unroll 1
  for (int k = 0; k < 6; ++k) {
    const DevBoxFace& F = bx->f[k];
    const d3 n = mk(F.nx, F.ny, F.nz), pt = mk(F.px, F.py, F.pz);
    const double den = dot(n, d);
    if (den == 0) continue;                    // plane.rb:41
    const double t = dot(pt - o, n) / den;
    const d3 ip = o + d * t;
    if (t < 0) continue;                       // plane.rb:46
    const d3 rel = ip - pt;
    const double u = dot(rel, mk(F.lx, F.ly, F.lz)) / F.u_unit;  // plane.rb:82
    const double v = dot(rel, mk(F.ux, F.uy, F.uz)) / F.v_unit;  // plane.rb:83
    if (-0.5 <= u && u <= 0.5 && -0.5 <= v && v <= 0.5) {
      const double dd = norm(ip - o);
      if (dd < nearest) { nearest = dd; face = k; }
    }
  }
  return face;
}
__device__ __forceinline__ bool box_intersect(const DevBox& bx, d3 o, d3 d, HitRec& h) {
  const int face = box_nearest_face(&bx, o.x, o.y, o.z, d.x, d.y, d.z);
  if (face < 0) return false;
  // the winning face's Plane#intersect again: same expressions, same bits
  const DevBoxFace& F = bx.f[face];
  const d3 n = mk(F.nx, F.ny, F.nz);
  const double den = dot(n, d);
  const double t = dot(mk(F.px, F.py, F.pz) - o, n) / den;
  h.p = o + d * t;
  h.dir_in = den < 0;
  h.face = (uint8_t)face;
  return true;
}

// Probe ray of cover_area (world_object.rb:42): from `target` towards the light, unnormalised.
struct CoverRay {
  d3 target, lp, lt, ltn, tl;
  double lt_r;
};
__device__ __forceinline__ CoverRay make_cover_ray(d3 target, const DevLight& L) {
  CoverRay c;
  c.target = target;
  c.lp = mk(L.px, L.py, L.pz);
  c.lt = c.lp - target;
  c.lt_r = norm(c.lt);
  c.ltn = mk(c.lt.x / c.lt_r, c.lt.y / c.lt_r, c.lt.z / c.lt_r);
  c.tl = target - c.lp;
  return c;
}

// cover_area of ONE object, exactly as the reference evaluates it:
// Sphere#cover_area (sphere.rb:28-57) / WorldObject#cover_area (world_object.rb:41-49).
// BOX = false compiles the box branch out (kernels specialised for sphere/plane scenes).
// KIND tells what the caller already knows about g.type, so that the other branches are not compiled into its loop
// (the sphere branch with its penumbra code is 800 instructions): 0 = anything (the reference's own scan), 1 = a sphere
// or a box (an entry of the sphere filter), 2 = a plane.
template <bool BOX = true, int KIND = 0>
__device__ __forceinline__ double cover_object_exact(const FrameParams& P, const DevGeom& g, const CoverRay& c,
                                                     double light_radius, ThreadCtx& ctx) {
  if (KIND != 2 && g.type == RTRB_OBJ_SPHERE) {
    RTRB_COUNT(ctx, RTRB_CNT_COV_SPH);
    HitRec h;
    int factor = 0;
    if (sphere_intersect(g, c.target, c.lt, c.lt_r, c.ltn, h) && dot(h.p - c.lp, c.tl) > 0) factor = 1;
    d3 cc = mk(g.px, g.py, g.pz);
    double t = dot(cc - c.target, c.lt) / (c.lt_r * c.lt_r);
    d3 x1 = c.target + c.lt * t;
    double r1 = light_radius * (norm(x1 - c.target) / c.lt_r);
    double dd = norm(x1 - cc);
    if (dd >= r1 + g.radius) return 0;
    double s1 = RTRB_PI * r1 * r1;
#ifdef RTRB_SCENE_LEAN
    // light_radius == 0.0, so r1 is a zero or a NaN (0 * inf).  Past the test above dd < R (or a comparison with a NaN
    // failed); then `dd > |R - r1|` = `dd > |R|` cannot hold (R >= 0: dd < R; R < 0: no dd >= 0 is < R; NaN: false), so
    // the partial-overlap branch (sphere.rb:37-47) is unreachable and its acos / sin / divisions are not compiled in
    if (false) {
#else
    if (dd > fabs(g.radius - r1)) {
#endif
      RTRB_COUNT(ctx, RTRB_CNT_COV_SPH_PEN);
      double ct1 = fmin((r1 * r1 + dd * dd - g.radius * g.radius) / (2 * r1 * dd), 1.0);
      double ct2 = fmin((g.radius * g.radius + dd * dd - r1 * r1) / (2 * g.radius * dd), 1.0);
      double th1 = rb_acos(ct1, ctx), th2 = rb_acos(ct2, ctx);
      double delta_s = ((th1 - m_sin(th1)) * r1 * r1 + (th2 - m_sin(th2)) * g.radius * g.radius) / 2;
      return factor * delta_s / s1;
    }
    RTRB_COUNT(ctx, RTRB_CNT_COV_SPH_FULL);
    if (r1 > g.radius) return factor * RTRB_PI * g.radius * g.radius / s1;
    return factor;
  }
  if constexpr (BOX && KIND != 2) {
    if (g.type == RTRB_OBJ_BOX) {  // WorldObject#cover_area (world_object.rb:41-49) through Box#intersect
      RTRB_COUNT(ctx, RTRB_CNT_COV_BOX);
      HitRec h;
      if (box_intersect(P.boxes[g.aux], c.target, c.lt, h) && dot(h.p - c.lp, c.tl) > 0) {
        RTRB_COUNT(ctx, RTRB_CNT_COV_BOX_ACC);
        return 1;
      }
      return 0;
    }
  }
  if constexpr (KIND == 1) return 0;  // (unreachable: the sphere filter only lists spheres and boxes)
  RTRB_COUNT(ctx, RTRB_CNT_COV_PL);
  HitRec h;
  double den;
  if (plane_intersect(g, c.target, c.lt, h, den) && dot(h.p - c.lp, c.tl) > 0) {
    RTRB_COUNT(ctx, RTRB_CNT_COV_PL_ACC);
    return 1;
  }
  return 0;
}

// World#lit_area (world.rb:62-69) for one light seen from `target`:
// max(1 - sum over ALL objects of cover_area, 0), subtraction in world_objects order.
template <bool BOX = true>
static __device__ __noinline__ double lit_area(const FrameParams& P, d3 target, const DevLight& L, ThreadCtx& ctx) {
  const CoverRay c = make_cover_ray(target, L);
  double total = 1;
  for (int i = 0; i < P.n_objects; ++i) {
    const DevGeom g = P.geom[i];
    total -= cover_object_exact<BOX>(P, g, c, L.radius, ctx);
  }
  return fmax(total, 0.0);
}

// get_a_random_vertical_vector (world_object.rb:105-120)
__device__ __forceinline__ d3 a_vertical_vector(d3 n, ThreadCtx& ctx) {
  if (norm(n) == 0) ctx.status |= RTRB_ST_ZERO_VECTOR;
  if (n.x == 0) {
    if (n.y == 0) return mk(1.0, 0.0, 0.0);
    return mk(0.0, -n.z / n.y, 1.0);
  }
  return mk(-(n.y + n.z) / n.x, 1.0, 1.0);
}

// Float#to_i then Integer#% (texture.rb:24-25): truncation toward zero, then FLOORED modulo by a positive size.
// `t` is already truncated.  Values that fit 32 bits take the integer path (the remainder of an integer division is
// exact either way, so the result equals fmod's: -2.6 % instructions on config 3); anything larger goes through fmod.
__device__ __forceinline__ int trunc_mod(double t, int n) {
  if (fabs(t) < 2147483648.0) {
    int m = (int)t % n;
    return m < 0 ? m + n : m;
  }
  double m = m_fmod(t, (double)n);
  if (m < 0) m += n;
  return (int)m;
}

// Texture#color (texture.rb:23-28)
__device__ __forceinline__ d3 texture_color(const DevMat& m, double uu, double vv, ThreadCtx& ctx) {
  double fu = (uu + m.uoff) / m.hscale, fv = (vv + m.voff) / m.vscale;
  if (!isfinite(fu) || !isfinite(fv)) { ctx.status |= RTRB_ST_NAN_TO_INT; return mk(0, 0, 0); }
  const int u = trunc_mod(trunc(fu), m.tex_w), v = trunc_mod(trunc(fv), m.tex_h);
  const uint8_t* p = m.tex + ((size_t)v * m.tex_w + u) * 3;
  return mk(__ldg(p) / 256.0, __ldg(p + 1) / 256.0, __ldg(p + 2) / 256.0);
}

// RTRB_RNG_MT: a thread's position in the host-generated MT19937 stream (Random.rand, camera.rb:135 and
// world_object.rb:84, consumed in program order).
struct MtCursor {
  const double* s;
  uint32_t pos, len;
  bool overflow;
  __device__ __forceinline__ double next() {
    double v = 0.0;
    if (pos < len) v = s[pos]; else overflow = true;
    pos++;
    return v;
  }
};

struct StackItem {
  double ox, oy, oz, dx, dy, dz, ax, ay, az;
  int32_t depth;
  uint32_t path;
};

// RayTracer#trace_sync for one sample.  Returns the summed colour; *primary_hit gets the root ray's
// World#intersect winner (-1 miss, -2 highlight-terminated).
template <int MAXS, bool MT = false>
__device__ __forceinline__ d3 trace_sample(const FrameParams& P, d3 ro, d3 rd, uint32_t pixel, uint32_t sample,
                                           ThreadCtx& ctx, int* primary_hit, MtCursor* mt = nullptr) {
  StackItem stack[MAXS];
  int sp = 0;
  stack[0].ox = ro.x; stack[0].oy = ro.y; stack[0].oz = ro.z;
  stack[0].dx = rd.x; stack[0].dy = rd.y; stack[0].dz = rd.z;
  stack[0].ax = 1.0; stack[0].ay = 1.0; stack[0].az = 1.0;
  stack[0].depth = P.trace_depth; stack[0].path = 1u;
  sp = 1;
  d3 sum = mk(0.0, 0.0, 0.0);
  bool first = true;
  const uint32_t K = (uint32_t)(P.mc + 2);
  *primary_hit = -1;

  while (sp > 0) {
    if ((uint32_t)sp > ctx.max_stack) ctx.max_stack = (uint32_t)sp;
    const StackItem it = stack[--sp];
    const bool is_first = first;
    first = false;
    d3 o = mk(it.ox, it.oy, it.oz), d = mk(it.dx, it.dy, it.dz), att = mk(it.ax, it.ay, it.az);
    // rt_map :52
    if (it.depth <= 0 || norm(att) < 0.0001) continue;
    ctx.rays++;

    // ---- World#high_lights (world.rb:83-98); occluders never block (lit_area is truthy) ----
    unsigned long long hl_mask = 0ull;
    int hl_n = 0;
    for (int l = 0; l < P.n_lights; ++l) {
      const DevLight& L = P.lights[l];
      d3 a = mk(L.px, L.py, L.pz) - o;
      double ct = vcos(d, a, ctx);
      double ang = rb_acos(ct, ctx);
      if (ang < L.hl_threshold) { hl_mask |= 1ull << l; hl_n++; }
    }
    if (hl_n > 0) {
      for (int l = 0; l < P.n_lights; ++l) {
        if (!((hl_mask >> l) & 1ull)) continue;
        d3 c = (att * ld3(P.lights[l].color_hl)) / (double)hl_n;  // ray_tracer.rb:65
        sum = sum + c;
        if (sum.x > 1 || sum.y > 1 || sum.z > 1) ctx.status |= RTRB_ST_COLOR_GT_1;
      }
      RTRB_COUNT(ctx, RTRB_CNT_HIGHLIGHT);
      if (is_first) *primary_hit = -2;
      continue;
    }

    // ---- World#intersect (world.rb:37-59): linear scan, strict `<`, lowest index wins ties ----
    const double d_r = norm(d);
    const d3 dn = mk(d.x / d_r, d.y / d_r, d.z / d_r);
    double best = P.max_distance;
    int best_i = -1;
    HitRec bh; bh.p = mk(0, 0, 0); bh.dir_in = false; bh.face = 0;
    for (int i = 0; i < P.n_objects; ++i) {
      const DevGeom g = P.geom[i];
      HitRec h;
      bool ok;
      if (g.type == RTRB_OBJ_SPHERE) {
        RTRB_COUNT(ctx, RTRB_CNT_SPH_TEST);
        ok = sphere_intersect(g, o, d, d_r, dn, h);
        if (ok) RTRB_COUNT(ctx, RTRB_CNT_SPH_ACC);
      } else if (g.type == RTRB_OBJ_BOX) {
        RTRB_COUNT(ctx, RTRB_CNT_BOX_TEST);
        ok = box_intersect(P.boxes[g.aux], o, d, h);
        if (ok) RTRB_COUNT(ctx, RTRB_CNT_BOX_ACC);
      } else {
        RTRB_COUNT(ctx, RTRB_CNT_PL_TEST);
        double den;
        ok = plane_intersect(g, o, d, h, den);
        if (ok) RTRB_COUNT(ctx, RTRB_CNT_PL_ACC);
      }
      if (ok) {
        double new_dis = norm(o - h.p);  // Ray#distance (algebra.rb:10-12)
        if (new_dis < best) { best = new_dis; best_i = i; bh = h; }
      }
    }
    if (best_i < 0) continue;  // light_dead
    if (is_first) *primary_hit = best_i;
    RTRB_COUNT(ctx, RTRB_CNT_HITS);

    const DevGeom g = P.geom[best_i];
    const DevMat& M = P.mat[best_i];
    // ---- intersect_parameters (sphere.rb:88-101 / plane.rb:54-67) ----
    d3 n, delta;
    double rate;
    bool can_refract;
    if (g.type == RTRB_OBJ_SPHERE) {
      d3 c = mk(g.px, g.py, g.pz);
      delta = ((bh.p - c) * RTRB_EPSILON) * (bh.dir_in ? 1.0 : -1.0);  // sphere.rb:84
      n = bh.dir_in ? (bh.p - c) : (c - bh.p);
      rate = bh.dir_in ? M.refractive_rate : 1.0 / M.refractive_rate;
      can_refract = true;
    } else {
      // a plane, or the face of a box that was hit (Box#intersect_parameters delegates, box.rb:102-107)
      d3 f = mk(g.nx, g.ny, g.nz);
      if (g.type == RTRB_OBJ_BOX) { const DevBoxFace& F = P.boxes[g.aux].f[bh.face]; f = mk(F.nx, F.ny, F.nz); }
      double fd = dot(f, d);
      double nfd = -fd;
      double sgn = nfd > 0 ? 1.0 : (nfd < 0 ? -1.0 : 0.0);
      delta = (f * RTRB_EPSILON) * sgn;  // plane.rb:50
      n = fd > 0 ? -f : f;               // plane.rb:55
      rate = M.refractive_rate;
      can_refract = M.has_refraction != 0;
    }
    const d3 nn = normalize(n, ctx);
    // get_reflection_by_ray_and_n (world_object.rb:121-125)
    double cos_theta = vcos(d, -n, ctx);
    d3 refl_dir = normalize(nn * (2 * cos_theta * d_r) + d, ctx);
    d3 refl_org = bh.p + delta;
    // get_refraction_by_ray_and_n (world_object.rb:127-137)
    bool has_refr = false;
    d3 refr_dir = mk(0, 0, 0), refr_org = mk(0, 0, 0);
    if (can_refract) {
      double cc = vcos(d, n, ctx);
      double sin_i = rb_sqrt(1 - cc * cc, ctx);
      double sin_r = sin_i / rate;
      if (!(sin_r >= 1)) {
        if (sin_r < -1 || sin_r > 1) ctx.status |= RTRB_ST_MATH_DOMAIN;
        double r = m_asin(sin_r);
        refr_dir = nn * (-m_cos(r)) + normalize(refl_dir + d, ctx) * sin_r;
        refr_org = bh.p - nn * RTRB_EPSILON;
        has_refr = true;
      }
    }
    // push children: reflection first, refraction second (popped first)  ray_tracer.rb:87-121
    if (sp + 2 + P.mc > MAXS) { ctx.status |= RTRB_ST_STACK_OVERFLOW; continue; }
    {
      StackItem& s = stack[sp++];
      d3 a2 = att * ld3(M.refl);
      s.ox = refl_org.x; s.oy = refl_org.y; s.oz = refl_org.z;
      s.dx = refl_dir.x; s.dy = refl_dir.y; s.dz = refl_dir.z;
      s.ax = a2.x; s.ay = a2.y; s.az = a2.z;
      s.depth = it.depth - 1; s.path = it.path * K + 0u;
    }
    if (has_refr) {
      RTRB_COUNT(ctx, RTRB_CNT_REFR);
      StackItem& s = stack[sp++];
      d3 a2 = att * ld3(M.refr);
      s.ox = refr_org.x; s.oy = refr_org.y; s.oz = refr_org.z;
      s.dx = refr_dir.x; s.dy = refr_dir.y; s.dz = refr_dir.z;
      s.ax = a2.x; s.ay = a2.y; s.az = a2.z;
      s.depth = it.depth - 1; s.path = it.path * K + 1u;
    }

    // ---- World#local_lights (world.rb:72-80) fused with WorldObject#local_lighting (:51-74) ----
    const d3 shade_from = bh.p + delta;
    d3 contrib = mk(0.0, 0.0, 0.0);
    int n_lit = 0;
    for (int l = 0; l < P.n_lights; ++l) {
      const DevLight& L = P.lights[l];
      ctx.shadow++;
      double area = lit_area(P, shade_from, L, ctx);
      if (area > 0) {
        d3 lc = ld3(L.color) * (rb_pow(area, P.soft_shadow_exponent) / (double)P.n_lights);
        d3 lv = normalize(mk(L.px, L.py, L.pz) - bh.p, ctx);  // shading point is WITHOUT delta (:152)
        double ldn = dot(lv, nn);
        if (ldn > 1) ldn = 1.0; else if (ldn < 0) ldn = 0.0;
        contrib = contrib + lc * ldn;
        n_lit++;
      }
    }
    if (n_lit == 0) {
      // WorldObject#path_tracing (world_object.rb:76-90): mc rays, no local/ambient colour
      if (P.mc > 0) {
        d3 att_pt = ld3(M.diffuse) / (double)P.mc;
        d3 leftv = normalize(a_vertical_vector(n, ctx), ctx);
        d3 upv = cross(nn, leftv);
        for (int m = 0; m < P.mc; ++m) {
          uint32_t child = it.path * K + (uint32_t)(2 + m);
          double u_theta, u_phi;  // theta's draw first (world_object.rb:84)
          if constexpr (MT) {
            u_theta = mt->next(); u_phi = mt->next();
          } else {
            uint32_t c0 = pixel, c1 = sample, c2 = child, c3 = 1u;
            philox4x32_10(P.key0, P.key1, c0, c1, c2, c3);
            u_theta = res53(c0, c1); u_phi = res53(c2, c3);
          }
          double theta = u_theta * RTRB_PI / 2, phi = u_phi * RTRB_PI * 2;
          d3 dir = nn * m_sin(theta) + (leftv * m_cos(phi) + upv * m_sin(phi)) * m_cos(theta);
          RTRB_COUNT(ctx, RTRB_CNT_MC);
          StackItem& s = stack[sp++];
          d3 a2 = att * att_pt;
          s.ox = shade_from.x; s.oy = shade_from.y; s.oz = shade_from.z;
          s.dx = dir.x; s.dy = dir.y; s.dz = dir.z;
          s.ax = a2.x; s.ay = a2.y; s.az = a2.z;
          s.depth = it.depth - 1; s.path = child;
        }
      }
    } else {
      RTRB_COUNT(ctx, RTRB_CNT_LOCAL);
      if (ctx.detail) ctx.c[RTRB_CNT_LIT] += n_lit;
      contrib = contrib / (double)n_lit;  // world_object.rb:66-68
      d3 filter = mk(1.0, 1.0, 1.0);
      if (M.tex != nullptr) {
        double u, v;
        if (g.type == RTRB_OBJ_SPHERE) {  // Sphere#get_uv (sphere.rb:111-120)
          d3 vec = bh.p - mk(g.px, g.py, g.pz);
          double x = dot(vec, ld3(M.e0)) / g.radius;
          double y = dot(vec, ld3(M.e1)) / g.radius;
          double z = dot(vec, ld3(M.e2)) / g.radius;
          double mm = rb_sqrt(x * x + y * y + z * z + 2 * x + 1, ctx);
          u = (y / mm + 1) / 2;
          v = (-z / mm + 1) / 2;
        } else {  // Plane#get_uv (plane.rb:81-85)
          d3 rel = bh.p - mk(g.px, g.py, g.pz);
          u = dot(rel, ld3(M.e0)) / M.u_unit;
          v = dot(rel, ld3(M.e1)) / M.v_unit;
        }
        RTRB_COUNT(ctx, RTRB_CNT_TEXEL);
        filter = texture_color(M, u, v, ctx) * filter;
      }
      d3 local = ((contrib * ld3(M.diffuse)) * filter) + ld3(M.ambient);
      d3 c = att * local;  // ray_tracer.rb:152
      sum = sum + c;
      if (sum.x > 1 || sum.y > 1 || sum.z > 1) ctx.status |= RTRB_ST_COLOR_GT_1;
    }
  }
  return sum;
}

// Camera#lens_func (camera.rb:129-151) for pixel (x, y) and aperture angle theta.
__device__ __forceinline__ void lens_ray(const FrameParams& P, int x, int y, double theta, d3& ro, d3& rd) {
  d3 pos = ld3(P.pos), left = ld3(P.left), upn = ld3(P.up_n), front = ld3(P.front);
  // host-built tables of 2.0 * (x.to_f / width - 0.5) * retina_width and the y analogue (camera.rb:133-134)
  const double sx = P.lens_sx[x];
  const double sy = P.lens_sy[y];
  d3 retina_position = (ld3(P.retina_center) + left * sx) + upn * sy;
  // aperture_radius == 0.0 (pinhole): the reference still evaluates (left*cos + up*sin) * 0.0 = (+-0, +-0, +-0);
  // adding a zero of either sign to `pos` gives `pos` (up to the sign of an exact zero, which no later
  // expression can observe), so the two FP64 transcendentals are skipped.
  d3 rand_vector = mk(0.0, 0.0, 0.0);
  if (P.aperture_radius != 0.0) rand_vector = (ld3(P.left_n) * m_cos(theta) + upn * m_sin(theta)) * P.aperture_radius;
  d3 aperture = pos + rand_vector;
  d3 rf = pos - retina_position;  // ray retina -> lens centre
  double t = dot(ld3(P.focal_point) - retina_position, front) / dot(front, rf);  // intersect_plane :123-127
  d3 target = retina_position + rf * t;
  ro = aperture;
  rd = target - aperture;
}

// array_to_color (camera.rb:153-156) + the byte truncation of PNG::Color.new
__device__ __forceinline__ uint8_t quantise_u8(double c) {
  // [c*256, 255].min truncated to a byte; negatives and NaN give 0.  c*256 is an exact scaling, and the
  // saturating round-toward-zero conversion followed by an integer clamp is the same function.
  int v = __double2int_rz(c * 256.0);  // NaN -> 0, +-inf / out of range saturate
  v = v < 0 ? 0 : (v > 255 ? 255 : v);
  return (uint8_t)v;
}
// render_at's result for pixel (x, y): row = y, column = x (camera.rb:98,105)
__device__ __forceinline__ void write_pixel(const FrameParams& P, int x, int y, double r, double g, double b) {
  size_t px = (size_t)y * P.width + x;
  if (P.rgb) { P.rgb[px * 3 + 0] = r; P.rgb[px * 3 + 1] = g; P.rgb[px * 3 + 2] = b; }
  if (P.pixel_format == RTRB_FMT_PNG_RGB8) {
    // PNG scanline: filter-type byte 0, then the row's RGB8 pixels (the stream a PNG encoder deflates into IDAT)
    uint8_t* row = P.rgba + (size_t)y * ((size_t)P.width * 3 + 1);
    if (x == 0) row[0] = 0;
    uint8_t* o = row + 1 + (size_t)x * 3;
    o[0] = quantise_u8(r); o[1] = quantise_u8(g); o[2] = quantise_u8(b);
    return;
  }
  if (P.pixel_format == RTRB_FMT_RGB8) {
    // alpha is the constant 255 (camera.rb:155): three byte stores; a warp's 8x4 pixel block is four
    // 24-byte runs which the L2 merges before the frame leaves over PCIe / NVLink
    uint8_t* o = P.rgba + px * 3;
    o[0] = quantise_u8(r); o[1] = quantise_u8(g); o[2] = quantise_u8(b);
    return;
  }
  uchar4 q = make_uchar4(quantise_u8(r), quantise_u8(g), quantise_u8(b), 255);
  reinterpret_cast<uchar4*>(P.rgba)[px] = q;
}

// Decodes work item -> (tile slot k, in-tile q, x, y); returns false when the pixel is outside the window.
__device__ __forceinline__ bool decode_pixel(const FrameParams& P, uint32_t slot, int& x, int& y) {
  uint32_t k = slot / RTRB_SUPER_PIXELS, q = slot % RTRB_SUPER_PIXELS;
  uint32_t tx, ty;
  if (P.tiles_magic != 0u) {  // whole frame, one renderer: row-major tiles, no dependent load
    ty = __umulhi(k, P.tiles_magic);
    tx = k - ty * (uint32_t)P.stx_count;
  } else {
    const uint32_t tile = (uint32_t)P.tiles[k];  // tx | ty << 16
    tx = tile & 0xffffu; ty = tile >> 16;
  }
  int qx, qy;
  rtrb_morton_decode(q, &qx, &qy);
  x = (int)tx * RTRB_SUPER + qx;
  y = (int)ty * RTRB_SUPER + qy;
  return x >= P.x0 && x < P.x1 && y >= P.y0 && y < P.y1;
}

// Records a pixel whose sample set a status bit (the reference would have raised there): status word and the
// first offending pixel in the reference's pixel order (x outer, y inner).  Clears ctx.status so that a thread
// which goes on to other pixels (grid-stride / persistent loops) attributes later bits to the right pixel.
__device__ __forceinline__ void report_status(const FrameParams& P, ThreadCtx& ctx, int x, int y) {
  if (ctx.status) {
    atomicOr(&P.status[0], ctx.status);
    // smallest x*H + y wins; stored complemented (atomicMax) so the control block can be zero-initialised
    atomicMax(P.first_bad, ~((unsigned long long)x * (unsigned long long)P.height + (unsigned long long)y));
    ctx.status = 0;
  }
}

__device__ __forceinline__ void flush_ctx(const FrameParams& P, ThreadCtx& ctx, int x, int y, bool active) {
  // warp totals by hardware reduction (REDUX), then one fire-and-forget RED per warp and counter, spread over
  // RTRB_HOT_SLICES address pairs (the host sums the slices): no block barrier, and no single L2 line taking
  // 65 K atomics per frame
  const unsigned full = 0xffffffffu;
  const uint32_t rays = __reduce_add_sync(full, ctx.rays), shadow = __reduce_add_sync(full, ctx.shadow),
                 ms = __reduce_max_sync(full, ctx.max_stack);
  const int lane = threadIdx.x & 31;
  if (lane == 0) {
    unsigned long long* hot = P.hot + 2u * ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (RTRB_HOT_SLICES - 1));
    if (rays) atomicAdd(&hot[0], (unsigned long long)rays);
    if (shadow) atomicAdd(&hot[1], (unsigned long long)shadow);
    if (ms > 1u) atomicMax(&P.status[1], ms);  // 1 (just the root) is the host-side default
  }
  if (ctx.detail) {
#pragma unroll
    for (int i = 0; i < RTRB_CNT_N; ++i) {
      if (i == RTRB_CNT_RAYS || i == RTRB_CNT_SHADOW) continue;
      const uint32_t v = __reduce_add_sync(full, ctx.c[i]);
      if (lane == 0 && v) atomicAdd(&P.counters[i], (unsigned long long)v);
    }
  }
  if (active) report_status(P, ctx, x, y);
}

__device__ __forceinline__ void init_ctx(ThreadCtx& ctx, bool detail) {
  ctx.status = 0; ctx.rays = 0; ctx.shadow = 0; ctx.max_stack = 0; ctx.detail = detail;
  if (detail) {  // the lean kernels never touch c[]: no zeroing, no local-memory array
#pragma unroll
    for (int i = 0; i < RTRB_CNT_N; ++i) ctx.c[i] = 0;
  }
}

// Folds the counters / status a callee collected in its own block into the thread's.
__device__ __forceinline__ void merge_ctx(ThreadCtx& ctx, const ThreadCtx& t) {
  ctx.status |= t.status;
  ctx.rays += t.rays; ctx.shadow += t.shadow;
  if (t.max_stack > ctx.max_stack) ctx.max_stack = t.max_stack;
  if (ctx.detail) {
#pragma unroll
    for (int i = 0; i < RTRB_CNT_N; ++i) ctx.c[i] += t.c[i];
  }
}

// FAST64 entry point, defined in rtrb_trace_fast.cuh (only instantiated by that translation unit).
template <int MAXS, bool BVH, bool BOX>
__device__ __forceinline__ d3 trace_sample_fast(const FrameParams& P, d3 ro, d3 rd, uint32_t pixel, uint32_t sample,
                                                ThreadCtx& ctx, int* primary_hit);

// MODE: 0 = STRICT, 1 = FAST64 with the linear filter, 2 = FAST64 with the sphere BVH.
// BOX: the FAST64 kernels carry the Box code only in their full-counter (DETAIL) variants; scenes with a
// box are always dispatched there (rtrb_api.cu), so the lean hot kernels stay sphere/plane-only.
template <int MAXS, int MODE, bool BOX>
__device__ __forceinline__ d3 trace_dispatch(const FrameParams& P, d3 ro, d3 rd, uint32_t pixel, uint32_t sample,
                                             ThreadCtx& ctx, int* primary_hit) {
  if constexpr (MODE == 2) return trace_sample_fast<MAXS, true, BOX>(P, ro, rd, pixel, sample, ctx, primary_hit);
  else if constexpr (MODE == 1) return trace_sample_fast<MAXS, false, BOX>(P, ro, rd, pixel, sample, ctx, primary_hit);
  else return trace_sample<MAXS>(P, ro, rd, pixel, sample, ctx, primary_hit);
}

// One thread per (pixel, sample j < pre_sample_times): the first loop of render_at (camera.rb:73-78).
template <int MAXS, bool DETAIL, int MODE = 0>
__device__ __forceinline__ void trace_pre_body(const FrameParams& P) {
  const uint32_t S = (uint32_t)P.pre;
  const unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long total = (unsigned long long)P.n_tiles * RTRB_SUPER_PIXELS * S;
  ThreadCtx ctx;
  init_ctx(ctx, DETAIL);
  int x = 0, y = 0;
  bool active = false;
  if (w < total) {
    uint32_t slot, j;
    if (S == 1u) { slot = (uint32_t)w; j = 0u; }
    else { slot = (uint32_t)(w / S); j = (uint32_t)(w - (unsigned long long)slot * S); }
    active = decode_pixel(P, slot, x, y);
    if (active) {
      const uint32_t pixel = (uint32_t)y * (uint32_t)P.width + (uint32_t)x;
      double theta = 0.0;
      if (P.aperture_radius != 0.0) {
        uint32_t c0 = pixel, c1 = j, c2 = 0u, c3 = 0u;
        philox4x32_10(P.key0, P.key1, c0, c1, c2, c3);
        theta = res53(c0, c1);  // Random.rand, camera.rb:135
      }
      d3 ro, rd;
      lens_ray(P, x, y, theta, ro, rd);
      int ph;
      d3 col = trace_dispatch<MAXS, MODE, DETAIL>(P, ro, rd, pixel, j, ctx, &ph);
      RTRB_COUNT(ctx, RTRB_CNT_SAMPLES);
      if (P.fuse_resolve) {
        // one sample, positive threshold: mean = s / 1.0 = s and variance = 0 < threshold (camera.rb:80-87)
        write_pixel(P, x, y, col.x, col.y, col.z);
      } else {
        double* out = P.samples + w * 3ull;
        out[0] = col.x; out[1] = col.y; out[2] = col.z;
      }
      if (j == 0 && P.hit) P.hit[(size_t)y * P.width + x] = ph;
    }
  }
  flush_ctx(P, ctx, x, y, active);
}

// Extra samples j in [pre, max) of the pixels that failed the variance test (camera.rb:88-93);
// persistent grid-stride loop because the pixel count is only known on the device.
template <int MAXS, bool DETAIL, int MODE = 0>
__device__ __forceinline__ void trace_extra_body(const FrameParams& P) {
  const uint32_t E = (uint32_t)(P.max_samples - P.pre);
  const unsigned long long total = (unsigned long long)(*P.extra_count) * E;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  ThreadCtx ctx;
  init_ctx(ctx, DETAIL);
  int x = 0, y = 0;
  bool any = false;
  for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
    const uint32_t e = (uint32_t)(w / E), j = (uint32_t)P.pre + (uint32_t)(w % E);
    const uint32_t slot = P.extra_list[e];
    if (!decode_pixel(P, slot, x, y)) continue;
    any = true;
    const uint32_t pixel = (uint32_t)y * (uint32_t)P.width + (uint32_t)x;
    double theta = 0.0;
    if (P.aperture_radius != 0.0) {
      uint32_t c0 = pixel, c1 = j, c2 = 0u, c3 = 0u;
      philox4x32_10(P.key0, P.key1, c0, c1, c2, c3);
      theta = res53(c0, c1);
    }
    d3 ro, rd;
    lens_ray(P, x, y, theta, ro, rd);
    int ph;
    d3 col = trace_dispatch<MAXS, MODE, DETAIL>(P, ro, rd, pixel, j, ctx, &ph);
    RTRB_COUNT(ctx, RTRB_CNT_SAMPLES);
    double* out = P.extra_samples + w * 3ull;
    out[0] = col.x; out[1] = col.y; out[2] = col.z;
    report_status(P, ctx, x, y);  // this pixel's bits, before the loop moves to another pixel
  }
  flush_ctx(P, ctx, x, y, any);
}

// mean of the pre samples in order, then the variance test (camera.rb:72-87); `s` = the pixel's S sample colours
__device__ __forceinline__ void pre_mean_of(const double* s, const int S, double& ax, double& ay, double& az,
                                            double& variance) {
  ax = 0.0; ay = 0.0; az = 0.0;
  for (int j = 0; j < S; ++j) { ax += s[j * 3 + 0]; ay += s[j * 3 + 1]; az += s[j * 3 + 2]; }
  ax = ax / (double)S; ay = ay / (double)S; az = az / (double)S;
  variance = 0;
  for (int j = 0; j < S; ++j) {
    double dx = s[j * 3 + 0] - ax, dy = s[j * 3 + 1] - ay, dz = s[j * 3 + 2] - az;
    double m = fmax(dx, fmax(dy, dz));  // (sample - mean).to_a.max, signed
    variance += m * m;
  }
  variance /= (double)S;
}
__device__ __forceinline__ void pre_mean(const FrameParams& P, uint32_t slot, double& ax, double& ay, double& az,
                                         double& variance) {
  pre_mean_of(P.samples + (size_t)slot * P.pre * 3, P.pre, ax, ay, az, variance);
}

// RTRB_RNG_MT: Camera#render_at (camera.rb:70-99) for ONE pixel in ONE thread, because the reference's
// MT19937 draws are consumed in program order: every sample's lens draw (camera.rb:135, drawn even when the
// aperture is 0), its Monte-Carlo draws in LIFO ray order, then the next sample, then the adaptive samples.
// The thread reads its draws from mt_stream starting at mt_offset[pixel order] and reports how many it used;
// the host iterates offsets = prefix sums of the counts until they stop changing (rtrb_api.cu).
template <int MAXS, bool DETAIL>
__device__ __forceinline__ void trace_mt_body(const FrameParams& P) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  ThreadCtx ctx;
  init_ctx(ctx, DETAIL);
  int x = 0, y = 0;
  bool active = false;
  if (slot < (uint32_t)P.n_tiles * RTRB_SUPER_PIXELS) active = decode_pixel(P, slot, x, y);
  if (active) {
    const uint32_t pixel = (uint32_t)y * (uint32_t)P.width + (uint32_t)x;
    const uint32_t order = (uint32_t)(x - P.x0) * (uint32_t)(P.y1 - P.y0) + (uint32_t)(y - P.y0);
    MtCursor cur;
    cur.s = P.mt_stream; cur.pos = P.mt_offset[order]; cur.len = P.mt_len; cur.overflow = false;
    const uint32_t start = cur.pos;
    const int S = P.pre;
    double* smp = P.samples + (size_t)slot * S * 3;
    for (int j = 0; j < S; ++j) {
      const double theta = cur.next();
      d3 ro, rd;
      lens_ray(P, x, y, theta, ro, rd);
      int ph;
      d3 col = trace_sample<MAXS, true>(P, ro, rd, pixel, (uint32_t)j, ctx, &ph, &cur);
      RTRB_COUNT(ctx, RTRB_CNT_SAMPLES);
      smp[j * 3 + 0] = col.x; smp[j * 3 + 1] = col.y; smp[j * 3 + 2] = col.z;
      if (j == 0 && P.hit) P.hit[(size_t)y * P.width + x] = ph;
    }
    double ax, ay, az, variance;
    pre_mean(P, slot, ax, ay, az, variance);
    if (variance >= P.variant_threshold) {
      atomicAdd(&P.counters[RTRB_CNT_ADAPTIVE], 1ull);
      double cx = 0.0, cy = 0.0, cz = 0.0;
      for (int j = S; j < P.max_samples; ++j) {
        const double theta = cur.next();
        d3 ro, rd;
        lens_ray(P, x, y, theta, ro, rd);
        int ph;
        d3 col = trace_sample<MAXS, true>(P, ro, rd, pixel, (uint32_t)j, ctx, &ph, &cur);
        RTRB_COUNT(ctx, RTRB_CNT_SAMPLES);
        cx += col.x; cy += col.y; cz += col.z;
      }
      const double fp = (double)S, fm = (double)P.max_samples;
      ax = (ax * fp + cx) / fm; ay = (ay * fp + cy) / fm; az = (az * fp + cz) / fm;  // camera.rb:93
    }
    write_pixel(P, x, y, ax, ay, az);
    P.mt_count[order] = cur.overflow ? 0xFFFFFFFFu : cur.pos - start;
  }
  flush_ctx(P, ctx, x, y, active);
}

}  // namespace rtrb
