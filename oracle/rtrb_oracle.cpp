// rtrb_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (FP64, scalar, literal) of raytracing_rb's per-pixel hot path, used only by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
// CHECKER and the timed CPU baseline.  Nothing under raytracing_rb_b200/ may link or call it.
//
// PARITY PINNING: the reference ships no renderer test, golden image or fixture
// (spec/fast_4d_matrix_spec.rb covers Vec3 only) and no Ruby interpreter exists in this image, so
//   * the Vec3 arithmetic below IS pinned: tests/test_vec3_reference.py runs the reference's own
//     ext/fast_4d_matrix/fast_4d_matrix.c (compiled unmodified against oracle/ruby_shim/ruby.h
//     into oracle/_ref/) against these functions and against the 19 RSpec known answers;
//   * the renderer above Vec3 is "parity unpinned": it is a line-by-line restatement checked by
//     the hand-derived known answers of SURVEY.md 8c (tests/test_oracle_kat.py, test_oracle_box_kat.py),
//     cross-checked bit for bit against a second, independently written restatement
//     (oracle/restate_py.py, tests/test_restatements_agree.py), and frozen in tests/golden/.
//
// Every function cites the reference lines it follows (paths relative to /root/reference).
// The style is deliberately literal: a Vec3 carries its cached norm exactly like the C ext, and
// expressions keep the reference's evaluation order so FP64 results are reproducible bit for bit.
// Compile WITHOUT fp contraction (-ffp-contract=off): the reference ext is built -mavx (no FMA).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../include/rtrb_b200.h"

namespace {

// ------------------------------------------------------------------------------------------------
// Status plumbing: the reference raises; we flag and carry on with IEEE semantics.
// ------------------------------------------------------------------------------------------------
struct Ctx;
static thread_local Ctx* g_ctx = nullptr;
static void flag(uint32_t bit);

// ------------------------------------------------------------------------------------------------
// Vec3 — ext/fast_4d_matrix/fast_4d_matrix.c:57-305
// ------------------------------------------------------------------------------------------------
struct V3 {
  double v[3];
  double r;  // cached norm, fast_4d_matrix.c:59,67
};

// Vec3_c_create, fast_4d_matrix.c:62-73
static inline V3 vmk(double x, double y, double z) {
  V3 a;
  a.v[0] = x; a.v[1] = y; a.v[2] = z;
  a.r = std::sqrt(x * x + y * y + z * z);
  return a;
}
// Vec3_method_dot, :98-108  (ret = 0; ret += ...)
static inline double vdot(const V3& a, const V3& b) {
  double ret = 0;
  ret += a.v[0] * b.v[0];
  ret += a.v[1] * b.v[1];
  ret += a.v[2] * b.v[2];
  return ret;
}
// Vec3_method_cos, :109-129 — returns |cos|, clamped to <= 1, raises on a zero vector
static inline double vcos(const V3& a, const V3& b) {
  double ret = 0, r1, r2;
  ret += a.v[0] * b.v[0];
  ret += a.v[1] * b.v[1];
  ret += a.v[2] * b.v[2];
  r1 = a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2];
  r2 = b.v[0] * b.v[0] + b.v[1] * b.v[1] + b.v[2] * b.v[2];
  if (r1 == 0 || r2 == 0) flag(RTRB_ST_ZERO_VECTOR);
  double v = std::sqrt(ret * ret / r1 / r2);
  if (v > 1) v = 1;
  return v;
}
// Vec3_method_cross, :131-141
static inline V3 vcross(const V3& a, const V3& b) {
  return vmk(a.v[1] * b.v[2] - a.v[2] * b.v[1],
             a.v[2] * b.v[0] - a.v[0] * b.v[2],
             a.v[0] * b.v[1] - a.v[1] * b.v[0]);
}
static inline V3 vneg(const V3& a) { return vmk(-a.v[0], -a.v[1], -a.v[2]); }               // :154-164
static inline V3 vadd(const V3& a, const V3& b) { return vmk(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }  // :166-176
static inline V3 vsub(const V3& a, const V3& b) { return vmk(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }  // :178-188
static inline V3 vmul(const V3& a, const V3& b) { return vmk(a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2]); }  // :200-206
static inline V3 vmul(const V3& a, double s) { return vmk(a.v[0] * s, a.v[1] * s, a.v[2] * s); }                     // :194-199
static inline V3 vdiv(const V3& a, double s) { return vmk(a.v[0] / s, a.v[1] / s, a.v[2] / s); }                     // :213-218
static inline double vr2(const V3& a) { return a.r * a.r; }                                                          // :280-284
// Vec3_method_normalize, :286-293
static inline V3 vnormalize(const V3& a) {
  double r = std::sqrt(a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2]);
  if (r == 0) flag(RTRB_ST_ZERO_VECTOR);
  return vmk(a.v[0] / r, a.v[1] / r, a.v[2] / r);
}
static inline V3 v3(const double* p) { return vmk(p[0], p[1], p[2]); }

// Ruby Math.sqrt / Math.acos / Math.asin raise Math::DomainError outside their domain.
static inline double rb_sqrt(double x) { if (x < 0) flag(RTRB_ST_MATH_DOMAIN); return std::sqrt(x); }
static inline double rb_acos(double x) { if (x < -1 || x > 1) flag(RTRB_ST_MATH_DOMAIN); return std::acos(x); }
static inline double rb_asin(double x) { if (x < -1 || x > 1) flag(RTRB_ST_MATH_DOMAIN); return std::asin(x); }
// Float ** Integer / Float ** Float -> libm pow() (Ruby rb_float_pow).  The reference's own platform
// (Dockerfile:1 ruby:2.3 = Debian jessie, glibc 2.19) has a correctly rounded pow, for which
// pow(x, 2.0) == x*x exactly.  This image's glibc 2.39 pow is only 0.52-ULP accurate and differs
// from x*x for ~0.09% of inputs (measured), so `** 2` is restated as the correctly rounded x*x.
static inline double rb_pow(double x, double y) { return (y == 2.0) ? x * x : std::pow(x, y); }

const double EPSILON = 1e-5;                       // src/libs/algebra.rb:2
const double RB_PI = 3.141592653589793;            // Math::PI

// Alex::Ray, src/libs/algebra.rb:3-12
struct Ray {
  V3 front, position;
  double distance(const V3& pos) const { return vsub(position, pos).r; }
};

// ------------------------------------------------------------------------------------------------
// RNG: counter mode (shared definition with the device) and MT19937 (reference stream)
// ------------------------------------------------------------------------------------------------
struct Philox {
  // Philox4x32-10 (Salmon et al. 2011), restated from the published round function.
  static void run(uint32_t k0, uint32_t k1, uint32_t c[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int i = 0; i < 10; ++i) {
      uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
      uint32_t n1 = (uint32_t)p1;
      uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
      uint32_t n3 = (uint32_t)p0;
      c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
      k0 += W0; k1 += W1;
    }
  }
};
static inline double res53(uint32_t a, uint32_t b) {  // genrand_res53 bit recipe
  a >>= 5; b >>= 6;
  return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

struct MT19937 {
  uint32_t mt[624]; int mti;
  void seed(uint32_t s) {  // init_genrand
    mt[0] = s;
    for (mti = 1; mti < 624; mti++) mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
  }
  uint32_t next() {
    if (mti >= 624) {
      for (int k = 0; k < 624; ++k) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  }
  double res53_() { uint32_t a = next(), b = next(); return res53(a, b); }
};

// ------------------------------------------------------------------------------------------------
// Scene objects
// ------------------------------------------------------------------------------------------------
struct Texture {  // src/objects/texture.rb:8-28
  int width = 0, height = 0;
  double hscale = 1, vscale = 1, u_off = 0, v_off = 0;
  const uint8_t* rgb8 = nullptr;
  // Texture#color, :23-28 — Float#to_i truncates toward zero, Integer#% is floored.
  V3 color(double uu, double vv) const {
    double fu = (uu + u_off) / hscale, fv = (vv + v_off) / vscale;
    if (!std::isfinite(fu) || !std::isfinite(fv)) { flag(RTRB_ST_NAN_TO_INT); return vmk(0, 0, 0); }
    double tu = std::trunc(fu), tv = std::trunc(fv);
    double mu = std::fmod(tu, (double)width), mv = std::fmod(tv, (double)height);
    if (mu < 0) mu += width;
    if (mv < 0) mv += height;
    int u = (int)mu, v = (int)mv;
    const uint8_t* p = rgb8 + ((size_t)v * width + u) * 3;
    // :19  (q16 >> 8) / 256.0 == v8 / 256.0 for an 8-bit source
    return vmk(p[0] / 256.0, p[1] / 256.0, p[2] / 256.0);
  }
};

struct Light {  // src/lights/light.rb, spot_light.rb
  V3 position, color;
  double radius, high_light_rate, high_light_angle;
};

struct Hit {  // what obj.intersect returns: [intersection, direction, delta]
  bool ok = false;
  V3 intersection;
  bool dir_in = false;  // :in / :out
  V3 delta;
  int data_index = -1;  // Box#intersect's `data = { index: index }` (box.rb:89): the face that was hit
};

struct Params {  // intersect_parameters result
  V3 n;
  Ray reflection;
  bool has_refraction = false;
  Ray refraction;
};

struct LitLight { const Light* light; V3 color; };

struct Counters {
  uint64_t samples = 0, rays = 0, shadow_queries = 0, highlight_hits = 0, hits = 0, local_shaded = 0,
           lit_lights = 0, mc_rays = 0, refractions = 0, texel_fetches = 0, sphere_tests = 0,
           sphere_accepts = 0, plane_tests = 0, plane_accepts = 0, cover_sphere = 0, cover_sphere_full = 0,
           cover_sphere_penumbra = 0, cover_plane = 0, cover_plane_accepts = 0, adaptive_pixels = 0,
           box_tests = 0, box_accepts = 0, cover_box = 0, cover_box_accepts = 0;
  uint32_t max_stack = 0;
  void add(const Counters& o) {
    samples += o.samples; rays += o.rays; shadow_queries += o.shadow_queries; highlight_hits += o.highlight_hits;
    hits += o.hits; local_shaded += o.local_shaded; lit_lights += o.lit_lights; mc_rays += o.mc_rays;
    refractions += o.refractions; texel_fetches += o.texel_fetches; sphere_tests += o.sphere_tests;
    sphere_accepts += o.sphere_accepts; plane_tests += o.plane_tests; plane_accepts += o.plane_accepts;
    cover_sphere += o.cover_sphere; cover_sphere_full += o.cover_sphere_full;
    cover_sphere_penumbra += o.cover_sphere_penumbra; cover_plane += o.cover_plane;
    cover_plane_accepts += o.cover_plane_accepts; adaptive_pixels += o.adaptive_pixels;
    box_tests += o.box_tests; box_accepts += o.box_accepts; cover_box += o.cover_box; cover_box_accepts += o.cover_box_accepts;
    max_stack = std::max(max_stack, o.max_stack);
  }
};

struct Ctx {
  Counters cnt;
  uint32_t status = 0;
  int64_t first_bad = -1;  // x*H + y of the first flagged pixel in reference order (x outer, y inner)
  int cur_x = 0, cur_y = 0, H = 0;
  bool counting_world_intersect = false;  // intersect() is shared by World#intersect and cover_area
};
static void flag(uint32_t bit) {
  if (!g_ctx) return;
  g_ctx->status |= bit;
  int64_t key = (int64_t)g_ctx->cur_x * g_ctx->H + g_ctx->cur_y;
  if (g_ctx->first_bad < 0 || key < g_ctx->first_bad) g_ctx->first_bad = key;
}

struct WorldObject {  // src/objects/world_object.rb
  int index = 0;
  bool has_refraction = false;
  double refractive_rate = 0;
  V3 diffuse_rate, reflective_attenuation, refractive_attenuation, ambient;
  bool has_texture = false;
  Texture texture;
  virtual ~WorldObject() {}
  virtual Hit intersect(const Ray& ray) const = 0;
  virtual Params intersect_parameters(const Ray& ray, const V3& intersection, bool dir_in, const V3& delta,
                                      int data_index = -1) const = 0;
  virtual double cover_area(const V3& light_position, double light_radius, const V3& target_position) const = 0;
  virtual V3 local_lighting(const V3& position, const std::vector<LitLight>& lights, const V3& normal_vector) const = 0;

  // WorldObject#cover_area, world_object.rb:41-49 — hard 0/1 factor
  int base_cover_area(const V3& light_position, const V3& target_position) const {
    Ray ray{vsub(light_position, target_position), target_position};
    Hit h = intersect(ray);
    if (h.ok && vdot(vsub(h.intersection, light_position), vsub(target_position, light_position)) > 0) return 1;
    return 0;
  }
  // WorldObject#local_lighting, world_object.rb:51-74
  V3 base_local_lighting(const V3& position, const std::vector<LitLight>& lights, const V3& normal_vector,
                         const V3* color_filter) const {
    V3 light_contribution = vmk(0.0, 0.0, 0.0);
    for (const LitLight& ll : lights) {
      V3 n = vnormalize(normal_vector);
      V3 l = vnormalize(vsub(ll.light->position, position));
      double l_dot_n = vdot(l, n);
      if (l_dot_n > 1) l_dot_n = 1.0;
      else if (l_dot_n < 0) l_dot_n = 0.0;
      light_contribution = vadd(light_contribution, vmul(ll.color, l_dot_n));
    }
    if (lights.size() > 0) light_contribution = vdiv(light_contribution, (double)lights.size());
    if (color_filter) return vadd(vmul(vmul(light_contribution, diffuse_rate), *color_filter), ambient);
    return vadd(vmul(light_contribution, diffuse_rate), ambient);
  }
  // get_a_random_vertical_vector, world_object.rb:105-120
  static V3 a_vertical_vector(const V3& n) {
    if (n.r == 0) flag(RTRB_ST_ZERO_VECTOR);
    const double* a = n.v;
    if (a[0] == 0) {
      if (a[1] == 0) return vmk(1.0, 0.0, 0.0);
      return vmk(0.0, -a[2] / a[1], 1.0);
    }
    return vmk(-(a[1] + a[2]) / a[0], 1.0, 1.0);
  }
  // get_reflection_by_ray_and_n, world_object.rb:121-125
  static Ray reflection_by_ray_and_n(const Ray& ray, const V3& n, const V3& intersection, const V3& delta) {
    double cos_theta = vcos(ray.front, vneg(n));
    V3 front = vnormalize(vadd(vmul(vnormalize(n), 2 * cos_theta * ray.front.r), ray.front));
    return Ray{front, vadd(intersection, delta)};
  }
  // get_refraction_by_ray_and_n, world_object.rb:127-137
  static bool refraction_by_ray_and_n(const Ray& ray, const V3& n, const V3& intersection, const V3& reflection,
                                      double refraction_rate, Ray* out) {
    double c = vcos(ray.front, n);
    double sin_i = rb_sqrt(1 - rb_pow(c, 2.0));
    double sin_r = sin_i / refraction_rate;
    if (sin_r >= 1) return false;  // total internal reflection
    double r = rb_asin(sin_r);
    V3 refraction_direction =
        vadd(vmul(vnormalize(n), -std::cos(r)), vmul(vnormalize(vadd(reflection, ray.front)), sin_r));
    *out = Ray{refraction_direction, vsub(intersection, vmul(vnormalize(n), EPSILON))};
    return true;
  }
};

struct Sphere : WorldObject {  // src/objects/sphere.rb
  V3 center; double radius = 0;
  V3 greenwich_vec, north_pole_vec, ninety_degree_east_vec;

  bool inner(const V3& position) const { return vsub(position, center).r <= radius; }  // :103-105

  // Sphere#intersect, :60-85
  Hit intersect(const Ray& ray) const override {
    Hit h;
    if (g_ctx && g_ctx->counting_world_intersect) g_ctx->cnt.sphere_tests++;
    double t = vdot(vsub(center, ray.position), ray.front) / vr2(ray.front);
    V3 v = vmul(ray.front, t);
    V3 nearest_point = vadd(ray.position, v);
    if (!inner(nearest_point)) return h;
    double nearest_dis = vsub(nearest_point, center).r;
    double nearest_point_to_intersection = rb_sqrt(rb_pow(radius, 2.0) - rb_pow(nearest_dis, 2.0));
    V3 vec = vmul(vnormalize(ray.front), nearest_point_to_intersection);
    bool from_inner = inner(ray.position);
    bool dir_in = !from_inner;
    V3 intersection = dir_in ? vsub(nearest_point, vec) : vadd(nearest_point, vec);
    if (!from_inner && t < 0) return h;
    h.ok = true;
    h.intersection = intersection;
    h.dir_in = dir_in;
    h.delta = vmul(vmul(vsub(intersection, center), EPSILON), dir_in ? 1.0 : -1.0);
    if (g_ctx && g_ctx->counting_world_intersect) g_ctx->cnt.sphere_accepts++;
    return h;
  }
  // Sphere#intersect_parameters, :88-101
  Params intersect_parameters(const Ray& ray, const V3& intersection, bool dir_in, const V3& delta, int = -1) const override {
    Params p;
    p.n = dir_in ? vsub(intersection, center) : vsub(center, intersection);
    p.reflection = reflection_by_ray_and_n(ray, p.n, intersection, delta);
    p.has_refraction = refraction_by_ray_and_n(ray, p.n, intersection, p.reflection.front,
                                               dir_in ? refractive_rate : 1.0 / refractive_rate, &p.refraction);
    return p;
  }
  // Sphere#cover_area, :28-57
  double cover_area(const V3& light_position, double light_radius, const V3& target_position) const override {
    if (g_ctx) g_ctx->cnt.cover_sphere++;
    int factor = base_cover_area(light_position, target_position);
    V3 lt = vsub(light_position, target_position);
    double t = vdot(vsub(center, target_position), lt) / vr2(lt);
    V3 x1 = vadd(target_position, vmul(lt, t));
    double r1 = light_radius * (vsub(x1, target_position).r / lt.r);
    double d = vsub(x1, center).r;
    if (d >= r1 + radius) return 0;
    double s1 = RB_PI * r1 * r1;
    if (d > std::fabs(radius - r1)) {
      if (g_ctx) g_ctx->cnt.cover_sphere_penumbra++;
      double cos_theta1 = std::min((r1 * r1 + d * d - radius * radius) / (2 * r1 * d), 1.0);
      double cos_theta2 = std::min((radius * radius + d * d - r1 * r1) / (2 * radius * d), 1.0);
      double theta1 = rb_acos(cos_theta1);
      double theta2 = rb_acos(cos_theta2);
      double delta_s = ((theta1 - std::sin(theta1)) * r1 * r1 + (theta2 - std::sin(theta2)) * radius * radius) / 2;
      return factor * delta_s / s1;
    }
    if (g_ctx) g_ctx->cnt.cover_sphere_full++;
    if (r1 > radius) return factor * RB_PI * radius * radius / s1;
    return factor;
  }
  // Sphere#get_uv, :111-120
  void get_uv(const V3& position, double* u, double* v) const {
    V3 vec = vsub(position, center);
    double x = vdot(vec, vnormalize(greenwich_vec)) / radius;
    double y = vdot(vec, vnormalize(ninety_degree_east_vec)) / radius;
    double z = vdot(vec, vnormalize(north_pole_vec)) / radius;
    double m = rb_sqrt(x * x + y * y + z * z + 2 * x + 1);
    *u = (y / m + 1) / 2;
    *v = (-z / m + 1) / 2;
  }
  // Sphere#local_lighting, :122-129
  V3 local_lighting(const V3& position, const std::vector<LitLight>& lights, const V3& normal_vector) const override {
    V3 color_filter = vmk(1.0, 1.0, 1.0);
    if (has_texture) {
      double u, v;
      get_uv(position, &u, &v);
      if (g_ctx) g_ctx->cnt.texel_fetches++;
      V3 f = vmul(texture.color(u, v), color_filter);
      return base_local_lighting(position, lights, normal_vector, &f);
    }
    return base_local_lighting(position, lights, normal_vector, &color_filter);
  }
};

struct Plane : WorldObject {  // src/objects/plane.rb
  V3 point, front, up, left;
  double u_unit = 1, v_unit = 1;

  void reinit() { left = vnormalize(vcross(front, up)); }  // :21-23

  // Plane#intersect, :38-51
  Hit intersect(const Ray& ray) const override {
    Hit h;
    if (g_ctx && g_ctx->counting_world_intersect) g_ctx->cnt.plane_tests++;
    double denominator = vdot(front, ray.front);
    if (denominator == 0) return h;
    double t = vdot(vsub(point, ray.position), front) / denominator;
    V3 intersection = vadd(ray.position, vmul(ray.front, t));
    if (t < 0) return h;
    h.ok = true;
    h.intersection = intersection;
    double fd = vdot(front, ray.front);
    h.dir_in = fd < 0;
    double nfd = -fd;
    double sgn = nfd > 0 ? 1.0 : (nfd < 0 ? -1.0 : 0.0);  // (-x <=> 0).to_f
    h.delta = vmul(vmul(front, EPSILON), sgn);
    if (g_ctx && g_ctx->counting_world_intersect) g_ctx->cnt.plane_accepts++;
    return h;
  }
  // Plane#intersect_parameters, :54-67
  Params intersect_parameters(const Ray& ray, const V3& intersection, bool, const V3& delta, int = -1) const override {
    Params p;
    p.n = vdot(front, ray.front) > 0 ? vneg(front) : front;
    p.reflection = reflection_by_ray_and_n(ray, p.n, intersection, delta);
    if (has_refraction)
      p.has_refraction = refraction_by_ray_and_n(ray, p.n, intersection, p.reflection.front, refractive_rate, &p.refraction);
    return p;
  }
  double cover_area(const V3& light_position, double, const V3& target_position) const override {
    if (g_ctx) g_ctx->cnt.cover_plane++;
    int f = base_cover_area(light_position, target_position);
    if (f && g_ctx) g_ctx->cnt.cover_plane_accepts++;
    return f;
  }
  // Plane#get_uv, :81-85
  void get_uv(const V3& position, double* u, double* v) const {
    *u = vdot(vsub(position, point), vnormalize(left)) / u_unit;
    *v = vdot(vsub(position, point), vnormalize(up)) / v_unit;
  }
  // Plane#local_lighting, :87-94
  V3 local_lighting(const V3& position, const std::vector<LitLight>& lights, const V3& normal_vector) const override {
    V3 light_filter = vmk(1.0, 1.0, 1.0);
    if (has_texture) {
      double u, v;
      get_uv(position, &u, &v);
      if (g_ctx) g_ctx->cnt.texel_fetches++;
      V3 f = vmul(texture.color(u, v), light_filter);
      return base_local_lighting(position, lights, normal_vector, &f);
    }
    return base_local_lighting(position, lights, normal_vector, &light_filter);
  }
};


// src/objects/box.rb — six bounded faces built from Plane.create_from_scratch (plane.rb:17-19: no
// attributes but the ones Box#initialize assigns), intersect = nearest face hit whose (u, v) lies in
// [-0.5, 0.5]^2.  cover_area, local_lighting and path_tracing are WorldObject's (no override), so a
// box casts hard shadows only and never samples its texture.
struct Box : WorldObject {
  V3 point, front, up;
  double width_front = 0, width_up = 0, width_left = 0;
  Plane planes[6];  // up, bottom, front, back, left, right (box.rb:60-65)

  // Box#initialize, box.rb:15-74
  void init() {
    V3 left = vnormalize(vcross(front, up));  // :23
    auto face = [&](int k, const V3& f, const V3& u, const V3& pt, double uu, double vu) {
      Plane& p = planes[k];
      p.front = f; p.up = u; p.point = pt; p.u_unit = uu; p.v_unit = vu;
      p.reflective_attenuation = reflective_attenuation;  // :66-72
      p.refractive_attenuation = refractive_attenuation;
      p.refractive_rate = refractive_rate;
      p.has_refraction = has_refraction;                  // `if self.refractive_rate` on the face (plane.rb:57)
      p.diffuse_rate = diffuse_rate;
      p.reinit();
    };
    face(0, up, left, vadd(point, vmul(vmul(up, width_up), 0.5)), width_front, width_left);            // :25-29
    face(1, vneg(up), left, vsub(point, vmul(vmul(up, width_up), 0.5)), width_front, width_left);      // :30-34
    face(2, front, up, vadd(point, vmul(vmul(front, width_front), 0.5)), width_left, width_up);        // :37-41
    face(3, vneg(front), up, vsub(point, vmul(vmul(front, width_front), 0.5)), width_left, width_up);  // :42-46
    V3 left2 = vnormalize(vcross(front, up));  // :49
    face(4, left2, up, vadd(point, vmul(vmul(left2, width_left), 0.5)), width_front, width_up);        // :50-54
    face(5, vneg(left2), up, vsub(point, vmul(vmul(left2, width_left), 0.5)), width_front, width_up);  // :55-59
  }
  // Box#intersect, box.rb:80-99
  Hit intersect(const Ray& ray) const override {
    const bool counting = g_ctx && g_ctx->counting_world_intersect;
    if (counting) { g_ctx->cnt.box_tests++; g_ctx->counting_world_intersect = false; }  // faces are not world planes
    double nearest_dis = INFINITY;
    Hit nearest;
    for (int index = 0; index < 6; ++index) {
      const Plane& plane = planes[index];
      Hit h = plane.intersect(ray);
      if (h.ok) {
        double u, v;
        plane.get_uv(h.intersection, &u, &v);
        if (-0.5 <= u && u <= 0.5 && -0.5 <= v && v <= 0.5) {
          double d = vsub(h.intersection, ray.position).r;
          if (d < nearest_dis) {
            nearest_dis = d;
            nearest = h;
            nearest.data_index = index;
          }
        }
      }
    }
    if (counting) { g_ctx->counting_world_intersect = true; if (nearest.ok) g_ctx->cnt.box_accepts++; }
    return nearest;
  }
  // Box#intersect_parameters, box.rb:102-107
  Params intersect_parameters(const Ray& ray, const V3& intersection, bool dir_in, const V3& delta, int data_index) const override {
    return planes[data_index].intersect_parameters(ray, intersection, dir_in, delta);
  }
  double cover_area(const V3& light_position, double, const V3& target_position) const override {  // world_object.rb:41-49
    if (g_ctx) g_ctx->cnt.cover_box++;
    int f = base_cover_area(light_position, target_position);
    if (f && g_ctx) g_ctx->cnt.cover_box_accepts++;
    return f;
  }
  V3 local_lighting(const V3& position, const std::vector<LitLight>& lights, const V3& normal_vector) const override {
    return base_local_lighting(position, lights, normal_vector, nullptr);  // world_object.rb:51-74, color_filter = nil
  }
};

// ------------------------------------------------------------------------------------------------
// World — src/world.rb
// ------------------------------------------------------------------------------------------------
struct World {
  double max_distance = 0, soft_shadow_exponent = 0;
  std::vector<std::unique_ptr<WorldObject>> world_objects;
  std::vector<Light> lights;

  // World#intersect, :37-59
  const WorldObject* intersect(const Ray& ray, Hit* out) const {
    const WorldObject* nearest_obj = nullptr;
    double nearest_dis = max_distance;
    g_ctx->counting_world_intersect = true;
    for (const auto& obj : world_objects) {
      Hit h = obj->intersect(ray);
      if (h.ok) {
        double new_dis = ray.distance(h.intersection);
        if (new_dis < nearest_dis) {
          nearest_dis = new_dis;
          nearest_obj = obj.get();
          *out = h;
        }
      }
    }
    g_ctx->counting_world_intersect = false;
    return nearest_obj;
  }
  // World#lit_area, :62-69
  double lit_area(const V3& target, const V3& light_pos, double radius) const {
    double total_area = 1;
    for (const auto& obj : world_objects) {
      double covered_area = obj->cover_area(light_pos, radius, target);
      total_area -= covered_area;
    }
    return std::max(total_area, 0.0);
  }
  // World#local_lights, :72-80
  std::vector<LitLight> local_lights(const V3& position) const {
    std::vector<LitLight> ret;
    for (const Light& light : lights) {
      g_ctx->cnt.shadow_queries++;
      double area = lit_area(position, light.position, light.radius);
      if (area > 0) {
        ret.push_back(LitLight{&light, vmul(light.color, rb_pow(area, soft_shadow_exponent) / (double)lights.size())});
      }
    }
    return ret;
  }
  // World#high_lights, :83-98. The `&& lit_area(...)` term is always truthy in Ruby (numbers are
  // truthy) so it never blocks a highlight; its cover tests are wasted work and are skipped here.
  std::vector<LitLight> high_lights(const Ray& ray) const {
    std::vector<LitLight> ret;
    for (const Light& light : lights) {
      V3 a = vsub(light.position, ray.position);
      double cos_theta = vcos(ray.front, a);
      if (cos_theta < -1) cos_theta = -1;
      if (cos_theta > 1) cos_theta = 1;
      double ang = rb_acos(cos_theta);
      if (ang < (light.high_light_angle / 180.0 * RB_PI)) {
        ret.push_back(LitLight{&light, vmul(light.color, light.high_light_rate)});
      }
    }
    return ret;
  }
};

// ------------------------------------------------------------------------------------------------
// RayTracer — src/ray_tracer.rb:7-164, 292-304
// ------------------------------------------------------------------------------------------------
struct RayItem {
  Ray ray;
  int trace_depth;
  V3 attenuation;
  uint32_t path;  // counter-RNG ray id: root 1, child = parent*(mc+2) + slot (slot 0 refl, 1 refr, 2+m mc)
};

struct RngState {
  int mode = RTRB_RNG_CTR;
  uint32_t k0 = 0, k1 = 0;
  MT19937* mt = nullptr;
  uint32_t pixel = 0, sample = 0;
  // one uniform for the lens (camera.rb:135)
  double lens() {
    if (mode == RTRB_RNG_MT) return mt->res53_();
    uint32_t c[4] = {pixel, sample, 0u, 0u};
    Philox::run(k0, k1, c);
    return res53(c[0], c[1]);
  }
  // two uniforms for one MC ray (world_object.rb:84), theta's draw first
  void mc(uint32_t path, double* u_theta, double* u_phi) {
    if (mode == RTRB_RNG_MT) { *u_theta = mt->res53_(); *u_phi = mt->res53_(); return; }
    uint32_t c[4] = {pixel, sample, path, 1u};
    Philox::run(k0, k1, c);
    *u_theta = res53(c[0], c[1]);
    *u_phi = res53(c[2], c[3]);
  }
};

struct RayTracer {
  const World* world;
  int trace_depth, mc_times;
  RngState* rng;

  // WorldObject#path_tracing, world_object.rb:76-90
  void path_tracing(const WorldObject* obj, const V3& intersection, const V3& n, const RayItem& parent,
                    std::vector<RayItem>* rays) const {
    V3 att = vdiv(obj->diffuse_rate, (double)mc_times);
    for (int m = 0; m < mc_times; ++m) {
      V3 front = vnormalize(n);
      V3 left = vnormalize(WorldObject::a_vertical_vector(n));
      V3 up = vcross(front, left);
      uint32_t child_path = parent.path * (uint32_t)(mc_times + 2) + (uint32_t)(2 + m);
      double ut, up_;
      rng->mc(child_path, &ut, &up_);
      double theta = ut * RB_PI / 2, phi = up_ * RB_PI * 2;
      V3 direction = vadd(vmul(front, std::sin(theta)),
                          vmul(vadd(vmul(left, std::cos(phi)), vmul(up, std::sin(phi))), std::cos(theta)));
      g_ctx->cnt.mc_rays++;
      rays->push_back(RayItem{Ray{direction, intersection}, parent.trace_depth - 1, vmul(parent.attenuation, att), child_path});
    }
  }

  // rt_map, ray_tracer.rb:50-164
  void rt_map(const RayItem& it, std::vector<RayItem>* rays, std::vector<V3>* colors, int* primary_hit) const {
    if (it.trace_depth <= 0 || it.attenuation.r < 0.0001) return;  // :52
    g_ctx->cnt.rays++;
    std::vector<LitLight> hl = world->high_lights(it.ray);  // :60
    for (const LitLight& l : hl) colors->push_back(vdiv(vmul(it.attenuation, l.color), (double)hl.size()));  // :65
    if (!hl.empty()) {  // :75
      g_ctx->cnt.highlight_hits++;
      if (primary_hit) *primary_hit = -2;
      return;
    }
    Hit h;
    const WorldObject* object = world->intersect(it.ray, &h);  // :78
    if (!object) { if (primary_hit) *primary_hit = -1; return; }
    if (primary_hit) *primary_hit = object->index;
    g_ctx->cnt.hits++;
    Params p = object->intersect_parameters(it.ray, h.intersection, h.dir_in, h.delta, h.data_index);  // :80
    const uint32_t K = (uint32_t)(mc_times + 2);
    // reflection child, :87-103 (always exists)
    rays->push_back(RayItem{p.reflection, it.trace_depth - 1, vmul(it.attenuation, object->reflective_attenuation), it.path * K + 0u});
    if (p.has_refraction) {  // :105-121
      g_ctx->cnt.refractions++;
      rays->push_back(RayItem{p.refraction, it.trace_depth - 1, vmul(it.attenuation, object->refractive_attenuation), it.path * K + 1u});
    }
    V3 shade_from = vadd(h.intersection, h.delta);
    std::vector<LitLight> lights = world->local_lights(shade_from);  // :123
    if (lights.empty()) {
      path_tracing(object, shade_from, p.n, it, rays);  // :131
    } else {
      g_ctx->cnt.local_shaded++;
      g_ctx->cnt.lit_lights += lights.size();
      colors->push_back(vmul(it.attenuation, object->local_lighting(h.intersection, lights, p.n)));  // :152
    }
  }

  // trace_sync, ray_tracer.rb:16-46 + rt_reduce :292-298
  V3 trace_sync(const Ray& ray, int* primary_hit) const {
    g_ctx->cnt.samples++;
    std::vector<RayItem> queue;
    std::vector<V3> light_queue;
    queue.push_back(RayItem{ray, trace_depth, vmk(1.0, 1.0, 1.0), 1u});
    bool first = true;
    while (!queue.empty()) {
      g_ctx->cnt.max_stack = std::max<uint32_t>(g_ctx->cnt.max_stack, (uint32_t)queue.size());
      RayItem item = queue.back();
      queue.pop_back();
      std::vector<RayItem> rays;
      rt_map(item, &rays, &light_queue, first ? primary_hit : nullptr);
      first = false;
      for (const RayItem& r : rays) queue.push_back(r);
    }
    V3 sum = vmk(0.0, 0.0, 0.0);
    for (const V3& c : light_queue) {
      sum = vadd(sum, c);  // mix_color, :300-303
      if (sum.v[0] > 1 || sum.v[1] > 1 || sum.v[2] > 1) flag(RTRB_ST_COLOR_GT_1);  // :294-296
    }
    return sum;
  }
};

// ------------------------------------------------------------------------------------------------
// Camera — src/camera.rb:70-99, 123-156
// ------------------------------------------------------------------------------------------------
struct Camera {
  V3 position, up, front;
  double retina_width, retina_height, aperture_radius, image_distance, focal_distance, variant_threshold;
  int width, height, pre_sample_times, max_sample_times;

  // intersect_plane, :123-127
  static V3 intersect_plane(const Ray& ray, const V3& point, const V3& front) {
    double t = vdot(vsub(point, ray.position), front) / vdot(front, ray.front);
    return vadd(ray.position, vmul(ray.front, t));
  }
  // lens_func, :129-151
  Ray lens_func(int x, int y, double theta) const {
    V3 left = vnormalize(vcross(up, front));
    V3 retina_center = vsub(position, vmul(vnormalize(front), image_distance));
    V3 retina_position = vadd(vadd(retina_center, vmul(left, 2.0 * ((double)x / width - 0.5) * retina_width)),
                              vmul(vnormalize(up), 2 * ((double)y / height - 0.5) * retina_height));
    V3 rand_vector = vmul(vadd(vmul(vnormalize(left), std::cos(theta)), vmul(vnormalize(up), std::sin(theta))), aperture_radius);
    V3 aperture_position = vadd(position, rand_vector);
    double object_distance = focal_distance * image_distance / (image_distance - focal_distance);
    V3 point_on_focal_plane = vadd(position, vmul(vnormalize(front), object_distance));
    V3 normal_vector_focal_plane = front;
    Ray r{vsub(position, retina_position), retina_position};
    V3 target_point = intersect_plane(r, point_on_focal_plane, normal_vector_focal_plane);
    return Ray{vsub(target_point, aperture_position), aperture_position};
  }
  // render_at, :70-99 — returns the unclamped colour
  V3 render_at(int x, int y, const RayTracer& rt, RngState* rng, int* primary_hit) const {
    g_ctx->cur_x = x; g_ctx->cur_y = y;
    rng->pixel = (uint32_t)((uint64_t)y * width + x);
    std::vector<V3> pre_samples;
    V3 average = vmk(0.0, 0.0, 0.0);
    for (int j = 0; j < pre_sample_times; ++j) {
      rng->sample = (uint32_t)j;
      Ray ray = lens_func(x, y, rng->lens());
      V3 v = rt.trace_sync(ray, j == 0 ? primary_hit : nullptr);
      pre_samples.push_back(v);
      average = vadd(average, v);
    }
    double variance = 0;
    average = vdiv(average, (double)pre_sample_times);
    for (int j = 0; j < pre_sample_times; ++j) {
      V3 d = vsub(pre_samples[j], average);
      double m = std::max(d.v[0], std::max(d.v[1], d.v[2]));  // .to_a.max (signed)
      variance += rb_pow(m, 2.0);
    }
    variance /= pre_sample_times;
    if (variance >= variant_threshold) {
      g_ctx->cnt.adaptive_pixels++;
      V3 color_vec = vmk(0.0, 0.0, 0.0);
      for (int j = pre_sample_times; j < max_sample_times; ++j) {
        rng->sample = (uint32_t)j;
        Ray ray = lens_func(x, y, rng->lens());
        color_vec = vadd(color_vec, rt.trace_sync(ray, nullptr));
      }
      average = vdiv(vadd(vmul(average, (double)pre_sample_times), color_vec), (double)max_sample_times);
    }
    return average;
  }
};

// array_to_color, camera.rb:153-156 + PNG::Color.new truncation (parity is defined here)
static inline uint8_t quantise(double c) {
  double x = c * 256.0;
  double m = std::min(x, 255.0);  // [x, 255].min
  if (!(m > 0)) return 0;        // negative / NaN -> 0 (byte packing of the png gem is unpinned)
  return (uint8_t)(int)m;
}

struct OracleScene {
  World world;
  std::vector<std::vector<uint8_t>> tex_store;
};

static OracleScene* build_scene(const rtrb_scene_desc* sd) {
  auto* s = new OracleScene();
  s->world.max_distance = sd->max_distance;
  s->world.soft_shadow_exponent = sd->soft_shadow_exponent;
  for (int i = 0; i < sd->n_textures; ++i) {
    const rtrb_texture_desc& t = sd->textures[i];
    s->tex_store.emplace_back(t.rgb8, t.rgb8 + (size_t)t.width * t.height * 3);
  }
  for (int i = 0; i < sd->n_lights; ++i) {
    const rtrb_light_desc& l = sd->lights[i];
    s->world.lights.push_back(Light{v3(l.position), v3(l.color), l.radius, l.high_light_rate, l.high_light_angle});
  }
  for (int i = 0; i < sd->n_objects; ++i) {
    const rtrb_object_desc& o = sd->objects[i];
    std::unique_ptr<WorldObject> wo;
    if (o.type == RTRB_OBJ_SPHERE) {
      auto sp = std::make_unique<Sphere>();
      sp->center = v3(o.point); sp->radius = o.radius;
      if (o.texture >= 0) {
        sp->greenwich_vec = v3(o.greenwich_vec); sp->north_pole_vec = v3(o.north_pole_vec);
        sp->ninety_degree_east_vec = vcross(sp->north_pole_vec, sp->greenwich_vec);  // sphere.rb:19
      }
      wo = std::move(sp);
    } else if (o.type == RTRB_OBJ_BOX) {
      auto bx = std::make_unique<Box>();
      bx->point = v3(o.point); bx->front = v3(o.front); bx->up = v3(o.up);
      bx->width_front = o.width_front; bx->width_up = o.width_up; bx->width_left = o.width_left;
      wo = std::move(bx);
    } else {
      auto pl = std::make_unique<Plane>();
      pl->point = v3(o.point); pl->front = v3(o.front); pl->up = v3(o.up);
      pl->u_unit = o.u_unit; pl->v_unit = o.v_unit;
      pl->reinit();
      wo = std::move(pl);
    }
    wo->index = i;
    wo->has_refraction = o.has_refraction != 0;
    wo->refractive_rate = o.refractive_rate;
    wo->diffuse_rate = v3(o.diffuse_rate);
    wo->reflective_attenuation = v3(o.reflective_attenuation);
    wo->refractive_attenuation = v3(o.refractive_attenuation);
    wo->ambient = v3(o.ambient);
    if (o.type == RTRB_OBJ_BOX) static_cast<Box*>(wo.get())->init();  // needs the material (box.rb:66-72)
    if (o.texture >= 0 && o.type != RTRB_OBJ_BOX) {
      wo->has_texture = true;
      const rtrb_texture_desc& t = sd->textures[o.texture];
      wo->texture.width = t.width; wo->texture.height = t.height;
      wo->texture.hscale = o.texture_horizontal_scale; wo->texture.vscale = o.texture_vertical_scale;
      wo->texture.u_off = o.texture_u_offset; wo->texture.v_off = o.texture_v_offset;
      wo->texture.rgb8 = s->tex_store[o.texture].data();
    }
    s->world.world_objects.push_back(std::move(wo));
  }
  return s;
}

static Camera build_camera(const rtrb_camera_desc* c) {
  Camera cam;
  cam.position = v3(c->position); cam.up = v3(c->up); cam.front = v3(c->front);
  cam.retina_width = c->retina_width; cam.retina_height = c->retina_height;
  cam.aperture_radius = c->aperture_radius; cam.image_distance = c->image_distance;
  cam.focal_distance = c->focal_distance; cam.variant_threshold = c->variant_threshold;
  cam.width = c->width; cam.height = c->height;
  cam.pre_sample_times = c->pre_sample_times; cam.max_sample_times = c->max_sample_times;
  return cam;
}

}  // namespace

// ================================================================================================
// C interface for the Python test harness (ctypes)
// ================================================================================================
extern "C" {

void* rtrb_oracle_scene_create(const rtrb_scene_desc* sd) { return build_scene(sd); }
void rtrb_oracle_scene_destroy(void* s) { delete (OracleScene*)s; }

// Renders the window [x0,x1) x [y0,y1) (all zero = full frame) with `threads` column strips
// (camera.rb:54 strip formula).  In MT mode every strip restarts from the same seeded state, as
// forked children do.  rgba: H*W*4, rgb: H*W*3 doubles, hit: H*W int32 (any may be NULL).
int rtrb_oracle_render(void* scene, const rtrb_camera_desc* cd, const rtrb_render_opts* opts, int threads,
                       uint8_t* rgba, double* rgb, int32_t* hit, rtrb_stats* stats) {
  const OracleScene* s = (const OracleScene*)scene;
  Camera cam = build_camera(cd);
  const int W = cam.width, H = cam.height;
  int x0 = opts->x0, y0 = opts->y0, x1 = opts->x1, y1 = opts->y1;
  if (x0 == 0 && y0 == 0 && x1 == 0 && y1 == 0) { x1 = W; y1 = H; }
  if (threads < 1) threads = 1;
  std::vector<Ctx> ctxs(threads);
  auto work = [&](int i) {
    Ctx& ctx = ctxs[i];
    ctx.H = H;
    g_ctx = &ctx;
    MT19937 mt;
    mt.seed((uint32_t)opts->seed);
    RngState rng;
    rng.mode = opts->rng_mode;
    rng.k0 = (uint32_t)opts->seed; rng.k1 = (uint32_t)(opts->seed >> 32);
    rng.mt = &mt;
    RayTracer rt{&s->world, cd->trace_depth, cd->monte_carlo_diffusion_times, &rng};
    const int wx = x1 - x0;
    int sx = x0 + (int)((double)i / threads * wx), ex = x0 + (int)((double)(i + 1) / threads * wx);  // camera.rb:54
    for (int x = sx; x < ex; ++x) {
      for (int y = y0; y < y1; ++y) {  // x outer, y inner (camera.rb:59-63)
        int ph = -1;
        V3 c = cam.render_at(x, y, rt, &rng, &ph);
        size_t px = (size_t)y * W + x;  // row = y, column = x
        if (rgb) { rgb[px * 3 + 0] = c.v[0]; rgb[px * 3 + 1] = c.v[1]; rgb[px * 3 + 2] = c.v[2]; }
        if (rgba) {
          rgba[px * 4 + 0] = quantise(c.v[0]); rgba[px * 4 + 1] = quantise(c.v[1]);
          rgba[px * 4 + 2] = quantise(c.v[2]); rgba[px * 4 + 3] = 255;
        }
        if (hit) hit[px] = ph;
      }
    }
    g_ctx = nullptr;
  };
  if (threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < threads; ++i) th.emplace_back(work, i);
    for (auto& t : th) t.join();
  }
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    Counters c;
    uint32_t status = 0; int64_t first_bad = -1;
    for (const Ctx& x : ctxs) {
      c.add(x.cnt);
      status |= x.status;
      if (x.first_bad >= 0 && (first_bad < 0 || x.first_bad < first_bad)) first_bad = x.first_bad;
    }
    stats->samples = c.samples; stats->rays = c.rays; stats->shadow_queries = c.shadow_queries;
    stats->highlight_hits = c.highlight_hits; stats->hits = c.hits; stats->local_shaded = c.local_shaded;
    stats->lit_lights = c.lit_lights; stats->mc_rays = c.mc_rays; stats->refractions = c.refractions;
    stats->texel_fetches = c.texel_fetches; stats->sphere_tests = c.sphere_tests; stats->sphere_accepts = c.sphere_accepts;
    stats->plane_tests = c.plane_tests; stats->plane_accepts = c.plane_accepts; stats->cover_sphere = c.cover_sphere;
    stats->cover_sphere_full = c.cover_sphere_full; stats->cover_sphere_penumbra = c.cover_sphere_penumbra;
    stats->cover_plane = c.cover_plane; stats->cover_plane_accepts = c.cover_plane_accepts;
    stats->adaptive_pixels = c.adaptive_pixels; stats->max_stack = c.max_stack;
    stats->box_tests = c.box_tests; stats->box_accepts = c.box_accepts;
    stats->cover_box = c.cover_box; stats->cover_box_accepts = c.cover_box_accepts;
    stats->status = status;
    stats->first_bad_x = first_bad < 0 ? -1 : (int32_t)(first_bad / H);
    stats->first_bad_y = first_bad < 0 ? -1 : (int32_t)(first_bad % H);
  }
  return 0;
}

// ---- fine-grained probes for the known-answer tests ------------------------------
// out6 = ray front(3), position(3)
void rtrb_oracle_lens_ray(const rtrb_camera_desc* cd, int x, int y, double theta, double* out6) {
  Camera cam = build_camera(cd);
  Ray r = cam.lens_func(x, y, theta);
  for (int i = 0; i < 3; ++i) { out6[i] = r.front.v[i]; out6[3 + i] = r.position.v[i]; }
}
double rtrb_oracle_object_distance(const rtrb_camera_desc* cd) {
  return cd->focal_distance * cd->image_distance / (cd->image_distance - cd->focal_distance);
}
// obj.intersect(ray): returns 1 on hit; out = intersection(3), delta(3); dir_in
int rtrb_oracle_intersect(void* scene, int obj, const double* o, const double* d, double* out6, int* dir_in) {
  Ctx ctx; g_ctx = &ctx;
  const OracleScene* s = (const OracleScene*)scene;
  Ray ray{v3(d), v3(o)};
  Hit h = s->world.world_objects[obj]->intersect(ray);
  g_ctx = nullptr;
  if (!h.ok) return 0;
  for (int i = 0; i < 3; ++i) { out6[i] = h.intersection.v[i]; out6[3 + i] = h.delta.v[i]; }
  *dir_in = h.dir_in;
  return 1;
}
// Box#intersect's data[:index] (box.rb:89) for object `obj`: the face hit (0 up, 1 bottom, 2 front, 3 back,
// 4 left, 5 right), -1 on a miss or when the object is not a box
int rtrb_oracle_box_face(void* scene, int obj, const double* o, const double* d) {
  Ctx ctx; g_ctx = &ctx;
  const OracleScene* s = (const OracleScene*)scene;
  Ray ray{v3(d), v3(o)};
  Hit h = s->world.world_objects[obj]->intersect(ray);
  g_ctx = nullptr;
  return h.ok ? h.data_index : -1;
}
// World#intersect: returns object index or -1; out = intersection(3)
int rtrb_oracle_world_intersect(void* scene, const double* o, const double* d, double* out3) {
  Ctx ctx; g_ctx = &ctx;
  const OracleScene* s = (const OracleScene*)scene;
  Ray ray{v3(d), v3(o)};
  Hit h;
  const WorldObject* w = s->world.intersect(ray, &h);
  g_ctx = nullptr;
  if (!w) return -1;
  for (int i = 0; i < 3; ++i) out3[i] = h.intersection.v[i];
  return w->index;
}
double rtrb_oracle_cover_area(void* scene, int obj, const double* light_pos, double light_radius, const double* target) {
  Ctx ctx; g_ctx = &ctx;
  const OracleScene* s = (const OracleScene*)scene;
  double r = s->world.world_objects[obj]->cover_area(v3(light_pos), light_radius, v3(target));
  g_ctx = nullptr;
  return r;
}
double rtrb_oracle_lit_area(void* scene, const double* light_pos, double light_radius, const double* target) {
  Ctx ctx; g_ctx = &ctx;
  const OracleScene* s = (const OracleScene*)scene;
  double r = s->world.lit_area(v3(target), v3(light_pos), light_radius);
  g_ctx = nullptr;
  return r;
}
// number of lights matched by World#high_lights for this ray
int rtrb_oracle_high_lights(void* scene, const double* o, const double* d) {
  Ctx ctx; g_ctx = &ctx;
  const OracleScene* s = (const OracleScene*)scene;
  Ray ray{v3(d), v3(o)};
  int n = (int)s->world.high_lights(ray).size();
  g_ctx = nullptr;
  return n;
}
// Texture#color for object `obj` at (u,v) -> out3; also returns col,row
void rtrb_oracle_texture_color(void* scene, int obj, double u, double v, double* out3) {
  Ctx ctx; g_ctx = &ctx;
  const OracleScene* s = (const OracleScene*)scene;
  V3 c = s->world.world_objects[obj]->texture.color(u, v);
  g_ctx = nullptr;
  for (int i = 0; i < 3; ++i) out3[i] = c.v[i];
}
// get_uv of object `obj` at position p -> out2
void rtrb_oracle_get_uv(void* scene, int obj, const double* p, double* out2) {
  Ctx ctx; g_ctx = &ctx;
  const OracleScene* s = (const OracleScene*)scene;
  const WorldObject* w = s->world.world_objects[obj].get();
  if (auto sp = dynamic_cast<const Sphere*>(w)) sp->get_uv(v3(p), &out2[0], &out2[1]);
  else if (auto pl = dynamic_cast<const Plane*>(w)) pl->get_uv(v3(p), &out2[0], &out2[1]);
  g_ctx = nullptr;
}
// reflection / refraction rays for a hit: out12 = refl front, refl pos, refr front, refr pos; returns has_refraction
int rtrb_oracle_intersect_parameters(void* scene, int obj, const double* o, const double* d, double* n3, double* out12) {
  Ctx ctx; g_ctx = &ctx;
  const OracleScene* s = (const OracleScene*)scene;
  Ray ray{v3(d), v3(o)};
  const WorldObject* w = s->world.world_objects[obj].get();
  Hit h = w->intersect(ray);
  if (!h.ok) { g_ctx = nullptr; return -1; }
  Params p = w->intersect_parameters(ray, h.intersection, h.dir_in, h.delta, h.data_index);
  g_ctx = nullptr;
  for (int i = 0; i < 3; ++i) {
    n3[i] = p.n.v[i];
    out12[i] = p.reflection.front.v[i]; out12[3 + i] = p.reflection.position.v[i];
    out12[6 + i] = p.has_refraction ? p.refraction.front.v[i] : 0.0;
    out12[9 + i] = p.has_refraction ? p.refraction.position.v[i] : 0.0;
  }
  return p.has_refraction ? 1 : 0;
}
// Vec3 probes: op codes used by tests/test_vec3_reference.py
//  0 dot 1 cos 2 cross 3 add 4 sub 5 mul(vec) 6 mul(scalar b[0]) 7 div(scalar b[0]) 8 r 9 r2 10 normalize 11 neg
void rtrb_oracle_vec3(int op, const double* a, const double* b, double* out3) {
  Ctx ctx; g_ctx = &ctx;
  V3 A = v3(a), B = b ? v3(b) : vmk(0, 0, 0), R = vmk(0, 0, 0);
  double s = 0; bool scalar = false;
  switch (op) {
    case 0: s = vdot(A, B); scalar = true; break;
    case 1: s = vcos(A, B); scalar = true; break;
    case 2: R = vcross(A, B); break;
    case 3: R = vadd(A, B); break;
    case 4: R = vsub(A, B); break;
    case 5: R = vmul(A, B); break;
    case 6: R = vmul(A, b[0]); break;
    case 7: R = vdiv(A, b[0]); break;
    case 8: s = A.r; scalar = true; break;
    case 9: s = vr2(A); scalar = true; break;
    case 10: R = vnormalize(A); break;
    case 11: R = vneg(A); break;
  }
  g_ctx = nullptr;
  if (scalar) { out3[0] = s; out3[1] = 0; out3[2] = 0; }
  else { out3[0] = R.v[0]; out3[1] = R.v[1]; out3[2] = R.v[2]; }
}
// RNG probes (the counter RNG definition shared with the device)
void rtrb_oracle_philox(uint32_t k0, uint32_t k1, const uint32_t* ctr4, uint32_t* out4) {
  uint32_t c[4] = {ctr4[0], ctr4[1], ctr4[2], ctr4[3]};
  Philox::run(k0, k1, c);
  for (int i = 0; i < 4; ++i) out4[i] = c[i];
}
void rtrb_oracle_mt_res53(uint32_t seed, int n, double* out) {
  MT19937 mt; mt.seed(seed);
  for (int i = 0; i < n; ++i) out[i] = mt.res53_();
}

}  // extern "C"
