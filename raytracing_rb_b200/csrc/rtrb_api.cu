// rtrb_api.cu — C ABI (include/rtrb_b200.h): scene baking, buffers, frame orchestration, the
// resolve kernels (Camera#render_at's mean / variance test / final average, camera.rb:80-98, and
// array_to_color, camera.rb:153-156), multi-GPU tile gather through peer mappings, and the FMA
// issue microbenchmark used as the roofline denominator.
//
// Compiled with -fmad=false: the host-side baking and the resolve arithmetic must round exactly
// like the reference's scalar FP64 code.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtrb_b200.h"
#include "rtrb_launch.h"
#include "rtrb_trace.cuh"
#include "rtrb_types.h"

// Scenes with at most this many spheres use the linear-scan filter kernels and keep cull_sph[] in
// world order; larger scenes get the BVH kernels (measured crossover, see rtrb_trace_fast.cuh).
#define RTRB_BVH_MIN_SPHERES 32

namespace {

thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(RTRB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ---- host FP64 vector helpers in the reference's evaluation order (fast_4d_matrix.c) -------------
struct H3 { double x, y, z; };
inline H3 h3(const double* p) { return H3{p[0], p[1], p[2]}; }
inline double hnorm(H3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline H3 hnormalize(H3 a) { double r = hnorm(a); return H3{a.x / r, a.y / r, a.z / r}; }
inline H3 hcross(H3 a, H3 b) { return H3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline H3 hadd(H3 a, H3 b) { return H3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline H3 hsub(H3 a, H3 b) { return H3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline H3 hmul(H3 a, double s) { return H3{a.x * s, a.y * s, a.z * s}; }
inline void put(double* d, H3 a) { d[0] = a.x; d[1] = a.y; d[2] = a.z; }

// Apex table entry for sphere (C, R) seen from apex A with positional uncertainty rho (see FrameParams).
// Margins: rounding of v and of the FP32 direction (8 eps |v| on b, hence 16 eps |v|^2 on b^2), the
// apex uncertainty, and the filter's usual scale term; Kq is rounded DOWN so the test only ever errs
// towards keeping a sphere.
float4 apex_entry(const double A[3], double rho, const double C[3], double R, double m_scale) {
  const double eps = 5.9604644775390625e-8;  // 2^-24
  const double vx = C[0] - A[0], vy = C[1] - A[1], vz = C[2] - A[2];
  const double v2 = vx * vx + vy * vy + vz * vz, vlen = sqrt(v2);
  const double Rp = fabs(R) + rho + 96.0 * eps * (m_scale + vlen) + 8.0 * eps * vlen;
  const double K = Rp * Rp + 32.0 * eps * v2;
  const double Kq = v2 - K;
  float kf = (float)Kq;
  if ((double)kf > Kq) kf = nextafterf(kf, -INFINITY);
  if (!(Kq == Kq)) kf = -INFINITY;  // NaN: always survive
  return make_float4((float)vx, (float)vy, (float)vz, kf);
}

// MT19937 (Matsumoto & Nishimura 1998) exactly as Ruby drives it: Random.srand(seed) with a seed that fits 32
// bits runs init_genrand(seed) (random.c rand_init), and every Random.rand is genrand_res53.  Used only by the
// RTRB_RNG_MT validation mode: the stream is generated here and the device indexes into it.
struct HostMT {
  uint32_t mt[624];
  int idx = 625;
  void seed(uint32_t s) {
    mt[0] = s;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next32() {
    if (idx >= 624) {
      for (int k = 0; k < 624; ++k) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  }
  double res53() {
    const uint32_t a = next32() >> 5, b = next32() >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
  }
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t want) {
    if (want <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    if (want == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

// Frame control block: every small per-frame device word in ONE buffer (one memset, one D2H copy).
// u64 words: [0, RTRB_CNT_N) counters | [N] first_bad | [N+1, N+2] work claims | [N+3] = {status, max_stack}
// | [N+4] = {extra_count, pad} | [N+5, N+5+2*RTRB_HOT_SLICES) sliced (rays, shadow) counters
#define RTRB_FCB_WORDS (RTRB_CNT_N + 5 + 2 * RTRB_HOT_SLICES)
#define RTRB_PIPE_SLOTS 4  // frames that may be in flight between rtrb_submit and rtrb_wait
struct FrameCtl {
  DevBuf<unsigned long long> d;
  unsigned long long* h = nullptr;  // pinned host mirror (mapped into the device's address space)
  unsigned long long* h_dev = nullptr;  // ... and the address the device stores to (publish_ctl_kernel)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr, copied = nullptr, ctl_copied = nullptr;
  DevBuf<uint8_t> rgba;             // pipeline slots own a framebuffer; the main slot uses rtrb_renderer::rgba
  // facts of the frame in flight, needed to finish its stats later
  int W = 0, H = 0, S = 0, E = 0, n_tiles = 0;
  bool detail = false, in_flight = false, timed = false;
  unsigned ticket = 0;              // the rtrb_submit ticket occupying this slot while in_flight
  size_t px_count = 0;
  int init() {
    cudaError_t e;
    if ((e = d.ensure(RTRB_FCB_WORDS)) != cudaSuccess) return (int)e;
    if ((e = cudaHostAlloc((void**)&h, RTRB_FCB_WORDS * sizeof(unsigned long long), cudaHostAllocMapped)) != cudaSuccess) return (int)e;
    if ((e = cudaHostGetDevicePointer((void**)&h_dev, h, 0)) != cudaSuccess) return (int)e;
    cudaEvent_t* evs[6] = {&ev0, &ev1, &evt0, &evt1, &copied, &ctl_copied};
    for (auto ev : evs)
      if ((e = cudaEventCreate(ev)) != cudaSuccess) return (int)e;
    return 0;
  }
  void destroy() {
    d.release(); rgba.release();
    if (h) cudaFreeHost(h);
    h = nullptr; h_dev = nullptr;
    cudaEvent_t* evs[6] = {&ev0, &ev1, &evt0, &evt1, &copied, &ctl_copied};
    for (auto ev : evs) { if (*ev) cudaEventDestroy(*ev); *ev = nullptr; }
  }
};

struct rtrb_renderer {
  int device = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t push_ev = nullptr, push_done_ev = nullptr;  // rtrb_peer_push ordering (no timing)
  FrameCtl main_ctl;       // synchronous calls
  FrameCtl pipe_ctl[RTRB_PIPE_SLOTS];  // rtrb_submit / rtrb_wait frame slots
  bool pipe_ready = false;
  unsigned next_ticket = 0;
  // baked scene
  int n_objects = 0, n_lights = 0, n_boxes = 0;
  double max_distance = 0, soft_shadow_exponent = 0;
  DevBuf<DevGeom> geom;
  DevBuf<DevMat> mat;
  DevBuf<DevLight> lights;
  DevBuf<DevBox> boxes;
  // FP32 filter view (FAST64)
  DevBuf<float4> cull_sph, cull_pl;
  DevBuf<BvhNode> bvh;
  DevBuf<float4> light_tab;            // apex tables of the lights (linear-filter scenes)
  // small scenes: host copy of the per-ray tables that travel in the kernel parameters (FrameParams::k_*)
  struct SmallTables {
    float4 light_tab[RTRB_K_LIGHTS][RTRB_APEX_MAX];
    float4 cull_sph[RTRB_APEX_MAX];
    float4 cull_pl[2 * RTRB_K_PLANES];
    DevLight lights[RTRB_K_LIGHTS];
    DevLightF lights_f[RTRB_K_LIGHTS];
    int32_t pl_index[RTRB_K_PLANES];
    int32_t has_light_tab;
  } k;
  int scene_class = 0;                 // RTRB_SCENE_CLASS_* bits (FrameParams::scene_class)
  bool small_scene = false;            // <= 32 spheres/boxes, <= 8 planes, <= 2 lights: linear-filter kernels
  std::vector<double> sph_world;       // (cx, cy, cz, R) in cull_sph[] order, FP64, for the per-frame camera table
  DevBuf<int32_t> sph_index, pl_index;
  DevBuf<DevLightF> lights_f;
  int n_sph = 0, n_pl = 0;
  float m_scene = 0.0f, max_distance_f = 0.0f;
  std::vector<uint8_t*> textures;
  // per-frame scratch
  DevBuf<int32_t> tiles;
  std::vector<int32_t> tiles_host;     // global ids ty * stx_count + tx
  size_t tiles_px_count = 0;           // pixels of the window covered by tiles_host
  uint32_t tiles_magic = 0;            // see FrameParams::tiles_magic (0 when the list is not plain row-major)
  DevBuf<double> lens_tab;             // [W + H] per-column / per-row retina offsets
  double lens_key[4] = {0, 0, 0, 0};   // (W, H, retina_width, retina_height) the table was built for
  int tiles_key[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  DevBuf<double> samples, extra_samples, pre_avg, rgb;
  DevBuf<uint32_t> extra_list;
  // RTRB_RNG_MT validation mode
  DevBuf<double> mt_stream;
  DevBuf<uint32_t> mt_offset, mt_count;
  std::vector<double> mt_host;         // the first mt_host.size() draws after init_genrand(mt_seed)
  HostMT mt_gen;
  uint64_t mt_seed = ~0ull;
  size_t mt_uploaded = 0;
  int mt_iterations = 0;               // iterations the last MT frame needed to reach its fixed point
  DevBuf<int32_t> hit;
  DevBuf<uint8_t> rgba;
  bool fb_exported = false;            // its address / IPC handle has been handed out: it may no longer move
  int fb_w = 0, fb_h = 0;
  // last frame
  int last_w = 0, last_h = 0;
  bool last_has_rgb = false, last_has_hit = false, last_rgba_own = true;
  int last_format = RTRB_FMT_RGBA8;
  cudaStream_t last_stream = nullptr;
  bool timing_valid = false;
};

namespace {

// ---- resolve kernels ---------------------------------------------------------------------------
using rtrb::write_pixel;

__device__ __forceinline__ bool slot_to_xy(const FrameParams& P, uint32_t slot, int& x, int& y) {
  return rtrb::decode_pixel(P, slot, x, y);
}

using rtrb::pre_mean;

__global__ void __launch_bounds__(256) resolve_kernel(const __grid_constant__ FrameParams P) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  int x = 0, y = 0;
  const bool valid = slot < (uint32_t)P.n_tiles * RTRB_SUPER_PIXELS && slot_to_xy(P, slot, x, y);
  double ax = 0, ay = 0, az = 0, variance = 0;
  if (valid) pre_mean(P, slot, ax, ay, az, variance);
  const bool adaptive = valid && variance >= P.variant_threshold;
  // warp-level compaction of the pixels that take the extra-sample branch (camera.rb:87-93): one ballot over
  // the whole warp, one atomic per warp for the counter and one for the list, each lane's list position from a
  // prefix popcount of the ballot
  const unsigned amask = __ballot_sync(0xffffffffu, adaptive);
  uint32_t base = 0;
  if (amask != 0u) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(amask) - 1;
    if (lane == leader) {
      const uint32_t n = __popc(amask);
      atomicAdd(&P.counters[RTRB_CNT_ADAPTIVE], (unsigned long long)n);
      if (P.max_samples > P.pre) base = atomicAdd(P.extra_count, n);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (adaptive) {
      if (P.max_samples > P.pre) {
        P.extra_list[base + __popc(amask & ((1u << lane) - 1u))] = slot;
        return;  // finished by resolve_extra_kernel
      }
      // empty extra loop: (average * pre + 0) / max  (camera.rb:93)
      const double fp = (double)P.pre, fm = (double)P.max_samples;
      ax = (ax * fp + 0.0) / fm; ay = (ay * fp + 0.0) / fm; az = (az * fp + 0.0) / fm;
    }
  }
  if (valid) write_pixel(P, x, y, ax, ay, az);
}

__global__ void __launch_bounds__(256) resolve_extra_kernel(const __grid_constant__ FrameParams P) {
  const uint32_t n = *P.extra_count;
  const int E = P.max_samples - P.pre;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const uint32_t slot = P.extra_list[e];
    int x, y;
    if (!slot_to_xy(P, slot, x, y)) continue;
    double ax, ay, az, variance;
    if (P.fuse_resolve == 2) { ax = P.pre_avg[(size_t)e * 3]; ay = P.pre_avg[(size_t)e * 3 + 1]; az = P.pre_avg[(size_t)e * 3 + 2]; }
    else pre_mean(P, slot, ax, ay, az, variance);
    const double* s = P.extra_samples + (size_t)e * E * 3;
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (int j = 0; j < E; ++j) { cx += s[j * 3 + 0]; cy += s[j * 3 + 1]; cz += s[j * 3 + 2]; }
    const double fp = (double)P.pre, fm = (double)P.max_samples;
    ax = (ax * fp + cx) / fm; ay = (ay * fp + cy) / fm; az = (az * fp + cz) / fm;
    write_pixel(P, x, y, ax, ay, az);
  }
}

// Frame control block -> its pinned host mirror by direct stores (zero-copy over PCIe): see rtrb_submit.
__global__ void publish_ctl_kernel(unsigned long long* host, const unsigned long long* dev, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) host[i] = dev[i];
}

__global__ void fill_i32_kernel(int32_t* p, size_t n, int32_t v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- FMA issue microbenchmark ------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
  T r0 = (T)threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
      r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
  }
  T s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
  if (s == (T)-12345.678) out[0] = s;  // never true; keeps the chains alive
}

template <typename T>
int measure_fma(int device, double* tflops_out) {
  CUDA_TRY(cudaSetDevice(device));
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  T* d = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d, sizeof(T)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  fma_peak_kernel<T><<<blocks, threads>>>(d, 64, (T)1.0000001, (T)1e-9);  // warm-up
  g_launches++;
  CUDA_TRY(cudaDeviceSynchronize());
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0));
    fma_peak_kernel<T><<<blocks, threads>>>(d, iters, (T)1.0000001, (T)1e-9);
    g_launches++;
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_out = best;
  return RTRB_OK;
}

// ---- scene baking ------------------------------------------------------------------------------
int validate_scene(const rtrb_scene_desc* s) {
  if (!s) return fail(RTRB_ERR_INVALID, "scene is NULL");
  if (s->n_objects < 0 || s->n_lights < 0 || s->n_textures < 0) return fail(RTRB_ERR_INVALID, "negative count");
  if (s->n_lights > RTRB_MAX_LIGHTS) return fail(RTRB_ERR_UNSUPPORTED, "more than %d lights", RTRB_MAX_LIGHTS);
  if (s->n_objects > 0 && !s->objects) return fail(RTRB_ERR_INVALID, "objects is NULL");
  if (s->n_lights > 0 && !s->lights) return fail(RTRB_ERR_INVALID, "lights is NULL");
  for (int i = 0; i < s->n_objects; ++i) {
    const rtrb_object_desc& o = s->objects[i];
    if (o.type != RTRB_OBJ_PLANE && o.type != RTRB_OBJ_SPHERE && o.type != RTRB_OBJ_BOX)
      return fail(RTRB_ERR_UNSUPPORTED, "object %d: unknown type %d", i, o.type);
    if (o.type == RTRB_OBJ_BOX) {
      H3 c = hcross(h3(o.front), h3(o.up));
      if (hnorm(c) == 0)  // box.rb:23 `front.cross(up).normalize` raises 'zero vector' in World#initialize
        return fail(RTRB_ERR_INVALID, "object %d: box front x up is the zero vector (fast_4d_matrix.c:290 would raise)", i);
    }
    if (o.texture >= s->n_textures) return fail(RTRB_ERR_INVALID, "object %d: texture index out of range", i);
    if (o.type == RTRB_OBJ_SPHERE && !o.has_refraction)
      return fail(RTRB_ERR_INVALID, "object %d: a sphere needs refractive_rate (sphere.rb:93 divides unconditionally)", i);
  }
  for (int i = 0; i < s->n_textures; ++i)
    if (!s->textures[i].rgb8 || s->textures[i].width <= 0 || s->textures[i].height <= 0)
      return fail(RTRB_ERR_INVALID, "texture %d is empty", i);
  return RTRB_OK;
}

int bake_scene(rtrb_renderer* r, const rtrb_scene_desc* s) {
  r->n_objects = s->n_objects;
  r->n_lights = s->n_lights;
  r->max_distance = s->max_distance;
  r->soft_shadow_exponent = s->soft_shadow_exponent;
  {
    bool textured = false;
    for (int i = 0; i < s->n_objects; ++i) textured = textured || (s->objects[i].texture >= 0 && s->objects[i].type != RTRB_OBJ_BOX);
    r->scene_class = 0;
    if (s->n_lights == 1 && s->soft_shadow_exponent == 2.0) {
      r->scene_class |= RTRB_SCENE_CLASS_ONE_LIGHT;
      if (s->lights[0].radius == 0.0 && !textured) r->scene_class |= RTRB_SCENE_CLASS_LEAN;
    }
  }
  for (int i = 0; i < s->n_textures; ++i) {
    const rtrb_texture_desc& t = s->textures[i];
    uint8_t* d = nullptr;
    size_t bytes = (size_t)t.width * t.height * 3;
    CUDA_TRY(cudaMalloc((void**)&d, bytes));
    r->textures.push_back(d);
    CUDA_TRY(cudaMemcpy(d, t.rgb8, bytes, cudaMemcpyHostToDevice));
  }
  std::vector<DevGeom> geom(s->n_objects);
  std::vector<DevMat> mat(s->n_objects);
  std::vector<DevBox> boxes;
  std::vector<double> box_radius;  // filter bound per box (INFINITY = none)
  for (int i = 0; i < s->n_objects; ++i) {
    const rtrb_object_desc& o = s->objects[i];
    DevGeom& g = geom[i];
    DevMat& m = mat[i];
    memset(&g, 0, sizeof(g));
    memset(&m, 0, sizeof(m));
    g.type = o.type;
    g.px = o.point[0]; g.py = o.point[1]; g.pz = o.point[2];
    g.radius = o.type == RTRB_OBJ_SPHERE ? o.radius : 0.0;
    g.nx = o.front[0]; g.ny = o.front[1]; g.nz = o.front[2];
    if (o.type == RTRB_OBJ_BOX) {
      // Box#initialize (box.rb:22-73) in the reference's evaluation order
      const H3 pt = h3(o.point), fr = h3(o.front), up = h3(o.up);
      const H3 left = hnormalize(hcross(fr, up));
      const H3 nfr = H3{-fr.x, -fr.y, -fr.z}, nup = H3{-up.x, -up.y, -up.z}, nleft = H3{-left.x, -left.y, -left.z};
      struct Spec { H3 front, up, point; double uu, vu; };
      const Spec spec[6] = {
          {up, left, hadd(pt, hmul(hmul(up, o.width_up), 0.5)), o.width_front, o.width_left},
          {nup, left, hsub(pt, hmul(hmul(up, o.width_up), 0.5)), o.width_front, o.width_left},
          {fr, up, hadd(pt, hmul(hmul(fr, o.width_front), 0.5)), o.width_left, o.width_up},
          {nfr, up, hsub(pt, hmul(hmul(fr, o.width_front), 0.5)), o.width_left, o.width_up},
          {left, up, hadd(pt, hmul(hmul(left, o.width_left), 0.5)), o.width_front, o.width_up},
          {nleft, up, hsub(pt, hmul(hmul(left, o.width_left), 0.5)), o.width_front, o.width_up}};
      DevBox bx;
      double bound = 0;
      // front _|_ up makes every face frame (left^, up^) orthonormal and in-plane, so a face's accepted
      // region is the rectangle |u|,|v| <= 1/2 around its point; otherwise the filter gets no bound
      const double fu = (fr.x * up.x + fr.y * up.y + fr.z * up.z) / (hnorm(fr) * hnorm(up));
      bool bounded = fabs(fu) < 1e-9;
      for (int k = 0; k < 6; ++k) {
        DevBoxFace& F = bx.f[k];
        const H3 fl = hnormalize(hcross(spec[k].front, spec[k].up));  // Plane#reinit, plane.rb:21-23
        const H3 l2 = hnormalize(fl), u2 = hnormalize(spec[k].up);     // Plane#get_uv, plane.rb:82-83
        F.px = spec[k].point.x; F.py = spec[k].point.y; F.pz = spec[k].point.z;
        F.nx = spec[k].front.x; F.ny = spec[k].front.y; F.nz = spec[k].front.z;
        F.lx = l2.x; F.ly = l2.y; F.lz = l2.z;
        F.ux = u2.x; F.uy = u2.y; F.uz = u2.z;
        F.u_unit = spec[k].uu; F.v_unit = spec[k].vu;
        const double half = 0.5 * sqrt(spec[k].uu * spec[k].uu + spec[k].vu * spec[k].vu);
        const double reach = hnorm(hsub(spec[k].point, pt)) + half;
        if (!(reach == reach) || reach > 1e30) bounded = false;
        bound = fmax(bound, reach);
      }
      box_radius.push_back(bounded ? bound * 1.001 : (double)INFINITY);
      g.aux = (int32_t)boxes.size();
      boxes.push_back(bx);
    }
    for (int k = 0; k < 3; ++k) {
      m.diffuse[k] = o.diffuse_rate[k]; m.refl[k] = o.reflective_attenuation[k];
      m.refr[k] = o.refractive_attenuation[k]; m.ambient[k] = o.ambient[k];
    }
    m.refractive_rate = o.refractive_rate;
    m.has_refraction = o.has_refraction;
    if (o.type == RTRB_OBJ_PLANE) {
      H3 f = h3(o.front);
      m.plane_nn_valid = hnorm(f) != 0 ? 1 : 0;
      if (m.plane_nn_valid) put(m.plane_nn, hnormalize(f));
    }
    m.u_unit = o.u_unit; m.v_unit = o.v_unit;
    m.hscale = o.texture_horizontal_scale; m.vscale = o.texture_vertical_scale;
    m.uoff = o.texture_u_offset; m.voff = o.texture_v_offset;
    if (o.texture >= 0 && o.type != RTRB_OBJ_BOX) {  // a box never samples its texture (box.rb has no local_lighting)
      m.tex = r->textures[o.texture];
      m.tex_w = s->textures[o.texture].width;
      m.tex_h = s->textures[o.texture].height;
      if (o.type == RTRB_OBJ_SPHERE) {
        H3 gw = h3(o.greenwich_vec), np = h3(o.north_pole_vec);
        H3 east = hcross(np, gw);  // sphere.rb:19
        put(m.e0, hnormalize(gw)); put(m.e1, hnormalize(east)); put(m.e2, hnormalize(np));  // sphere.rb:113-115
      } else {
        H3 left = hnormalize(hcross(h3(o.front), h3(o.up)));  // plane.rb:21-23
        put(m.e0, hnormalize(left));                          // plane.rb:82 normalises again
        put(m.e1, hnormalize(h3(o.up)));
      }
    }
  }
  std::vector<DevLight> lights(s->n_lights);
  for (int i = 0; i < s->n_lights; ++i) {
    const rtrb_light_desc& l = s->lights[i];
    DevLight& d = lights[i];
    d.px = l.position[0]; d.py = l.position[1]; d.pz = l.position[2];
    for (int k = 0; k < 3; ++k) {
      d.color[k] = l.color[k];
      d.color_hl[k] = l.color[k] * l.high_light_rate;  // world.rb:94
    }
    d.radius = l.radius;
    d.hl_threshold = l.high_light_angle / 180.0 * 3.141592653589793;  // world.rb:92
  }
  // ---- FP32 filter view: conversions round to nearest; the filter's margins cover that error ----
  std::vector<float4> csph, cpl;
  std::vector<int32_t> isph, ipl;
  std::vector<BvhBuildSphere> bsph;
  float m_scene = 0.0f;
  for (int i = 0; i < s->n_objects; ++i) {
    const rtrb_object_desc& o = s->objects[i];
    if (o.type == RTRB_OBJ_SPHERE || o.type == RTRB_OBJ_BOX) {
      // a box enters the sphere filter through its bounding sphere (flagged: line test only)
      const bool is_box = o.type == RTRB_OBJ_BOX;
      const double rad = is_box ? box_radius[geom[i].aux] : o.radius;
      BvhBuildSphere bs;
      bs.c[0] = o.point[0]; bs.c[1] = o.point[1]; bs.c[2] = o.point[2]; bs.r = rad; bs.world_index = i;
      bs.bound_only = is_box ? 1 : 0;
      bsph.push_back(bs);
      double cm = fmax(fabs(o.point[0]), fmax(fabs(o.point[1]), fabs(o.point[2]))) + (fabs(rad) < 1e30 ? fabs(rad) : 0.0);
      m_scene = fmaxf(m_scene, nextafterf((float)cm, INFINITY));
    } else {
      double n1 = fabs(o.front[0]) + fabs(o.front[1]) + fabs(o.front[2]);
      double pm = fmax(fabs(o.point[0]), fmax(fabs(o.point[1]), fabs(o.point[2])));
      cpl.push_back(make_float4((float)o.front[0], (float)o.front[1], (float)o.front[2], nextafterf((float)n1, INFINITY)));
      cpl.push_back(make_float4((float)o.point[0], (float)o.point[1], (float)o.point[2], nextafterf((float)pm, INFINITY)));
      ipl.push_back(i);
    }
  }
  // the BVH build reorders the spheres so that every leaf is a contiguous run of cull_sph[]
  std::vector<BvhNode> nodes;
  r->small_scene = (int)bsph.size() <= RTRB_BVH_MIN_SPHERES && (int)ipl.size() <= RTRB_K_PLANES && s->n_lights <= RTRB_K_LIGHTS;
  if (!r->small_scene) nodes = rtrb_bvh::build_tree(bsph);
  for (const BvhBuildSphere& bs : bsph) {
    float w = (float)bs.r;
    if (bs.bound_only) {  // rounded up, sign bit set
      w = fabsf(w);
      if ((double)w < fabs(bs.r)) w = nextafterf(w, INFINITY);
      w = -w;
      if (w == 0.0f) w = -0.0f;
    }
    csph.push_back(make_float4((float)bs.c[0], (float)bs.c[1], (float)bs.c[2], w));
    isph.push_back(bs.world_index);
  }
  CUDA_TRY(r->bvh.ensure(std::max<size_t>(1, nodes.size())));
  if (!nodes.empty()) CUDA_TRY(cudaMemcpy(r->bvh.p, nodes.data(), nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice));
  r->n_sph = (int)isph.size(); r->n_pl = (int)ipl.size();
  r->sph_world.clear();
  for (const BvhBuildSphere& bs : bsph) { for (int k = 0; k < 3; ++k) r->sph_world.push_back(bs.c[k]); r->sph_world.push_back(bs.r); }
  memset(&r->k, 0, sizeof(r->k));
  if (r->small_scene && !bsph.empty() && s->n_lights > 0) {
    std::vector<float4> lt((size_t)s->n_lights * bsph.size());
    for (int l = 0; l < s->n_lights; ++l) {
      double lm = fmax(fabs(s->lights[l].position[0]), fmax(fabs(s->lights[l].position[1]), fabs(s->lights[l].position[2])));
      for (size_t k = 0; k < bsph.size(); ++k)
        lt[(size_t)l * bsph.size() + k] = apex_entry(s->lights[l].position, 0.0, bsph[k].c, bsph[k].r, (double)m_scene + lm);
    }
    CUDA_TRY(r->light_tab.ensure(lt.size()));
    CUDA_TRY(cudaMemcpy(r->light_tab.p, lt.data(), lt.size() * sizeof(float4), cudaMemcpyHostToDevice));
    for (int l = 0; l < RTRB_K_LIGHTS; ++l)
      for (int k = 0; k < RTRB_APEX_MAX; ++k)
        r->k.light_tab[l][k] = (l < s->n_lights && k < (int)bsph.size()) ? lt[(size_t)l * bsph.size() + k]
                                                                        : make_float4(0.0f, 0.0f, 0.0f, INFINITY);
    r->k.has_light_tab = 1;
  }
  r->m_scene = m_scene;
  r->max_distance_f = nextafterf((float)s->max_distance, INFINITY);
  std::vector<DevLightF> lf(s->n_lights);
  for (int i = 0; i < s->n_lights; ++i) {
    const rtrb_light_desc& l = s->lights[i];
    double thr = l.high_light_angle / 180.0 * 3.141592653589793;
    lf[i].px = (float)l.position[0]; lf[i].py = (float)l.position[1]; lf[i].pz = (float)l.position[2];
    lf[i].pmax = nextafterf((float)fmax(fabs(l.position[0]), fmax(fabs(l.position[1]), fabs(l.position[2]))), INFINITY);
    lf[i].mode = (thr > 1e-4 && thr < 1.5) ? 1 : 0;
    double c = cos(thr);
    lf[i].cos2_thr = (float)(c * c);
  }
  if (r->small_scene) {
    for (size_t k = 0; k < csph.size(); ++k) r->k.cull_sph[k] = csph[k];
    for (size_t k = 0; k < cpl.size(); ++k) r->k.cull_pl[k] = cpl[k];
    for (size_t k = 0; k < ipl.size(); ++k) r->k.pl_index[k] = ipl[k];
    for (int l = 0; l < s->n_lights; ++l) { r->k.lights[l] = lights[l]; r->k.lights_f[l] = lf[l]; }
  }
  CUDA_TRY(r->cull_sph.ensure(std::max<size_t>(1, csph.size())));
  CUDA_TRY(r->sph_index.ensure(std::max<size_t>(1, isph.size())));
  CUDA_TRY(r->cull_pl.ensure(std::max<size_t>(1, cpl.size())));
  CUDA_TRY(r->pl_index.ensure(std::max<size_t>(1, ipl.size())));
  CUDA_TRY(r->lights_f.ensure(std::max<size_t>(1, lf.size())));
  if (!csph.empty()) {
    CUDA_TRY(cudaMemcpy(r->cull_sph.p, csph.data(), csph.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r->sph_index.p, isph.data(), isph.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  if (!ipl.empty()) {
    CUDA_TRY(cudaMemcpy(r->cull_pl.p, cpl.data(), cpl.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r->pl_index.p, ipl.data(), ipl.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  if (!lf.empty()) CUDA_TRY(cudaMemcpy(r->lights_f.p, lf.data(), lf.size() * sizeof(DevLightF), cudaMemcpyHostToDevice));
  CUDA_TRY(r->geom.ensure(std::max(1, s->n_objects)));
  CUDA_TRY(r->mat.ensure(std::max(1, s->n_objects)));
  CUDA_TRY(r->lights.ensure(std::max(1, s->n_lights)));
  CUDA_TRY(r->boxes.ensure(std::max<size_t>(1, boxes.size())));
  r->n_boxes = (int)boxes.size();
  if (!boxes.empty()) CUDA_TRY(cudaMemcpy(r->boxes.p, boxes.data(), boxes.size() * sizeof(DevBox), cudaMemcpyHostToDevice));
  if (s->n_objects) {
    CUDA_TRY(cudaMemcpy(r->geom.p, geom.data(), geom.size() * sizeof(DevGeom), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r->mat.p, mat.data(), mat.size() * sizeof(DevMat), cudaMemcpyHostToDevice));
  }
  if (s->n_lights)
    CUDA_TRY(cudaMemcpy(r->lights.p, lights.data(), lights.size() * sizeof(DevLight), cudaMemcpyHostToDevice));
  return RTRB_OK;
}

// Camera#lens_func's per-frame constants in the reference's evaluation order (camera.rb:129-151).
void bake_camera(FrameParams& P, const rtrb_camera_desc* c) {
  H3 pos = h3(c->position), up = h3(c->up), front = h3(c->front);
  H3 left = hnormalize(hcross(up, front));
  H3 front_n = hnormalize(front);
  put(P.pos, pos);
  put(P.front, front);
  put(P.left, left);
  put(P.left_n, hnormalize(left));
  put(P.up_n, hnormalize(up));
  put(P.retina_center, hsub(pos, hmul(front_n, c->image_distance)));
  double object_distance = c->focal_distance * c->image_distance / (c->image_distance - c->focal_distance);
  put(P.focal_point, hadd(pos, hmul(front_n, object_distance)));
  P.retina_width = c->retina_width; P.retina_height = c->retina_height;
  P.aperture_radius = c->aperture_radius;
  P.variant_threshold = c->variant_threshold;
  P.width = c->width; P.height = c->height;
  P.pre = c->pre_sample_times; P.max_samples = c->max_sample_times;
  P.trace_depth = c->trace_depth; P.mc = c->monte_carlo_diffusion_times;
}

void enumerate_tiles(int W, int x0, int y0, int x1, int y1, int rank, int world, std::vector<int32_t>& out) {
  const int stx_count = (W + RTRB_SUPER - 1) / RTRB_SUPER;
  out.clear();
  // deal the window's super-tiles round-robin in row-major order: shares differ by at most one tile
  int i = 0;
  for (int ty = y0 / RTRB_SUPER; ty <= (y1 - 1) / RTRB_SUPER; ++ty)
    for (int tx = x0 / RTRB_SUPER; tx <= (x1 - 1) / RTRB_SUPER; ++tx, ++i)
      if (i % world == rank) out.push_back(ty * stx_count + tx);
}

struct FrameTargets {  // where this renderer writes (own buffers, or rank 0's through a peer mapping)
  uint8_t* rgba = nullptr;
  double* rgb = nullptr;
  int32_t* hit = nullptr;
  bool want_rgb = true, want_hit = true;
  bool no_fill = false;  // multi-GPU: the caller pre-fills, ranks must not race on the shared buffer
};

// Own framebuffers for a (w, h) frame.  `rgba` = false when the 8-bit frame goes to a caller-supplied device buffer
// (rgba_device_out, a peer mapping, a pipeline slot).  Once the 8-bit framebuffer's address or IPC handle has been
// handed out (rtrb_framebuffer_device_ptr / _ipc_export) it is pinned: peers may be storing into it, so a frame that
// would need a larger one fails instead of freeing memory other ranks still write to.
int ensure_framebuffers(rtrb_renderer* r, int w, int h, bool rgba, bool rgb, bool hit) {
  size_t px = (size_t)w * h;
  CUDA_TRY(cudaSetDevice(r->device));
  if (rgba && r->rgba.n < px * 4) {
    if (r->fb_exported)
      return fail(RTRB_ERR_INVALID, "the %dx%d frame needs a larger framebuffer than the one whose device pointer / IPC handle "
                                    "was exported (%zu bytes); it cannot be reallocated while peers may write to it", w, h, r->rgba.n);
    CUDA_TRY(r->rgba.ensure(px * 4));
    CUDA_TRY(cudaMemset(r->rgba.p, 0, px * 4));
  }
  if (rgb) CUDA_TRY(r->rgb.ensure(px * 3));
  if (hit) CUDA_TRY(r->hit.ensure(px));
  r->fb_w = w; r->fb_h = h;
  return RTRB_OK;
}

int finish_stats(rtrb_renderer* r, FrameCtl& fc, rtrb_stats* stats_out);

int render_impl(rtrb_renderer* r, const rtrb_camera_desc* cam, const rtrb_render_opts* opts_in,
                const FrameTargets* targets, rtrb_stats* stats_out, FrameCtl* ctl = nullptr) {
  if (!r || !cam) return fail(RTRB_ERR_INVALID, "renderer/camera is NULL");
  rtrb_render_opts opts;
  memset(&opts, 0, sizeof(opts));
  if (opts_in) opts = *opts_in;
  else { opts.seed = 1; opts.precision = RTRB_PREC_DEFAULT; }
  if (cam->width <= 0 || cam->height <= 0) return fail(RTRB_ERR_INVALID, "bad frame size %dx%d", cam->width, cam->height);
  if (cam->pre_sample_times < 1) return fail(RTRB_ERR_INVALID, "pre_sample_times must be >= 1");
  if (cam->max_sample_times < 0 || cam->monte_carlo_diffusion_times < 0 || cam->trace_depth < 0)
    return fail(RTRB_ERR_INVALID, "negative sampling parameter");
  if (opts.rng_mode != RTRB_RNG_CTR && opts.rng_mode != RTRB_RNG_MT) return fail(RTRB_ERR_INVALID, "unknown rng_mode %d", opts.rng_mode);
  const bool mt_mode = opts.rng_mode == RTRB_RNG_MT;
  if (mt_mode) {
    // stream-exact validation mode: serial by construction (every pixel's first draw depends on how many draws
    // all earlier pixels made), so it is blocking, single-GPU, STRICT arithmetic, and iterates to a fixed point
    if (ctl != nullptr) return fail(RTRB_ERR_UNSUPPORTED, "RTRB_RNG_MT frames cannot be pipelined (rtrb_submit)");
    if (opts.tile_world > 1) return fail(RTRB_ERR_UNSUPPORTED, "RTRB_RNG_MT does not combine with a tile partition");
    if ((opts.seed >> 32) != 0) return fail(RTRB_ERR_UNSUPPORTED, "RTRB_RNG_MT: seeds above 32 bits take Ruby's init_by_array path");
    opts.precision = RTRB_PREC_STRICT;
  }
  if (opts.precision != RTRB_PREC_STRICT && opts.precision != RTRB_PREC_FAST64)
    return fail(RTRB_ERR_INVALID, "unknown precision mode %d", opts.precision);
  if (opts.pixel_format != RTRB_FMT_RGBA8 && opts.pixel_format != RTRB_FMT_RGB8 && opts.pixel_format != RTRB_FMT_PNG_RGB8)
    return fail(RTRB_ERR_INVALID, "unknown pixel format %d", opts.pixel_format);
  const int W = cam->width, H = cam->height;
  int x0 = opts.x0, y0 = opts.y0, x1 = opts.x1, y1 = opts.y1;
  if (x0 == 0 && y0 == 0 && x1 == 0 && y1 == 0) { x1 = W; y1 = H; }
  if (x0 < 0 || y0 < 0 || x1 > W || y1 > H || x0 >= x1 || y0 >= y1) return fail(RTRB_ERR_INVALID, "bad window");
  int world = opts.tile_world <= 1 ? 1 : opts.tile_world;
  int rank = world == 1 ? 0 : opts.tile_rank;
  if (rank < 0 || rank >= world) return fail(RTRB_ERR_INVALID, "tile_rank %d outside tile_world %d", rank, world);
  const int stack_need = cam->trace_depth * (1 + cam->monte_carlo_diffusion_times) + 1;
  if (stack_need > rtrb_max_stack_supported())
    return fail(RTRB_ERR_UNSUPPORTED, "trace_depth*(1+mc)+1 = %d exceeds the device stack limit %d", stack_need,
                rtrb_max_stack_supported());

  CUDA_TRY(cudaSetDevice(r->device));
  cudaStream_t stream = opts.stream ? (cudaStream_t)opts.stream : r->stream;

  FrameTargets tg;
  if (targets) tg = *targets;
  else {
    tg.want_rgb = !(opts.skip_outputs & RTRB_SKIP_RGB);
    tg.want_hit = !(opts.skip_outputs & RTRB_SKIP_HIT);
  }
  if (!tg.rgba && opts.rgba_device_out) tg.rgba = (uint8_t*)opts.rgba_device_out;
  {
    int rc = ensure_framebuffers(r, W, H, tg.rgba == nullptr, tg.want_rgb && !tg.rgb, tg.want_hit && !tg.hit);
    if (rc) return rc;
  }
  if (!tg.rgba) tg.rgba = r->rgba.p;
  if (tg.want_rgb && !tg.rgb) tg.rgb = r->rgb.p;
  if (tg.want_hit && !tg.hit) tg.hit = r->hit.p;
  if (!tg.want_rgb) tg.rgb = nullptr;
  if (!tg.want_hit) tg.hit = nullptr;

  // ---- tile list (cached) ----
  const int stx_count = (W + RTRB_SUPER - 1) / RTRB_SUPER;
  int key[8] = {W, H, x0, y0, x1, y1, rank, world};
  if (memcmp(key, r->tiles_key, sizeof(key)) != 0) {
    enumerate_tiles(W, x0, y0, x1, y1, rank, world, r->tiles_host);
    CUDA_TRY(r->tiles.ensure(std::max<size_t>(1, r->tiles_host.size())));
    if (!r->tiles_host.empty()) {
      std::vector<int32_t> packed(r->tiles_host.size());
      for (size_t i = 0; i < packed.size(); ++i)
        packed[i] = (r->tiles_host[i] % stx_count) | ((r->tiles_host[i] / stx_count) << 16);
      CUDA_TRY(cudaMemcpyAsync(r->tiles.p, packed.data(), packed.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    memcpy(r->tiles_key, key, sizeof(key));
    r->tiles_magic = 0;
    if (stx_count > 0 && stx_count < 65536 && r->tiles_host.size() < 65536) {
      const uint32_t magic = (uint32_t)((1ull << 32) / (unsigned)stx_count) + 1u;
      bool ok = true;
      for (size_t k = 0; k < r->tiles_host.size() && ok; ++k) {
        const uint32_t ty = (uint32_t)(((unsigned long long)k * magic) >> 32);
        ok = r->tiles_host[k] == (int32_t)k && ty == (uint32_t)k / (unsigned)stx_count;
      }
      if (ok) r->tiles_magic = magic;
    }
    r->tiles_px_count = 0;
    for (int b : r->tiles_host) {
      int tx = b % stx_count, ty = b / stx_count;
      int ax0 = std::max(x0, tx * RTRB_SUPER), ax1 = std::min(x1, (tx + 1) * RTRB_SUPER);
      int ay0 = std::max(y0, ty * RTRB_SUPER), ay1 = std::min(y1, (ty + 1) * RTRB_SUPER);
      if (ax1 > ax0 && ay1 > ay0) r->tiles_px_count += (size_t)(ax1 - ax0) * (ay1 - ay0);
    }
  }
  // ---- lens tables (cached): the per-column / per-row scalars of camera.rb:133-134, evaluated on the
  // host in the reference's order so the device does two loads instead of two FP64 divisions per sample
  {
    double lk[4] = {(double)W, (double)H, cam->retina_width, cam->retina_height};
    if (memcmp(lk, r->lens_key, sizeof(lk)) != 0 || !r->lens_tab.p) {
      std::vector<double> tab((size_t)W + H);
      for (int x = 0; x < W; ++x) tab[x] = 2.0 * ((double)x / W - 0.5) * cam->retina_width;
      for (int y = 0; y < H; ++y) tab[(size_t)W + y] = 2 * ((double)y / H - 0.5) * cam->retina_height;
      CUDA_TRY(r->lens_tab.ensure(tab.size()));
      CUDA_TRY(cudaMemcpyAsync(r->lens_tab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      memcpy(r->lens_key, lk, sizeof(lk));
    }
  }
  const int n_tiles = (int)r->tiles_host.size();
  const size_t n_slots = (size_t)n_tiles * RTRB_SUPER_PIXELS;
  const int S = cam->pre_sample_times;
  const int E = cam->max_sample_times > S ? cam->max_sample_times - S : 0;
  // Who finishes render_at (FrameParams::fuse_resolve).  The FAST64 ray-tree kernels resolve a pixel inside the CTA
  // that traced its samples whenever the sample count divides the CTA size: no FP64 sample buffer (24 B per sample:
  // 12.7 GB for one 4K / 64 spp frame) and no resolve launch.
  const bool strict_mode = opts.precision == RTRB_PREC_STRICT;
  int fuse = (S == 1 && cam->variant_threshold > 0) ? 1 : 0;
#ifndef RTRB_NO_FUSE2  // (build-time A/B switch, profiles/README.md)
  if (!strict_mode && !mt_mode && cam->trace_depth > 1 && S <= RTRB_TREE_MIN_BLOCK && RTRB_TREE_MIN_BLOCK % S == 0) fuse = 2;
#endif
  const bool need_samples = fuse == 0;
  if ((need_samples && n_slots * (size_t)S * 3 * sizeof(double) > ((size_t)96 << 30)) || n_slots * (size_t)E * 3 * sizeof(double) > ((size_t)64 << 30))
    return fail(RTRB_ERR_UNSUPPORTED, "sample buffer would exceed the per-frame memory budget");
  if (need_samples) CUDA_TRY(r->samples.ensure(std::max<size_t>(1, n_slots * S * 3)));
  CUDA_TRY(r->extra_list.ensure(std::max<size_t>(1, n_slots)));
  if (E > 0) CUDA_TRY(r->extra_samples.ensure(n_slots * E * 3));
  if (E > 0 && fuse == 2) CUDA_TRY(r->pre_avg.ensure(n_slots * 3));
  FrameCtl& fc = ctl ? *ctl : r->main_ctl;

  FrameParams P;
  memset(&P, 0, sizeof(P));
  bake_camera(P, cam);
  P.max_distance = r->max_distance; P.soft_shadow_exponent = r->soft_shadow_exponent;
  P.n_objects = r->n_objects; P.n_lights = r->n_lights;
  P.geom = r->geom.p; P.mat = r->mat.p; P.lights = r->lights.p; P.boxes = r->boxes.p;
  P.cull_sph = r->cull_sph.p; P.sph_index = r->sph_index.p; P.cull_pl = r->cull_pl.p; P.pl_index = r->pl_index.p;
  P.lights_f = r->lights_f.p; P.n_sph = r->n_sph; P.n_pl = r->n_pl; P.bvh = r->bvh.p;
  P.m_scene = r->m_scene; P.max_distance_f = r->max_distance_f;
  P.use_bvh = r->small_scene ? 0 : 1;
  P.scene_class = r->scene_class;
  if (r->small_scene) {
    static_assert(sizeof(r->k.light_tab) == sizeof(P.k_light_tab) && sizeof(r->k.lights) == sizeof(P.k_lights), "table layout");
    memcpy(P.k_light_tab, r->k.light_tab, sizeof(P.k_light_tab));
    memcpy(P.k_cull_sph, r->k.cull_sph, sizeof(P.k_cull_sph));
    memcpy(P.k_cull_pl, r->k.cull_pl, sizeof(P.k_cull_pl));
    memcpy(P.k_lights, r->k.lights, sizeof(P.k_lights));
    memcpy(P.k_lights_f, r->k.lights_f, sizeof(P.k_lights_f));
    memcpy(P.k_pl_index, r->k.pl_index, sizeof(P.k_pl_index));
    P.k_has_light_tab = r->k.has_light_tab;
  }
  P.light_tab = r->light_tab.p;  // nullptr unless the scene uses the linear filter
  P.cam_tab_valid = 0;
  if (!P.use_bvh && r->n_sph > 0 && r->n_sph <= RTRB_APEX_MAX) {
    const double cm = fmax(fabs(cam->position[0]), fmax(fabs(cam->position[1]), fabs(cam->position[2])));
    for (int k = 0; k < r->n_sph; ++k)
      P.cam_tab[k] = apex_entry(cam->position, fabs(cam->aperture_radius), &r->sph_world[4 * (size_t)k],
                                r->sph_world[4 * (size_t)k + 3], (double)r->m_scene + cm + fabs(cam->aperture_radius));
    for (int k = r->n_sph; k < RTRB_APEX_MAX; ++k) P.cam_tab[k] = make_float4(0.0f, 0.0f, 0.0f, INFINITY);  // never survives
    P.cam_tab_valid = 1;
  }
  P.key0 = (uint32_t)opts.seed; P.key1 = (uint32_t)(opts.seed >> 32);
  P.x0 = x0; P.y0 = y0; P.x1 = x1; P.y1 = y1;
  P.n_tiles = n_tiles; P.stx_count = stx_count; P.tiles = r->tiles.p;
  P.lens_sx = r->lens_tab.p; P.lens_sy = r->lens_tab.p + W;
  P.samples = need_samples ? r->samples.p : nullptr; P.rgb = tg.rgb; P.hit = tg.hit; P.rgba = tg.rgba;
  P.pre_avg = r->pre_avg.p;
  P.counters = fc.d.p; P.first_bad = fc.d.p + RTRB_CNT_N;
  P.work_counter = fc.d.p + RTRB_CNT_N + 1;
  P.status = reinterpret_cast<uint32_t*>(fc.d.p + RTRB_CNT_N + 3);
  P.extra_count = reinterpret_cast<uint32_t*>(fc.d.p + RTRB_CNT_N + 4);
  P.hot = fc.d.p + RTRB_CNT_N + 5;
  P.tiles_magic = (world == 1 && x0 == 0 && y0 == 0 && x1 == W && y1 == H) ? r->tiles_magic : 0u;
  P.extra_list = r->extra_list.p; P.extra_samples = r->extra_samples.p;
  // the Box code lives only in the full-counter kernel variants (rtrb_trace.cuh, trace_dispatch)
  P.count_detail = (opts.count_detail || r->n_boxes > 0) ? 1 : 0;
  P.pixel_format = opts.pixel_format;
  // (1: a single sample with a positive threshold can never take the adaptive branch, variance == 0)
  P.fuse_resolve = fuse;

  const bool strict = strict_mode;
  // timing events only for the blocking calls that return stats: a timestamp between two kernels keeps
  // frame i+1 from starting under frame i's tail, so pipelined frames (rtrb_submit) are not timed and
  // report device_ms = trace_ms = 0
  const bool timed = stats_out != nullptr;
  fc.timed = timed;
  if (timed) CUDA_TRY(cudaEventRecord(fc.ev0, stream));
  CUDA_TRY(cudaMemsetAsync(fc.d.p, 0, RTRB_FCB_WORDS * sizeof(unsigned long long), stream));
  const bool partial = !(x0 == 0 && y0 == 0 && x1 == W && y1 == H && world == 1);
  if (partial && !tg.no_fill && tg.hit == r->hit.p && tg.hit) {
    size_t n = (size_t)W * H;
    fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tg.hit, n, -3);
    g_launches++;
  }
  if (n_tiles > 0 && mt_mode) {
    // ---- RTRB_RNG_MT: offsets = exclusive prefix sums of the per-pixel draw counts, iterated until stable.
    // The first pixel whose offset is wrong is right after the next pass (its predecessors no longer move), so
    // the loop ends after at most one pass per pixel; in practice a handful of passes per data-dependent pixel.
    const size_t npx = (size_t)(x1 - x0) * (size_t)(y1 - y0);
    if (npx * (size_t)std::max(S, cam->max_sample_times) > 0x3fffffffull) return fail(RTRB_ERR_UNSUPPORTED, "window too large for RTRB_RNG_MT");
    CUDA_TRY(r->mt_offset.ensure(npx));
    CUDA_TRY(r->mt_count.ensure(npx));
    std::vector<uint32_t> off(npx), cnt(npx), nxt(npx);
    for (size_t i = 0; i < npx; ++i) off[i] = (uint32_t)(i * (size_t)S);
    if (r->mt_seed != opts.seed) {
      r->mt_host.clear(); r->mt_gen.seed((uint32_t)opts.seed); r->mt_seed = opts.seed; r->mt_uploaded = 0;
    }
    size_t want_len = npx * (size_t)(S + 1) + 4096;
    const int max_iter = 200000;
    int it = 0;
    for (;; ++it) {
      if (it >= max_iter) return fail(RTRB_ERR_UNSUPPORTED, "RTRB_RNG_MT did not reach its fixed point in %d passes", max_iter);
      if (r->mt_host.size() < want_len) {
        r->mt_host.reserve(want_len);
        while (r->mt_host.size() < want_len) r->mt_host.push_back(r->mt_gen.res53());
      }
      if (r->mt_uploaded < r->mt_host.size()) {
        CUDA_TRY(r->mt_stream.ensure(r->mt_host.size()));
        CUDA_TRY(cudaMemcpyAsync(r->mt_stream.p, r->mt_host.data(), r->mt_host.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        r->mt_uploaded = r->mt_host.size();
      }
      P.mt_stream = r->mt_stream.p; P.mt_len = (uint32_t)std::min<size_t>(r->mt_host.size(), 0xfffffff0u);
      P.mt_offset = r->mt_offset.p; P.mt_count = r->mt_count.p;
      CUDA_TRY(cudaMemcpyAsync(r->mt_offset.p, off.data(), npx * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
      CUDA_TRY(cudaMemsetAsync(fc.d.p, 0, RTRB_FCB_WORDS * sizeof(unsigned long long), stream));
      if (timed) CUDA_TRY(cudaEventRecord(fc.evt0, stream));
      CUDA_TRY(rtrb_launch_trace_mt_strict(P, stack_need, stream));
      if (timed) CUDA_TRY(cudaEventRecord(fc.evt1, stream));
      g_launches++;
      CUDA_TRY(cudaMemcpyAsync(cnt.data(), r->mt_count.p, npx * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      bool overflow = false;
      for (size_t i = 0; i < npx && !overflow; ++i) overflow = cnt[i] == 0xFFFFFFFFu;
      if (overflow) {  // some pixel ran past the end of the stream: generate more and repeat the pass
        want_len = r->mt_host.size() * 2;
        if (want_len > 0xfffffff0ull) return fail(RTRB_ERR_UNSUPPORTED, "RTRB_RNG_MT stream would exceed 2^32 draws");
        continue;
      }
      size_t acc = 0;
      bool same = true;
      for (size_t i = 0; i < npx; ++i) {
        nxt[i] = (uint32_t)acc;
        same = same && nxt[i] == off[i];
        acc += cnt[i];
      }
      if (acc > 0xfffffff0ull) return fail(RTRB_ERR_UNSUPPORTED, "RTRB_RNG_MT stream would exceed 2^32 draws");
      if (same) break;
      off.swap(nxt);
      want_len = std::max(want_len, acc + 4096);
    }
    r->mt_iterations = it + 1;
  } else if (n_tiles > 0) {
    if (timed) CUDA_TRY(cudaEventRecord(fc.evt0, stream));
    CUDA_TRY(strict ? rtrb_launch_trace_pre_strict(P, stack_need, stream) : rtrb_launch_trace_pre_fast(P, stack_need, stream));
    if (timed) CUDA_TRY(cudaEventRecord(fc.evt1, stream));
    g_launches++;
    if (!P.fuse_resolve) {
      resolve_kernel<<<(unsigned)((n_slots + 255) / 256), 256, 0, stream>>>(P);
      CUDA_TRY(cudaGetLastError());
      g_launches++;
    }
    if (E > 0 && P.fuse_resolve != 1) {
      CUDA_TRY(strict ? rtrb_launch_trace_extra_strict(P, stack_need, stream) : rtrb_launch_trace_extra_fast(P, stack_need, stream));
      g_launches++;
      resolve_extra_kernel<<<296, 256, 0, stream>>>(P);
      CUDA_TRY(cudaGetLastError());
      g_launches++;
    }
  }
  CUDA_TRY(cudaEventRecord(fc.ev1, stream));
  r->last_w = W; r->last_h = H;
  r->last_rgba_own = tg.rgba == r->rgba.p;
  r->last_has_rgb = tg.rgb == r->rgb.p && tg.rgb != nullptr;
  r->last_has_hit = tg.hit == r->hit.p && tg.hit != nullptr;
  r->last_stream = stream;
  r->last_format = opts.pixel_format;

  // facts needed to finish the stats once the control block has been copied back
  fc.W = W; fc.H = H; fc.S = S; fc.E = E; fc.n_tiles = n_tiles; fc.detail = P.count_detail != 0;
  fc.px_count = r->tiles_px_count;
  if (stats_out) {
    CUDA_TRY(cudaMemcpyAsync(fc.h, fc.d.p, RTRB_FCB_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return finish_stats(r, fc, stats_out);
  }
  return RTRB_OK;
}

// Turns a copied-back control block into rtrb_stats (+ RTRB_ERR_RAISED when a status bit is set).
int finish_stats(rtrb_renderer* r, FrameCtl& fc, rtrb_stats* stats_out) {
  (void)r;
  const unsigned long long* c = fc.h;
  const uint32_t* st = reinterpret_cast<const uint32_t*>(fc.h + RTRB_CNT_N + 3);
  memset(stats_out, 0, sizeof(*stats_out));
  stats_out->rays = c[RTRB_CNT_RAYS]; stats_out->shadow_queries = c[RTRB_CNT_SHADOW];
  for (int sl = 0; sl < RTRB_HOT_SLICES; ++sl) {
    stats_out->rays += c[RTRB_CNT_N + 5 + 2 * sl];
    stats_out->shadow_queries += c[RTRB_CNT_N + 5 + 2 * sl + 1];
  }
  stats_out->highlight_hits = c[RTRB_CNT_HIGHLIGHT]; stats_out->hits = c[RTRB_CNT_HITS];
  stats_out->local_shaded = c[RTRB_CNT_LOCAL]; stats_out->lit_lights = c[RTRB_CNT_LIT];
  stats_out->mc_rays = c[RTRB_CNT_MC]; stats_out->refractions = c[RTRB_CNT_REFR];
  stats_out->texel_fetches = c[RTRB_CNT_TEXEL];
  stats_out->sphere_tests = c[RTRB_CNT_SPH_TEST]; stats_out->sphere_accepts = c[RTRB_CNT_SPH_ACC];
  stats_out->plane_tests = c[RTRB_CNT_PL_TEST]; stats_out->plane_accepts = c[RTRB_CNT_PL_ACC];
  stats_out->cover_sphere = c[RTRB_CNT_COV_SPH]; stats_out->cover_sphere_full = c[RTRB_CNT_COV_SPH_FULL];
  stats_out->cover_sphere_penumbra = c[RTRB_CNT_COV_SPH_PEN];
  stats_out->cover_plane = c[RTRB_CNT_COV_PL]; stats_out->cover_plane_accepts = c[RTRB_CNT_COV_PL_ACC];
  stats_out->adaptive_pixels = c[RTRB_CNT_ADAPTIVE]; stats_out->exact_tests = c[RTRB_CNT_EXACT];
  stats_out->box_tests = c[RTRB_CNT_BOX_TEST]; stats_out->box_accepts = c[RTRB_CNT_BOX_ACC];
  stats_out->cover_box = c[RTRB_CNT_COV_BOX]; stats_out->cover_box_accepts = c[RTRB_CNT_COV_BOX_ACC];
  // samples: counted on the device with count_detail; otherwise a pure function of the window
  stats_out->samples = fc.detail ? c[RTRB_CNT_SAMPLES]
                                 : (uint64_t)fc.px_count * fc.S + (uint64_t)(fc.E > 0 ? c[RTRB_CNT_ADAPTIVE] * (uint64_t)fc.E : 0);
  stats_out->status = st[0];
  stats_out->max_stack = st[1] > 1u ? st[1] : (stats_out->rays ? 1u : 0u);  // blocks only report depths > 1
  if (st[0] && c[RTRB_CNT_N] != 0ull) {  // stored complemented so that an all-zero block means "none"
    const unsigned long long key = ~c[RTRB_CNT_N];
    stats_out->first_bad_x = (int32_t)(key / (unsigned long long)fc.H);
    stats_out->first_bad_y = (int32_t)(key % (unsigned long long)fc.H);
  } else {
    stats_out->first_bad_x = stats_out->first_bad_y = -1;
  }
  float ms = 0;
  if (fc.timed) {
    CUDA_TRY(cudaEventElapsedTime(&ms, fc.ev0, fc.ev1));
    stats_out->device_ms = ms;
  }
  if (fc.timed && fc.n_tiles > 0) {
    CUDA_TRY(cudaEventElapsedTime(&ms, fc.evt0, fc.evt1));
    stats_out->trace_ms = ms;
  }
  if (st[0]) {
    fail(RTRB_ERR_RAISED, "the reference would have raised: status 0x%x at pixel (%d, %d)", st[0],
         stats_out->first_bad_x, stats_out->first_bad_y);
    return RTRB_ERR_RAISED;
  }
  return RTRB_OK;
}

}  // namespace

// ================================================================================================
extern "C" {
#pragma GCC visibility push(default)

int rtrb_abi_version(void) { return RTRB_ABI_VERSION; }
const char* rtrb_last_error(void) { return g_last_error.c_str(); }
uint64_t rtrb_launch_count(void) { return g_launches.load(); }

int rtrb_device_count(int* count_out) {
  if (!count_out) return fail(RTRB_ERR_INVALID, "count_out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count_out = 0; return fail(RTRB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
  *count_out = n;
  return RTRB_OK;
}

int rtrb_renderer_create(const rtrb_scene_desc* scene, int device, rtrb_renderer** out) {
  if (!out) return fail(RTRB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = validate_scene(scene);
  if (rc) return rc;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(RTRB_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(RTRB_ERR_INVALID, "device %d out of range (%d devices)", device, n);
  CUDA_TRY(cudaSetDevice(device));
  rtrb_renderer* r = new rtrb_renderer();
  r->device = device;
  cudaError_t ce;
  if ((ce = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (ce = cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (ce = cudaEventCreateWithFlags(&r->push_ev, cudaEventDisableTiming)) != cudaSuccess ||
      (ce = cudaEventCreateWithFlags(&r->push_done_ev, cudaEventDisableTiming)) != cudaSuccess ||
      (ce = (cudaError_t)r->main_ctl.init()) != cudaSuccess) {
    rtrb_renderer_destroy(r);
    return fail(RTRB_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(ce));
  }
  rc = bake_scene(r, scene);
  if (rc) { rtrb_renderer_destroy(r); return rc; }
  *out = r;
  return RTRB_OK;
}

int rtrb_renderer_destroy(rtrb_renderer* r) {
  if (!r) return RTRB_OK;
  cudaSetDevice(r->device);
  cudaDeviceSynchronize();
  r->mt_stream.release(); r->mt_offset.release(); r->mt_count.release();
  r->geom.release(); r->mat.release(); r->lights.release(); r->lens_tab.release(); r->boxes.release();
  r->bvh.release(); r->light_tab.release();
  r->cull_sph.release(); r->cull_pl.release(); r->sph_index.release(); r->pl_index.release(); r->lights_f.release(); r->tiles.release(); r->samples.release();
  r->extra_samples.release(); r->pre_avg.release(); r->rgb.release(); r->extra_list.release(); r->hit.release(); r->rgba.release();
  r->main_ctl.destroy();
  for (int i = 0; i < RTRB_PIPE_SLOTS; ++i) r->pipe_ctl[i].destroy();
  for (uint8_t* t : r->textures) cudaFree(t);
  if (r->push_ev) cudaEventDestroy(r->push_ev);
  if (r->push_done_ev) cudaEventDestroy(r->push_done_ev);
  if (r->stream) cudaStreamDestroy(r->stream);
  if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
  delete r;
  return RTRB_OK;
}

int rtrb_render_device(rtrb_renderer* r, const rtrb_camera_desc* cam, const rtrb_render_opts* opts,
                       rtrb_stats* stats_out) {
  return render_impl(r, cam, opts, nullptr, stats_out);
}

int rtrb_download(rtrb_renderer* r, uint8_t* rgba, double* rgb_or_null, int32_t* hit_or_null) {
  if (!r) return fail(RTRB_ERR_INVALID, "renderer is NULL");
  if (r->last_w == 0) return fail(RTRB_ERR_INVALID, "nothing rendered yet");
  CUDA_TRY(cudaSetDevice(r->device));
  cudaStream_t s = r->last_stream ? r->last_stream : r->stream;
  size_t px = (size_t)r->last_w * r->last_h;
  if (rgba) {
    if (!r->last_rgba_own) return fail(RTRB_ERR_INVALID, "the last frame was written to rgba_device_out, not to the renderer's framebuffer");
    CUDA_TRY(cudaMemcpyAsync(rgba, r->rgba.p, RTRB_FRAME_BYTES(r->last_w, r->last_h, r->last_format), cudaMemcpyDeviceToHost, s));
  }
  if (rgb_or_null) {
    if (!r->last_has_rgb) return fail(RTRB_ERR_INVALID, "the last frame kept no float RGB");
    CUDA_TRY(cudaMemcpyAsync(rgb_or_null, r->rgb.p, px * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  if (hit_or_null) {
    if (!r->last_has_hit) return fail(RTRB_ERR_INVALID, "the last frame kept no hit ids");
    CUDA_TRY(cudaMemcpyAsync(hit_or_null, r->hit.p, px * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return RTRB_OK;
}

int rtrb_render(rtrb_renderer* r, const rtrb_camera_desc* cam, const rtrb_render_opts* opts, uint8_t* rgba,
                double* rgb_or_null, int32_t* hit_or_null, rtrb_stats* stats_out) {
  if (!rgba) return fail(RTRB_ERR_INVALID, "rgba is NULL");
  if (opts && opts->rgba_device_out)
    return fail(RTRB_ERR_INVALID, "rgba_device_out does not combine with a host-buffer call (use rtrb_render_device)");
  FrameTargets tg;
  tg.want_rgb = rgb_or_null != nullptr;
  tg.want_hit = hit_or_null != nullptr;
  rtrb_stats local;
  int rc = render_impl(r, cam, opts, &tg, stats_out ? stats_out : &local);
  if (rc != RTRB_OK && rc != RTRB_ERR_RAISED) return rc;
  std::string keep = g_last_error;
  int rc2 = rtrb_download(r, rgba, rgb_or_null, hit_or_null);
  if (rc2) return rc2;
  if (rc == RTRB_ERR_RAISED) g_last_error = keep;
  return rc;
}

int rtrb_submit(rtrb_renderer* r, const rtrb_camera_desc* cam, const rtrb_render_opts* opts, uint8_t* rgba_host,
                int* ticket_out) {
  if (!r || !cam || !rgba_host || !ticket_out) return fail(RTRB_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(r->device));
  if (!r->pipe_ready) {
    for (int i = 0; i < RTRB_PIPE_SLOTS; ++i) {
      cudaError_t e = (cudaError_t)r->pipe_ctl[i].init();
      if (e != cudaSuccess) return fail(RTRB_ERR_CUDA, "pipeline slot init failed: %s", cudaGetErrorString(e));
    }
    r->pipe_ready = true;
  }
  const unsigned ticket = r->next_ticket;
  FrameCtl& fc = r->pipe_ctl[ticket % RTRB_PIPE_SLOTS];
  if (fc.in_flight) return fail(RTRB_ERR_INVALID, "%d frames are already in flight: call rtrb_wait first", RTRB_PIPE_SLOTS);
  const size_t bytes = RTRB_FRAME_BYTES(cam->width, cam->height, opts ? opts->pixel_format : RTRB_FMT_RGBA8);
  const size_t slot_bytes = (size_t)cam->width * cam->height * 4;
  if (cam->width > 0 && cam->height > 0 && fc.rgba.n < slot_bytes) {
    CUDA_TRY(fc.rgba.ensure(slot_bytes));
    CUDA_TRY(cudaMemsetAsync(fc.rgba.p, 0, slot_bytes, r->stream));
  }
  rtrb_render_opts o;
  memset(&o, 0, sizeof(o));
  if (opts) o = *opts;
  else { o.seed = 1; o.precision = RTRB_PREC_DEFAULT; }
  o.stream = nullptr;  // the pipeline owns its streams
  FrameTargets tg;
  tg.rgba = o.rgba_device_out ? (uint8_t*)o.rgba_device_out : fc.rgba.p;
  tg.want_rgb = false; tg.want_hit = false;
  int rc = render_impl(r, cam, &o, &tg, nullptr, &fc);
  if (rc) return rc;
  // the copy stream picks the frame up as soon as its kernels are done; the render stream is free
  // to start the next frame into the other slot meanwhile
  CUDA_TRY(cudaStreamWaitEvent(r->copy_stream, fc.ev1, 0));
  CUDA_TRY(cudaMemcpyAsync(rgba_host, tg.rgba, bytes, cudaMemcpyDeviceToHost, r->copy_stream));
  CUDA_TRY(cudaEventRecord(fc.copied, r->copy_stream));
  // The 1.3 KB control block does not go through the copy engine at all: a one-block kernel behind the frame's kernels
  // stores it straight into the pinned (device-mapped) host mirror.  As a DMA of its own - even on its own stream - it
  // sat between the 6.2 MB frame copies of a PCIe-bound sequence and cost 6 % of the frame rate (8 057 -> 8 540 frames/s
  // on config 2); behind the frame on the copy stream it cost 14 % (round 1).
  publish_ctl_kernel<<<1, 256, 0, r->stream>>>(fc.h_dev, fc.d.p, RTRB_FCB_WORDS);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  CUDA_TRY(cudaEventRecord(fc.ctl_copied, r->stream));
  // (the slot is reused only after rtrb_wait has host-synchronised on `copied`, so no stream wait is needed)
  fc.in_flight = true;
  fc.ticket = ticket;
  r->next_ticket = ticket + 1;
  *ticket_out = (int)ticket;
  return RTRB_OK;
}

int rtrb_wait(rtrb_renderer* r, int ticket, rtrb_stats* stats_out) {
  if (!r || !r->pipe_ready) return fail(RTRB_ERR_INVALID, "nothing submitted");
  FrameCtl& fc = r->pipe_ctl[(unsigned)ticket % RTRB_PIPE_SLOTS];
  if (!fc.in_flight || fc.ticket != (unsigned)ticket) return fail(RTRB_ERR_INVALID, "ticket %d is not in flight", ticket);
  CUDA_TRY(cudaSetDevice(r->device));
  CUDA_TRY(cudaEventSynchronize(fc.copied));
  CUDA_TRY(cudaEventSynchronize(fc.ctl_copied));
  fc.in_flight = false;
  rtrb_stats local;
  return finish_stats(r, fc, stats_out ? stats_out : &local);
}

int rtrb_framebuffer_device_ptr(rtrb_renderer* r, int width, int height, void** ptr_out) {
  if (!r || !ptr_out || width <= 0 || height <= 0) return fail(RTRB_ERR_INVALID, "bad argument");
  int rc = ensure_framebuffers(r, width, height, true, false, false);
  if (rc) return rc;
  r->fb_exported = true;
  *ptr_out = r->rgba.p;
  return RTRB_OK;
}

int rtrb_framebuffer_download(rtrb_renderer* r, int width, int height, uint8_t* rgba_host) {
  if (!r || !rgba_host || width <= 0 || height <= 0) return fail(RTRB_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(r->device));
  const size_t bytes = (size_t)width * height * 4;
  if (r->rgba.n < bytes) return fail(RTRB_ERR_INVALID, "framebuffer is smaller than %dx%d", width, height);
  CUDA_TRY(cudaMemcpyAsync(rgba_host, r->rgba.p, bytes, cudaMemcpyDeviceToHost, r->copy_stream));
  CUDA_TRY(cudaStreamSynchronize(r->copy_stream));
  return RTRB_OK;
}

int rtrb_framebuffer_copy_async(rtrb_renderer* r, size_t bytes, uint8_t* host, void* stream) {
  if (!r || !host) return fail(RTRB_ERR_INVALID, "bad argument");
  if (bytes == 0) return RTRB_OK;
  CUDA_TRY(cudaSetDevice(r->device));
  if (r->rgba.n < bytes) return fail(RTRB_ERR_INVALID, "framebuffer is smaller than %zu bytes", bytes);
  CUDA_TRY(cudaMemcpyAsync(host, r->rgba.p, bytes, cudaMemcpyDeviceToHost, stream ? (cudaStream_t)stream : r->stream));
  return RTRB_OK;
}

int rtrb_framebuffer_ipc_export(rtrb_renderer* r, int width, int height, uint8_t handle_out[64]) {
  if (!r || !handle_out) return fail(RTRB_ERR_INVALID, "bad argument");
  int rc = ensure_framebuffers(r, width, height, true, false, false);
  if (rc) return rc;
  r->fb_exported = true;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, r->rgba.p));
  memcpy(handle_out, &h, 64);
  return RTRB_OK;
}

int rtrb_ipc_open(int device, const uint8_t handle[64], void** ptr_out) {
  if (!handle || !ptr_out) return fail(RTRB_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return RTRB_OK;
}

int rtrb_ipc_close(int device, void* ptr) {
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return RTRB_OK;
}

int rtrb_peer_push(rtrb_renderer* r, const void* src, void* dst_peer, size_t bytes, void* after_stream) {
  if (!r || !src || !dst_peer) return fail(RTRB_ERR_INVALID, "bad argument");
  if (bytes == 0) return RTRB_OK;
  CUDA_TRY(cudaSetDevice(r->device));
  cudaStream_t after = after_stream ? (cudaStream_t)after_stream : r->stream;
  CUDA_TRY(cudaEventRecord(r->push_ev, after));
  CUDA_TRY(cudaStreamWaitEvent(r->copy_stream, r->push_ev, 0));
  CUDA_TRY(cudaMemcpyAsync(dst_peer, src, bytes, cudaMemcpyDefault, r->copy_stream));  // copy engine over NVLink
  return RTRB_OK;
}

int rtrb_peer_push_join(rtrb_renderer* r, void* stream) {
  if (!r) return fail(RTRB_ERR_INVALID, "renderer is NULL");
  CUDA_TRY(cudaSetDevice(r->device));
  CUDA_TRY(cudaEventRecord(r->push_done_ev, r->copy_stream));
  CUDA_TRY(cudaStreamWaitEvent(stream ? (cudaStream_t)stream : r->stream, r->push_done_ev, 0));
  return RTRB_OK;
}

int rtrb_render_multi(rtrb_renderer* const* renderers, int n, const rtrb_camera_desc* cam,
                      const rtrb_render_opts* opts_in, uint8_t* rgba, double* rgb_or_null, int32_t* hit_or_null,
                      rtrb_stats* stats_out) {
  if (!renderers || n < 1 || !cam || !rgba) return fail(RTRB_ERR_INVALID, "bad argument");
  if (opts_in && opts_in->rgba_device_out)
    return fail(RTRB_ERR_INVALID, "rgba_device_out does not combine with a host-buffer call (use rtrb_render_device)");
  if (n == 1) return rtrb_render(renderers[0], cam, opts_in, rgba, rgb_or_null, hit_or_null, stats_out);
  rtrb_renderer* root = renderers[0];
  const bool want_rgb = rgb_or_null != nullptr, want_hit = hit_or_null != nullptr;
  int rc = ensure_framebuffers(root, cam->width, cam->height, true, want_rgb, want_hit);
  if (rc) return rc;
  // peer access root <- others (writes land in root's framebuffer over NVLink; no collective)
  for (int i = 1; i < n; ++i) {
    int can = 0;
    CUDA_TRY(cudaDeviceCanAccessPeer(&can, renderers[i]->device, root->device));
    if (!can) return fail(RTRB_ERR_CUDA, "device %d cannot access device %d", renderers[i]->device, root->device);
    CUDA_TRY(cudaSetDevice(renderers[i]->device));
    cudaError_t e = cudaDeviceEnablePeerAccess(root->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
      return fail(RTRB_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
    cudaGetLastError();
  }
  rtrb_render_opts base;
  memset(&base, 0, sizeof(base));
  if (opts_in) base = *opts_in;
  else { base.seed = 1; base.precision = RTRB_PREC_DEFAULT; }
  {
    const rtrb_render_opts& b = base;
    const bool full = (b.x0 == 0 && b.y0 == 0 && b.x1 == 0 && b.y1 == 0) ||
                      (b.x0 == 0 && b.y0 == 0 && b.x1 == cam->width && b.y1 == cam->height);
    if (!full && want_hit) {
      CUDA_TRY(cudaSetDevice(root->device));
      size_t npx = (size_t)cam->width * cam->height;
      fill_i32_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, root->stream>>>(root->hit.p, npx, -3);
      g_launches++;
      CUDA_TRY(cudaStreamSynchronize(root->stream));
    }
  }
  std::vector<rtrb_stats> st(n);
  std::vector<int> rcs(n, 0);
  std::vector<std::string> errs(n);
  std::vector<std::thread> th;
  for (int i = 0; i < n; ++i) {
    th.emplace_back([&, i]() {
      rtrb_render_opts o = base;
      o.tile_rank = i; o.tile_world = n; o.stream = nullptr; o.rgba_device_out = nullptr;
      FrameTargets tg;
      tg.rgba = root->rgba.p;
      tg.rgb = want_rgb ? root->rgb.p : nullptr;
      tg.hit = want_hit ? root->hit.p : nullptr;
      tg.want_rgb = want_rgb; tg.want_hit = want_hit; tg.no_fill = true;
      rcs[i] = render_impl(renderers[i], cam, &o, &tg, &st[i]);
      errs[i] = g_last_error;
    });
  }
  for (auto& t : th) t.join();
  int worst = RTRB_OK;
  for (int i = 0; i < n; ++i) {
    if (rcs[i] != RTRB_OK && rcs[i] != RTRB_ERR_RAISED) { g_last_error = errs[i]; return rcs[i]; }
    if (rcs[i] == RTRB_ERR_RAISED) { worst = RTRB_ERR_RAISED; g_last_error = errs[i]; }
  }
  rtrb_stats agg;
  memset(&agg, 0, sizeof(agg));
  agg.first_bad_x = agg.first_bad_y = -1;
  long long best_key = -1;
  {
    uint64_t* a = &agg.samples;
    for (int i = 0; i < n; ++i) {
      const uint64_t* b = &st[i].samples;
      for (int k = 0; k < 25; ++k) a[k] += b[k];  // samples .. cover_box_accepts
      agg.status |= st[i].status;
      agg.max_stack = std::max(agg.max_stack, st[i].max_stack);
      agg.device_ms = std::max(agg.device_ms, st[i].device_ms);
      agg.trace_ms = std::max(agg.trace_ms, st[i].trace_ms);
      if (st[i].first_bad_x >= 0) {
        long long key = (long long)st[i].first_bad_x * cam->height + st[i].first_bad_y;
        if (best_key < 0 || key < best_key) { best_key = key; agg.first_bad_x = st[i].first_bad_x; agg.first_bad_y = st[i].first_bad_y; }
      }
    }
  }
  if (stats_out) *stats_out = agg;
  root->last_w = cam->width; root->last_h = cam->height;
  root->last_has_rgb = want_rgb; root->last_has_hit = want_hit; root->last_rgba_own = true;
  root->last_stream = root->stream;
  root->last_format = base.pixel_format;
  std::string keep = g_last_error;
  rc = rtrb_download(root, rgba, rgb_or_null, hit_or_null);
  if (rc) return rc;
  if (worst == RTRB_ERR_RAISED) g_last_error = keep;
  return worst;
}

int rtrb_tile_partition(int width, int height, const int32_t* window, int tile_rank, int tile_world,
                        int32_t* tiles_out, int capacity, int* count_out) {
  if (width <= 0 || height <= 0 || !count_out) return fail(RTRB_ERR_INVALID, "bad argument");
  int x0 = 0, y0 = 0, x1 = width, y1 = height;
  if (window && !(window[0] == 0 && window[1] == 0 && window[2] == 0 && window[3] == 0)) {
    x0 = window[0]; y0 = window[1]; x1 = window[2]; y1 = window[3];
  }
  if (x0 < 0 || y0 < 0 || x1 > width || y1 > height || x0 >= x1 || y0 >= y1) return fail(RTRB_ERR_INVALID, "bad window");
  int world = tile_world <= 1 ? 1 : tile_world;
  int rank = world == 1 ? 0 : tile_rank;
  if (rank < 0 || rank >= world) return fail(RTRB_ERR_INVALID, "tile_rank %d outside tile_world %d", rank, world);
  std::vector<int32_t> t;
  enumerate_tiles(width, x0, y0, x1, y1, rank, world, t);
  *count_out = (int)t.size();
  if (tiles_out)
    for (int i = 0; i < (int)t.size() && i < capacity; ++i) tiles_out[i] = t[i];
  return RTRB_OK;
}

int rtrb_last_mt_passes(rtrb_renderer* r) { return r ? r->mt_iterations : 0; }

int rtrb_measure_fma_peak(int device, int which, double* tflops_out) {
  if (!tflops_out) return fail(RTRB_ERR_INVALID, "tflops_out is NULL");
  return which == 0 ? measure_fma<float>(device, tflops_out) : measure_fma<double>(device, tflops_out);
}

#pragma GCC visibility pop
}  // extern "C"
