/*
 * rtrb_b200.h — C ABI of the B200 tracing core for raytracing_rb's per-pixel hot path.
 *
 * One call per FRAME replaces the reference's per-pixel loop
 *   Camera#render_sync / render_fork -> Camera#render_at -> RayTracer#trace_sync
 *   (reference src/camera.rb:41-110, src/ray_tracer.rb:16-164, src/world.rb:37-98,
 *    src/objects/{world_object,sphere,plane,texture}.rb, ext/fast_4d_matrix/fast_4d_matrix.c:57-305).
 *
 * The reference has no FFI for this path; its only native-extension convention is
 * ext/fast_4d_matrix (extconf.rb:5,16 + fast_4d_matrix.c:29 `Init_fast_4d_matrix`).  The Ruby
 * shim that follows the same convention and binds exactly these entry points is in
 * ext/rtrb_b200/rtrb_b200.c; INTEGRATION.md shows how Camera/World hand their ivars over.
 *
 * Conventions (mirroring fast_4d_matrix.c's rb_raise style at :124,:220,:291):
 *   - every function returns an int status, 0 = RTRB_OK; rtrb_last_error() gives the message
 *     for the calling thread.
 *   - plain pointers and sizes only; the caller owns every output buffer.
 *   - all scene numbers are FP64 because the reference is FP64 end to end.
 *   - there is NO CPU fallback: every entry point that computes fails with RTRB_ERR_CUDA when
 *     no CUDA device is usable.
 */
#ifndef RTRB_B200_H
#define RTRB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTRB_ABI_VERSION 3

/* ---- status codes ------------------------------------------------------------------------- */
enum {
  RTRB_OK = 0,
  RTRB_ERR_INVALID = 1,   /* bad argument / scene the reference itself would reject (TypeError etc.) */
  RTRB_ERR_CUDA = 2,      /* CUDA runtime failure or no device */
  RTRB_ERR_RAISED = 3,    /* frame finished, but the reference would have raised (see rtrb_stats.status) */
  RTRB_ERR_UNSUPPORTED = 4
};

/* ---- per-frame status word: the reference's raise sites, one bit each ---------------------- */
enum {
  RTRB_ST_COLOR_GT_1 = 1u << 0,   /* ray_tracer.rb:294-296 'color greater than 1' */
  RTRB_ST_ZERO_VECTOR = 1u << 1,  /* fast_4d_matrix.c:123-124,290-291 and world_object.rb:106 */
  RTRB_ST_MATH_DOMAIN = 1u << 2,  /* Math.sqrt / Math.acos / Math.asin outside their domain */
  RTRB_ST_STACK_OVERFLOW = 1u << 3, /* device bounce stack exceeded (not a reference condition) */
  RTRB_ST_NAN_TO_INT = 1u << 4    /* texture.rb:24-25 Float#to_i on NaN/Infinity (FloatDomainError) */
};

/* ---- object kinds (world.yml `type:`; world.rb:28-34) -------------------------------------- */
enum { RTRB_OBJ_PLANE = 0, RTRB_OBJ_SPHERE = 1, RTRB_OBJ_BOX = 2 };

/* ---- RNG modes ----------------------------------------------------------------------------- */
enum {
  RTRB_RNG_CTR = 0,  /* Philox4x32-10 keyed by (seed; pixel, sample, ray path, purpose); device + oracle */
  RTRB_RNG_MT = 1    /* MT19937 genrand_res53 in the reference's consumption order (Random.srand(seed), main.rb:10;
                        one draw per lens_func, camera.rb:135, two per Monte-Carlo ray, world_object.rb:84; pixels x
                        outer / y inner, rays LIFO).  Stream-exact VALIDATION mode: blocking calls only, one GPU,
                        STRICT arithmetic, one thread per pixel, iterated to the fixed point of the per-pixel
                        stream offsets (the draw counts are data dependent).  Seeds must fit 32 bits. */
};

/* ---- rtrb_render_opts.pixel_format: layout of the 8-bit frame -------------------------------- */
enum {
  RTRB_FMT_RGBA8 = 0, /* H*W*4 bytes, alpha = 255: what array_to_color builds (camera.rb:153-156) */
  RTRB_FMT_RGB8 = 1,  /* H*W*3 bytes: alpha is the constant 255, so it need not cross PCIe; the caller
                         (Camera#render_cuda) re-inserts it when it fills the PNG canvas */
  RTRB_FMT_PNG_RGB8 = 2 /* H rows of 1 + 3*W bytes: the PNG filter-type byte 0 ("None") followed by the row's RGB8
                         pixels, i.e. exactly the byte stream a PNG encoder deflates into IDAT for an 8-bit RGB image
                         (what Camera#save_image, camera.rb:36-39, produces through the png gem): the host only runs
                         zlib over the buffer and frames the chunks */
};
/* bytes of one 8-bit frame in `format` */
#define RTRB_FRAME_BYTES(width, height, format) \
  ((format) == RTRB_FMT_RGBA8 ? (size_t)(width) * (height) * 4 : (format) == RTRB_FMT_RGB8 ? (size_t)(width) * (height) * 3 \
                                                               : (size_t)(height) * ((size_t)(width) * 3 + 1))

/* ---- rtrb_render_opts.skip_outputs ----------------------------------------------------------- */
enum { RTRB_SKIP_RGB = 1, RTRB_SKIP_HIT = 2 };

/* ---- arithmetic modes of the device path ---------------------------------------------------- */
enum {
  RTRB_PREC_STRICT = 0,   /* FP64, no FMA contraction, every div/sqrt as the reference writes it */
  RTRB_PREC_FAST64 = 1,   /* FP64 decisions identical to STRICT (FP32 cull + exact FP64 refine) */
  RTRB_PREC_DEFAULT = 1
};

/* One entry of world.yml `world_objects` (plane.rb:9-15, sphere.rb:8-14, box.rb:10-14), order preserved:
 * World#intersect's strict `<` (world.rb:48) makes the lowest index win distance ties. */
typedef struct rtrb_object_desc {
  int32_t type;            /* RTRB_OBJ_* */
  int32_t texture;         /* index into rtrb_scene_desc.textures, -1 = untextured */
  int32_t has_refraction;  /* plane / box: refractive_rate key present (plane.rb:57); sphere: must be 1 */
  int32_t reserved0;
  double point[3];         /* sphere: center ; plane: point ; box: point (its centre) */
  double radius;           /* sphere only */
  double front[3];         /* plane normal / box front axis, NOT normalised by the reference */
  double up[3];            /* plane / box `up` */
  double u_unit, v_unit;   /* plane uv units (plane.rb:81-85) */
  double greenwich_vec[3]; /* sphere texture frame (sphere.rb:111-120) */
  double north_pole_vec[3];
  double texture_horizontal_scale, texture_vertical_scale;
  double texture_u_offset, texture_v_offset; /* 0.0 when absent (texture.rb:15-16) */
  double refractive_rate;
  double diffuse_rate[3];
  double reflective_attenuation[3];
  double refractive_attenuation[3];
  double ambient[3];
  /* box only (box.rb:10,22-58): extents along front / up / normalize(front x up).  The six bounded
   * faces are derived by the library exactly as Box#initialize does.  A box ignores `texture`
   * (Box never overrides local_lighting, world_object.rb:51-74 runs without a colour filter). */
  double width_front, width_up, width_left;
} rtrb_object_desc;

/* One entry of world.yml `lights` (lights/light.rb:3-4, spot_light.rb:5). */
typedef struct rtrb_light_desc {
  double position[3];
  double color[3];
  double radius;
  double high_light_rate;
  double high_light_angle; /* degrees */
} rtrb_light_desc;

/* Decoded texture (texture.rb:12-20): 8-bit RGB rows, texel = v8/256.0, alpha ignored. */
typedef struct rtrb_texture_desc {
  int32_t width, height;   /* columns, rows */
  const uint8_t* rgb8;     /* height*width*3, row-major */
} rtrb_texture_desc;

typedef struct rtrb_scene_desc {
  double max_distance;          /* world.yml:1  */
  double soft_shadow_exponent;  /* world.yml:2  */
  int32_t n_objects, n_lights, n_textures, reserved0;
  const rtrb_object_desc* objects;
  const rtrb_light_desc* lights;
  const rtrb_texture_desc* textures;
} rtrb_scene_desc;

/* camera.yml (camera.rb:17-24). retina_* are HALF extents (camera.rb:133-134). */
typedef struct rtrb_camera_desc {
  double position[3], up[3], front[3];
  double retina_width, retina_height;
  double aperture_radius, image_distance, focal_distance;
  double variant_threshold;
  int32_t width, height;
  int32_t pre_sample_times, max_sample_times;
  int32_t trace_depth, monte_carlo_diffusion_times;
} rtrb_camera_desc;

typedef struct rtrb_render_opts {
  int32_t rng_mode;     /* RTRB_RNG_* */
  int32_t precision;    /* RTRB_PREC_* */
  uint64_t seed;        /* main.rb:10 uses Random.srand(1) */
  /* window of pixels to render, x in [x0,x1), y in [y0,y1); all zero = whole frame.
   * A column strip is exactly render_fork's child_work (camera.rb:53-65). */
  int32_t x0, y0, x1, y1;
  /* image-tile partition (replaces fork_jobs row farming): this call renders the 32x32-pixel
   * super-tiles number tile_rank, tile_rank + tile_world, ... of the window (row-major order; see
   * rtrb_tile_partition). tile_world <= 1 = everything. */
  int32_t tile_rank, tile_world;
  int32_t count_detail; /* 1 = fill every rtrb_stats counter (slower); 0 = rays/shadow/samples only */
  int32_t skip_outputs; /* rtrb_render_device only: RTRB_SKIP_RGB | RTRB_SKIP_HIT drop the optional float RGB
                           (24 B/pixel) / hit-id (4 B/pixel) frames; 0 keeps both for rtrb_download */
  void* stream;         /* cudaStream_t to launch on; NULL = the renderer's own stream */
  void* rgba_device_out;/* rtrb_render_device / rtrb_submit only: device pointer (possibly a peer mapping) the 8-bit
                           frame is written to; NULL = the renderer's own framebuffer.  The host-buffer calls
                           (rtrb_render, rtrb_render_multi) reject it with RTRB_ERR_INVALID, and rtrb_download refuses
                           to fetch an 8-bit frame that went elsewhere */
  int32_t pixel_format; /* RTRB_FMT_*: layout of the 8-bit frame (device buffer and host copies alike) */
  int32_t reserved0;
} rtrb_render_opts;

typedef struct rtrb_stats {
  uint64_t samples;           /* trace_sync calls (camera.rb:75,91) */
  uint64_t rays;              /* work-stack items passing the cut at ray_tracer.rb:52 */
  uint64_t shadow_queries;    /* lit_area calls made from local_lights (world.rb:75) */
  uint64_t highlight_hits;    /* items terminated by the highlight test (ray_tracer.rb:60-75) */
  uint64_t hits;              /* items whose World#intersect found an object */
  uint64_t local_shaded;      /* local_lighting evaluations (ray_tracer.rb:152) */
  uint64_t lit_lights;        /* lights kept by local_lights (area > 0) */
  uint64_t mc_rays;           /* path_tracing rays spawned (world_object.rb:76-90) */
  uint64_t refractions;       /* refraction children created */
  uint64_t texel_fetches;     /* Texture#color calls */
  uint64_t sphere_tests, sphere_accepts;      /* Sphere#intersect from World#intersect */
  uint64_t plane_tests, plane_accepts;        /* Plane#intersect from World#intersect */
  uint64_t cover_sphere, cover_sphere_full, cover_sphere_penumbra; /* Sphere#cover_area by branch */
  uint64_t cover_plane, cover_plane_accepts;  /* WorldObject#cover_area on planes */
  uint64_t adaptive_pixels;   /* pixels taking the extra-sample branch (camera.rb:87-93) */
  uint64_t exact_tests;       /* FAST64 only: objects re-evaluated in strict FP64 after the FP32 cull */
  uint64_t box_tests, box_accepts;            /* Box#intersect from World#intersect (box.rb:78-97) */
  uint64_t cover_box, cover_box_accepts;      /* WorldObject#cover_area on boxes */
  uint32_t status;            /* RTRB_ST_* bits */
  int32_t first_bad_x, first_bad_y; /* lowest (y*W+x) pixel that set a status bit, -1 if none */
  uint32_t max_stack;         /* deepest work stack seen */
  float device_ms;            /* CUDA-event time of the frame's kernels on the launch stream (0 for rtrb_submit frames) */
  float trace_ms;             /* CUDA-event time of the dominant kernel alone (trace over the pre samples); 0 likewise */
} rtrb_stats;

typedef struct rtrb_renderer rtrb_renderer; /* opaque: scene SoA + scratch + framebuffers on ONE device */

/* -- lifecycle ------------------------------------------------------------------------------- */
int rtrb_abi_version(void);
const char* rtrb_last_error(void);
int rtrb_device_count(int* count_out);

/* Bakes the scene (derived vectors plane `left`, sphere `ninety_degree_east_vec`, unit frames) into
 * flat SoA device buffers once; replaces World#initialize's object graph (world.rb:15-34). */
int rtrb_renderer_create(const rtrb_scene_desc* scene, int device, rtrb_renderer** out);
int rtrb_renderer_destroy(rtrb_renderer* r);

/* -- the hot path ---------------------------------------------------------------------------- */
/* Renders one frame into DEVICE buffers (async on opts->stream unless stats_out != NULL, which
 * synchronises).  Replaces the pixel loops camera.rb:59-63 and :102-106. */
int rtrb_render_device(rtrb_renderer* r, const rtrb_camera_desc* cam, const rtrb_render_opts* opts,
                       rtrb_stats* stats_out);

/* Copies the last frame to HOST buffers. rgba: RTRB_FRAME_BYTES(W, H, pixel_format) bytes (H*W*4 for the default
 * RTRB_FMT_RGBA8), row = y, column = x (camera.rb:98,105);
 * rgb_or_null: H*W*3 doubles = render_at's unclamped `color`; hit_or_null: H*W int32 primary hit
 * ids (index in world_objects, -1 miss, -2 highlight-terminated, -3 pixel not rendered). */
int rtrb_download(rtrb_renderer* r, uint8_t* rgba, double* rgb_or_null, int32_t* hit_or_null);

/* render_device + download in one blocking call with HOST buffers: what Camera#render_cuda binds. */
int rtrb_render(rtrb_renderer* r, const rtrb_camera_desc* cam, const rtrb_render_opts* opts,
                uint8_t* rgba, double* rgb_or_null, int32_t* hit_or_null, rtrb_stats* stats_out);

/* Pipelined form of rtrb_render for frame sequences: rtrb_submit launches the frame and queues its
 * device->host copy into rgba_host (pinned memory recommended) on a separate copy stream, then
 * returns; rtrb_wait blocks until that frame's bytes and stats are in host memory.  Up to four frames
 * may be in flight per renderer (tickets are consecutive integers; wait for them in order), so
 * later frames render while earlier ones cross PCIe.  8-bit frame only (opts->pixel_format).  rtrb_wait fails with
 * RTRB_ERR_INVALID for a ticket that is not the one occupying its slot (stale or duplicate). */
int rtrb_submit(rtrb_renderer* r, const rtrb_camera_desc* cam, const rtrb_render_opts* opts, uint8_t* rgba_host,
                int* ticket_out);
int rtrb_wait(rtrb_renderer* r, int ticket, rtrb_stats* stats_out);

/* -- multi-GPU tile gather without a collective ----------------------------------------------- */
/* Device pointer of the renderer's RGBA8 framebuffer for (width,height) (allocates it if needed).  From this
 * call (or rtrb_framebuffer_ipc_export) on the framebuffer is PINNED: a later frame or request that needs a larger
 * one fails with RTRB_ERR_INVALID instead of reallocating memory that peers may still be writing to. */
int rtrb_framebuffer_device_ptr(rtrb_renderer* r, int width, int height, void** ptr_out);
/* Copies width*height*4 bytes of that framebuffer to host memory (blocking).  With a framebuffer
 * allocated for (width, height * n) this fetches n frames that were rendered into consecutive slots
 * through rgba_device_out. */
int rtrb_framebuffer_download(rtrb_renderer* r, int width, int height, uint8_t* rgba_host);
/* Queues a copy of the first `bytes` bytes of that framebuffer to host memory on `stream` (NULL = the renderer's own)
 * and returns without synchronising: the last step of a multi-process tile gather, ordered on the caller's stream
 * behind whatever tells rank 0 that every rank's tiles have landed (render_fork's parent, camera.rb:42-52). */
int rtrb_framebuffer_copy_async(rtrb_renderer* r, size_t bytes, uint8_t* host, void* stream);
/* CUDA IPC handle (64 bytes) of that framebuffer, to hand to the other per-GPU processes. */
int rtrb_framebuffer_ipc_export(rtrb_renderer* r, int width, int height, uint8_t handle_out[64]);
/* Maps a peer process's framebuffer into this process; pass the result as rgba_device_out. */
int rtrb_ipc_open(int device, const uint8_t handle[64], void** ptr_out);
int rtrb_ipc_close(int device, void* ptr);
/* In-process variant: renders tile_world = n renderers (one per GPU, one host thread each) into
 * renderers[0]'s framebuffer through peer mappings, then downloads from renderers[0]. */
int rtrb_render_multi(rtrb_renderer* const* renderers, int n, const rtrb_camera_desc* cam,
                      const rtrb_render_opts* opts, uint8_t* rgba, double* rgb_or_null,
                      int32_t* hit_or_null, rtrb_stats* stats_out);

/* Copy-engine form of the gather (whole frames dealt to ranks): queues a
 * device-to-device copy of `bytes` from `src` (a buffer of this renderer's device) to `dst_peer` (a
 * peer mapping, e.g. a frame slot of rank 0's framebuffer) on the renderer's copy stream, ordered
 * after everything queued so far on `after_stream` (NULL = the renderer's own stream).  The render
 * stream does not wait: the next frame renders while this one crosses NVLink. */
int rtrb_peer_push(rtrb_renderer* r, const void* src, void* dst_peer, size_t bytes, void* after_stream);
/* Makes `stream` (NULL = the renderer's own) wait for every rtrb_peer_push queued so far. */
int rtrb_peer_push_join(rtrb_renderer* r, void* stream);

/* Pure host helper (no GPU): the 32x32-pixel super-tiles of a width x height frame (optionally
 * clipped to window {x0,y0,x1,y1}, NULL = whole frame) that tile_rank renders out of tile_world,
 * as global ids ty * ceil(width/32) + tx in ascending order.  count_out receives the number of
 * tiles even when it exceeds capacity.  This is the partition rtrb_render_device applies. */
int rtrb_tile_partition(int width, int height, const int32_t* window_or_null, int tile_rank, int tile_world,
                        int32_t* tiles_out, int capacity, int* count_out);

/* -- measurement helpers ---------------------------------------------------------------------- */
/* Dependent-free FMA issue microbenchmark on `device`: the roofline denominator SURVEY.md 8d asks
 * for. which: 0 = FP32 FFMA, 1 = FP64 DFMA. Result in TFLOP/s (2 flops per FMA). */
int rtrb_measure_fma_peak(int device, int which, double* tflops_out);
/* Passes the last RTRB_RNG_MT frame of this renderer needed to reach its fixed point (0 if none yet). */
int rtrb_last_mt_passes(rtrb_renderer* r);
/* Number of kernel launches issued by this library in this process so far. */
uint64_t rtrb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RTRB_B200_H */
