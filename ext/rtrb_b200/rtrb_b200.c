/*
 * rtrb_b200.c — Ruby C extension: the thin shim between the reference's Ruby classes and the CUDA
 * tracing core (include/rtrb_b200.h).  Same shape as the reference's only native component
 * (ext/fast_4d_matrix/fast_4d_matrix.c): `Init_<name>` registers methods on a module; objects are
 * wrapped C pointers (Data_Wrap_Struct with a dfree hook, fast_4d_matrix.c:68); errors surface as
 * rb_raise(rb_eRuntimeError, ...) (fast_4d_matrix.c:124).
 *
 * It only FLATTENS: it reads World's @world_objects / @lights / scalars and Camera's accessors
 * (camera.rb:17-24), turns every Vec3 into three doubles through #to_a (fast_4d_matrix.c:86-96) and
 * calls the C ABI.  No tracing arithmetic lives here.  NOT compiled in this repository's image (no
 * ruby.h); INTEGRATION.md lists the build steps for a machine with Ruby.
 *
 *   Rtrb::Renderer.new(world)                      -> bakes the scene on GPU 0 (rtrb_renderer_create)
 *   renderer.render(camera, seed = 1, precision = 1) -> String of H*W*3 RGB8 bytes, row = y, col = x
 *                                                     (RTRB_FMT_RGB8: alpha is always 255, camera.rb:155)
 *   renderer.stats                                 -> Hash of the last frame's counters
 */
#include <ruby.h>
#include <ruby/thread.h>
#include <stdlib.h>
#include <string.h>

#include "rtrb_b200.h"

static VALUE mRtrb, cRenderer;

typedef struct {
  rtrb_renderer* r;
  rtrb_stats last;
} shim_renderer;

static void shim_free(void* p) {
  shim_renderer* s = (shim_renderer*)p;
  if (s->r) rtrb_renderer_destroy(s->r);
  free(s);
}

static void vec3_into(VALUE v, double out[3], const char* what) {
  if (NIL_P(v)) rb_raise(rb_eTypeError, "%s is nil", what);
  VALUE a = rb_funcall(v, rb_intern("to_a"), 0);
  for (int i = 0; i < 3; ++i) out[i] = NUM2DBL(rb_ary_entry(a, i));
}
static VALUE ivar(VALUE obj, const char* name) { return rb_iv_get(obj, name); }
static double ivar_f(VALUE obj, const char* name, const char* what) {
  VALUE v = rb_iv_get(obj, name);
  if (NIL_P(v)) rb_raise(rb_eTypeError, "%s is nil", what);
  return NUM2DBL(v);
}
static int is_a(VALUE obj, const char* klass_path) { return RTEST(rb_obj_is_kind_of(obj, rb_path2class(klass_path))); }

/* Texture -> 8-bit RGB rows: Texture#to_a holds rows of Vec3(v8/256.0) (texture.rb:12-20,34-36).  The buffer is
 * entered into `td` BEFORE it is filled, so the rb_ensure cleanup frees it even when a texel conversion raises. */
static void texture_bytes(VALUE tex, rtrb_texture_desc* td) {
  const int w = NUM2INT(rb_funcall(tex, rb_intern("width"), 0));
  const int h = NUM2INT(rb_funcall(tex, rb_intern("height"), 0));
  VALUE rows = rb_funcall(tex, rb_intern("to_a"), 0);
  uint8_t* px = (uint8_t*)malloc((size_t)w * h * 3);
  if (!px) rb_raise(rb_eNoMemError, "rtrb_b200: texture buffer");
  td->rgb8 = px; td->width = w; td->height = h;
  for (int y = 0; y < h; ++y) {
    VALUE row = rb_ary_entry(rows, y);
    for (int x = 0; x < w; ++x) {
      double c[3];
      vec3_into(rb_ary_entry(row, x), c, "texel");
      for (int k = 0; k < 3; ++k) px[((size_t)y * w + x) * 3 + k] = (uint8_t)(c[k] * 256.0 + 0.5);
    }
  }
}

/* Flattening World -> rtrb_scene_desc can raise half way (a nil ivar, a non-numeric value: vec3_into / ivar_f call
 * rb_raise).  The scratch arrays therefore live in a struct that rb_ensure frees whether the body returns or raises. */
typedef struct {
  VALUE self, world;
  rtrb_object_desc* od;
  rtrb_light_desc* ld;
  rtrb_texture_desc* td;
  int n_tex;
} init_call;

static VALUE init_cleanup(VALUE p) {
  init_call* c = (init_call*)p;
  for (int i = 0; i < c->n_tex; ++i) free((void*)c->td[i].rgb8);
  free(c->od); free(c->ld); free(c->td);
  return Qnil;
}

static VALUE init_body(VALUE p) {
  init_call* c = (init_call*)p;
  VALUE world = c->world;
  VALUE objs = ivar(world, "@world_objects"), lights = ivar(world, "@lights");
  long n_obj = RARRAY_LEN(objs), n_li = RARRAY_LEN(lights);
  rtrb_object_desc* od = c->od = (rtrb_object_desc*)calloc(n_obj > 0 ? n_obj : 1, sizeof(*od));
  rtrb_light_desc* ld = c->ld = (rtrb_light_desc*)calloc(n_li > 0 ? n_li : 1, sizeof(*ld));
  rtrb_texture_desc* td = c->td = (rtrb_texture_desc*)calloc(n_obj > 0 ? n_obj : 1, sizeof(*td));
  if (!od || !ld || !td) rb_raise(rb_eNoMemError, "rtrb_b200: scene scratch");
  for (long i = 0; i < n_obj; ++i) {
    VALUE o = rb_ary_entry(objs, i);
    rtrb_object_desc* d = &od[i];
    d->texture = -1;
    VALUE tex = ivar(o, "@texture");
    if (!NIL_P(tex)) {
      d->texture = c->n_tex++;          /* counted first: init_cleanup then frees a half-filled buffer too */
      texture_bytes(tex, &td[d->texture]);
      /* what Texture#color really uses (texture.rb:9-10,15-16,24-25): the Texture object's own scales and offsets.
       * Sphere passes its texture_u/v_offset on (sphere.rb:25); Plane and Box do not (plane.rb:35, box.rb:76), so
       * their textures keep 0.0 whatever the YAML says. */
      d->texture_horizontal_scale = ivar_f(tex, "@horizontal_scale", "texture horizontal_scale");
      d->texture_vertical_scale = ivar_f(tex, "@vertical_scale", "texture vertical_scale");
      d->texture_u_offset = ivar_f(tex, "@u_off", "texture u_off");
      d->texture_v_offset = ivar_f(tex, "@v_off", "texture v_off");
    }
    if (is_a(o, "Alex::Objects::Sphere")) {
      d->type = RTRB_OBJ_SPHERE;
      d->has_refraction = 1;
      vec3_into(ivar(o, "@center"), d->point, "center");
      d->radius = ivar_f(o, "@radius", "radius");
      d->refractive_rate = ivar_f(o, "@refractive_rate", "refractive_rate");   /* sphere.rb:93 */
      if (d->texture >= 0) {
        vec3_into(ivar(o, "@greenwich_vec"), d->greenwich_vec, "greenwich_vec");
        vec3_into(ivar(o, "@north_pole_vec"), d->north_pole_vec, "north_pole_vec");
      }
    } else if (is_a(o, "Alex::Objects::Plane")) {
      d->type = RTRB_OBJ_PLANE;
      vec3_into(ivar(o, "@point"), d->point, "point");
      vec3_into(ivar(o, "@front"), d->front, "front");
      vec3_into(ivar(o, "@up"), d->up, "up");
      VALUE uu = ivar(o, "@u_unit"), vu = ivar(o, "@v_unit"), rr = ivar(o, "@refractive_rate");
      d->u_unit = NIL_P(uu) ? 1.0 : NUM2DBL(uu);
      d->v_unit = NIL_P(vu) ? 1.0 : NUM2DBL(vu);
      d->has_refraction = RTEST(rr) ? 1 : 0;                                    /* plane.rb:57 */
      if (d->has_refraction) d->refractive_rate = NUM2DBL(rr);
    } else if (is_a(o, "Alex::Objects::Box")) {
      /* box.rb:10: the library derives the six faces exactly as Box#initialize does (box.rb:22-73) */
      d->type = RTRB_OBJ_BOX;
      d->texture = -1;                                                          /* Box never samples its texture */
      vec3_into(ivar(o, "@point"), d->point, "point");
      vec3_into(ivar(o, "@front"), d->front, "front");
      vec3_into(ivar(o, "@up"), d->up, "up");
      d->width_front = ivar_f(o, "@width_front", "width_front");
      d->width_up = ivar_f(o, "@width_up", "width_up");
      d->width_left = ivar_f(o, "@width_left", "width_left");
      VALUE rr = ivar(o, "@refractive_rate");
      d->has_refraction = RTEST(rr) ? 1 : 0;                                    /* plane.rb:57 on every face */
      if (d->has_refraction) d->refractive_rate = NUM2DBL(rr);
    } else {
      rb_raise(rb_eNotImpError, "object %ld: only Sphere, Plane and Box run on the GPU path", i);
    }
    vec3_into(ivar(o, "@diffuse_rate"), d->diffuse_rate, "diffuse_rate");
    vec3_into(ivar(o, "@reflective_attenuation"), d->reflective_attenuation, "reflective_attenuation");
    vec3_into(ivar(o, "@ambient"), d->ambient, "ambient");
    if (!NIL_P(ivar(o, "@refractive_attenuation")))
      vec3_into(ivar(o, "@refractive_attenuation"), d->refractive_attenuation, "refractive_attenuation");
  }
  for (long i = 0; i < n_li; ++i) {
    VALUE l = rb_ary_entry(lights, i);
    vec3_into(ivar(l, "@position"), ld[i].position, "light position");
    vec3_into(ivar(l, "@color"), ld[i].color, "light color");
    ld[i].radius = ivar_f(l, "@radius", "radius");
    ld[i].high_light_rate = ivar_f(l, "@high_light_rate", "high_light_rate");
    ld[i].high_light_angle = ivar_f(l, "@high_light_angle", "high_light_angle");
  }
  rtrb_scene_desc sd;
  memset(&sd, 0, sizeof(sd));
  sd.max_distance = ivar_f(world, "@max_distance", "max_distance");
  sd.soft_shadow_exponent = ivar_f(world, "@soft_shadow_exponent", "soft_shadow_exponent");
  sd.n_objects = (int32_t)n_obj; sd.n_lights = (int32_t)n_li; sd.n_textures = c->n_tex;
  sd.objects = od; sd.lights = ld; sd.textures = td;

  shim_renderer* s;
  Data_Get_Struct(c->self, shim_renderer, s);
  int rc = rtrb_renderer_create(&sd, 0, &s->r);
  if (rc != RTRB_OK) rb_raise(rb_eRuntimeError, "%s", rtrb_last_error());
  return c->self;
}

static VALUE renderer_initialize(VALUE self, VALUE world) {
  init_call c;
  memset(&c, 0, sizeof(c));
  c.self = self; c.world = world;
  return rb_ensure(init_body, (VALUE)&c, init_cleanup, (VALUE)&c);
}

static VALUE renderer_alloc(VALUE klass) {
  shim_renderer* s = (shim_renderer*)calloc(1, sizeof(*s));
  return Data_Wrap_Struct(klass, 0, shim_free, s);
}

typedef struct {
  shim_renderer* s;
  rtrb_camera_desc cam;
  rtrb_render_opts opts;
  uint8_t* rgba;
  int rc;
} render_call;

static void* render_without_gvl(void* p) {   /* one blocking call per frame; the GVL is released */
  render_call* c = (render_call*)p;
  c->rc = rtrb_render(c->s->r, &c->cam, &c->opts, c->rgba, NULL, NULL, &c->s->last);
  return NULL;
}

static VALUE renderer_render(int argc, VALUE* argv, VALUE self) {
  VALUE camera, seed, precision;
  rb_scan_args(argc, argv, "12", &camera, &seed, &precision);
  shim_renderer* s;
  Data_Get_Struct(self, shim_renderer, s);
  render_call c;
  memset(&c, 0, sizeof(c));
  c.s = s;
  vec3_into(rb_funcall(camera, rb_intern("position"), 0), c.cam.position, "position");
  vec3_into(rb_funcall(camera, rb_intern("up"), 0), c.cam.up, "up");
  vec3_into(rb_funcall(camera, rb_intern("front"), 0), c.cam.front, "front");
#define CAMF(field) c.cam.field = NUM2DBL(rb_funcall(camera, rb_intern(#field), 0))
#define CAMI(field) c.cam.field = NUM2INT(rb_funcall(camera, rb_intern(#field), 0))
  CAMF(retina_width); CAMF(retina_height); CAMF(aperture_radius); CAMF(image_distance); CAMF(focal_distance);
  CAMF(variant_threshold);
  CAMI(width); CAMI(height); CAMI(pre_sample_times); CAMI(max_sample_times); CAMI(trace_depth);
  CAMI(monte_carlo_diffusion_times);
  c.opts.rng_mode = RTRB_RNG_CTR;
  c.opts.seed = NIL_P(seed) ? 1 : NUM2ULL(seed);                 /* main.rb:10 Random.srand(1) */
  c.opts.precision = NIL_P(precision) ? RTRB_PREC_DEFAULT : NUM2INT(precision);
  c.opts.pixel_format = RTRB_FMT_RGB8;                           /* alpha = 255 never crosses PCIe */
  size_t bytes = (size_t)c.cam.width * c.cam.height * 3;
  VALUE out = rb_str_new(NULL, (long)bytes);
  c.rgba = (uint8_t*)RSTRING_PTR(out);
  rb_thread_call_without_gvl(render_without_gvl, &c, RUBY_UBF_IO, NULL);
  if (c.rc == RTRB_ERR_RAISED) rb_raise(rb_eRuntimeError, "%s", rtrb_last_error());  /* ray_tracer.rb:295 et al. */
  if (c.rc != RTRB_OK) rb_raise(rb_eRuntimeError, "%s", rtrb_last_error());
  return out;
}

static VALUE renderer_stats(VALUE self) {
  shim_renderer* s;
  Data_Get_Struct(self, shim_renderer, s);
  VALUE h = rb_hash_new();
#define PUT(k) rb_hash_aset(h, ID2SYM(rb_intern(#k)), ULL2NUM(s->last.k))
  PUT(samples); PUT(rays); PUT(shadow_queries); PUT(highlight_hits); PUT(hits); PUT(local_shaded); PUT(texel_fetches);
  PUT(adaptive_pixels); PUT(box_tests); PUT(box_accepts);
  rb_hash_aset(h, ID2SYM(rb_intern("device_ms")), DBL2NUM(s->last.device_ms));
  rb_hash_aset(h, ID2SYM(rb_intern("status")), UINT2NUM(s->last.status));
  return h;
}

void Init_rtrb_b200(void) {
  mRtrb = rb_define_module("Rtrb");
  cRenderer = rb_define_class_under(mRtrb, "Renderer", rb_cObject);
  rb_define_alloc_func(cRenderer, renderer_alloc);
  rb_define_method(cRenderer, "initialize", renderer_initialize, 1);
  rb_define_method(cRenderer, "render", renderer_render, -1);
  rb_define_method(cRenderer, "stats", renderer_stats, 0);
  rb_define_const(mRtrb, "ABI_VERSION", INT2NUM(rtrb_abi_version()));
}
