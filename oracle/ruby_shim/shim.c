/*
 * shim.c — TEST INFRASTRUCTURE.  Runtime behind oracle/ruby_shim/ruby.h plus a small C driver so
 * Python (ctypes) can call the methods that the reference's fast_4d_matrix.c registers in
 * Init_fast_4d_matrix (fast_4d_matrix.c:29-55) by their Ruby names.
 */
#include "ruby.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

VALUE rb_cObject = 0;
VALUE rb_eRuntimeError = 0;
VALUE rb_eArgError = 0;

typedef struct method_entry {
  char name[32];
  VALUE (*fn)();
  int argc;
  int singleton;
} method_entry;

static method_entry g_methods[64];
static int g_n_methods = 0;
static char g_last_raise[256];
static int g_raised = 0;

static rtrb_box* new_box(int type) {
  rtrb_box* b = (rtrb_box*)calloc(1, sizeof(rtrb_box));
  b->type = type;
  return b;
}

VALUE rb_define_module(const char* name) { (void)name; return (VALUE)new_box(T_MODULE); }
VALUE rb_define_class_under(VALUE outer, const char* name, VALUE super) {
  (void)outer; (void)name; (void)super;
  return (VALUE)new_box(T_CLASS);
}
static void add_method(const char* name, VALUE (*fn)(), int argc, int singleton) {
  method_entry* m = &g_methods[g_n_methods++];
  snprintf(m->name, sizeof(m->name), "%s", name);
  m->fn = fn; m->argc = argc; m->singleton = singleton;
}
void rb_define_singleton_method(VALUE klass, const char* name, VALUE (*fn)(), int argc) {
  (void)klass; add_method(name, fn, argc, 1);
}
void rb_define_method(VALUE klass, const char* name, VALUE (*fn)(), int argc) {
  (void)klass; add_method(name, fn, argc, 0);
}
void rb_define_alias(VALUE klass, const char* new_name, const char* old_name) {
  (void)klass;
  for (int i = 0; i < g_n_methods; ++i)
    if (!g_methods[i].singleton && strcmp(g_methods[i].name, old_name) == 0) {
      add_method(new_name, g_methods[i].fn, g_methods[i].argc, 0);
      return;
    }
}
VALUE rb_float_new(double d) { rtrb_box* b = new_box(T_FLOAT); b->f = d; return (VALUE)b; }
VALUE rb_ary_new(void) { return (VALUE)new_box(T_ARRAY); }
VALUE rb_ary_push(VALUE ary, VALUE item) {
  rtrb_box* b = (rtrb_box*)ary;
  if (b->n_items < 8) b->items[b->n_items++] = item;
  return ary;
}
void rb_raise(VALUE exc, const char* fmt, ...) {
  (void)exc;
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_last_raise, sizeof(g_last_raise), fmt, ap);
  va_end(ap);
  g_raised = 1;
}
VALUE rtrb_data_wrap(VALUE klass, void (*dfree)(void*), void* ptr) {
  rtrb_box* b = new_box(T_DATA);
  b->klass = klass; b->dfree = dfree; b->data = ptr;
  return (VALUE)b;
}
static void release(VALUE v) {  /* what the GC would do through the dfree hook */
  if (v == Qnil) return;
  rtrb_box* b = (rtrb_box*)v;
  if (b->type == T_DATA && b->dfree && b->data) b->dfree(b->data);
  if (b->type == T_ARRAY) for (int i = 0; i < b->n_items; ++i) release(b->items[i]);
  free(b);
}

extern void Init_fast_4d_matrix(void);
static int g_inited = 0;

static method_entry* find(const char* name, int singleton) {
  for (int i = 0; i < g_n_methods; ++i)
    if (g_methods[i].singleton == singleton && strcmp(g_methods[i].name, name) == 0) return &g_methods[i];
  return NULL;
}

/* number of methods Init_fast_4d_matrix registered (incl. aliases) */
int rtrb_ref_init(void) {
  if (!g_inited) { Init_fast_4d_matrix(); g_inited = 1; }
  return g_n_methods;
}
int rtrb_ref_has_method(const char* name) { rtrb_ref_init(); return find(name, 0) != NULL || find(name, 1) != NULL; }
const char* rtrb_ref_last_raise(void) { return g_last_raise; }

/*
 * Calls Vec3.from_a(a).<method>(arg) through the reference's compiled functions.
 *   b_kind: 0 = no argument, 1 = Vec3 argument b[0..2], 2 = Float argument b[0]
 *   out[0..3]: Float result -> out[0]; Vec3 result -> values[0..2] and the cached r in out[3];
 *              Array result -> items.  `self_after` (4 doubles, may be NULL) receives self's
 *              {values, r} after the call, for the bang methods.
 * Returns: 1 Float, 2 Vec3, 3 Array, 0 nil, -1 unknown method; *raised set if rb_raise was hit.
 */
int rtrb_ref_vec3_call(const char* method, const double* a, int b_kind, const double* b, double* out,
                       double* self_after, int* raised) {
  rtrb_ref_init();
  method_entry* from_a = find("from_a", 1);
  method_entry* m = find(method, 0);
  if (!from_a || !m) return -1;
  g_raised = 0; g_last_raise[0] = 0;
  VALUE fa[3] = {rb_float_new(a[0]), rb_float_new(a[1]), rb_float_new(a[2])};
  VALUE self = from_a->fn((VALUE)0, fa[0], fa[1], fa[2]);
  VALUE arg = Qnil, fb[3] = {Qnil, Qnil, Qnil};
  if (b_kind == 1) {
    for (int i = 0; i < 3; ++i) fb[i] = rb_float_new(b[i]);
    arg = from_a->fn((VALUE)0, fb[0], fb[1], fb[2]);
  } else if (b_kind == 2) {
    arg = rb_float_new(b[0]);
  }
  VALUE res = (m->argc == 0) ? m->fn(self) : m->fn(self, arg);
  int kind = 0;
  if (res != Qnil) {
    rtrb_box* rb = (rtrb_box*)res;
    if (rb->type == T_FLOAT) { out[0] = rb->f; kind = 1; }
    else if (rb->type == T_DATA) { memcpy(out, rb->data, 4 * sizeof(double)); kind = 2; }
    else if (rb->type == T_ARRAY) { for (int i = 0; i < rb->n_items; ++i) out[i] = RFLOAT_VALUE(rb->items[i]); kind = 3; }
  }
  if (self_after) memcpy(self_after, ((rtrb_box*)self)->data, 4 * sizeof(double));
  if (raised) *raised = g_raised;
  if (res != self) release(res);
  release(self);
  if (arg != Qnil) release(arg);
  for (int i = 0; i < 3; ++i) { release(fa[i]); release(fb[i]); }
  return kind;
}
