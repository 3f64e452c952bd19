/*
 * ruby.h — TEST INFRASTRUCTURE.  A minimal stand-in for the Ruby C API, just large enough to
 * compile the reference's ONLY native component, ext/fast_4d_matrix/fast_4d_matrix.c, UNMODIFIED
 * from where it lies under /root/reference (no Ruby interpreter or ruby.h exists in this image).
 * oracle/Makefile builds it into oracle/_ref/libfast_4d_matrix_ref.so; tests use that library to
 * pin the oracle's Vec3 arithmetic against the reference's own compiled code.
 *
 * Only the API surface that file touches is provided (fast_4d_matrix.c:1-305): VALUE boxes,
 * Data_Wrap_Struct / Data_Get_Struct, floats, arrays, method registration and rb_raise.
 * rb_raise records the message and RETURNS (real Ruby longjmps); every raise site in the
 * reference file is followed by code that is safe to run, so this only means a flagged result.
 */
#ifndef RTRB_FAKE_RUBY_H
#define RTRB_FAKE_RUBY_H

#include <stdint.h>
#include <stdlib.h>

typedef uintptr_t VALUE;

enum { T_NIL = 0, T_FLOAT = 1, T_DATA = 2, T_ARRAY = 3, T_MODULE = 4, T_CLASS = 5 };

typedef struct rtrb_box {
  int type;
  double f;                 /* T_FLOAT */
  void* data;               /* T_DATA */
  void (*dfree)(void*);
  VALUE klass;
  VALUE items[8];           /* T_ARRAY */
  int n_items;
} rtrb_box;

#define Qnil ((VALUE)0)

extern VALUE rb_cObject;
extern VALUE rb_eRuntimeError;
extern VALUE rb_eArgError;

VALUE rb_define_module(const char* name);
VALUE rb_define_class_under(VALUE outer, const char* name, VALUE super);
void rb_define_singleton_method(VALUE klass, const char* name, VALUE (*fn)(), int argc);
void rb_define_method(VALUE klass, const char* name, VALUE (*fn)(), int argc);
void rb_define_alias(VALUE klass, const char* new_name, const char* old_name);

VALUE rb_float_new(double d);
VALUE rb_ary_new(void);
VALUE rb_ary_push(VALUE ary, VALUE item);
void rb_raise(VALUE exc, const char* fmt, ...);

VALUE rtrb_data_wrap(VALUE klass, void (*dfree)(void*), void* ptr);

#define TYPE(v) ((v) == Qnil ? T_NIL : ((rtrb_box*)(v))->type)
#define RFLOAT_VALUE(v) (((rtrb_box*)(v))->f)
#define NUM2DBL(v) (((rtrb_box*)(v))->f)
#define Data_Wrap_Struct(klass, mark, dfree, ptr) rtrb_data_wrap((klass), (void (*)(void*))(dfree), (ptr))
#define Data_Get_Struct(obj, type, sval) ((sval) = (type*)(((rtrb_box*)(obj))->data))

#endif
