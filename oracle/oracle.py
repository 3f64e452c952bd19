"""oracle.py — TEST INFRASTRUCTURE (see oracle/rtrb_oracle.cpp header).  ctypes wrapper around
oracle/liboracle.so, the FP64 CPU restatement of the reference hot path, and around
oracle/_ref/libfast_4d_matrix_ref.so, the reference's own Vec3 C code.  Only tests/, bench.py's CPU
baseline legs and __graft_entry__.smoke() may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

from raytracing_rb_b200 import _abi
from raytracing_rb_b200.renderer import Frame, make_opts

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libfast_4d_matrix_ref.so")

_lib = None
_ref = None


def build(force=False):
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "rtrb_oracle.cpp")):
        subprocess.check_call(["make", "-C", HERE, "-s", "all"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        P = C.POINTER
        L.rtrb_oracle_scene_create.restype = C.c_void_p
        L.rtrb_oracle_scene_create.argtypes = [P(_abi.SceneDesc)]
        L.rtrb_oracle_scene_destroy.argtypes = [C.c_void_p]
        L.rtrb_oracle_render.argtypes = [C.c_void_p, P(_abi.CameraDesc), P(_abi.RenderOpts), C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, P(_abi.Stats)]
        L.rtrb_oracle_lens_ray.argtypes = [P(_abi.CameraDesc), C.c_int, C.c_int, C.c_double, P(C.c_double)]
        L.rtrb_oracle_object_distance.restype = C.c_double
        L.rtrb_oracle_object_distance.argtypes = [P(_abi.CameraDesc)]
        L.rtrb_oracle_intersect.argtypes = [C.c_void_p, C.c_int, P(C.c_double), P(C.c_double), P(C.c_double), P(C.c_int)]
        L.rtrb_oracle_box_face.argtypes = [C.c_void_p, C.c_int, P(C.c_double), P(C.c_double)]
        L.rtrb_oracle_world_intersect.argtypes = [C.c_void_p, P(C.c_double), P(C.c_double), P(C.c_double)]
        L.rtrb_oracle_cover_area.restype = C.c_double
        L.rtrb_oracle_cover_area.argtypes = [C.c_void_p, C.c_int, P(C.c_double), C.c_double, P(C.c_double)]
        L.rtrb_oracle_lit_area.restype = C.c_double
        L.rtrb_oracle_lit_area.argtypes = [C.c_void_p, P(C.c_double), C.c_double, P(C.c_double)]
        L.rtrb_oracle_high_lights.argtypes = [C.c_void_p, P(C.c_double), P(C.c_double)]
        L.rtrb_oracle_texture_color.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, P(C.c_double)]
        L.rtrb_oracle_get_uv.argtypes = [C.c_void_p, C.c_int, P(C.c_double), P(C.c_double)]
        L.rtrb_oracle_intersect_parameters.argtypes = [C.c_void_p, C.c_int, P(C.c_double), P(C.c_double),
                                                       P(C.c_double), P(C.c_double)]
        L.rtrb_oracle_vec3.argtypes = [C.c_int, P(C.c_double), P(C.c_double), P(C.c_double)]
        L.rtrb_oracle_philox.argtypes = [C.c_uint32, C.c_uint32, P(C.c_uint32), P(C.c_uint32)]
        L.rtrb_oracle_mt_res53.argtypes = [C.c_uint32, C.c_int, P(C.c_double)]
        _lib = L
    return _lib


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


class OracleScene:
    def __init__(self, scene_holder):
        self._holder = scene_holder
        self._h = lib().rtrb_oracle_scene_create(C.byref(scene_holder.desc))

    def __del__(self):
        try:
            if self._h:
                lib().rtrb_oracle_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def render(self, cam, opts=None, threads=None, want_rgb=True, want_hit=True):
        opts = opts or make_opts()
        threads = threads or (os.cpu_count() or 1)
        H, W = cam.height, cam.width
        rgba = np.zeros((H, W, 4), np.uint8)
        rgb = np.zeros((H, W, 3), np.float64) if want_rgb else None
        hit = np.full((H, W), -3, np.int32) if want_hit else None
        st = _abi.Stats()
        lib().rtrb_oracle_render(self._h, C.byref(cam), C.byref(opts), threads, rgba.ctypes.data,
                                 rgb.ctypes.data if want_rgb else None, hit.ctypes.data if want_hit else None,
                                 C.byref(st))
        code = _abi.RTRB_ERR_RAISED if st.status else _abi.RTRB_OK
        return Frame(rgba, rgb, hit, st.as_dict(), code)

    # ---- probes ----
    def intersect(self, obj, o, d):
        out, din = (C.c_double * 6)(), C.c_int()
        ok = lib().rtrb_oracle_intersect(self._h, obj, _d3(o), _d3(d), out, C.byref(din))
        if not ok:
            return None
        return list(out[:3]), ("in" if din.value else "out"), list(out[3:])

    def box_face(self, obj, o, d):
        return lib().rtrb_oracle_box_face(self._h, obj, _d3(o), _d3(d))

    def world_intersect(self, o, d):
        out = (C.c_double * 3)()
        idx = lib().rtrb_oracle_world_intersect(self._h, _d3(o), _d3(d), out)
        return idx, list(out)

    def cover_area(self, obj, light_pos, light_radius, target):
        return lib().rtrb_oracle_cover_area(self._h, obj, _d3(light_pos), float(light_radius), _d3(target))

    def lit_area(self, light_pos, light_radius, target):
        return lib().rtrb_oracle_lit_area(self._h, _d3(light_pos), float(light_radius), _d3(target))

    def high_lights(self, o, d):
        return lib().rtrb_oracle_high_lights(self._h, _d3(o), _d3(d))

    def texture_color(self, obj, u, v):
        out = (C.c_double * 3)()
        lib().rtrb_oracle_texture_color(self._h, obj, float(u), float(v), out)
        return list(out)

    def get_uv(self, obj, p):
        out = (C.c_double * 2)()
        lib().rtrb_oracle_get_uv(self._h, obj, _d3(p), out)
        return list(out)

    def intersect_parameters(self, obj, o, d):
        n, out = (C.c_double * 3)(), (C.c_double * 12)()
        r = lib().rtrb_oracle_intersect_parameters(self._h, obj, _d3(o), _d3(d), n, out)
        if r < 0:
            return None
        return {"n": list(n), "reflection": (list(out[0:3]), list(out[3:6])),
                "refraction": (list(out[6:9]), list(out[9:12])) if r else None}


def lens_ray(cam, x, y, theta):
    out = (C.c_double * 6)()
    lib().rtrb_oracle_lens_ray(C.byref(cam), x, y, float(theta), out)
    return list(out[:3]), list(out[3:])


def object_distance(cam):
    return lib().rtrb_oracle_object_distance(C.byref(cam))


VEC3_OPS = {"dot": 0, "cos": 1, "cross": 2, "add": 3, "sub": 4, "mul": 5, "mul_scalar": 6, "div": 7, "r": 8,
            "r2": 9, "normalize": 10, "neg": 11}


def vec3(op, a, b=None):
    out = (C.c_double * 3)()
    bb = None
    if b is not None:
        bb = _d3(b) if hasattr(b, "__len__") else _d3([b, 0, 0])
    lib().rtrb_oracle_vec3(VEC3_OPS[op], _d3(a), bb, out)
    return list(out)


def philox(k0, k1, ctr):
    c, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 4)()
    lib().rtrb_oracle_philox(k0, k1, c, o)
    return list(o)


def mt_res53(seed, n):
    out = (C.c_double * n)()
    lib().rtrb_oracle_mt_res53(seed, n, out)
    return list(out)


# ---- the reference's own compiled Vec3 code (oracle/_ref) --------------------------------------
def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.isfile(REF_LIB):
            build(force=True)
        if not os.path.isfile(REF_LIB):
            return None
        R = C.CDLL(REF_LIB)
        R.rtrb_ref_init.restype = C.c_int
        R.rtrb_ref_has_method.argtypes = [C.c_char_p]
        R.rtrb_ref_last_raise.restype = C.c_char_p
        R.rtrb_ref_vec3_call.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        _ref = R
    return _ref


def ref_vec3_call(method, a, b=None):
    """Calls Vec3.from_a(*a).<method>(b) in the reference's compiled fast_4d_matrix.c.
    Returns (kind, values, self_after, raised): kind 'float'|'vec3'|'array'|'nil'."""
    R = ref_lib()
    out, selfa, raised = (C.c_double * 8)(), (C.c_double * 4)(), C.c_int()
    if b is None:
        kind = R.rtrb_ref_vec3_call(method.encode(), _d3(a), 0, None, out, selfa, C.byref(raised))
    elif hasattr(b, "__len__"):
        kind = R.rtrb_ref_vec3_call(method.encode(), _d3(a), 1, _d3(b), out, selfa, C.byref(raised))
    else:
        kind = R.rtrb_ref_vec3_call(method.encode(), _d3(a), 2, _d3([b, 0, 0]), out, selfa, C.byref(raised))
    if kind < 0:
        raise KeyError(method)
    names = {0: "nil", 1: "float", 2: "vec3", 3: "array"}
    n = {0: 0, 1: 1, 2: 4, 3: 3}[kind]
    return names[kind], list(out[:n]), list(selfa), bool(raised.value)
