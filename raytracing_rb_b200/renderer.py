"""Thin object wrapper over the C ABI (include/rtrb_b200.h): one Renderer = one baked scene on one
GPU.  Frames come back as numpy arrays in the layout Camera#render_at defines (row = y, col = x)."""
import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, lib


class Frame:
    """One rendered frame: rgba uint8 [H,W,4]; rgb float64 [H,W,3] (render_at's unclamped colour) or
    None; hit int32 [H,W] primary hit ids or None; stats dict; status word."""

    def __init__(self, rgba, rgb, hit, stats, code):
        self.rgba, self.rgb, self.hit, self.stats, self.code = rgba, rgb, hit, stats, code

    @property
    def status(self):
        return self.stats["status"]

    @property
    def raised(self):
        return self.code == _abi.RTRB_ERR_RAISED


def make_opts(rng_mode=_abi.RNG_CTR, precision=_abi.PREC_DEFAULT, seed=1, window=None, tile_rank=0, tile_world=1,
              count_detail=False, stream=None, rgba_device_out=None, skip_outputs=0, pixel_format=_abi.FMT_RGBA8):
    o = _abi.RenderOpts()
    o.rng_mode, o.precision, o.seed = rng_mode, precision, seed
    if window:
        o.x0, o.y0, o.x1, o.y1 = window
    o.tile_rank, o.tile_world = tile_rank, tile_world
    o.count_detail = 1 if count_detail else 0
    o.skip_outputs = skip_outputs
    o.stream = stream
    o.rgba_device_out = rgba_device_out
    o.pixel_format = pixel_format
    return o


class Renderer:
    def __init__(self, scene, device=0):
        self._scene = scene  # SceneDescHolder; only needed during create, kept for introspection
        self._h = C.c_void_p()
        self.device = device
        check(lib().rtrb_renderer_create(C.byref(scene.desc), device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().rtrb_renderer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def render_device(self, cam, opts=None, want_stats=True):
        """Frame stays in device memory.  Returns (stats dict or None, status code)."""
        opts = opts or make_opts()
        st = _abi.Stats()
        code = lib().rtrb_render_device(self._h, C.byref(cam), C.byref(opts), C.byref(st) if want_stats else None)
        check(code, allow_raised=True)
        return (st.as_dict() if want_stats else None), code

    def download(self, width, height, want_rgb=True, want_hit=True, channels=4):
        rgba = np.empty((height, width, channels), np.uint8)
        rgb = np.empty((height, width, 3), np.float64) if want_rgb else None
        hit = np.empty((height, width), np.int32) if want_hit else None
        check(lib().rtrb_download(self._h, rgba.ctypes.data, rgb.ctypes.data if want_rgb else None,
                                  hit.ctypes.data if want_hit else None))
        return rgba, rgb, hit

    def render(self, cam, opts=None, want_rgb=True, want_hit=True, out_rgba=None):
        """The frame-level call the Ruby shim binds (host buffers in, host buffers out)."""
        opts = opts or make_opts()
        H, W = cam.height, cam.width
        if opts.pixel_format == _abi.FMT_PNG_RGB8:   # H scanlines of (filter byte 0, W x RGB8): IDAT's payload before deflate
            shape = (H, W * 3 + 1)
        else:
            shape = (H, W, 3 if opts.pixel_format == _abi.FMT_RGB8 else 4)
        rgba = out_rgba if out_rgba is not None else np.empty(shape, np.uint8)
        rgb = np.empty((H, W, 3), np.float64) if want_rgb else None
        hit = np.empty((H, W), np.int32) if want_hit else None
        st = _abi.Stats()
        code = lib().rtrb_render(self._h, C.byref(cam), C.byref(opts), rgba.ctypes.data,
                                 rgb.ctypes.data if want_rgb else None, hit.ctypes.data if want_hit else None,
                                 C.byref(st))
        check(code, allow_raised=True)
        return Frame(rgba, rgb, hit, st.as_dict(), code)

    def submit(self, cam, out_rgba, opts=None):
        """Pipelined frame: returns a ticket at once; the RGBA8 frame lands in `out_rgba` (a numpy view
        of pinned host memory for full speed) by the time `wait(ticket)` returns.  Up to four in flight."""
        opts = opts or make_opts()
        t = C.c_int()
        check(lib().rtrb_submit(self._h, C.byref(cam), C.byref(opts), out_rgba.ctypes.data, C.byref(t)))
        return t.value

    def wait(self, ticket):
        st = _abi.Stats()
        code = lib().rtrb_wait(self._h, ticket, C.byref(st))
        check(code, allow_raised=True)
        return st.as_dict(), code

    def peer_push(self, src_ptr, dst_peer_ptr, nbytes, after_stream=None):
        """Copy-engine gather: queue a D2D copy to a peer mapping behind `after_stream` (see rtrb_peer_push)."""
        check(lib().rtrb_peer_push(self._h, C.c_void_p(src_ptr), C.c_void_p(dst_peer_ptr), nbytes,
                                   C.c_void_p(after_stream) if after_stream else None))

    def peer_push_join(self, stream=None):
        check(lib().rtrb_peer_push_join(self._h, C.c_void_p(stream) if stream else None))

    def last_mt_passes(self):
        """Passes the last RNG_MT frame needed to reach the fixed point of its stream offsets."""
        return int(lib().rtrb_last_mt_passes(self._h))

    def framebuffer_ptr(self, width, height):
        p = C.c_void_p()
        check(lib().rtrb_framebuffer_device_ptr(self._h, width, height, C.byref(p)))
        return p.value

    def framebuffer_download(self, width, height, out):
        check(lib().rtrb_framebuffer_download(self._h, width, height, out.ctypes.data))
        return out

    def framebuffer_copy_async(self, nbytes, out, stream=None):
        """Queues framebuffer[:nbytes] -> `out` (pinned host memory) on `stream`; the caller synchronises."""
        check(lib().rtrb_framebuffer_copy_async(self._h, nbytes, out.ctypes.data, C.c_void_p(stream) if stream else None))

    def framebuffer_ipc_export(self, width, height):
        buf = (C.c_uint8 * 64)()
        check(lib().rtrb_framebuffer_ipc_export(self._h, width, height, buf))
        return bytes(buf)


def ipc_open(device, handle_bytes):
    buf = (C.c_uint8 * 64).from_buffer_copy(handle_bytes)
    p = C.c_void_p()
    check(lib().rtrb_ipc_open(device, buf, C.byref(p)))
    return p.value


def ipc_close(device, ptr):
    check(lib().rtrb_ipc_close(device, C.c_void_p(ptr)))


def render_multi(renderers, cam, opts=None, want_rgb=True, want_hit=True):
    """In-process multi-GPU frame: image tiles interleaved over len(renderers) GPUs, written through
    peer mappings into renderers[0]'s framebuffer (no collective)."""
    opts = opts or make_opts()
    n = len(renderers)
    arr = (C.c_void_p * n)(*[r.handle for r in renderers])
    H, W = cam.height, cam.width
    rgba = np.empty((H, W, 4), np.uint8)
    rgb = np.empty((H, W, 3), np.float64) if want_rgb else None
    hit = np.empty((H, W), np.int32) if want_hit else None
    st = _abi.Stats()
    code = lib().rtrb_render_multi(arr, n, C.byref(cam), C.byref(opts), rgba.ctypes.data,
                                   rgb.ctypes.data if want_rgb else None, hit.ctypes.data if want_hit else None,
                                   C.byref(st))
    check(code, allow_raised=True)
    return Frame(rgba, rgb, hit, st.as_dict(), code)


def tile_partition(width, height, tile_rank=0, tile_world=1, window=None):
    """The super-tile ids (ty * ceil(width/32) + tx) `tile_rank` of `tile_world` renders.  Pure host logic."""
    win = (C.c_int32 * 4)(*(window or (0, 0, 0, 0)))
    n = C.c_int()
    check(lib().rtrb_tile_partition(width, height, win, tile_rank, tile_world, None, 0, C.byref(n)))
    out = (C.c_int32 * max(1, n.value))()
    check(lib().rtrb_tile_partition(width, height, win, tile_rank, tile_world, out, n.value, C.byref(n)))
    return list(out[:n.value])


def deal_frames(rank, world, frames_per_gpu):
    """Whole-frame dealing of a batch of small frames: rank r renders frames
    [r*B, (r+1)*B) of the step and frame f lands in slot f of rank 0's framebuffer.  Returns the list of
    (frame index, slot) pairs of `rank`.  Pure host logic."""
    if world < 1 or not 0 <= rank < world or frames_per_gpu < 0:
        raise ValueError("bad rank/world/frames_per_gpu")
    first = rank * frames_per_gpu
    return [(first + f, first + f) for f in range(frames_per_gpu)]


def measure_fma_peak(device=0, fp64=True):
    v = C.c_double()
    check(lib().rtrb_measure_fma_peak(device, 1 if fp64 else 0, C.byref(v)))
    return v.value


def device_count():
    n = C.c_int()
    check(lib().rtrb_device_count(C.byref(n)))
    return n.value
