#!/usr/bin/env python
"""Where does the pipelined frame path (rtrb_submit / rtrb_wait) lose against raw back-to-back D2H copies?
Frames per second of config 2 through the API with (a) the real frame, (b) an 8x8-pixel window (the kernel is trivial,
the 6.2 MB copy is the same), (c) want_stats dict conversion skipped."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from raytracing_rb_b200 import Camera, World, scenes, make_opts, _abi
from raytracing_rb_b200._lib import lib
w, c = scenes.build(2)
cam = Camera(World(w), c)
r, cd = cam.renderer(), cam.camera_desc()
bufs = [torch.empty((1080, 1920, 3), dtype=torch.uint8).pin_memory().numpy() for _ in range(4)]
def run(opts, n=600, depth=3, raw=False):
    pend = []
    L = lib()
    t = C.c_int()
    st = _abi.Stats()
    t0 = time.perf_counter()
    for i in range(n):
        if len(pend) == depth:
            if raw: L.rtrb_wait(r.handle, pend.pop(0), C.byref(st))
            else: r.wait(pend.pop(0))
        if raw:
            L.rtrb_submit(r.handle, C.byref(cd), C.byref(opts), bufs[i % depth].ctypes.data, C.byref(t)); pend.append(t.value)
        else:
            pend.append(r.submit(cd, bufs[i % depth], opts))
    for p in pend:
        r.wait(p)
    return n / (time.perf_counter() - t0)
full = make_opts(seed=1, pixel_format=_abi.FMT_RGB8)
tiny = make_opts(seed=1, pixel_format=_abi.FMT_RGB8, window=(0, 0, 8, 8))
run(full, 50)
print("full frame           : %.0f frames/s" % run(full))
print("full frame, raw ctypes: %.0f frames/s" % run(full, raw=True))
print("8x8 window (copy only): %.0f frames/s" % run(tiny))
print("8x8 window, raw ctypes: %.0f frames/s" % run(tiny, raw=True))
print("full frame, depth 4   : %.0f frames/s" % run(full, depth=4, raw=True))
