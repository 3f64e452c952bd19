#!/usr/bin/env python
"""Renders a few frames of one BASELINE config on cuda:0 (profiling / ncu target; not a bench)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracing_rb_b200 import Camera, World, make_opts, scenes, PREC_FAST64, PREC_STRICT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--precision", default="fast64")
ap.add_argument("--width", type=int, default=0)
ap.add_argument("--height", type=int, default=0)
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--lean", action="store_true", help="RGBA8 only (what bench.py renders): skip the optional float RGB / hit-id frames")
a = ap.parse_args()
kw = {}
if a.width:
    kw.update(width=a.width, height=a.height)
if a.spp:
    kw.update(spp=a.spp)
w, c = scenes.build(a.config, **kw)
cam = Camera(World(w), c)
prec = PREC_STRICT if a.precision == "strict" else PREC_FAST64
for i in range(a.frames):
    st, _ = cam.renderer().render_device(cam.camera_desc(), make_opts(seed=1, precision=prec, skip_outputs=3 if a.lean else 0))
    print("frame %d: device %.3f ms trace %.3f ms rays %d shadow %d" % (i, st["device_ms"], st["trace_ms"], st["rays"],
                                                                      st["shadow_queries"]))
