#!/usr/bin/env python
"""Config 5 (1024 spheres, depth 8) at 480x270 for several sample counts: per-sample cost vs warp coherence
(at S >= 32 a warp holds samples of ONE pixel, whose ray trees are nearly identical)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracing_rb_b200 import Camera, World, make_opts, scenes  # noqa: E402

for spp in (1, 4, 16, 64):
    w, c = scenes.build(5, width=480, height=270, spp=spp)
    cam = Camera(World(w), c)
    r, cd = cam.renderer(), cam.camera_desc()
    best = None
    for i in range(3):
        st, _ = r.render_device(cd, make_opts(seed=1, skip_outputs=3))
        if i and (best is None or st["trace_ms"] < best["trace_ms"]):
            best = st
    q = best["rays"] + best["shadow_queries"]
    print("spp %3d: trace %9.3f ms  device %9.3f ms  %8.2f Mq  %8.1f Mrays/s  %.3f us per sample" % (
        spp, best["trace_ms"], best["device_ms"], q / 1e6, q / best["device_ms"] / 1e3,
        best["trace_ms"] * 1e3 / (480 * 270 * spp)))
