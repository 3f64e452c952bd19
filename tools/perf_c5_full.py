#!/usr/bin/env python
"""BASELINE.json configs[4] at full size on one GPU: 3840x2160, 1024 spheres, depth 8, 64 spp (device time only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracing_rb_b200 import Camera, World, make_opts, scenes  # noqa: E402

w, c = scenes.build(5)
cam = Camera(World(w), c)
r, cd = cam.renderer(), cam.camera_desc()
for i in range(3):
    st, _ = r.render_device(cd, make_opts(seed=1, skip_outputs=3))
    q = st["rays"] + st["shadow_queries"]
    print("frame %d: device %.1f ms  trace %.1f ms  %.0f M ray queries  %.1f Mrays/s" % (
        i, st["device_ms"], st["trace_ms"], q / 1e6, q / st["device_ms"] / 1e3))
