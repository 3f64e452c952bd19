#!/usr/bin/env python
"""Small frames through every kernel family (for compute-sanitizer): lean depth-1, depth-1 generic, ray-tree linear with
in-CTA resolve at 4 and 16 spp, ray-tree BVH at 4 and 64 spp, adaptive pass, sample-buffer fallback (3 spp), boxes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracing_rb_b200 import Camera, World, make_opts, scenes
cases = [(2, dict(width=96, height=54), {}), (1, {}, dict(width=64, height=36)), (3, dict(width=64, height=36), {}),
         (4, dict(width=48, height=27), {}), (5, dict(width=32, height=18, spp=4), {}), (5, dict(width=16, height=10, spp=64), {}),
         (6, {}, dict(width=48, height=27)), (7, dict(width=48, height=27), {})]
for cid, kw, cam_over in cases:
    w, c = scenes.build(cid, **kw)
    c = dict(c, **cam_over)
    cam = Camera(World(w), c)
    f = cam.render_frame(seed=1)
    print("config %d %dx%d spp %d: rays %d status %d" % (cid, c["width"], c["height"], c["pre_sample_times"], f.stats["rays"], f.stats["status"]), flush=True)
