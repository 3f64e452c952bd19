#!/usr/bin/env python
"""Quick device-time table over the BASELINE configs (development aid; bench.py is the contract)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracing_rb_b200 import Camera, World, make_opts, scenes, PREC_FAST64, PREC_STRICT  # noqa: E402

CASES = [(1, {}), (2, {}), (3, {}), (4, dict(width=960, height=540)), (5, dict(width=480, height=270, spp=4))]
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fast64", "strict"]
for cid, kw in CASES:
    w, c = scenes.build(cid, **kw)
    cam = Camera(World(w), c)
    r = cam.renderer()
    cd = cam.camera_desc()
    for mode in modes:
        prec = PREC_STRICT if mode == "strict" else PREC_FAST64
        best = None
        for i in range(4):
            st, _ = r.render_device(cd, make_opts(seed=1, precision=prec))
            if i and (best is None or st["device_ms"] < best["device_ms"]):
                best = st
        q = best["rays"] + best["shadow_queries"]
        print("config %d %-7s %dx%d: device %8.3f ms  trace %8.3f ms  %7.2f Mq  -> %8.1f Mrays/s" % (
            cid, mode, cd.width, cd.height, best["device_ms"], best["trace_ms"], q / 1e6, q / best["device_ms"] / 1e3))
