#!/usr/bin/env python
"""Renders N frames of one BASELINE config on cuda:0 and prints device times (development aid; also the
command `ncu` wraps for the per-kernel captures under profiles/).
Usage: run_config.py <config> [--width W --height H --spp S --frames N --strict --detail]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracing_rb_b200 import Camera, World, make_opts, scenes, PREC_FAST64, PREC_STRICT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("config", type=int)
ap.add_argument("--width", type=int, default=0)
ap.add_argument("--height", type=int, default=0)
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--strict", action="store_true")
ap.add_argument("--detail", action="store_true")
a = ap.parse_args()
kw = {}
if a.width:
    kw.update(width=a.width, height=a.height)
if a.spp and a.config in (5, 7):
    kw["spp"] = a.spp
w, c = scenes.build(a.config, **kw)
if a.spp and a.config not in (5, 7):
    c = dict(c, pre_sample_times=a.spp, max_sample_times=a.spp)
cam = Camera(World(w), c)
r, cd = cam.renderer(), cam.camera_desc()
for i in range(a.frames):
    st, _ = r.render_device(cd, make_opts(seed=1, skip_outputs=3, count_detail=a.detail,
                                          precision=PREC_STRICT if a.strict else PREC_FAST64))
    q = st["rays"] + st["shadow_queries"]
    print("config %d %dx%d spp %d frame %d: device %.3f ms  trace %.3f ms  %.2f M ray queries  %.1f Mrays/s" % (
        a.config, cd.width, cd.height, cd.pre_sample_times, i, st["device_ms"], st["trace_ms"], q / 1e6,
        q / st["device_ms"] / 1e3), flush=True)
    if a.detail:
        print("   detail: rays %d hits %d exact %d box_tests %d box_accepts %d cover_box %d" % (
            st["rays"], st["hits"], st["exact_tests"], st["box_tests"], st["box_accepts"], st["cover_box"]), flush=True)
