#!/usr/bin/env python
"""Generates the committed fixtures of tests/golden/ (run in the build container, where /root/reference exists):

  vec3_reference_ops.json   outputs of the REFERENCE's own ext/fast_4d_matrix/fast_4d_matrix.c (compiled unmodified
                            into oracle/_ref by oracle/Makefile) on seeded random inputs: doubles as 16-hex-digit
                            bit patterns.  These pin the oracle's and the host mirror's Vec3 arithmetic on machines
                            where neither /root/reference nor oracle/_ref exists.
  frames.json + *.png       frames of the CPU ORACLE (not of the Ruby reference: no Ruby exists here) for small
                            scenes, with the SHA-256 of their float RGB bytes and their ray counters.
                            `config1_mt_seed1.png` is the frame a Ruby run of the reference's default scene
                            (`ruby src/main.rb s out.png config/world.yml config/camera.yml`, Random.srand(1),
                            ground texture lines dropped) is predicted to produce at render_at's output: the file to
                            diff against the day such a run exists.

Usage: python tests/golden/make_golden.py"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle  # noqa: E402
from raytracing_rb_b200 import RNG_CTR, RNG_MT, Camera, World, make_opts, scenes, write_png  # noqa: E402

FRAMES = [  # name, config id, scene kwargs, rng, seed
    ("config1_mt_seed1", 1, {}, "mt", 1),
    ("config1_ctr_seed1", 1, {}, "ctr", 1),
    ("config2_96x54_ctr_seed1", 2, dict(width=96, height=54), "ctr", 1),
    ("config3_96x54_ctr_seed1", 3, dict(width=96, height=54), "ctr", 1),
    ("config4_64x36_mt_seed1", 4, dict(width=64, height=36), "mt", 1),
    ("config6_boxes_96x54_ctr_seed1", 6, dict(width=96, height=54), "ctr", 1),
]
COUNTERS = ("samples", "rays", "shadow_queries", "highlight_hits", "hits", "local_shaded", "lit_lights", "mc_rays",
            "refractions", "texel_fetches", "adaptive_pixels", "status")


def hexd(x):
    return "%016x" % struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def vec3_cases(n=300, seed=20261018):
    rs = np.random.RandomState(seed)
    methods = [("dot", 1), ("cos", 1), ("cross", 1), ("+", 1), ("-", 1), ("*", 1), ("*", 2), ("/", 2), ("r", 0),
               ("r2", 0), ("normalize", 0), ("-@", 0)]
    out = []
    for i in range(n):
        m, arity = methods[i % len(methods)]
        scale = 10.0 ** rs.randint(-3, 4)
        a = [float(v) for v in rs.uniform(-1, 1, 3) * scale]
        b = None
        if arity == 1:
            b = [float(v) for v in rs.uniform(-1, 1, 3) * scale]
        elif arity == 2:
            b = float(rs.uniform(0.1, 3.0) * scale)
        kind, vals, _, raised = oracle.ref_vec3_call(m, a, b)
        assert not raised
        out.append({"method": m, "a": [hexd(v) for v in a],
                    "b": None if b is None else ([hexd(v) for v in b] if isinstance(b, list) else hexd(b)),
                    "kind": kind, "out": [hexd(v) for v in (vals[:1] if kind == "float" else vals[:3])]})
    return out


def render(config_id, kw, rng, seed):
    wdoc, cdoc = scenes.build(config_id, **kw)
    world = World(wdoc)
    cam = Camera(world, cdoc)
    opts = make_opts(seed=seed, rng_mode=RNG_MT if rng == "mt" else RNG_CTR)
    return oracle.OracleScene(world.to_scene_desc()).render(cam.camera_desc(), opts, threads=1)


def main():
    if oracle.ref_lib() is None:
        raise SystemExit("oracle/_ref is not built: run `make -C oracle` where /root/reference exists")
    with open(os.path.join(HERE, "vec3_reference_ops.json"), "w") as f:
        json.dump({"source": "reference ext/fast_4d_matrix/fast_4d_matrix.c compiled unmodified (oracle/_ref)",
                   "cases": vec3_cases()}, f, indent=0)
    meta = {}
    for name, cid, kw, rng, seed in FRAMES:
        fr = render(cid, kw, rng, seed)
        write_png(os.path.join(HERE, name + ".png"), fr.rgba)
        meta[name] = {"config": cid, "kwargs": kw, "rng": rng, "seed": seed,
                      "rgb_sha256": hashlib.sha256(np.ascontiguousarray(fr.rgb).tobytes()).hexdigest(),
                      "hit_sha256": hashlib.sha256(np.ascontiguousarray(fr.hit).tobytes()).hexdigest(),
                      "counters": {k: int(fr.stats[k]) for k in COUNTERS}}
        print(name, meta[name]["counters"]["rays"], meta[name]["rgb_sha256"][:12])
    with open(os.path.join(HERE, "frames.json"), "w") as f:
        json.dump({"source": "oracle/rtrb_oracle.cpp (CPU restatement; NOT a Ruby run)", "frames": meta}, f, indent=1)


if __name__ == "__main__":
    main()
