"""The oracle is "parity unpinned" above Vec3 (no Ruby here, no golden image upstream).  This test holds it
against a SECOND restatement written independently from the Ruby text (oracle/restate_py.py: pure Python,
class for class after the Ruby sources): both must produce bit-identical float colours, primary hit ids and
ray counters - in the reference's own MT19937 consumption order (x outer, y inner, LIFO rays) and with the
counter RNG - on every scene family, including the Box scene and the adaptive-sampling default scene."""
import numpy as np
import pytest

from helpers_rtrb import load_scene
from raytracing_rb_b200 import RNG_CTR, RNG_MT, make_opts

SHARED = ("samples", "rays", "shadow_queries", "highlight_hits", "hits", "local_shaded", "lit_lights", "mc_rays",
          "refractions", "texel_fetches", "adaptive_pixels", "max_stack")

CASES = [
    (1, dict(), (96, 40, 132, 60)),                        # default scene: wall texture, glass sphere, adaptive 3..10
    (2, dict(width=96, height=54), None),                   # hard shadows, depth 1
    (3, dict(width=64, height=36), None),                   # depth 8, glass, textured sphere + wall
    (4, dict(width=48, height=27), (8, 6, 40, 22)),         # soft shadows, 16 spp, MC rays
    (5, dict(width=32, height=18, spp=1, grid=6), None),    # many spheres
    (6, dict(width=64, height=36), None),                   # boxes
]


@pytest.mark.parametrize("rng", ["mt", "ctr"])
@pytest.mark.parametrize("config_id,kw,window", CASES)
def test_cpp_oracle_equals_python_restatement(oracle_mod, config_id, kw, window, rng):
    from oracle import restate_py
    world, cam = load_scene(config_id, **kw)
    seed = 1
    rgb, hit, cnt = restate_py.render(world, cam, rng_mode=rng, seed=seed, window=window)
    opts = make_opts(seed=seed, rng_mode=RNG_MT if rng == "mt" else RNG_CTR, window=window)
    ref = oracle_mod.OracleScene(world.to_scene_desc()).render(cam.camera_desc(), opts, threads=1)
    if window:
        x0, y0, x1, y1 = window
        sel = (slice(y0, y1), slice(x0, x1))
    else:
        sel = (slice(None), slice(None))
    assert np.array_equal(hit[sel], ref.hit[sel])
    assert np.array_equal(rgb[sel], ref.rgb[sel]), "max abs diff %g" % np.abs(rgb[sel] - ref.rgb[sel]).max()
    for k in SHARED:
        assert cnt[k] == ref.stats[k], k
    assert ref.stats["status"] == 0
    assert cnt["rays"] > 0


def test_mt_stream_is_rubys_documented_one(oracle_mod):
    """Ruby's Random docs: `Random.new(42).rand # => 0.3745401188473625`, and Random.new(1).rand =>
    0.417022004702574: init_genrand(seed) + genrand_res53, i.e. numpy's RandomState stream."""
    assert oracle_mod.mt_res53(42, 1)[0] == 0.3745401188473625
    assert oracle_mod.mt_res53(1, 1)[0] == 0.417022004702574
    assert oracle_mod.mt_res53(1, 5) == [float(v) for v in np.random.RandomState(1).random_sample(5)]


def _fuzz_cases():
    import test_gpu_fuzz as fz
    cases = []
    for seed in (1, 2, 5, 8, 10, 13):
        cases.append(("random", fz.random_scene, seed))
    for seed in (2, 3, 9, 14, 15, 21):
        cases.append(("structural", fz.structural_scene, seed))
    for seed in (0, 3, 7, 8, 14, 18):
        cases.append(("material", fz.material_scene, seed))
    return cases


@pytest.mark.parametrize("family,builder,seed", _fuzz_cases(), ids=lambda v: v if isinstance(v, (str, int)) else "")
def test_restatements_agree_on_randomised_scenes(oracle_mod, family, builder, seed):
    """The two independent restatements on the randomised scene families of tests/test_gpu_fuzz.py (several lights,
    boxes, textures with offsets, tilted refractive planes, the libm pow path, short max_distance, skewed cameras):
    a centred window, MT19937 order, bit-identical colours / hit ids / counters."""
    from oracle import restate_py
    from raytracing_rb_b200 import Camera, World
    wdoc, cdoc = builder(seed)
    world = World(wdoc)
    cam = Camera(world, cdoc)
    W, H = cam.width, cam.height
    window = (W // 2 - 8, H // 2 - 5, W // 2 + 8, H // 2 + 5)
    try:
        rgb, hit, cnt = restate_py.render(world, cam, rng_mode="mt", seed=1, window=window)
        raised = False
    except restate_py.Raised:
        raised = True
    ref = oracle_mod.OracleScene(world.to_scene_desc()).render(cam.camera_desc(),
                                                              make_opts(seed=1, rng_mode=RNG_MT, window=window), threads=1)
    if raised:  # the Python restatement stops where the reference raises; the oracle flags and carries on
        assert ref.stats["status"] != 0
        return
    assert ref.stats["status"] == 0
    x0, y0, x1, y1 = window
    assert np.array_equal(hit[y0:y1, x0:x1], ref.hit[y0:y1, x0:x1])
    assert np.array_equal(rgb[y0:y1, x0:x1], ref.rgb[y0:y1, x0:x1])
    for k in SHARED:
        assert cnt[k] == ref.stats[k], k
