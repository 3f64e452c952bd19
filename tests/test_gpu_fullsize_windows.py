"""Parity at the sizes BASELINE.json states: the CUDA path renders the FULL-SIZE frame of
configs 3 / 4 / 5 (1920x1080 x 4 spp depth 8; 1920x1080 x 16 spp mc 1; 3840x2160 x 64 spp, 1 024 spheres) and
64x64-pixel windows of it are held to the CPU oracle rendering the same window of the same frame
(`OracleScene.render(window=...)`, counter RNG keyed by the pixel's FULL-frame index): 8-bit output, primary hit ids
and the ray / shadow-query counters of the window.  The pixel loop being windowed is the reference's own
(`render_fork`'s column strips, camera.rb:53-65; `render_at`, camera.rb:70-99).

Windows: the four corners, the centre, one that is not aligned to the 32-pixel super-tile or the 8x4 warp block, one
straddling a super-tile row past y = 2048 (4K only), and the two 64x64 cells of the frame with the most distinct
primary hit ids (so that glass, texture and penumbra pixels are always covered).
"""
import numpy as np
import pytest

from helpers_rtrb import compare_u8, load_scene
from raytracing_rb_b200 import PREC_FAST64, make_opts

pytestmark = pytest.mark.gpu

WIN = 64
COUNTERS = ("samples", "rays", "shadow_queries", "highlight_hits", "hits", "local_shaded", "lit_lights", "mc_rays",
            "texel_fetches", "adaptive_pixels")


def fixed_windows(W, H):
    wins = [(0, 0), (W - WIN, 0), (0, H - WIN), (W - WIN, H - WIN), ((W - WIN) // 2, (H - WIN) // 2),
            (W // 3 + 13, (2 * H) // 3 + 5)]
    if H > 2048 + WIN:
        wins.append((W // 2 + 7, 2048 - 18))  # rows 2030..2093: crosses the super-tile row boundary at y = 2048
    return [(x, y, x + WIN, y + WIN) for x, y in wins]


def busiest_windows(hit, n):
    """The n cells of the 64-pixel grid with the most distinct primary hit ids (ties: first in row-major order)."""
    H, W = hit.shape
    cells = []
    for y in range(0, H - WIN + 1, WIN):
        for x in range(0, W - WIN + 1, WIN):
            cells.append((-len(np.unique(hit[y:y + WIN, x:x + WIN])), y, x))
    cells.sort()
    return [(x, y, x + WIN, y + WIN) for _, y, x in cells[:n]]


@pytest.mark.parametrize("config_id", [3, 4, 5])
def test_full_size_frame_windows_match_oracle(oracle_mod, config_id):
    world, cam = load_scene(config_id)  # the size and sample count BASELINE.json states
    W, H = cam.width, cam.height
    assert (W, H) == ((3840, 2160) if config_id == 5 else (1920, 1080))
    assert cam.pre_sample_times == {3: 4, 4: 16, 5: 64}[config_id]
    full = cam.render_frame(seed=1, precision=PREC_FAST64, want_rgb=False, want_hit=True)
    assert full.stats["status"] == 0
    sc = oracle_mod.OracleScene(world.to_scene_desc())
    cd = cam.camera_desc()
    wins = fixed_windows(W, H) + busiest_windows(full.hit, 2)
    worst, ndiff_total = 0, 0
    for win in wins:
        x0, y0, x1, y1 = win
        ref = sc.render(cd, make_opts(seed=1, window=win), want_rgb=False)
        got = cam.render_frame(seed=1, precision=PREC_FAST64, window=win, count_detail=True, want_rgb=False)
        crop = full.rgba[y0:y1, x0:x1]
        # the window rendered alone and the same pixels of the whole frame: identical bytes (tile decode, lens
        # tables, sample ordering do not depend on how the frame is cut)
        assert np.array_equal(got.rgba[y0:y1, x0:x1], crop), win
        assert np.array_equal(got.hit[y0:y1, x0:x1], full.hit[y0:y1, x0:x1]), win
        frac, maxdiff, ndiff = compare_u8(crop, ref.rgba[y0:y1, x0:x1])
        worst, ndiff_total = max(worst, maxdiff), ndiff_total + ndiff
        assert frac >= 0.999 and maxdiff <= 1, (win, frac, maxdiff)   # BASELINE tier: +-1 LSB on >= 99.9 % of pixels
        assert np.array_equal(full.hit[y0:y1, x0:x1], ref.hit[y0:y1, x0:x1]), win  # hit ids bit-exact
        for k in COUNTERS:
            assert got.stats[k] == ref.stats[k], (win, k)
        assert got.stats["status"] == ref.stats["status"] == 0
    print("config %d at %dx%dx%d spp: %d windows of %dx%d px vs oracle: max abs diff %d, differing pixels %d" % (
        config_id, W, H, cam.pre_sample_times, len(wins), WIN, WIN, worst, ndiff_total))
