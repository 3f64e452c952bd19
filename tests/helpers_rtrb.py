"""Shared helpers for the rtrb test-suite (kept out of conftest so tests can import them by name)."""
import numpy as np


def load_scene(config_id, **kw):
    from raytracing_rb_b200 import Camera, World, scenes
    w, c = scenes.build(config_id, **kw)
    world = World(w)
    cam = Camera(world, c)
    return world, cam


def compare_u8(a, b):
    """Returns (fraction of pixels whose RGB differs by <= 1 LSB in every channel, max abs diff,
    number of pixels that differ at all)."""
    d = np.abs(a[..., :3].astype(np.int16) - b[..., :3].astype(np.int16))
    per_px = d.max(axis=-1)
    return float((per_px <= 1).mean()), int(d.max()), int((per_px > 0).sum())


def psnr_u8(a, b):
    mse = np.mean((a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)
