"""The five BASELINE.json workloads made concrete exactly as SURVEY.md 8d specifies them (shapes,
seeds, materials).  Each builder returns (world_doc, camera_doc): plain dicts in the reference's
YAML scene format, loadable by World(doc)/Camera(world, doc) here and — once dumped with
`write_yaml` — by the reference's own ConfigurableObject (src/configurable_object.rb:43-49)."""
import copy
import os

import numpy as np
import yaml

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEXTURE = "./textures/RubyOnRails.png"

COMMON_CAMERA = {  # reference config/camera.yml:1-9
    "position": [0, 0, 0], "up": [0, 0, 1], "front": [1, 0, 0],
    "retina_width": 0.016, "retina_height": 0.009,
    "aperture_radius": 0.001, "image_distance": 0.01714573877962683, "focal_distance": 0.017,
    "width": 192, "height": 108, "pre_sample_times": 3, "max_sample_times": 10, "variant_threshold": 0.001,
    "trace_depth": 4, "monte_carlo_diffusion_times": 1,
}


def _c3(v):
    return [float(v)] * 3


def ground():
    return {"type": "Plane", "properties": {
        "name": "ground", "point": [0, 0, -1], "front": [0, 0, 1], "up": [1, 0, 0], "u_unit": 1, "v_unit": 1,
        "diffuse_rate": _c3(0.6), "reflective_attenuation": _c3(0.39), "ambient": _c3(0.01)}}


def wall(x0):
    return {"type": "Plane", "properties": {
        "name": "front wall", "point": [x0, 0, 0], "front": [-1, 0, 0], "up": [0, 0, -1], "u_unit": 1, "v_unit": 1,
        "diffuse_rate": _c3(0.6), "reflective_attenuation": _c3(0.39), "ambient": _c3(0.01),
        "texture_file_path": TEXTURE, "texture_horizontal_scale": 0.015, "texture_vertical_scale": 0.015}}


def matte(name, center, radius, tint):
    return {"type": "Sphere", "properties": {
        "name": name, "center": [float(c) for c in center], "radius": float(radius), "refractive_rate": 1.6,
        "reflective_attenuation": _c3(0.39), "refractive_attenuation": _c3(0.0),
        "diffuse_rate": [0.6 * float(t) for t in tint], "ambient": _c3(0.01)}}


def glass(name, center, radius):
    return {"type": "Sphere", "properties": {
        "name": name, "center": [float(c) for c in center], "radius": float(radius), "refractive_rate": 1.6,
        "reflective_attenuation": _c3(0.1), "refractive_attenuation": _c3(0.8),
        "diffuse_rate": _c3(0.09), "ambient": _c3(0.01)}}


def light(position, radius):
    return {"type": "Spot", "properties": {
        "name": "Main", "position": [float(p) for p in position], "radius": float(radius), "color": [1, 1, 1],
        "high_light_rate": 1, "high_light_angle": 3}}


def _world(objects, lights):
    return {"max_distance": 10000, "soft_shadow_exponent": 2, "lights": lights, "world_objects": objects}


def _camera(**kw):
    c = copy.deepcopy(COMMON_CAMERA)
    c.update(kw)
    return c


def _config2_spheres():
    rs = np.random.RandomState(20261018)
    out = []
    for i in range(16):
        x = rs.uniform(3, 12)
        y = rs.uniform(-5, 5)
        r = rs.uniform(0.2, 0.7)
        tint = rs.uniform(0.3, 1.0, size=3)
        out.append((i, (x, y, -1 + r), r, tint))
    return out


def config1():
    """Reference default scene (config/*.yml, ground texture dropped — file absent upstream)."""
    with open(os.path.join(REPO_ROOT, "config", "world.yml")) as f:
        w = yaml.safe_load(f)
    with open(os.path.join(REPO_ROOT, "config", "camera.yml")) as f:
        c = yaml.safe_load(f)
    return w, c


def config2(width=1920, height=1080):
    """Deterministic: ground + 16 matte spheres, hard shadows, 1 spp, root rays only."""
    objs = [ground()] + [matte("s%d" % i, c, r, t) for i, c, r, t in _config2_spheres()]
    return (_world(objs, [light([5, -4, 4], 0.0)]),
            _camera(width=width, height=height, aperture_radius=0.0, pre_sample_times=1, max_sample_times=1,
                    trace_depth=1, monte_carlo_diffusion_times=0))


def _config3_objects():
    tex_c, tex_r = (5.0, 0.0, -0.3), 0.7
    objs = [ground(), wall(15)]
    for i, c, r, t in _config2_spheres():
        d = float(np.sqrt(sum((a - b) ** 2 for a, b in zip(c, tex_c))))
        if d < r + tex_r:
            continue  # overlaps the textured sphere
        objs.append(glass("s%d" % i, c, r) if i % 2 == 1 else matte("s%d" % i, c, r, t))
    ts = matte("rails sphere", tex_c, tex_r, (1.0, 1.0, 1.0))
    ts["properties"].update({
        "greenwich_vec": [-1, 0, 0], "north_pole_vec": [0, 0, 1], "texture_file_path": TEXTURE,
        "texture_horizontal_scale": 0.0082, "texture_vertical_scale": 0.0063,
        "texture_u_offset": 0, "texture_v_offset": 0})
    objs.append(ts)
    return objs


def config3(width=1920, height=1080):
    """Recursion + texture: depth 8, 4 samples through the aperture, glass + matte + textured sphere."""
    return (_world(_config3_objects(), [light([5, -4, 4], 0.8)]),
            _camera(width=width, height=height, aperture_radius=0.001, pre_sample_times=4, max_sample_times=4,
                    trace_depth=8, monte_carlo_diffusion_times=0))


def config4(width=1920, height=1080):
    """Soft shadows from the area light, 16 samples, depth 4, 1 Monte-Carlo diffuse ray (counter RNG)."""
    return (_world(_config3_objects(), [light([5, -4, 4], 0.8)]),
            _camera(width=width, height=height, aperture_radius=0.001, pre_sample_times=16, max_sample_times=16,
                    trace_depth=4, monte_carlo_diffusion_times=1))


def config5(width=3840, height=2160, spp=64, grid=32):
    """Scale: ground + wall(45) + grid x grid spheres (odd = glass, even = matte), depth 8, 64 spp."""
    rs = np.random.RandomState(5)
    objs = [ground(), wall(45)]
    xs, ys = np.linspace(3, 40, grid), np.linspace(-18, 18, grid)
    k = 0
    for gx in xs:
        for gy in ys:
            jx, jy = rs.uniform(-0.3, 0.3, size=2)
            r = rs.uniform(0.15, 0.45)
            tint = rs.uniform(0.3, 1.0, size=3)
            c = (gx + jx, gy + jy, -1 + r)
            objs.append(glass("g%d" % k, c, r) if k % 2 == 1 else matte("m%d" % k, c, r, tint))
            k += 1
    return (_world(objs, [light([20, -4, 8], 0.8)]),
            _camera(width=width, height=height, aperture_radius=0.001, pre_sample_times=spp, max_sample_times=spp,
                    trace_depth=8, monte_carlo_diffusion_times=0))


def box(name, point, front, up, widths, glassy=True):
    """A Box in the reference's world.yml format (config/world.yml:114-129, commented out upstream)."""
    p = {"name": name, "point": [float(v) for v in point], "front": [float(v) for v in front],
         "up": [float(v) for v in up], "width_front": float(widths[0]), "width_up": float(widths[1]),
         "width_left": float(widths[2]), "ambient": _c3(0.01)}
    if glassy:  # the upstream "big box" material
        p.update({"refractive_rate": 1.1, "reflective_attenuation": _c3(0.1), "refractive_attenuation": _c3(0.4),
                  "diffuse_rate": _c3(0.4)})
    else:       # no refractive_rate key: the faces never refract (plane.rb:57)
        p.update({"reflective_attenuation": _c3(0.3), "diffuse_rate": [0.5, 0.6, 0.3]})
    return {"type": "Box", "properties": p}


def config6(width=192, height=108):
    """the reference default scene with its commented-out `big box`
    (config/world.yml:114-129) switched on, plus a matte box lying on the ground near the camera."""
    w, c = config1()
    w = copy.deepcopy(w)
    w["world_objects"].append(box("big box", [10, -2.2, 1], [1, 0, 0], [0, 0, 1], (0.4, 4.0, 0.4)))
    w["world_objects"].append(box("crate", [6, 1.5, -0.6], [1, 1, 0], [0, 0, 1], (0.9, 0.8, 1.3), glassy=False))
    c = dict(c, width=width, height=height)
    return w, c


def config7(width=480, height=270, spp=2, grid=8):
    """Boxes among many spheres (the BVH filter path, > 32 bounded objects): a glass box, a matte box and a
    skewed box whose `up` is not perpendicular to `front` (no filter bound: always exact-tested)."""
    w, c = config5(width=width, height=height, spp=spp, grid=grid)
    objs = w["world_objects"]
    objs.insert(5, box("glass slab", [9, -3, 0.2], [1, 0.3, 0], [0, 0, 1], (0.6, 2.4, 1.8)))
    objs.insert(20, box("crate", [14, 4, -0.5], [0, 1, 0], [0, 0, 1], (1.5, 1.0, 1.5), glassy=False))
    objs.append(box("skewed", [7, 2.5, -0.4], [1, 0, 0.2], [0, 0.1, 1], (1.0, 1.2, 0.8), glassy=False))
    return w, dict(c, trace_depth=5)


CONFIGS = {1: config1, 2: config2, 3: config3, 4: config4, 5: config5, 6: config6, 7: config7}
NAMES = {
    1: "config1: reference default scene 192x108 (adaptive 3..10 spp, depth 4, mc 1)",
    2: "config2: 1920x1080 ground + 16 matte spheres, hard shadows, 1 spp, depth 1",
    3: "config3: 1920x1080 depth 8, textured sphere + wall, glass, 4 spp",
    4: "config4: 1920x1080 soft shadows r=0.8, 16 spp, depth 4, mc 1",
    5: "config5: 3840x2160 1024 spheres, depth 8, 64 spp",
    6: "config6: reference default scene + its commented-out Box (world.yml:114-129) + a matte box",
    7: "config7: boxes among 64 spheres (BVH filter), depth 5",
}


def build(config_id, **kw):
    return CONFIGS[config_id](**kw)


def write_yaml(config_id, out_dir, **kw):
    w, c = build(config_id, **kw)
    os.makedirs(out_dir, exist_ok=True)
    wp, cp = os.path.join(out_dir, "world_%d.yml" % config_id), os.path.join(out_dir, "camera_%d.yml" % config_id)
    with open(wp, "w") as f:
        yaml.safe_dump(w, f, default_flow_style=None, sort_keys=False)
    with open(cp, "w") as f:
        yaml.safe_dump(c, f, default_flow_style=None, sort_keys=False)
    return wp, cp
