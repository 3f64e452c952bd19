"""Host logic on CPU: the Python mirror of the reference's configuration layer must follow the reference's
rules (src/configurable_object.rb:11-49, world.rb:15-34, plane.rb:57, texture.rb:23-28) because the scene the
CUDA core bakes is whatever this layer hands it."""
import os

import pytest

from raytracing_rb_b200 import Box, Camera, ConfigurableObject, Plane, Sphere, Vec3, World, _abi, scenes
from raytracing_rb_b200.texture import Texture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configurable_object_vector_rules():
    c = ConfigurableObject({
        "a": [1, 2, 3],                    # three numerics -> Vec3 of floats (:30-32)
        "b": [1, 2],                       # any other array -> array_parse_vector, which DROPS scalars (:13-22)
        "c": {"d": [0.5, 1, 2], "e": "s"},  # hashes recurse, other values pass through (:28-29, :35-36)
        "f": [[1, 2, 3], {"g": [4, 5, 6]}, [7, [8, 9, 10]]],
        "h": [True, 2, 3],                 # true is not Numeric in Ruby -> not a vector; scalars dropped
        "n": 7,
    })
    assert isinstance(c.a, Vec3) and c.a.to_a() == [1.0, 2.0, 3.0] and all(isinstance(v, float) for v in c.a.to_a())
    assert c.b == []
    assert isinstance(c.c["d"], Vec3) and c.c["e"] == "s"
    assert isinstance(c.f[0], Vec3) and isinstance(c.f[1]["g"], Vec3)
    assert len(c.f[2]) == 1 and isinstance(c.f[2][0], Vec3) and c.f[2][0].to_a() == [8.0, 9.0, 10.0]
    assert c.h == [] and c.n == 7


def test_shipped_configs_load_like_the_reference():
    world = World(os.path.join(ROOT, "config", "world.yml"))
    cam = Camera(world, os.path.join(ROOT, "config", "camera.yml"))
    assert [type(o).__name__ for o in world.world_objects] == ["Plane", "Plane", "Sphere"]
    assert world.max_distance == 10000 and world.soft_shadow_exponent == 2 and len(world.lights) == 1
    assert (cam.width, cam.height, cam.pre_sample_times, cam.max_sample_times) == (192, 108, 3, 10)
    assert cam.front.to_a() == [1.0, 0.0, 0.0] and cam.trace_depth == 4 and cam.monte_carlo_diffusion_times == 1
    sd = world.to_scene_desc().desc
    assert (sd.n_objects, sd.n_lights, sd.n_textures) == (3, 1, 1)
    wall = sd.objects[1]
    assert wall.type == _abi.OBJ_PLANE and wall.texture == 0 and wall.has_refraction == 0
    sph = sd.objects[2]
    assert sph.type == _abi.OBJ_SPHERE and sph.has_refraction == 1 and sph.refractive_rate == 1.6 and sph.radius == 0.7
    c = cam.camera_desc()
    assert c.image_distance == 0.01714573877962683 and c.variant_threshold == 0.001


def test_unknown_object_and_light_types_raise_like_eval_would():
    base = {"max_distance": 1, "soft_shadow_exponent": 2, "lights": [], "world_objects": []}
    with pytest.raises(NameError):
        World(dict(base, world_objects=[{"type": "Torus", "properties": {}}]))
    with pytest.raises(NameError):
        World(dict(base, lights=[{"type": "Area", "properties": {}}]))


def test_plane_refraction_follows_ruby_truthiness():
    g = scenes.ground()
    assert Plane(ConfigurableObject({"p": g["properties"]}).p).to_desc(-1).has_refraction == 0   # key absent -> nil
    g["properties"]["refractive_rate"] = 0      # 0 is truthy in Ruby (plane.rb:57 `if self.refractive_rate`)
    g["properties"]["refractive_attenuation"] = [0.1, 0.1, 0.1]
    d = Plane(ConfigurableObject({"p": g["properties"]}).p).to_desc(-1)
    assert d.has_refraction == 1 and d.refractive_rate == 0.0


def test_missing_mandatory_keys_raise():
    s = scenes.matte("s", (1, 2, 3), 0.5, (1, 1, 1))["properties"]
    del s["refractive_rate"]                    # sphere.rb:93 divides by it unconditionally
    with pytest.raises(TypeError):
        Sphere(ConfigurableObject({"p": s}).p).to_desc(-1)
    b = scenes.box("b", [0, 0, 0], [1, 0, 0], [0, 0, 1], (1, 1, 1))["properties"]
    del b["width_up"]
    with pytest.raises(TypeError):
        Box(ConfigurableObject({"p": b}).p)


def test_texture_index_rule_truncates_then_floors():
    t = Texture("./textures/RubyOnRails.png", 0.015, 0.015)
    assert (t.width, t.height) == (122, 158)
    assert t.texel_index(-0.5, 0.25) == (89, 16)     # SURVEY 8c KAT 6: trunc(-33.3) = -33; -33 mod 122 = 89
    assert t.texel_index(0.0, 0.0) == (0, 0) and t.texel_index(-0.014, 0.0) == (0, 0)  # texel 0 is double width around 0
    assert t.texel_index(-0.016, 0.0) == (121, 0)
    assert tuple(t.rgb8[40, 30]) == (166, 44, 57)


def test_scene_builders_match_survey_shapes():
    w2, c2 = scenes.build(2)
    assert len(w2["world_objects"]) == 17 and (c2["width"], c2["height"], c2["trace_depth"]) == (1920, 1080, 1)
    w5, c5 = scenes.build(5)
    assert len(w5["world_objects"]) == 2 + 1024 and c5["pre_sample_times"] == c5["max_sample_times"] == 64
    for cid in (2, 3, 4, 5, 6, 7):  # materials keep every channel's sum <= 1 (ray_tracer.rb:294-296 raises otherwise)
        w, _ = scenes.build(cid) if cid != 5 else scenes.build(5, grid=4)
        for o in w["world_objects"]:
            p = o["properties"]
            tot = [p["diffuse_rate"][k] + p["ambient"][k] + p["reflective_attenuation"][k] +
                   (p.get("refractive_attenuation") or [0, 0, 0])[k] for k in range(3)]
            assert max(tot) <= 1.0 + 1e-12, (cid, p["name"], tot)
