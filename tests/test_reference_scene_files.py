"""The reference's UNTOUCHED scene files through the host loader.

/root/reference/config/world.yml carries two things the shipped copy (config/world.yml) does not: duplicate keys in
the "front wall" entry (:40-42 and again :47-49 give diffuse_rate / reflective_attenuation / ambient; YAML lets the later key
win, as Ruby's Psych does) and texture keys on the ground plane that name a file absent from the reference repository
(:30-32, ./textures/floor.jpg).  Fed through ConfigurableObject's rules (src/configurable_object.rb:26-49) with only
that missing file substituted away, the reference's own files must produce exactly the scene description the shipped
files produce.  Skipped where /root/reference does not exist (the GPU box)."""
import ctypes as C
import os

import numpy as np
import pytest
import yaml

from raytracing_rb_b200 import Camera, World, _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "config", "world.yml")),
                                reason="/root/reference is not mounted here")


def _struct_bytes(s):
    return bytes(memoryview(s).cast("B")) if not isinstance(s, C.Structure) else C.string_at(C.addressof(s), C.sizeof(s))


def _scene_fingerprint(world):
    holder = world.to_scene_desc()
    sd = holder.desc
    objs = [_struct_bytes(sd.objects[i]) for i in range(sd.n_objects)]
    lights = [_struct_bytes(sd.lights[i]) for i in range(sd.n_lights)]
    texs = []
    for i in range(sd.n_textures):
        t = sd.textures[i]
        texs.append((t.width, t.height, C.string_at(t.rgb8, t.width * t.height * 3)))
    return (sd.max_distance, sd.soft_shadow_exponent, sd.n_objects, sd.n_lights, sd.n_textures, objs, lights, texs)


def _load_reference_world():
    with open(os.path.join(REF, "config", "world.yml")) as f:
        text = f.read()
    doc = yaml.safe_load(text)
    dropped = []
    for item in doc["world_objects"]:
        p = item["properties"]
        path = p.get("texture_file_path")
        if path and not os.path.isfile(os.path.join(REF, path)):
            # the ONE substitution: a texture file the reference repository does not contain
            dropped.append((p["name"], path))
            for k in ("texture_file_path", "texture_horizontal_scale", "texture_vertical_scale",
                      "texture_u_offset", "texture_v_offset"):
                p.pop(k, None)
    return doc, text, dropped


def test_reference_world_yml_gives_the_shipped_scene_byte_for_byte():
    doc, text, dropped = _load_reference_world()
    assert dropped == [("ground", "./textures/floor.jpg")]
    # the duplicate keys are really there, and the later occurrence is what a YAML loader keeps
    start = text.index("front wall")
    wall_block = text[start:text.index("\n#", start)]  # the active entry, up to the first commented-out block
    assert wall_block.count("diffuse_rate:") == 2 and wall_block.count("ambient:") == 2
    wall = [o for o in doc["world_objects"] if o["properties"]["name"] == "front wall"][0]["properties"]
    assert wall["diffuse_rate"] == [0.6, 0.6, 0.6] and wall["ambient"] == [0.01, 0.01, 0.01]
    cwd = os.getcwd()
    os.chdir(ROOT)  # the reference opens './textures/RubyOnRails.png' relative to the process CWD (texture.rb:12)
    try:
        ref_world = World(doc)
        own_world = World(os.path.join(ROOT, "config", "world.yml"))
        a, b = _scene_fingerprint(ref_world), _scene_fingerprint(own_world)
    finally:
        os.chdir(cwd)
    assert a[:5] == b[:5]
    assert a[5] == b[5], "object descriptors differ"
    assert a[6] == b[6], "light descriptors differ"
    assert a[7] == b[7], "decoded textures differ"
    # and the texture the reference ships is the one this repository ships
    with open(os.path.join(REF, "textures", "RubyOnRails.png"), "rb") as f1, open(os.path.join(ROOT, "textures", "RubyOnRails.png"), "rb") as f2:
        assert f1.read() == f2.read()


def test_reference_camera_yml_gives_the_shipped_camera_byte_for_byte():
    world = World(os.path.join(ROOT, "config", "world.yml"))
    ref_cam = Camera(world, os.path.join(REF, "config", "camera.yml")).camera_desc()
    own_cam = Camera(world, os.path.join(ROOT, "config", "camera.yml")).camera_desc()
    assert _struct_bytes(ref_cam) == _struct_bytes(own_cam)
    assert (ref_cam.width, ref_cam.height, ref_cam.pre_sample_times, ref_cam.max_sample_times) == (192, 108, 3, 10)
