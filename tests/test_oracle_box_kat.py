"""Known answers for the Box primitive (reference src/objects/box.rb:15-105), derived by hand from the
cited lines, held against the CPU oracle and against the host mirror's face construction.

Box used throughout: point (5,0,0), front (1,0,0), up (0,0,1), width_front 2, width_up 4, width_left 1.
box.rb:23 `left = front.cross(up).normalize` = (1,0,0) x (0,0,1) = (0,-1,0), so the box spans
x in [4,6], y in [-0.5,0.5], z in [-2,2] and its "left" is -y.  Faces in the order of box.rb:60-65:
0 up (z=2), 1 bottom (z=-2), 2 front (x=6), 3 back (x=4), 4 left (y=-0.5), 5 right (y=0.5)."""
import numpy as np
import pytest

from raytracing_rb_b200 import Box, Vec3, World, scenes


def _world(extra=None):
    objs = [scenes.box("kat box", [5, 0, 0], [1, 0, 0], [0, 0, 1], (2.0, 4.0, 1.0))]
    if extra:
        objs += extra
    return World({"max_distance": 10000, "soft_shadow_exponent": 2, "lights": [scenes.light([0, 0, 9], 0.0)],
                  "world_objects": objs})


@pytest.fixture(scope="module")
def sc(oracle_mod):
    return oracle_mod.OracleScene(_world().to_scene_desc())


def test_host_mirror_builds_the_six_faces_like_box_rb():
    b = _world().world_objects[0]
    assert isinstance(b, Box) and len(b.planes) == 6
    pts = [p.point.to_a() for p in b.planes]
    assert pts == [[5, 0, 2], [5, 0, -2], [6, 0, 0], [4, 0, 0], [5, -0.5, 0], [5, 0.5, 0]]
    fronts = [p.front.to_a() for p in b.planes]
    assert fronts == [[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, -1, 0], [0, 1, 0]]
    # u_unit / v_unit per face, box.rb:29,34,41,46,54,59
    assert [(p.u_unit, p.v_unit) for p in b.planes] == [(2, 1), (2, 1), (1, 4), (1, 4), (2, 4), (2, 4)]
    # Plane#reinit (plane.rb:21-23): left = front x up normalised; up face: (0,0,1) x (0,-1,0) = (1,0,0)
    assert b.planes[0].left.to_a() == [1, 0, 0]
    assert b.planes[3].left.to_a() == [0, 1, 0]
    # the faces carry the box's material but no ambient and no texture (box.rb:66-72)
    assert b.planes[2].diffuse_rate.to_a() == b.diffuse_rate.to_a() and b.planes[2].ambient is None


def test_ray_from_outside_hits_the_back_face(sc):
    hit, direction, delta = sc.intersect(0, [0, 0, 0], [1, 0, 0])
    assert hit == [4, 0, 0] and direction == "in"
    assert delta == [-1e-5, 0, 0]       # front * EPSILON * (-front.dot(ray.front) <=> 0), plane.rb:50
    assert sc.box_face(0, [0, 0, 0], [1, 0, 0]) == 3


def test_ray_from_above_hits_the_up_face(sc):
    hit, direction, delta = sc.intersect(0, [5, 0, 5], [0, 0, -1])
    assert hit == [5, 0, 2] and direction == "in" and delta == [0, 0, 1e-5]
    assert sc.box_face(0, [5, 0, 5], [0, 0, -1]) == 0


def test_ray_from_inside_leaves_through_the_right_face(sc):
    # left face (4): t = ((5,-.5,0)-(5,0,0)).(0,-1,0) / ((0,-1,0).(0,1,0)) = 0.5 / -1 < 0 -> rejected (plane.rb:46)
    hit, direction, delta = sc.intersect(0, [5, 0, 0], [0, 1, 0])
    assert hit == [5, 0.5, 0] and direction == "out"
    assert delta == [0, -1e-5, 0]       # inward: front (0,1,0) * EPSILON * sign(-1)
    assert sc.box_face(0, [5, 0, 0], [0, 1, 0]) == 5


def test_uv_bounds_are_inclusive(sc):
    # back face: left = (-1,0,0) x (0,0,1) = (0,1,0), u_unit = width_left = 1 -> u = 0.5 exactly at y = 0.5
    hit, _, _ = sc.intersect(0, [0, 0, 0], [4, 0.5, 0])
    assert hit == [4, 0.5, 0]
    assert sc.intersect(0, [0, 0, 0], [4, 0.5000001, 0]) is None


def test_ray_passing_over_the_box_misses(sc):
    # reaches x=4 at z=2.4 (outside the back face); meets z=2 at x=3.33 where u = (3.33-5)/2 = -0.83
    assert sc.intersect(0, [0, 0, 0], [1, 0, 0.6]) is None
    assert sc.box_face(0, [0, 0, 0], [1, 0, 0.6]) == -1


def test_nearest_face_wins_not_first_face(sc):
    # from far +x looking back: front face (index 2, x=6) is nearer than back face (index 3, x=4)
    hit, direction, _ = sc.intersect(0, [20, 0, 0], [-1, 0, 0])
    assert hit == [6, 0, 0] and direction == "in"
    assert sc.box_face(0, [20, 0, 0], [-1, 0, 0]) == 2


def test_cover_area_is_the_hard_shadow_factor(sc):
    # WorldObject#cover_area (world_object.rb:41-49): 1 iff the probe ray target->light hits and the hit is on
    # the target's side of the light
    assert sc.cover_area(0, [0, 0, 0], 0.8, [10, 0, 0]) == 1.0   # box between target and light
    assert sc.cover_area(0, [10, 0, 0], 0.8, [3, 0, 0]) == 1.0
    assert sc.cover_area(0, [0, 0, 0], 0.8, [3, 0, 0]) == 0.0    # box behind the target
    assert sc.cover_area(0, [3, 0, 0], 0.8, [2, 0, 0]) == 0.0    # hit lies beyond the light
    assert sc.cover_area(0, [0, 5, 0], 0.8, [10, 5, 0]) == 0.0   # probe ray passes beside the box


def test_reflection_uses_the_hit_face_normal(sc):
    p = sc.intersect_parameters(0, [0, 0, 1], [1, 0, -0.25])
    # hits the back face at x=4: n = face front (-1,0,0) (den < 0), mirror direction flips x
    assert p["n"] == [-1, 0, 0]
    d = np.array(p["reflection"][0])
    want = np.array([-1, 0, -0.25]) / np.sqrt(1 + 0.0625)
    assert np.allclose(d, want, atol=1e-15)
    assert p["refraction"] is not None   # the kat box has a refractive_rate


def test_world_intersect_orders_box_against_sphere(oracle_mod):
    w = _world([scenes.matte("s", (2.5, 0, 0), 0.5, (1, 1, 1))])
    sc2 = oracle_mod.OracleScene(w.to_scene_desc())
    assert sc2.world_intersect([0, 0, 0], [1, 0, 0])[0] == 1      # sphere at x=2 is nearer than the box at x=4
    assert sc2.world_intersect([0, 0, 1.5], [1, 0, 0])[0] == 0     # above the sphere only the box is hit


def test_zero_cross_product_is_rejected_like_the_reference():
    with pytest.raises(Exception):
        World({"max_distance": 10, "soft_shadow_exponent": 2, "lights": [],
               "world_objects": [scenes.box("bad", [0, 0, 0], [1, 0, 0], [2, 0, 0], (1, 1, 1))]})


def test_box_frame_counters_and_hits(oracle_mod):
    """A small frame of config 6: boxes are seen (primary hits on objects 3 and 4) and every World#intersect
    tests every box once: box_tests = rays that reached the scan x 2 boxes."""
    from raytracing_rb_b200 import Camera, make_opts
    wdoc, cdoc = scenes.build(6, width=96, height=54)
    world = World(wdoc)
    cam = Camera(world, cdoc)
    f = oracle_mod.OracleScene(world.to_scene_desc()).render(cam.camera_desc(), make_opts(seed=1))
    assert set(np.unique(f.hit)) >= {3, 4}
    scans = f.stats["rays"] - f.stats["highlight_hits"]
    assert f.stats["box_tests"] == 2 * scans
    assert f.stats["cover_box"] == 2 * f.stats["shadow_queries"]
    assert 0 < f.stats["box_accepts"] < f.stats["box_tests"]
    assert f.stats["status"] == 0
