"""restate_py.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A SECOND, independent restatement of raytracing_rb's per-pixel hot path: pure Python, class for class
and method for method after the Ruby sources (paths below are relative to /root/reference), written
from the Ruby text alone and sharing no code with oracle/rtrb_oracle.cpp.  Its only job is to pin the
C++ oracle a little harder while no Ruby interpreter exists in this image: tests/test_restatements_agree.py
renders small frames with both and requires bit-identical float colours, hit ids and counters, in the
reference's own MT19937 consumption order (RNG `mt`) as well as with the counter RNG.

Python floats are IEEE doubles and math.sqrt/sin/cos/acos/asin are this image's libm, i.e. exactly what
the C++ oracle calls, so agreement is expected to the last bit.  `x ** 2` is written x * x for the
reason given in rtrb_oracle.cpp (correctly rounded pow on the reference's platform).

Only small frames: this runs ~1e4 rays per second.
"""
import math

import numpy as np

EPSILON = 1e-5  # src/libs/algebra.rb:2


class Raised(Exception):
    """The reference's raise sites (ray_tracer.rb:295, fast_4d_matrix.c:124,291, Math::DomainError)."""


# ---- ext/fast_4d_matrix/fast_4d_matrix.c:57-305 ----------------------------------------------------
class Vec3:
    __slots__ = ("x", "y", "z", "r")

    def __init__(self, x, y, z):  # Vec3_c_create :62-73
        self.x, self.y, self.z = x, y, z
        self.r = math.sqrt(x * x + y * y + z * z)

    def to_a(self):
        return [self.x, self.y, self.z]

    def dot(self, o):  # :98-108
        ret = 0.0
        ret += self.x * o.x
        ret += self.y * o.y
        ret += self.z * o.z
        return ret

    def cos(self, o):  # :109-129 — |cos|, clamped to <= 1
        ret = self.dot(o)
        r1 = self.x * self.x + self.y * self.y + self.z * self.z
        r2 = o.x * o.x + o.y * o.y + o.z * o.z
        if r1 == 0 or r2 == 0:
            raise Raised("zero vector detected!")
        v = math.sqrt(ret * ret / r1 / r2)
        return 1.0 if v > 1 else v

    def cross(self, o):  # :131-141
        return Vec3(self.y * o.z - self.z * o.y, self.z * o.x - self.x * o.z, self.x * o.y - self.y * o.x)

    def __neg__(self):
        return Vec3(-self.x, -self.y, -self.z)

    def __add__(self, o):
        return Vec3(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return Vec3(self.x - o.x, self.y - o.y, self.z - o.z)

    def __mul__(self, o):  # :190-208 — element-wise for Vec3, scalar for Float
        if isinstance(o, Vec3):
            return Vec3(self.x * o.x, self.y * o.y, self.z * o.z)
        return Vec3(self.x * o, self.y * o, self.z * o)

    def __truediv__(self, s):  # :209-224 (C double division: x / 0.0 is +-Infinity or NaN, never an error)
        return Vec3(_fdiv(self.x, s), _fdiv(self.y, s), _fdiv(self.z, s))

    def r2(self):  # :280-284
        return self.r * self.r

    def normalize(self):  # :286-293
        r = math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)
        if r == 0:
            raise Raised("zero vector detected")
        return Vec3(self.x / r, self.y / r, self.z / r)


def _fdiv(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a == 0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


def v3(a):
    return Vec3(float(a[0]), float(a[1]), float(a[2]))


def rb_sqrt(x):
    if x < 0:
        raise Raised("Math::DomainError sqrt")
    return math.sqrt(x)


def rb_acos(x):
    if x < -1 or x > 1:
        raise Raised("Math::DomainError acos")
    return math.acos(x)


def rb_asin(x):
    if x < -1 or x > 1:
        raise Raised("Math::DomainError asin")
    return math.asin(x)


def rb_pow(x, y):
    return x * x if y == 2 else math.pow(x, y)


class Ray:  # src/libs/algebra.rb:3-17
    def __init__(self, f, pos):
        self.front, self.position = f, pos

    def distance(self, pos):
        return (self.position - pos).r


# ---- src/objects/texture.rb ---------------------------------------------------------------------------
class Texture:
    def __init__(self, rgb8, width, height, hscale, vscale, u_off, v_off):
        self.data, self.width, self.height = rgb8, width, height  # rows of (v8/256.0) triples, :19
        self.horizontal_scale, self.vertical_scale = hscale, vscale
        self.u_off, self.v_off = u_off or 0.0, v_off or 0.0

    def color(self, uu, vv):  # :23-28 — Float#to_i truncates, Integer#% is floored
        fu = (uu + self.u_off) / self.horizontal_scale
        fv = (vv + self.v_off) / self.vertical_scale
        if not (math.isfinite(fu) and math.isfinite(fv)):
            raise Raised("FloatDomainError")
        u = int(fu) % self.width
        v = int(fv) % self.height
        p = self.data[v, u]
        return Vec3(int(p[0]) / 256.0, int(p[1]) / 256.0, int(p[2]) / 256.0)


# ---- src/objects/world_object.rb ------------------------------------------------------------------------
class WorldObject:
    texture = None
    refractive_rate = None

    def reflect_refract_vector(self):  # :33-39
        return self.reflective_attenuation, self.refractive_attenuation

    def cover_area(self, light_position, light_radius, target_position):  # :41-49
        ray = Ray(light_position - target_position, target_position)
        res = self.intersect(ray)
        intersection = res[0] if res else None
        if intersection is not None and (intersection - light_position).dot(target_position - light_position) > 0:
            return 1
        return 0

    def local_lighting(self, position, lights, normal_vector, ray, color_filter=None):  # :51-74
        light_contribution = Vec3(0.0, 0.0, 0.0)
        for light, light_color in lights:
            n = normal_vector.normalize()
            l = (light.position - position).normalize()
            l_dot_n = l.dot(n)
            if l_dot_n > 1:
                l_dot_n = 1.0
            elif l_dot_n < 0:
                l_dot_n = 0.0
            light_contribution = light_contribution + light_color * l_dot_n
        if len(lights) > 0:
            light_contribution = light_contribution / float(len(lights))
        if color_filter is not None:
            return light_contribution * self.diffuse_rate * color_filter + self.ambient
        return light_contribution * self.diffuse_rate + self.ambient

    def path_tracing(self, intersection, n, pt_times, rng):  # :76-90
        ret = []
        att = self.diffuse_rate / float(pt_times)
        for m in range(pt_times):
            front = n.normalize()
            left = self.get_a_random_vertical_vector(n).normalize()
            up = front.cross(left)
            u_theta, u_phi = rng.mc(m)
            theta, phi = u_theta * math.pi / 2, u_phi * math.pi * 2
            direction = front * math.sin(theta) + (left * math.cos(phi) + up * math.sin(phi)) * math.cos(theta)
            ret.append((Ray(direction, intersection), att, m))
        return ret

    @staticmethod
    def get_a_random_vertical_vector(n):  # :105-120
        if n.r == 0:
            raise Raised("zero vector detected")
        if n.x == 0:
            if n.y == 0:
                return Vec3(1.0, 0.0, 0.0)
            return Vec3(0.0, -n.z / n.y, 1.0)
        return Vec3(-(n.y + n.z) / n.x, 1.0, 1.0)

    @staticmethod
    def get_reflection_by_ray_and_n(ray, n, intersection, delta):  # :121-125
        cos_theta = ray.front.cos(-n)
        front = (n.normalize() * (2 * cos_theta * ray.front.r) + ray.front).normalize()
        return Ray(front, intersection + delta)

    @staticmethod
    def get_refraction_by_ray_and_n(ray, n, intersection, reflection, refraction_rate, delta):  # :127-137
        c = ray.front.cos(n)
        sin_i = rb_sqrt(1 - rb_pow(c, 2))
        sin_r = sin_i / refraction_rate
        if sin_r >= 1:
            return None
        r = rb_asin(sin_r)
        refraction_direction = n.normalize() * (-math.cos(r)) + (reflection + ray.front).normalize() * sin_r
        return Ray(refraction_direction, intersection - n.normalize() * EPSILON)


class Sphere(WorldObject):  # src/objects/sphere.rb
    def inner(self, position):  # :103-105
        return (position - self.center).r <= self.radius

    def cover_area(self, light_position, light_radius, target_position):  # :28-57
        factor = WorldObject.cover_area(self, light_position, light_radius, target_position)
        t = (self.center - target_position).dot(light_position - target_position) / (light_position - target_position).r2()
        x1 = target_position + (light_position - target_position) * t
        r1 = light_radius * ((x1 - target_position).r / (light_position - target_position).r)
        d = (x1 - self.center).r
        if d >= r1 + self.radius:
            return 0
        s1 = math.pi * r1 * r1
        if d > abs(self.radius - r1):
            cos_theta1 = min((r1 * r1 + d * d - self.radius * self.radius) / (2 * r1 * d), 1.0)
            cos_theta2 = min((self.radius * self.radius + d * d - r1 * r1) / (2 * self.radius * d), 1.0)
            theta1 = rb_acos(cos_theta1)
            theta2 = rb_acos(cos_theta2)
            delta_s = ((theta1 - math.sin(theta1)) * r1 * r1 + (theta2 - math.sin(theta2)) * self.radius * self.radius) / 2
            return factor * delta_s / s1
        if r1 > self.radius:
            return factor * math.pi * self.radius * self.radius / s1
        return factor

    def intersect(self, ray):  # :60-85
        t = (self.center - ray.position).dot(ray.front) / ray.front.r2()
        v = ray.front * t
        nearest_point = ray.position + v
        if not self.inner(nearest_point):
            return None
        nearest_dis = (nearest_point - self.center).r
        nearest_point_to_intersection = rb_sqrt(rb_pow(self.radius, 2) - rb_pow(nearest_dis, 2))
        vec = ray.front.normalize() * nearest_point_to_intersection
        from_inner = self.inner(ray.position)
        direction = "out" if from_inner else "in"
        intersection = nearest_point - vec if direction == "in" else nearest_point + vec
        if not from_inner and t < 0:
            return None
        return intersection, direction, (intersection - self.center) * EPSILON * (1.0 if direction == "in" else -1.0), None

    def intersect_parameters(self, ray, intersection, direction, delta, data=None):  # :88-101
        n = (intersection - self.center) if direction == "in" else (self.center - intersection)
        reflection = self.get_reflection_by_ray_and_n(ray, n, intersection, delta)
        refraction = self.get_refraction_by_ray_and_n(
            ray, n, intersection, reflection.front,
            self.refractive_rate if direction == "in" else 1.0 / self.refractive_rate, delta)
        return n, reflection, refraction

    def get_uv(self, position):  # :111-120
        vec = position - self.center
        x = vec.dot(self.greenwich_vec.normalize()) / self.radius
        y = vec.dot(self.ninety_degree_east_vec.normalize()) / self.radius
        z = vec.dot(self.north_pole_vec.normalize()) / self.radius
        m = rb_sqrt(x * x + y * y + z * z + 2 * x + 1)
        return (y / m + 1) / 2, (-z / m + 1) / 2

    def local_lighting(self, position, lights, normal_vector, ray, counters=None):  # :122-129
        color_filter = Vec3(1.0, 1.0, 1.0)
        if self.texture:
            u, v = self.get_uv(position)
            if counters is not None:
                counters["texel_fetches"] += 1
            return WorldObject.local_lighting(self, position, lights, normal_vector, ray, self.texture.color(u, v) * color_filter)
        return WorldObject.local_lighting(self, position, lights, normal_vector, ray, color_filter)


class Plane(WorldObject):  # src/objects/plane.rb
    u_unit = v_unit = None
    diffuse_rate = ambient = reflective_attenuation = refractive_attenuation = None

    def reinit(self):  # :21-23
        self.left = self.front.cross(self.up).normalize()

    def intersect(self, ray):  # :38-51
        denominator = self.front.dot(ray.front)
        if denominator == 0:
            return None
        t = (self.point - ray.position).dot(self.front) / denominator
        intersection = ray.position + ray.front * t
        if t < 0:
            return None
        direction = "in" if self.front.dot(ray.front) < 0 else "out"
        x = -self.front.dot(ray.front)
        sign = 1.0 if x > 0 else (-1.0 if x < 0 else 0.0)  # (x <=> 0).to_f
        return intersection, direction, self.front * EPSILON * sign, None

    def intersect_parameters(self, ray, intersection, direction, delta, data=None):  # :54-67
        n = -self.front if self.front.dot(ray.front) > 0 else self.front
        reflection = self.get_reflection_by_ray_and_n(ray, n, intersection, delta)
        refraction = None
        if self.refractive_rate is not None:
            refraction = self.get_refraction_by_ray_and_n(ray, n, intersection, reflection.front, self.refractive_rate, delta)
        return n, reflection, refraction

    def get_uv(self, position):  # :81-85
        u = (position - self.point).dot(self.left.normalize()) / self.u_unit
        v = (position - self.point).dot(self.up.normalize()) / self.v_unit
        return u, v

    def local_lighting(self, position, lights, normal_vector, ray, counters=None):  # :87-94
        light_filter = Vec3(1.0, 1.0, 1.0)
        if self.texture:
            u, v = self.get_uv(position)
            if counters is not None:
                counters["texel_fetches"] += 1
            return WorldObject.local_lighting(self, position, lights, normal_vector, ray, self.texture.color(u, v) * light_filter)
        return WorldObject.local_lighting(self, position, lights, normal_vector, ray, light_filter)


class Box(WorldObject):  # src/objects/box.rb
    def init_planes(self):  # :22-73
        self.planes = []
        left = self.front.cross(self.up).normalize()

        def face(front, up, point, u_unit, v_unit):
            p = Plane()
            p.front, p.up, p.point, p.u_unit, p.v_unit = front, up, point, u_unit, v_unit
            return p
        up_plane = face(self.up, left, self.point + self.up * self.width_up * 0.5, self.width_front, self.width_left)
        bottom_plane = face(-self.up, left, self.point - self.up * self.width_up * 0.5, self.width_front, self.width_left)
        front_plane = face(self.front, self.up, self.point + self.front * self.width_front * 0.5, self.width_left, self.width_up)
        back_plane = face(-self.front, self.up, self.point - self.front * self.width_front * 0.5, self.width_left, self.width_up)
        left = self.front.cross(self.up).normalize()
        left_plane = face(left, self.up, self.point + left * self.width_left * 0.5, self.width_front, self.width_up)
        right_plane = face(-left, self.up, self.point - left * self.width_left * 0.5, self.width_front, self.width_up)
        self.planes = [up_plane, bottom_plane, front_plane, back_plane, left_plane, right_plane]
        for p in self.planes:
            p.reflective_attenuation = self.reflective_attenuation
            p.refractive_attenuation = self.refractive_attenuation
            p.refractive_rate = self.refractive_rate
            p.diffuse_rate = self.diffuse_rate
            p.reinit()

    def intersect(self, ray):  # :80-99
        nearest_dis = math.inf
        nearest_ret = None
        for index, plane in enumerate(self.planes):
            res = plane.intersect(ray)
            if res:
                intersection, direction, delta, _ = res
                u, v = plane.get_uv(intersection)
                if -0.5 <= u <= 0.5 and -0.5 <= v <= 0.5:
                    d = (intersection - ray.position).r
                    if d < nearest_dis:
                        nearest_dis = d
                        nearest_ret = (intersection, direction, delta, index)
        return nearest_ret

    def intersect_parameters(self, ray, intersection, direction, delta, data=None):  # :102-107
        return self.planes[data].intersect_parameters(ray, intersection, direction, delta)

    def local_lighting(self, position, lights, normal_vector, ray, counters=None):
        return WorldObject.local_lighting(self, position, lights, normal_vector, ray)  # color_filter = nil


class Light:  # src/lights/light.rb, spot_light.rb
    pass


# ---- src/world.rb ------------------------------------------------------------------------------------
class World:
    def __init__(self, world):
        """`world` is the host mirror raytracing_rb_b200.World (YAML already parsed by ConfigurableObject)."""
        from raytracing_rb_b200 import objects as O
        self.max_distance = float(world.max_distance)
        self.soft_shadow_exponent = world.soft_shadow_exponent
        self.world_objects, self.lights = [], []
        for i, o in enumerate(world.world_objects):
            if isinstance(o, O.Sphere):
                w = Sphere()
                w.center, w.radius = v3(o.center.to_a()), float(o.radius)
                w.refractive_rate = float(o.refractive_rate)
                if o.texture is not None:
                    w.greenwich_vec, w.north_pole_vec = v3(o.greenwich_vec.to_a()), v3(o.north_pole_vec.to_a())
                    w.ninety_degree_east_vec = w.north_pole_vec.cross(w.greenwich_vec)  # sphere.rb:19
            elif isinstance(o, O.Box):
                w = Box()
                w.point, w.front, w.up = v3(o.point.to_a()), v3(o.front.to_a()), v3(o.up.to_a())
                w.width_front, w.width_up, w.width_left = float(o.width_front), float(o.width_up), float(o.width_left)
                w.refractive_rate = None if o.refractive_rate is None else float(o.refractive_rate)
            else:
                w = Plane()
                w.point, w.front, w.up = v3(o.point.to_a()), v3(o.front.to_a()), v3(o.up.to_a())
                w.u_unit = None if o.u_unit is None else float(o.u_unit)
                w.v_unit = None if o.v_unit is None else float(o.v_unit)
                w.refractive_rate = None if o.refractive_rate is None else float(o.refractive_rate)
                w.reinit()
            w.index = i
            w.diffuse_rate = v3(o.diffuse_rate.to_a())
            w.reflective_attenuation = v3(o.reflective_attenuation.to_a())
            w.refractive_attenuation = None if o.refractive_attenuation is None else v3(o.refractive_attenuation.to_a())
            w.ambient = v3(o.ambient.to_a())
            if isinstance(o, O.Box):
                w.init_planes()
            elif o.texture is not None:
                t = o.texture
                w.texture = Texture(t.rgb8, t.width, t.height, float(t.horizontal_scale), float(t.vertical_scale), float(t.u_off), float(t.v_off))
            self.world_objects.append(w)
        for l in world.lights:
            x = Light()
            x.position, x.color = v3(l.position.to_a()), v3(l.color.to_a())
            x.radius, x.high_light_rate, x.high_light_angle = float(l.radius), l.high_light_rate, l.high_light_angle
            self.lights.append(x)

    def intersect(self, ray):  # :37-59
        nearest_obj = None
        nearest_dis = self.max_distance
        nearest = (None, None, None, None)
        for obj in self.world_objects:
            res = obj.intersect(ray)
            if res:
                intersection, direction, delta, data = res
                new_dis = ray.distance(intersection)
                if new_dis < nearest_dis:
                    nearest_dis = new_dis
                    nearest_obj = obj
                    nearest = (intersection, direction, delta, data)
        return (nearest_obj,) + nearest

    def lit_area(self, target, light_pos, radius):  # :62-69
        total_area = 1
        for obj in self.world_objects:
            total_area -= obj.cover_area(light_pos, radius, target)
        return max(total_area, 0)

    def local_lights(self, position, counters):  # :72-80
        ret = []
        for light in self.lights:
            counters["shadow_queries"] += 1
            area = self.lit_area(position, light.position, light.radius)
            if area > 0:
                ret.append((light, light.color * (rb_pow(float(area), self.soft_shadow_exponent) / len(self.lights))))
        return ret

    def high_lights(self, ray):  # :83-98 (the `&& lit_area(...)` term is a number: always truthy)
        ret = []
        for light in self.lights:
            a = light.position - ray.position
            cos_theta = ray.front.cos(a)
            cos_theta = max(-1, min(1, cos_theta))
            ang = rb_acos(cos_theta)
            if ang < (light.high_light_angle / 180.0 * math.pi):
                ret.append((light, light.color * float(light.high_light_rate)))
        return ret


# ---- RNG streams -----------------------------------------------------------------------------------------
def _philox(k0, k1, c):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = c
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _res53(a, b):  # genrand_res53
    return ((a >> 5) * 67108864.0 + (b >> 6)) * (1.0 / 9007199254740992.0)


class Rng:
    """mode 'mt': Ruby's global MT19937 (Random.srand(seed); genrand_res53 per Random.rand) consumed in
    program order.  mode 'ctr': the device's counter RNG, keyed by (pixel, sample, ray path)."""

    def __init__(self, mode, seed):
        self.mode, self.seed = mode, seed
        self.mt = np.random.RandomState(seed) if mode == "mt" else None
        self.pixel = self.sample = 0
        self.path, self.K = 1, 2

    def lens(self):  # camera.rb:135
        if self.mode == "mt":
            return float(self.mt.random_sample())
        c = _philox(self.seed & 0xFFFFFFFF, self.seed >> 32, (self.pixel, self.sample, 0, 0))
        return _res53(c[0], c[1])

    def mc(self, m):  # world_object.rb:84, theta's draw first
        if self.mode == "mt":
            return float(self.mt.random_sample()), float(self.mt.random_sample())
        child = (self.path * self.K + 2 + m) & 0xFFFFFFFF
        c = _philox(self.seed & 0xFFFFFFFF, self.seed >> 32, (self.pixel, self.sample, child, 1))
        return _res53(c[0], c[1]), _res53(c[2], c[3])


# ---- src/ray_tracer.rb ------------------------------------------------------------------------------------
class RayTracer:
    def __init__(self, world, trace_depth, monte_carlo_diffusion_times, rng, counters):
        self.world, self.trace_depth, self.mc = world, trace_depth, monte_carlo_diffusion_times
        self.rng, self.counters = rng, counters
        rng.K = monte_carlo_diffusion_times + 2

    def trace_sync(self, ray, hit_out=None):  # :16-46
        self.counters["samples"] += 1
        queue = [dict(ray=ray, trace_depth=self.trace_depth, attenuation=Vec3(1.0, 1.0, 1.0), path=1)]
        light_queue = []
        first = True
        while queue:
            self.counters["max_stack"] = max(self.counters["max_stack"], len(queue))
            item = queue.pop()
            rays, lights = self.rt_map(item, hit_out if first else None)
            first = False
            light_queue.extend(lights)   # Queue#<< ... Queue#pop is FIFO
            queue.extend(rays)
        s = Vec3(0.0, 0.0, 0.0)
        for color in light_queue:
            s = s + color                # mix_color :300-303
            if not (s.x <= 1 and s.y <= 1 and s.z <= 1):
                raise Raised("color greater than 1")
        return s

    def rt_map(self, rt_ray, hit_out):  # :50-164
        rays, colors = [], []
        cnt = self.counters
        if rt_ray["trace_depth"] <= 0 or rt_ray["attenuation"].r < 0.0001:
            return rays, colors
        cnt["rays"] += 1
        K = self.mc + 2
        lights = self.world.high_lights(rt_ray["ray"])
        for light, color in lights:
            colors.append(rt_ray["attenuation"] * color / float(len(lights)))
        if colors:
            cnt["highlight_hits"] += 1
            if hit_out is not None:
                hit_out[0] = -2
            return rays, colors
        obj, intersection, direction, delta, data = self.world.intersect(rt_ray["ray"])
        if obj is None:
            if hit_out is not None:
                hit_out[0] = -1
            return rays, colors
        if hit_out is not None:
            hit_out[0] = obj.index
        cnt["hits"] += 1
        n, reflection_ray, refraction_ray = obj.intersect_parameters(rt_ray["ray"], intersection, direction, delta, data)
        att_reflect, att_refract = obj.reflect_refract_vector()
        if reflection_ray:
            rays.append(dict(ray=reflection_ray, trace_depth=rt_ray["trace_depth"] - 1,
                             attenuation=rt_ray["attenuation"] * att_reflect, path=(rt_ray["path"] * K) & 0xFFFFFFFF))
        if refraction_ray:
            cnt["refractions"] += 1
            rays.append(dict(ray=refraction_ray, trace_depth=rt_ray["trace_depth"] - 1,
                             attenuation=rt_ray["attenuation"] * att_refract, path=(rt_ray["path"] * K + 1) & 0xFFFFFFFF))
        lit = self.world.local_lights(intersection + delta, cnt)
        if len(lit) == 0:
            self.rng.path = rt_ray["path"]
            for pt_ray, pt_att, m in obj.path_tracing(intersection + delta, n, self.mc, self.rng):
                cnt["mc_rays"] += 1
                rays.append(dict(ray=pt_ray, trace_depth=rt_ray["trace_depth"] - 1,
                                 attenuation=rt_ray["attenuation"] * pt_att, path=(rt_ray["path"] * K + 2 + m) & 0xFFFFFFFF))
        else:
            cnt["local_shaded"] += 1
            cnt["lit_lights"] += len(lit)
            colors.append(rt_ray["attenuation"] * obj.local_lighting(intersection, lit, n, rt_ray["ray"], cnt))
        return rays, colors


# ---- src/camera.rb ---------------------------------------------------------------------------------------
class Camera:
    def __init__(self, cam):
        """`cam` is the host mirror raytracing_rb_b200.Camera."""
        self.position, self.up, self.front = v3(cam.position.to_a()), v3(cam.up.to_a()), v3(cam.front.to_a())
        for k in ("retina_width", "retina_height", "aperture_radius", "image_distance", "focal_distance", "variant_threshold"):
            setattr(self, k, float(getattr(cam, k)))
        for k in ("width", "height", "pre_sample_times", "max_sample_times", "trace_depth", "monte_carlo_diffusion_times"):
            setattr(self, k, int(getattr(cam, k)))

    @staticmethod
    def intersect_plane(ray, point, front):  # :123-127
        t = float((point - ray.position).dot(front)) / (front.dot(ray.front))
        return ray.position + ray.front * t

    def lens_func(self, x, y, rng):  # :129-151
        left = self.up.cross(self.front).normalize()
        retina_center = self.position - self.front.normalize() * self.image_distance
        retina_position = retina_center + \
            left * (2.0 * (float(x) / self.width - 0.5) * self.retina_width) + \
            self.up.normalize() * (2 * (float(y) / self.height - 0.5) * self.retina_height)
        theta = rng.lens()
        rand_vector = (left.normalize() * math.cos(theta) + self.up.normalize() * math.sin(theta)) * self.aperture_radius
        aperture_position = self.position + rand_vector
        object_distance = self.focal_distance * self.image_distance / (self.image_distance - self.focal_distance)
        point_on_focal_plane = self.position + self.front.normalize() * object_distance
        normal_vector_focal_plane = self.front
        r = Ray(self.position - retina_position, retina_position)
        target_point = self.intersect_plane(r, point_on_focal_plane, normal_vector_focal_plane)
        return Ray(target_point - aperture_position, aperture_position)

    def render_at(self, x, y, tracer, rng, counters, hit_out):  # :70-99
        rng.pixel = y * self.width + x
        pre_samples = []
        average = Vec3(0.0, 0.0, 0.0)
        for j in range(self.pre_sample_times):
            rng.sample = j
            ray = self.lens_func(x, y, rng)
            v = tracer.trace_sync(ray, hit_out if j == 0 else None)
            pre_samples.append(v)
            average = average + v
        variance = 0
        average = average / float(self.pre_sample_times)
        for j in range(self.pre_sample_times):
            m = max((pre_samples[j] - average).to_a())
            variance += m * m
        variance /= self.pre_sample_times
        if variance >= self.variant_threshold:
            counters["adaptive_pixels"] += 1
            color_vec = Vec3(0.0, 0.0, 0.0)
            for j in range(self.pre_sample_times, self.max_sample_times):
                rng.sample = j
                ray = self.lens_func(x, y, rng)
                color_vec = color_vec + tracer.trace_sync(ray)
            average = (average * float(self.pre_sample_times) + color_vec) / float(self.max_sample_times)
        return average.to_a()


COUNTERS = ("samples", "rays", "shadow_queries", "highlight_hits", "hits", "local_shaded", "lit_lights", "mc_rays",
            "refractions", "texel_fetches", "adaptive_pixels", "max_stack")


def render(world_mirror, camera_mirror, rng_mode="ctr", seed=1, window=None):
    """Camera#render_sync's loop (camera.rb:101-110: x outer, y inner) over `window` = (x0, y0, x1, y1).
    Returns (rgb float64 [H,W,3], hit int32 [H,W], counters dict)."""
    world, cam = World(world_mirror), Camera(camera_mirror)
    W, H = cam.width, cam.height
    x0, y0, x1, y1 = window or (0, 0, W, H)
    counters = {k: 0 for k in COUNTERS}
    rng = Rng(rng_mode, seed)
    tracer = RayTracer(world, cam.trace_depth, cam.monte_carlo_diffusion_times, rng, counters)
    rgb = np.zeros((H, W, 3), np.float64)
    hit = np.full((H, W), -3, np.int32)
    for x in range(x0, x1):
        for y in range(y0, y1):
            h = [-1]
            rgb[y, x] = cam.render_at(x, y, tracer, rng, counters, h)
            hit[y, x] = h[0]
    return rgb, hit, counters
