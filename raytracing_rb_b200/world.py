"""Mirror of Alex::World (reference src/world.rb:10-34): YAML -> objects + lights.  The scene
queries World#intersect / lit_area / local_lights / high_lights (:37-98) are device code; this
class only builds the flat scene description handed to rtrb_renderer_create."""
import ctypes as C

from . import _abi
from .configurable_object import ConfigurableObject
from .lights import LIGHT_CLASSES
from .objects import OBJECT_CLASSES


class SceneDescHolder:
    """A SceneDesc plus the Python objects that own the memory it points to."""

    def __init__(self, desc, keepalive):
        self.desc = desc
        self._keepalive = keepalive


class World(ConfigurableObject):
    max_distance = None
    trace_depth = None          # world.rb:12, unused upstream (depth comes from the camera)
    soft_shadow_exponent = None

    def __init__(self, config_file):
        super().__init__(config_file)
        self.world_objects = self.parse_objects(self.world_objects)  # world.rb:17
        self.lights = self.parse_lights(self.lights)                  # world.rb:18

    def parse_lights(self, array):  # world.rb:21-27
        ret = []
        for item in array:
            klass = LIGHT_CLASSES.get(str(item["type"]))
            if klass is None:
                raise NameError("uninitialized constant Alex::Lights::%sLight" % item["type"])
            ret.append(klass(item["properties"]))
        return ret

    def parse_objects(self, array):  # world.rb:28-34
        ret = []
        for item in array:
            klass = OBJECT_CLASSES.get(str(item["type"]))
            if klass is None:
                raise NameError("uninitialized constant Alex::Objects::%s" % item["type"])
            ret.append(klass(item["properties"], config_path=self.config_path))
        return ret

    def to_scene_desc(self):
        textures, tex_index, keep = [], {}, []
        objs = (_abi.ObjectDesc * max(1, len(self.world_objects)))()
        for i, o in enumerate(self.world_objects):
            ti = -1
            if o.texture is not None and getattr(o, "TYPE", None) != _abi.OBJ_BOX:
                key = o.texture.file_name
                if key not in tex_index:
                    tex_index[key] = len(textures)
                    textures.append(o.texture)
                ti = tex_index[key]
            objs[i] = o.to_desc(ti)
        lights = (_abi.LightDesc * max(1, len(self.lights)))()
        for i, l in enumerate(self.lights):
            d = _abi.LightDesc()
            d.position = (_abi.D3)(*l.position.to_a())
            d.color = (_abi.D3)(*l.color.to_a())
            d.radius = float(l.radius)
            d.high_light_rate = float(l.high_light_rate)
            d.high_light_angle = float(l.high_light_angle)
            lights[i] = d
        texs = (_abi.TextureDesc * max(1, len(textures)))()
        for i, t in enumerate(textures):
            texs[i].width, texs[i].height = t.width, t.height
            texs[i].rgb8 = t.rgb8.ctypes.data_as(C.POINTER(C.c_uint8))
            keep.append(t.rgb8)
        sd = _abi.SceneDesc()
        sd.max_distance = float(self.max_distance)
        sd.soft_shadow_exponent = float(self.soft_shadow_exponent)
        sd.n_objects, sd.n_lights, sd.n_textures = len(self.world_objects), len(self.lights), len(textures)
        sd.objects = C.cast(objs, C.POINTER(_abi.ObjectDesc))
        sd.lights = C.cast(lights, C.POINTER(_abi.LightDesc))
        sd.textures = C.cast(texs, C.POINTER(_abi.TextureDesc))
        keep += [objs, lights, texs, textures]
        return SceneDescHolder(sd, keep)
