"""Mirror of Alex::ConfigurableObject (reference src/configurable_object.rb:5-50): every top-level
YAML key becomes an attribute; every array of exactly three numerics becomes a Vec3 of floats,
recursively through hashes and arrays (:11-41).  Hash keys keep their names (Ruby symbolises them)."""
import yaml

from .vec3 import Vec3


def _is_numeric(x):
    return isinstance(x, (int, float)) and not isinstance(x, bool)


def _is_vec3_array(value):
    return isinstance(value, list) and len(value) == 3 and all(_is_numeric(x) for x in value)


class ConfigurableObject:
    def __init__(self, config_file):
        self.parse_config_file(config_file)

    def array_parse_vector(self, array):  # configurable_object.rb:11-24
        new_array = []
        for value in array:
            if isinstance(value, dict):
                new_array.append(self.hash_value_parse_vector(value))
            elif _is_vec3_array(value):
                new_array.append(Vec3.from_a(*[float(x) for x in value]))
            elif isinstance(value, list):
                new_array.append(self.array_parse_vector(value))
            # scalars inside arrays are dropped, exactly as the reference does (:13-22)
        return new_array

    def hash_value_parse_vector(self, h):  # configurable_object.rb:26-41
        new_hash = {}
        for key, value in h.items():
            if isinstance(value, dict):
                new_hash[key] = self.hash_value_parse_vector(value)
            elif _is_vec3_array(value):
                new_hash[key] = Vec3.from_a(*[float(x) for x in value])
            elif isinstance(value, list):
                new_hash[key] = self.array_parse_vector(value)
            else:
                new_hash[key] = value
        return new_hash

    def parse_config_file(self, file):  # configurable_object.rb:43-49
        if isinstance(file, dict):  # convenience for procedural scenes: an already-loaded YAML document
            config = file
            self.config_path = None
        else:
            with open(file, "r") as f:
                config = yaml.safe_load(f.read())
            self.config_path = str(file)
        config = self.hash_value_parse_vector(config)
        for key, value in config.items():
            setattr(self, key, value)
