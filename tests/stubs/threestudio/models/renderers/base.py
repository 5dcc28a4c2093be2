from dataclasses import dataclass
from .._base import BaseModule


class Renderer(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        radius: float = 1.0

    def configure(self, geometry=None, material=None, background=None):
        # plain attributes (not registered sub-modules), as threestudio's BaseModule does for renderers
        object.__setattr__(self, "geometry", geometry)
        object.__setattr__(self, "material", material)
        object.__setattr__(self, "background", background)


class Rasterizer(Renderer):
    pass
