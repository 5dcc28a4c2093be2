from dataclasses import dataclass
from .._base import BaseModule


class BaseMaterial(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        pass
