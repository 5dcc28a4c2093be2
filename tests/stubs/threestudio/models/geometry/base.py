from dataclasses import dataclass
from .._base import BaseModule


class BaseGeometry(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        pass
