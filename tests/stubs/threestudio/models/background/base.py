from dataclasses import dataclass
from .._base import BaseModule


class BaseBackground(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        pass
