from dataclasses import dataclass, fields

import torch.nn as nn


class BaseModule(nn.Module):
    @dataclass
    class Config:
        pass

    def __init__(self, cfg=None, *args, **kwargs):
        super().__init__()
        cfg = dict(cfg or {})
        names = {f.name for f in fields(self.Config)}
        self.cfg = self.Config(**{k: v for k, v in cfg.items() if k in names})
        self.configure(*args, **kwargs)

    def configure(self, *args, **kwargs):
        pass
