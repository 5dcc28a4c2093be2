"""CPU oracle of the optimizer step on the Gaussian parameters.

TEST INFRASTRUCTURE ONLY -- the product path is csrc/adam.cu behind b200splat_adam_step.

Restates, with PyTorch's own autograd and ``torch.optim.Adam``, what the reference does between ``loss.backward()``
and ``opt.step()``: raw parameters -> activations (geometry/gaussian_base.py:240-248 ``exp`` / ``sigmoid`` /
``F.normalize``; :371-400 getters incl. ``features_dc.clip(-color_clip, color_clip)`` and
``cat(features_dc, features_rest)``) -> gradients of the activated values pulled back by autograd ->
``torch.optim.Adam(groups, lr=0.0, eps=1e-15)`` with the six named groups of ``training_setup`` (:470-525).
Pinned against the reference's ``GaussianBaseModel`` itself in tests/test_optim_cpu.py (where /root/reference exists).
"""
from __future__ import annotations

from typing import Dict

import torch

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


class OracleGaussianAdam:
    def __init__(self, params: Dict[str, torch.Tensor], lrs: Dict[str, float], eps: float = 1e-15,
                 color_clip: float = float("inf")):
        self.p = {k: torch.nn.Parameter(params[k].detach().clone().float().cpu()) for k in GROUPS}
        self.color_clip = color_clip
        self.optimizer = torch.optim.Adam([{"params": [self.p[k]], "lr": float(lrs[k]), "name": k} for k in GROUPS],
                                          lr=0.0, eps=eps)

    def set_lrs(self, lrs: Dict[str, float]) -> None:
        for grp in self.optimizer.param_groups:
            grp["lr"] = float(lrs[grp["name"]])

    def activated(self):
        p = self.p
        feats = torch.cat((p["f_dc"].clip(-self.color_clip, self.color_clip), p["f_rest"]), dim=1)
        return dict(means3D=p["xyz"], shs=feats, opacities=torch.sigmoid(p["opacity"]),
                    scales=torch.exp(p["scaling"]), rotations=torch.nn.functional.normalize(p["rotation"]))

    def step(self, grads: Dict[str, torch.Tensor]) -> None:
        """grads: gradients with respect to the activated values (means3D, shs, opacities, scales, rotations)."""
        self.optimizer.zero_grad(set_to_none=True)
        act = self.activated()
        keys = ("means3D", "shs", "opacities", "scales", "rotations")
        torch.autograd.backward([act[k] for k in keys], [grads[k].detach().float().cpu() for k in keys])
        self.optimizer.step()

    def state(self, group: str):
        st = self.optimizer.state[self.p[group]]
        return st["exp_avg"], st["exp_avg_sq"]
