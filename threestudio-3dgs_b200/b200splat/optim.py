"""Fused optimizer step for the Gaussian parameters: activation backward + Adam in one pass over the gradient buffer
the rasterizer's backward (and the multi-GPU all-reduce) left behind (b200splat_adam_step, csrc/adam.cu).

Mirrors what the reference does with ``torch.optim.Adam(l, lr=0.0, eps=1e-15)`` over the six parameter groups
``xyz, f_dc, f_rest, opacity, scaling, rotation`` (geometry/gaussian_base.py:470-525) after autograd has walked
``exp`` / ``sigmoid`` / ``F.normalize`` / ``clip`` back (geometry/gaussian_base.py:240-248, :371-400); learning rates
are set per step by the caller exactly like ``update_learning_rate`` does (:539-572).

CUDA only (no CPU path).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib, ops
from ._lib import check, lib

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
_FIELD = {"xyz": "xyz", "f_dc": "features_dc", "f_rest": "features_rest", "opacity": "opacity", "scaling": "scaling",
          "rotation": "rotation"}


class FusedGaussianAdam:
    """``params``: dict of the RAW parameter tensors (contiguous fp32 CUDA, updated in place) under the reference's
    group names: xyz (P,3), f_dc (P,1,3), f_rest (P,M-1,3), opacity (P,1), scaling (P,3), rotation (P,4).
    ``step(grads)`` takes the gradients with respect to the ACTIVATED values under the rasterizer's names
    (means3D, shs (P,M,3), opacities, scales, rotations) -- e.g. ``PackedGrads.grads()``."""

    def __init__(self, params: Dict[str, torch.Tensor], lrs: Dict[str, float], betas=(0.9, 0.999), eps: float = 1e-15,
                 color_clip: float = float("inf")):
        for k in GROUPS:
            t = params[k]
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError(f"FusedGaussianAdam: {k} must be a contiguous fp32 CUDA tensor (no CPU fallback)")
        self.params = params
        self.lrs = dict(lrs)
        self.betas, self.eps, self.color_clip = betas, eps, color_clip
        self.exp_avg = {k: torch.zeros_like(params[k]) for k in GROUPS}
        self.exp_avg_sq = {k: torch.zeros_like(params[k]) for k in GROUPS}
        self.steps = 0

    def step(self, grads: Dict[str, torch.Tensor]) -> None:
        P = self.params["xyz"].shape[0]
        M = 1 + self.params["f_rest"].shape[1]
        g = {k: grads[k] for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        if tuple(g["shs"].shape) != (P, M, 3):
            raise ValueError(f"shs gradient must be (P, M, 3) = {(P, M, 3)}; got {tuple(g['shs'].shape)}")
        for k, t in g.items():
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError(f"FusedGaussianAdam.step: gradient {k} must be contiguous fp32 CUDA")
        self.steps += 1
        a = _lib.AdamArgs()
        a.P, a.M = P, M
        for grp in GROUPS:
            f = _FIELD[grp]
            setattr(a, f, ops._ptr(self.params[grp]))
            setattr(a, "m_" + f, ops._ptr(self.exp_avg[grp]))
            setattr(a, "v_" + f, ops._ptr(self.exp_avg_sq[grp]))
        a.g_means3D, a.g_shs, a.g_opacities = g["means3D"].data_ptr(), g["shs"].data_ptr(), g["opacities"].data_ptr()
        a.g_scales, a.g_rotations = g["scales"].data_ptr(), g["rotations"].data_ptr()
        a.lr = (C.c_double * 6)(*[float(self.lrs[k]) for k in GROUPS])
        a.beta1, a.beta2, a.eps = float(self.betas[0]), float(self.betas[1]), float(self.eps)
        a.color_clip = float(min(self.color_clip, 3.0e38))
        a.step = self.steps
        a.stream = ops._stream(self.params["xyz"].device)
        with torch.cuda.device(self.params["xyz"].device):
            check(lib.b200splat_adam_step(C.byref(a)), "b200splat_adam_step")
