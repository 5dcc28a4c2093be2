"""Shared helpers for the parity tests: one scene -> oracle inputs and CUDA inputs."""
from __future__ import annotations

import torch

from b200splat import scenes
from oracle import torch_oracle as O
from oracle.checks import (borderline_bounds, borderline_pixels, check_grads_bounded, check_images,  # noqa: F401
                           cut_variants)


def oracle_settings(cam: scenes.Camera, sh_degree: int, bg=(1.0, 1.0, 1.0), scale_modifier=1.0) -> O.Settings:
    return O.Settings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                      torch.tensor(bg, dtype=torch.float32), scale_modifier, cam.viewmatrix, cam.projmatrix,
                      sh_degree, cam.campos, False, False)


def cuda_settings(s: O.Settings, device="cuda"):
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=s.image_height, image_width=s.image_width, tanfovx=s.tanfovx, tanfovy=s.tanfovy,
        bg=s.bg.to(device), scale_modifier=s.scale_modifier, viewmatrix=s.viewmatrix.to(device),
        projmatrix=s.projmatrix.to(device), sh_degree=s.sh_degree, campos=s.campos.to(device),
        prefiltered=False, debug=False)


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    """||g - g*||_inf / max(||g*||_inf, eps)  (SURVEY.md 8c tolerance definition)."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-20))
