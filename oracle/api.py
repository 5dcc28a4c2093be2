"""Operator surface of the oracle: the same names the reference imports.

TEST INFRASTRUCTURE ONLY.  Lets tests run the *unchanged* reference renderer files
(renderer/diff_gaussian_rasterizer.py:8-11 imports ``GaussianRasterizationSettings`` and
``GaussianRasterizer`` from ``diff_gaussian_rasterization``) on the CPU oracle by
registering this module under that name (tests/stubs), which is BASELINE.json configs[0].
The product package (threestudio-3dgs_b200/diff_gaussian_rasterization) never imports it.
"""
from __future__ import annotations

import torch
from torch import nn

from . import torch_oracle as O
from .knn import dist2_oracle

GaussianRasterizationSettings = O.Settings


class _OracleRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3D_precomp, settings):
        opt = lambda t: None if (t is None or t.numel() == 0) else t.detach().float().cpu()
        inputs = tuple(map(opt, (means3D, means2D, sh, colors_precomp, opacities, scales,
                                 rotations, cov3D_precomp)))
        s = settings._replace(bg=settings.bg.detach().float().cpu(),
                              viewmatrix=settings.viewmatrix.detach().float().cpu(),
                              projmatrix=settings.projmatrix.detach().float().cpu(),
                              campos=settings.campos.detach().float().cpu())
        m3 = inputs[0]
        if m3 is None or m3.shape[0] == 0:
            H, W = s.image_height, s.image_width
            color = s.bg.reshape(3, 1, 1).expand(3, H, W).clone()
            ctx.empty = True
            return color, torch.zeros(0, dtype=torch.int32), torch.zeros(1, H, W), torch.zeros(1, H, W)
        out, pre, binned = O.rasterize_forward(m3, inputs[1], inputs[2], inputs[3], inputs[4],
                                               inputs[5], inputs[6], inputs[7], s)
        ctx.empty = False
        ctx.s = s
        ctx.inputs = inputs
        ctx.state = (pre, binned, {"n_contrib": out["n_contrib"]})
        ctx.present = tuple(t is not None for t in inputs)
        ctx.mark_non_differentiable(pre["radii"])
        return out["color"], pre["radii"], out["depth"], out["alpha"]

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha):
        if ctx.empty:
            return (None,) * 9
        s = ctx.s
        H, W = s.image_height, s.image_width
        z3 = torch.zeros(3, H, W)
        z1 = torch.zeros(1, H, W)
        gc = z3 if g_color is None else g_color.contiguous().float()
        gd = z1 if g_depth is None else g_depth.contiguous().float()
        ga = z1 if g_alpha is None else g_alpha.contiguous().float()
        pre, binned, fwd = ctx.state
        res = O.rasterize_backward(ctx.inputs, s, pre, binned, fwd, gc, gd, ga)
        order = ("means3D", "means2D", "shs", "colors_precomp", "opacities", "scales", "rotations",
                 "cov3D_precomp")
        grads = [res[k] for k in order]
        # means2D is always given a dense gradient (geometry/gaussian_base.py:816-818 reads it)
        return (*grads, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            return O.mark_visible(positions.float().cpu(), self.raster_settings)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        return _OracleRasterize.apply(means3D, means2D, shs, colors_precomp, opacities, scales,
                                      rotations, cov3D_precomp, self.raster_settings)


def distCUDA2(points):
    return dist2_oracle(points.detach().float().cpu())
