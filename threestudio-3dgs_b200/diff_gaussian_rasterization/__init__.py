"""Drop-in ``diff_gaussian_rasterization`` backed by libb200splat.so (B200 / sm_100a).

Same operator surface as the package the reference imports
(renderer/diff_gaussian_rasterizer.py:8-11 and the 8 sibling renderers):

    GaussianRasterizationSettings   NamedTuple built at renderer/diff_gaussian_rasterizer.py:83-96
    GaussianRasterizer(raster_settings=...)(means3D=, means2D=, shs=, colors_precomp=, opacities=,
                                            scales=, rotations=, cov3D_precomp=)
        -> (color (3,H,W), radii (P,) int32, depth (1,H,W), alpha (1,H,W))     (4-tuple unpack at
           renderer/diff_gaussian_rasterizer_advanced.py:122)
    GaussianRasterizer.markVisible(positions)

Autograd contract kept: gradients are returned for every differentiable input in
the order (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3D_precomp);
means2D always receives a dense (P,3) gradient in NDC units because
geometry/gaussian_base.py:815-819 reads ``viewspace_points.grad[:, :2]``; outputs are freshly
allocated non-view tensors (callers write into them in place,
renderer/diff_gaussian_rasterizer_shading.py:209-213) and only ``alpha`` of the outputs is kept
for backward; backward can run twice on one forward (system/gaussian_splatting.py:129,137-138).

One optional, non-breaking addition: ``extra_features=(P, C')`` (C' <= 4) renders C' more per-Gaussian channels
with the SAME pass and alphas and appends their ``(C',H,W)`` image as a fifth output -- what the reference's
normal / shading variants obtain from a second full rasterizer call with the normals as colours
(renderer/diff_gaussian_rasterizer_shading.py:177-187).

CUDA only: there is no CPU fallback; a missing extension raises at import.
"""
from __future__ import annotations

import os
from typing import NamedTuple

import torch
from torch import nn

from b200splat import ops as _ops


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _opt(t):
    return None if (t is None or t.numel() == 0) else t


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, extra_features=None):
    # Fast path (the reference's call shape: scales + rotations, debug off): a pooled persistent one-view workspace --
    # no host round trip inside the forward, no per-call buffer allocation, no scratch memset (b200splat/batched.py).
    # cov3D_precomp, debug=True, P == 0 and more live graphs than the pool holds take the allocating entry point.
    P = means3D.shape[0]
    if (P > 0 and means3D.is_cuda and _opt(cov3Ds_precomp) is None and _opt(scales) is not None
            and not raster_settings.debug and _POOLED):
        from b200splat import batched as _batched
        rast = _batched.one_view_rasterizer(P, int(raster_settings.image_height), int(raster_settings.image_width),
                                            means3D.device)
        if rast is not None:
            cam = _ops.make_cam(raster_settings, means3D.device)
            out = _batched.one_view_forward(rast, cam, means3D, means2D, _opt(sh), _opt(colors_precomp), opacities,
                                            scales, rotations, extra_features)
            return out if extra_features is not None else out[:4]
    out = _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                    cov3Ds_precomp, raster_settings, extra_features)
    return out if extra_features is not None else out[:4]


_POOLED = os.environ.get("B200SPLAT_DROPIN_POOL", "1") != "0"


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, extra_features=None):
        dev = means3D.device
        if not means3D.is_cuda:
            raise RuntimeError("diff_gaussian_rasterization (b200splat): inputs must be CUDA tensors; "
                               "there is no CPU fallback")
        cam = _ops.make_cam(raster_settings, dev)
        P = means3D.shape[0]
        ctx.P = P
        n_extra = _ops._check_extra(extra_features, P)
        if P == 0:
            H, W = cam.H, cam.W
            color = cam.bg.reshape(3, 1, 1).expand(3, H, W).contiguous()
            z = lambda c=1: torch.zeros(c, H, W, dtype=torch.float32, device=dev)
            radii = torch.zeros(0, dtype=torch.int32, device=dev)
            ctx.mark_non_differentiable(radii)
            return color, radii, z(), z(), z(n_extra)
        f = _ops._f32c
        m3, sh_, cp_, op_ = f(means3D, "means3D"), f(_opt(sh), "shs"), f(_opt(colors_precomp), "colors_precomp"), \
            f(opacities, "opacities")
        sc_, ro_, c3_ = f(_opt(scales), "scales"), f(_opt(rotations), "rotations"), f(_opt(cov3Ds_precomp), "cov3D")
        ex_ = f(extra_features, "extra_features") if n_extra else None
        res = _ops.forward(cam, m3, sh_, cp_, op_, sc_, ro_, c3_, ex_)
        color, radii, depth, alpha = res[:4]
        st = res[-1]
        extra_img = res[4] if n_extra else torch.zeros(0, cam.H, cam.W, dtype=torch.float32, device=dev)
        ctx.n_extra = n_extra
        ctx.cam = cam
        ctx.state = (st.P, st.M, st.num_rendered)
        ctx.present = (sh_ is not None, cp_ is not None, sc_ is not None, c3_ is not None)
        e = lambda t: t if t is not None else torch.empty(0, device=dev)
        ctx.save_for_backward(m3, e(sh_), e(cp_), op_, e(sc_), e(ro_), e(c3_), radii, alpha, st.geom,
                              e(st.binning), st.image, e(ex_))
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha, extra_img

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha, g_extra):
        if ctx.P == 0:
            return (None,) * 10
        m3, sh_, cp_, op_, sc_, ro_, c3_, radii, alpha, geom, binning, image, ex_ = ctx.saved_tensors
        has_sh, has_cp, has_sr, has_c3 = ctx.present
        P, M, R = ctx.state
        st = _ops.ForwardState(P, M, R, geom, _opt(binning), image)
        gc = None if g_color is None else _ops._f32c(g_color, "grad_color")
        gd = None if g_depth is None else _ops._f32c(g_depth, "grad_depth")
        ga = None if g_alpha is None else _ops._f32c(g_alpha, "grad_alpha")
        n_extra = ctx.n_extra
        ge = None if (g_extra is None or not n_extra) else _ops._f32c(g_extra, "grad_extra")
        g = _ops.backward(ctx.cam, st, m3, sh_ if has_sh else None, cp_ if has_cp else None, op_,
                          sc_ if has_sr else None, ro_ if has_sr else None, c3_ if has_c3 else None,
                          radii, alpha, gc, gd, ga, extra_features=ex_ if n_extra else None, g_extra=ge)
        return (g["means3D"], g["means2D"], g.get("shs"), g.get("colors_precomp"), g["opacities"],
                g.get("scales"), g.get("rotations"), g.get("cov3D_precomp"), None, g.get("extra_features"))


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _ops.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, extra_features=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, shs if shs is not None else empty,
                                   colors_precomp if colors_precomp is not None else empty, opacities,
                                   scales if scales is not None else empty,
                                   rotations if rotations is not None else empty,
                                   cov3D_precomp if cov3D_precomp is not None else empty, self.raster_settings,
                                   extra_features)


__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
