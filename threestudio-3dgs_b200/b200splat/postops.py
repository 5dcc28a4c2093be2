"""Fused per-pixel post-ops of the reference's renderer variants for a batch of views (autograd surface over
b200splat_postprocess_forward / _backward, csrc/postops.cu).

``postprocess_views(mode, image (V,3,H,W), depth (V,1,H,W), alpha (V,1,H,W), ...)`` returns what the tail of the
reference renderer's ``forward`` returns per view, stacked over the views:

    mode "plain"       render = clamp(image)                                  renderer/diff_gaussian_rasterizer_advanced.py:139-146
    mode "background"  render = clamp(image + (1 - alpha) * bg)               ..._background.py:130-141
    mode "normal"      render, normal (from depth), depth (gradient masked)   ..._normal.py:172-201
    mode "shading"     Lambert point light + composite, normal, depth         ..._shading.py:174-222,
                                                                              material/gaussian_material.py:70-104

CUDA only; there is no CPU path (the CPU restatement lives in oracle/postops.py and is test infrastructure).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib, ops
from ._lib import check, lib

MODES = {"plain": 0, "background": 1, "normal": 2, "shading": 3}
SHADINGS = {"albedo": 0, "textureless": 1, "diffuse": 2}


def _args(mode, shading, image, depth, alpha, rays_o, rays_d, bg, light, pred_normal, ambient, diffuse):
    V, _, H, W = image.shape
    a = _lib.PostprocessArgs()
    if isinstance(shading, (list, tuple)):      # per-view modes
        if len(shading) != V:
            raise ValueError("one shading mode per view expected")
        a._keep = (C.c_int32 * V)(*shading)
        a.shading_per_view = a._keep
        shading = shading[0]
    a.V, a.H, a.W, a.mode, a.shading = V, H, W, mode, shading
    a.image, a.depth, a.alpha = image.data_ptr(), depth.data_ptr(), alpha.data_ptr()
    a.rays_o, a.rays_d, a.bg = ops._ptr(rays_o), ops._ptr(rays_d), ops._ptr(bg)
    a.light, a.pred_normal = ops._ptr(light), ops._ptr(pred_normal)
    if ambient and isinstance(ambient[0], (list, tuple)):     # per-view light colours (soft_shading)
        if len(ambient) != V or len(diffuse) != V:
            raise ValueError("one (ambient, diffuse) pair per view expected")
        flat = [float(x) for v in range(V) for x in (*ambient[v], *diffuse[v])]
        a._keep_l = (C.c_float * (6 * V))(*flat)
        a.lights_per_view = a._keep_l
        ambient, diffuse = ambient[0], diffuse[0]
    a.ambient = (C.c_float * 3)(*ambient)
    a.diffuse = (C.c_float * 3)(*diffuse)
    a.stream = ops._stream(image.device)
    return a


class _Postprocess(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, depth, alpha, bg, mode, shading, rays_o, rays_d, light, pred_normal, ambient, diffuse):
        if not image.is_cuda:
            raise RuntimeError("b200splat.postprocess_views: tensors must be on a CUDA device (no CPU fallback)")
        f = ops._f32c
        image, depth, alpha = f(image, "image"), f(depth, "depth"), f(alpha, "alpha")
        bg, rays_o, rays_d = f(bg, "bg"), f(rays_o, "rays_o"), f(rays_d, "rays_d")
        light, pred_normal = f(light, "light"), f(pred_normal, "pred_normal")
        V, _, H, W = image.shape
        dev = image.device
        render = torch.empty(V, 3, H, W, dtype=torch.float32, device=dev)
        has_normal = mode >= MODES["normal"]
        normal = torch.empty(V, 3 if has_normal else 0, H, W, dtype=torch.float32, device=dev)
        depth_out = torch.empty_like(depth)
        a = _args(mode, shading, image, depth, alpha, rays_o, rays_d, bg, light, pred_normal, ambient, diffuse)
        a.render, a.normal, a.depth_out = render.data_ptr(), ops._ptr(normal), depth_out.data_ptr()
        with torch.cuda.device(dev):
            check(lib.b200splat_postprocess_forward(C.byref(a)), "b200splat_postprocess_forward")
        e = lambda t: t if t is not None else torch.empty(0, device=dev)
        ctx.save_for_backward(image, depth, alpha, e(bg), e(rays_o), e(rays_d), e(light), e(pred_normal))
        ctx.cfg = (mode, shading, ambient, diffuse, bg is not None and bg.requires_grad)
        return render, normal, depth_out

    @staticmethod
    def backward(ctx, g_render, g_normal, g_depth):
        image, depth, alpha, bg, rays_o, rays_d, light, pred = ctx.saved_tensors
        mode, shading, ambient, diffuse, bg_grad = ctx.cfg
        o = lambda t: t if t.numel() else None
        bg, rays_o, rays_d, light, pred = o(bg), o(rays_o), o(rays_d), o(light), o(pred)
        V, _, H, W = image.shape
        dev = image.device
        g = lambda t: None if (t is None or t.numel() == 0) else ops._f32c(t, "grad")
        g_render, g_normal, g_depth = g(g_render), g(g_normal), g(g_depth)
        d_image, d_depth, d_alpha = torch.empty_like(image), torch.empty_like(depth), torch.empty_like(alpha)
        d_bg = torch.empty_like(bg) if (bg is not None and bg_grad) else None
        a = _args(mode, shading, image, depth, alpha, rays_o, rays_d, bg, light, pred, ambient, diffuse)
        a.g_render, a.g_normal, a.g_depth = ops._ptr(g_render), ops._ptr(g_normal), ops._ptr(g_depth)
        a.d_image, a.d_depth, a.d_alpha, a.d_bg = d_image.data_ptr(), d_depth.data_ptr(), d_alpha.data_ptr(), \
            ops._ptr(d_bg)
        scratch = None
        if mode >= MODES["normal"]:
            nbytes = int(lib.b200splat_postprocess_scratch_bytes(V, H, W))
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            a.scratch, a.scratch_bytes = scratch.data_ptr(), nbytes
        with torch.cuda.device(dev):
            check(lib.b200splat_postprocess_backward(C.byref(a)), "b200splat_postprocess_backward")
        return (d_image, d_depth, d_alpha, d_bg) + (None,) * 8


def _lights(c):
    """(3,) colour or a list of V (3,) colours -> plain floats (tuples: they are kept on the autograd ctx)."""
    if len(c) and isinstance(c[0], (list, tuple)):
        return tuple(tuple(float(x) for x in v) for v in c)
    return tuple(float(x) for x in c)


def postprocess_views(mode: str, image, depth, alpha, *, bg=None, rays_o=None, rays_d=None, light_positions=None,
                      pred_normal: Optional[torch.Tensor] = None, shading: str = "diffuse",
                      ambient: Sequence[float] = (0.1, 0.1, 0.1), diffuse: Sequence[float] = (0.9, 0.9, 0.9)):
    """image (V,3,H,W), depth / alpha (V,1,H,W) as returned by ``ViewBatchRasterizer``; bg / rays_o / rays_d
    (V,H,W,3); light_positions (V,3); pred_normal (V,3,H,W) rendered per-Gaussian normals (used detached);
    shading: one mode or a list of V modes (the reference's material draws it per view in training).
    Returns dict(render (V,3,H,W) clamped, normal (V,3,H,W) | None, depth (V,1,H,W))."""
    m = MODES[mode]
    s = [SHADINGS[x] for x in shading] if isinstance(shading, (list, tuple)) else SHADINGS[shading]
    if m in (1, 3) and bg is None:
        raise ValueError(f"mode {mode!r} needs bg")
    if m >= 2 and (rays_o is None or rays_d is None):
        raise ValueError(f"mode {mode!r} needs rays_o and rays_d")
    if m == 3 and light_positions is None:
        raise ValueError("mode 'shading' needs light_positions")
    render, normal, depth_out = _Postprocess.apply(
        image, depth, alpha, bg, m, s, rays_o, rays_d, light_positions,
        None if pred_normal is None else pred_normal.detach(), _lights(ambient), _lights(diffuse))
    return dict(render=render, normal=normal if m >= 2 else None, depth=depth_out)
