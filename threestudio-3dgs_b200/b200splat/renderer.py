"""Batched drop-in for the reference's ``GaussianBatchRenderer`` mixin (renderer/gaussian_batch_renderer.py:8-122).

The reference's ``batch_forward(batch)`` is a Python loop: per view it builds a camera, calls the renderer's
``forward`` (one rasterizer call -- two with ``pred_normal`` -- followed by ~25 per-pixel PyTorch kernels in the
normal / shading variants) and finally stacks the per-view results.  ``B200GaussianBatchRenderer.batch_forward`` takes
the same ``batch`` dict and returns the same ``outputs`` dict (same keys, shapes and autograd behaviour) from

  * ONE view-batched rasterizer call per <= 8 views (``ViewBatchRasterizer``; the per-Gaussian normals of
    ``pred_normal`` ride along as extra feature channels of the same pass instead of a second pass), and
  * ONE fused post-op launch for all views (``postprocess_views``).

Mix it into a renderer class that has what the reference's renderers have: ``self.geometry`` (the
``GaussianBaseModel`` getters, geometry/gaussian_base.py:371-411), ``self.background_tensor``, ``self.training`` and,
for the background / shading variants, ``self.background`` (called with ``dirs=(V,H,W,3)``) and ``self.material``
(``ambient_light_color``, ``diffuse_light_color``, ``ambient_only``, ``cfg.diffuse_prob``,
``cfg.textureless_prob``, ``cfg.soft_shading`` -- material/gaussian_material.py:13-104; all of them honoured).  ``variant`` selects which of
the reference's renderer files is mirrored:

    "plain"       renderer/diff_gaussian_rasterizer.py            render only, background inverted in eval
    "advanced"    renderer/diff_gaussian_rasterizer_advanced.py   + depth, mask
    "background"  renderer/diff_gaussian_rasterizer_background.py learned background composited after the pass
    "normal"      renderer/diff_gaussian_rasterizer_normal.py     + normal from depth, masked gradients
    "shading"     renderer/diff_gaussian_rasterizer_shading.py    + point-light shading and composite

Differences a caller can observe, all deliberate: (1) ``viewspace_points[v]`` is a small stand-in whose ``.grad`` is
view v's slice of ONE (V,P,3) gradient tensor -- enough for the unchanged ``GaussianBaseModel.update_states``
(geometry/gaussian_base.py:815-819, :845-851), which only reads ``.grad[filter, :2]``; (2) the material's random
shading mode is drawn per view with the same ``random.random()`` call sequence as the reference's per-view loop.
With ``pred_normal`` the normals ride along as extra channels of the colour pass; as in the reference (whose second pass
gets a gradient-free zeros tensor as means2D, renderer/diff_gaussian_rasterizer_shading.py:180) their gradient reaches
``means3D`` but not ``viewspace_points`` / the densification statistic (the library keeps the two apart, render.cu K7).
"""
from __future__ import annotations

import math
import random
from typing import Any, Dict, List

import numpy as np
import torch

from . import scenes
from .batched import MAX_VIEWS, ViewBatchRasterizer
from .postops import postprocess_views

SH_C0 = 0.28209479177387814


def _camera_batch(batch, znear: float = 0.1, zfar: float = 100.0):
    """All V cameras of ``batch`` at once, ON THE DEVICE: what the reference's loop builds per view with
    ``get_cam_info_gaussian(c2w, fovx=fovy, fovy, znear=0.1, zfar=100)`` (renderer/gaussian_batch_renderer.py:23-49;
    the same construction as geometry/sugar.py:891-896 + utils/sugar_utils.py:809-829) -- flip the y / z camera axes
    of c2w, invert, transpose -> world_view_transform; times the transposed projection -> full_proj_transform; camera
    centre = inverse(world_view_transform)[3, :3].  Nothing here reads a tensor back: the reference's
    ``math.tan(FoVx * 0.5)`` on a 0-dim CUDA tensor (renderer/diff_gaussian_rasterizer.py:80-81) is one D2H
    synchronisation per view; here tan(fovy / 2) stays a device tensor all the way into the kernels
    (b200splat_camera.scalars_dev).  The tangent is taken in float64 and rounded to fp32, as Python's math.tan does.
    Returns (world_view (V,4,4), full_proj (V,4,4), centre (V,3), tanfov (V,))."""
    c2w = batch["c2w"].detach().to(torch.float32).clone()
    V = c2w.shape[0]
    if c2w.shape[-2] == 3:
        last = torch.tensor([0.0, 0.0, 0.0, 1.0], device=c2w.device).expand(V, 1, 4)
        c2w = torch.cat([c2w, last], dim=1)
    c2w[:, :3, 1:3] *= -1
    # inv_ex without the error check: linalg.inv reads its info tensor back (a synchronisation)
    wvt = torch.linalg.inv_ex(c2w, check_errors=False).inverse.transpose(1, 2).contiguous()
    fovy = batch["fovy"].detach().reshape(-1).to(c2w.device)
    tan64 = torch.tan(fovy.double() * 0.5)
    # utils/sugar_utils.py:809-829 with fovx = fovy: P[0,0] = P[1,1] = 2 znear / (2 tan znear), P[3,2] = 1,
    # P[2,2] = zfar / (zfar - znear), P[2,3] = -zfar znear / (zfar - znear)
    p00 = (2.0 * znear / (2.0 * (tan64 * znear))).float()
    proj_t = torch.zeros(V, 4, 4, device=c2w.device)          # the TRANSPOSED projection
    proj_t[:, 0, 0] = p00
    proj_t[:, 1, 1] = p00
    proj_t[:, 2, 2] = zfar / (zfar - znear)
    proj_t[:, 3, 2] = -(zfar * znear) / (zfar - znear)
    proj_t[:, 2, 3] = 1.0
    full = torch.bmm(wvt, proj_t).contiguous()
    centre = torch.linalg.inv_ex(wvt, check_errors=False).inverse[:, 3, :3].contiguous()
    return wvt, full, centre, tan64.float()


def _settings(batch, v: int, bg: torch.Tensor, sh_degree: int, device, cams=None):
    """GaussianRasterizationSettings of view v; tanfovx / tanfovy are 0-dim DEVICE tensors (b200splat's operators take
    them as such and never read them back)."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    wvt, full, centre, tan = cams if cams is not None else _camera_batch(batch)
    return GaussianRasterizationSettings(
        image_height=int(batch["height"]), image_width=int(batch["width"]), tanfovx=tan[v], tanfovy=tan[v], bg=bg,
        scale_modifier=1.0, viewmatrix=wvt[v], projmatrix=full[v], sh_degree=sh_degree, campos=centre[v],
        prefiltered=False, debug=False)


class _ViewspacePoints:
    """Stand-in for one view's ``means2D`` tensor in ``outputs["viewspace_points"]``: ``.grad`` is that view's slice of
    the batch tensor's gradient, which is all the unchanged ``GaussianBaseModel.update_states`` reads
    (geometry/gaussian_base.py:815-819, :845-851)."""

    def __init__(self, owner: torch.Tensor, v: int):
        self._owner, self._v = owner, v

    @property
    def grad(self):
        g = self._owner.grad
        return None if g is None else g[self._v]

    @property
    def data(self):
        return self._owner.detach()[self._v]


class B200GaussianBatchRenderer:
    variant: str = "plain"
    invert_bg_prob: float = 1.0     # Config.invert_bg_prob of the plain / advanced / normal renderers

    # ---- what the reference's material.forward decides per call (material/gaussian_material.py:51-93) ----------
    def _light_and_shading(self):
        """One view's (ambient rgb, diffuse rgb, shading mode), drawn with the same ``random.random()`` call sequence as
        the reference's per-view ``material.forward``: first the ambient ratio of ``soft_shading`` (training only),
        then the shading-mode draws.  The material's two colour buffers are read back once and cached."""
        mat = self.material
        cache = self.__dict__.setdefault("_b200_light_cache", {})
        key = (id(mat.ambient_light_color), id(mat.diffuse_light_color))
        if key not in cache:
            cache.clear()
            cache[key] = (tuple(float(x) for x in mat.ambient_light_color.tolist()),
                          tuple(float(x) for x in mat.diffuse_light_color.tolist()))
        ambient, diffuse = cache[key]
        if mat.training and getattr(mat.cfg, "soft_shading", False):
            r = random.random()
            diffuse, ambient = (r, r, r), (1.0 - r, 1.0 - r, 1.0 - r)
        if mat.training:
            if mat.ambient_only or random.random() > mat.cfg.diffuse_prob:
                mode = "albedo"
            elif random.random() < mat.cfg.textureless_prob:
                mode = "textureless"
            else:
                mode = "diffuse"
        else:
            mode = "albedo" if mat.ambient_only else "diffuse"
        return ambient, diffuse, mode

    def _rasterizer(self, V: int, P: int, H: int, W: int, device) -> ViewBatchRasterizer:
        """A ViewBatchRasterizer (persistent workspace) for this shape whose last graph has been consumed.  Systems that
        render twice before ``backward`` (system/gaussian_zero123.py:212-235: random-camera batch + reference view)
        get a second instance instead of overwriting the first one's saved state."""
        key = (V, H, W, str(device))   # not P: a workspace survives densify / prune (it grows when it has to)
        pool = self.__dict__.setdefault("_b200_rasterizers", {}).setdefault(key, [])
        for rast in pool:
            if not rast.pending:
                return rast
        if len(pool) >= 4:
            # four live graphs of one shape: hand out the oldest workspace again; should its graph still run a
            # backward, that backward raises (generation check) instead of reading another forward's state
            pool.append(pool.pop(0))
            return pool[-1]
        pool.append(ViewBatchRasterizer(V, P, H, W, device))
        return pool[-1]

    def batch_forward(self, batch) -> Dict[str, Any]:
        pc = self.geometry
        variant = self.variant
        bs = batch["c2w"].shape[0]
        H, W = int(batch["height"]), int(batch["width"])
        means3D = pc.get_xyz
        dev, P = means3D.device, means3D.shape[0]
        # background colour of the raster pass: plain / advanced / normal invert it at random in training and always
        # in eval (renderer/diff_gaussian_rasterizer.py:59-64, ..._normal.py:92-97); background / shading rasterize
        # over black and composite afterwards (..._background.py:58, ..._shading.py:94)
        bgs: List[torch.Tensor] = []
        for _ in range(bs):
            if variant in ("plain", "advanced", "normal"):
                invert = (np.random.rand() > self.invert_bg_prob) if self.training else True
                bgs.append(1.0 - self.background_tensor if invert else self.background_tensor)
            else:
                bgs.append(self.background_tensor * 0)
        override = batch.get("override_color")
        shs = pc.get_features if override is None else None
        pred_normal = variant in ("normal", "shading") and bool(getattr(pc.cfg, "pred_normal", False))
        # the reference's second pass treats the normals as a degree-0 SH colour: max(C0 * n + 0.5, 0)
        extra = torch.clamp_min(SH_C0 * pc.get_normal + 0.5, 0.0) if pred_normal else None
        opac, scales, rots = pc.get_opacity, pc.get_scaling, pc.get_rotation
        vsp = torch.zeros(bs, P, 3, dtype=means3D.dtype, device=dev, requires_grad=True)
        images, depths, alphas, radiis, extras = [], [], [], [], []
        cams = _camera_batch(batch)     # all views' cameras with batched device ops: no per-view D2H / H2D round trips
        for v0 in range(0, bs, MAX_VIEWS):
            n = min(MAX_VIEWS, bs - v0)
            settings = [_settings(batch, v, bgs[v], pc.active_sh_degree, dev, cams) for v in range(v0, v0 + n)]
            rast = self._rasterizer(n, P, H, W, dev)
            with torch.autocast("cuda", enabled=False):
                out = rast(settings, means3D, vsp[v0:v0 + n], opac, shs=shs, colors_precomp=override, scales=scales,
                           rotations=rots, extra_features=extra)
            images.append(out[0]), radiis.append(out[1]), depths.append(out[2]), alphas.append(out[3])
            if pred_normal:
                extras.append(out[4])
        image, depth, alpha = torch.cat(images), torch.cat(depths), torch.cat(alphas)
        radii = torch.cat(radiis)
        pred_map = torch.cat(extras) if pred_normal else None
        if pred_normal and variant == "normal":
            # the reference's second pass shares the first one's settings, so its image carries the background term
            # T_final * bg (renderer/diff_gaussian_rasterizer_normal.py:175-185); sum(alpha_i T_i) = 1 - T_final
            pred_map = pred_map + (1.0 - alpha) * torch.stack(bgs).reshape(bs, 3, 1, 1)

        outputs: Dict[str, Any] = {
            "viewspace_points": [_ViewspacePoints(vsp, v) for v in range(bs)],
            "visibility_filter": list((radii > 0).unbind(0)),
            "radii": list(radii.unbind(0)),
        }
        kw = {}
        comp_bg = None
        if variant in ("background", "shading"):
            if batch.get("override_bg_color") is not None and variant == "shading":
                comp_bg = batch["override_bg_color"].expand(bs, H, W, -1)
            else:
                comp_bg = self.background(dirs=batch["rays_d"])
            kw["bg"] = comp_bg.reshape(bs, H, W, 3)
        if variant in ("normal", "shading"):
            kw.update(rays_o=batch["rays_o"], rays_d=batch["rays_d"])
        if variant == "shading":
            lights = [self._light_and_shading() for _ in range(bs)]
            kw.update(light_positions=batch["light_positions"], pred_normal=pred_map, shading=[m for _, _, m in lights],
                      ambient=[a for a, _, _ in lights], diffuse=[d for _, d, _ in lights])
        mode = {"plain": "plain", "advanced": "plain"}.get(variant, variant)
        post = postprocess_views(mode, image, depth, alpha, **kw)
        outputs["comp_rgb"] = post["render"].permute(0, 2, 3, 1)
        if variant in ("normal", "shading"):
            outputs["comp_normal"] = post["normal"].permute(0, 2, 3, 1)
            if pred_map is not None:
                outputs["comp_pred_normal"] = pred_map.permute(0, 2, 3, 1)
        if variant != "plain" and variant != "background":
            outputs["comp_depth"] = post["depth"].permute(0, 2, 3, 1)
            outputs["comp_mask"] = alpha.permute(0, 2, 3, 1)
        if variant == "shading":
            outputs["comp_rgb_bg"] = comp_bg.reshape(bs, H, W, 3)
        return outputs
