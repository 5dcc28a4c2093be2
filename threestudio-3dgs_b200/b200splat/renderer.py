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
``cfg.textureless_prob``, ``cfg.soft_shading`` -- material/gaussian_material.py:13-104).  ``variant`` selects which of
the reference's renderer files is mirrored:

    "plain"       renderer/diff_gaussian_rasterizer.py            render only, background inverted in eval
    "advanced"    renderer/diff_gaussian_rasterizer_advanced.py   + depth, mask
    "background"  renderer/diff_gaussian_rasterizer_background.py learned background composited after the pass
    "normal"      renderer/diff_gaussian_rasterizer_normal.py     + normal from depth, masked gradients
    "shading"     renderer/diff_gaussian_rasterizer_shading.py    + point-light shading and composite

Differences a caller can observe, all deliberate: (1) ``viewspace_points[v]`` is a small stand-in whose ``.grad`` is
view v's slice of ONE (V,P,3) gradient tensor -- enough for the unchanged ``GaussianBaseModel.update_states``
(geometry/gaussian_base.py:815-819, :845-851), which only reads ``.grad[filter, :2]``; (2) with ``pred_normal`` the
normals' gradient also reaches ``means2D`` (the reference feeds its second pass a gradient-free zeros tensor,
renderer/diff_gaussian_rasterizer_shading.py:180); (3) the material's random shading mode is drawn per view with the
same ``random.random()`` call sequence as the reference's per-view loop.
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


def _settings(batch, v: int, bg: torch.Tensor, sh_degree: int, device):
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    fovy = float(batch["fovy"][v])
    wvt, full, center = scenes.cam_info_gaussian(batch["c2w"][v].detach().float().cpu(), fovy, fovy, 0.1, 100.0)
    t = math.tan(fovy * 0.5)
    return GaussianRasterizationSettings(
        image_height=int(batch["height"]), image_width=int(batch["width"]), tanfovx=t, tanfovy=t, bg=bg,
        scale_modifier=1.0, viewmatrix=wvt.to(device), projmatrix=full.to(device), sh_degree=sh_degree,
        campos=center.to(device), prefiltered=False, debug=False)


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
    def _shading_mode(self) -> str:
        mat = self.material
        if self.training and getattr(mat.cfg, "soft_shading", False):
            raise NotImplementedError("soft_shading draws new light colours per view; use the per-view operator")
        if mat.training:
            if mat.ambient_only or random.random() > mat.cfg.diffuse_prob:
                return "albedo"
            if random.random() < mat.cfg.textureless_prob:
                return "textureless"
            return "diffuse"
        return "albedo" if mat.ambient_only else "diffuse"

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
        for v0 in range(0, bs, MAX_VIEWS):
            n = min(MAX_VIEWS, bs - v0)
            settings = [_settings(batch, v, bgs[v], pc.active_sh_degree, dev) for v in range(v0, v0 + n)]
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
            modes = [self._shading_mode() for _ in range(bs)]
            mat = self.material
            kw.update(light_positions=batch["light_positions"], pred_normal=pred_map, shading=modes,
                      ambient=mat.ambient_light_color.tolist(), diffuse=mat.diffuse_light_color.tolist())
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
