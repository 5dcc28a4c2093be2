"""CPU oracle of the per-pixel post-ops that follow the rasterizer in the reference's renderer variants.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg) -- the product path
(threestudio-3dgs_b200/csrc/postops.cu behind b200splat_postprocess_*) never imports this file.

Plain PyTorch restatement, statement by statement, of

  * ``Depth2Normal``                       renderer/diff_gaussian_rasterizer_shading.py:22-51
  * the shading variant's tail             renderer/diff_gaussian_rasterizer_shading.py:174-213, :222
  * the normal variant's tail              renderer/diff_gaussian_rasterizer_normal.py:172-173, :189-193, :201
  * the background variant's tail          renderer/diff_gaussian_rasterizer_background.py:130-132, :141
  * ``GaussianDiffuseWithPointLightMaterial.forward``   material/gaussian_material.py:41-104
    (``dot`` is threestudio.utils.ops.dot = (x*y).sum(-1, keepdim=True); restated)

Backward = autograd of this forward (the reference has no hand-written backward here).  Pinned against the
reference's own classes imported from /root/reference in tests/test_postops_cpu.py (skipped where the reference
is absent) and against tests/golden/postops_ref.npz, generated from those classes by tests/golden/make_postops_golden.py.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

MODE_PLAIN = 0        # render = clamp(image)                                   (…_advanced.py:139-146)
MODE_BACKGROUND = 1   # render = clamp(image + (1 - alpha) * bg)                (…_background.py:130-141)
MODE_NORMAL = 2       # plain render + normal map from depth + masked depth     (…_normal.py)
MODE_SHADING = 3      # normal map + Lambert point light + composite            (…_shading.py)

SHADING = {"albedo": 0, "textureless": 1, "diffuse": 2}


def depth2normal(xyz_chw: torch.Tensor) -> torch.Tensor:
    """Depth2Normal.forward for one (3,H,W) map (shading.py:40-51): central differences with zero padding,
    normal = -cross(d/dx, d/dy)."""
    x = xyz_chw.unsqueeze(0)
    B, C, H, W = x.shape
    kx = torch.tensor([[0.0, 0.0, 0.0], [-1.0, 0.0, 1.0], [0.0, 0.0, 0.0]], dtype=x.dtype).view(1, 1, 3, 3)
    ky = torch.tensor([[0.0, -1.0, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dtype=x.dtype).view(1, 1, 3, 3)
    dzdx = F.conv2d(x.reshape(B * C, 1, H, W), kx, padding=1).reshape(B, C, H, W)
    dzdy = F.conv2d(x.reshape(B * C, 1, H, W), ky, padding=1).reshape(B, C, H, W)
    return (-torch.cross(dzdx, dzdy, dim=1))[0]


def material(positions, shading_normal, light_positions, albedo, ambient, diffuse, shading: str):
    """GaussianDiffuseWithPointLightMaterial.forward with the light colours and the shading mode given
    (material/gaussian_material.py:70-104); all maps are (H,W,3)."""
    light_directions = F.normalize(light_positions - positions, dim=-1)
    dot = (shading_normal * light_directions).sum(-1, keepdim=True)
    diffuse_light = dot.clamp(min=0.0) * diffuse
    textureless_color = diffuse_light + ambient
    color = albedo.clamp(0.0, 1.0) * textureless_color
    if shading == "albedo":
        return albedo + textureless_color * 0
    if shading == "textureless":
        return albedo * 0 + textureless_color
    if shading == "diffuse":
        return color
    raise ValueError(shading)


def postprocess_view(mode: int, image, depth, alpha, rays_o=None, rays_d=None, bg=None, light_position=None,
                     ambient=None, diffuse=None, shading: str = "diffuse", pred_normal: Optional[torch.Tensor] = None):
    """One view.  image (3,H,W), depth (1,H,W), alpha (1,H,W), rays_o / rays_d / bg (H,W,3), light_position (3,).
    Returns dict(render (3,H,W), normal (3,H,W) | None, depth (1,H,W)) with the reference's gradient masking
    (normal and depth detached where alpha <= 0.99)."""
    _, H, W = image.shape
    out_depth = depth
    normal_out = None
    if mode in (MODE_NORMAL, MODE_SHADING):
        xyz_map = rays_o + depth.permute(1, 2, 0) * rays_d                       # shading.py:174
        normal_map = depth2normal(xyz_map.permute(2, 0, 1))                      # :175
        normal_map = F.normalize(normal_map, dim=0)                              # :176
    if mode == MODE_SHADING:
        light_positions = light_position[None, None, :].expand(H, W, -1)         # :191-193
        if pred_normal is not None:
            shading_normal = pred_normal.permute(1, 2, 0).detach() * 2 - 1       # :196
            shading_normal = F.normalize(shading_normal, dim=2)                  # :197
        else:
            shading_normal = normal_map.permute(1, 2, 0)                         # :199
        rgb_fg = material(xyz_map, shading_normal, light_positions,
                          (image / (alpha + 1e-6)).permute(1, 2, 0), ambient, diffuse, shading).permute(2, 0, 1)
        image = rgb_fg * alpha + (1 - alpha) * bg.reshape(H, W, 3).permute(2, 0, 1)   # :206-208
    elif mode == MODE_BACKGROUND:
        image = image + (1 - alpha) * bg.reshape(H, W, 3).permute(2, 0, 1)       # background.py:130-132
    if mode in (MODE_NORMAL, MODE_SHADING):
        normal_map = normal_map * 0.5 * alpha + 0.5                              # shading.py:209
        mask = alpha > 0.99                                                      # :210
        normal_mask = mask.repeat(3, 1, 1)
        # out-of-place form of `x[~m] = x[~m].detach()` (:212-213): same values, gradient only where m
        normal_out = torch.where(normal_mask, normal_map, normal_map.detach())
        out_depth = torch.where(mask, depth, depth.detach())
    return dict(render=image.clamp(0, 1), normal=normal_out, depth=out_depth)
