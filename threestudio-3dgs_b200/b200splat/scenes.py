"""Synthetic random-Gaussian scenes and cameras for the BASELINE.json workloads.

Host-side data generation only (CPU torch.Generator, so the oracle and the GPU see identical
bits); recipe in SURVEY.md 8d:

  * positions: uniform in a ball (r0 * cbrt(U), uniform direction) -- same recipe as the
    reference's random init, geometry/gaussian_base.py:350-359
  * scales: sqrt(clamp_min(mean 3-NN dist^2, 1e-7)) (what distCUDA2 gives at init,
    geometry/gaussian_base.py:434-438) times a per-axis anisotropy exp(U(-0.7, 0.7))
  * rotations: normalised N(0, I4); opacities: sigmoid(N(0, 1.5^2)) shaped (P,1)
  * SH: DC = RGB2SH(U(0,1)) (geometry/gaussian_base.py:35-36), rest N(0, 0.05^2), layout (P,M,3)
  * cameras: data/uncond.py:150-352 (SDS) or the 4-view MVDream rig; camera matrices as
    threestudio's get_cam_info_gaussian (restated: flip c2w y/z columns, invert, transpose)
"""
from __future__ import annotations

import math
from typing import NamedTuple

import numpy as np
import torch

SH_C0 = 0.28209479177387814


class Camera(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor     # (4,4) world_view_transform  (W2C transposed)
    projmatrix: torch.Tensor     # (4,4) full_proj_transform
    campos: torch.Tensor         # (3,)
    fovy: float


class Scene(NamedTuple):
    means3D: torch.Tensor        # (P,3)
    scales: torch.Tensor         # (P,3) activated
    rotations: torch.Tensor      # (P,4) unit (r,x,y,z)
    opacities: torch.Tensor      # (P,1) activated
    shs: torch.Tensor            # (P,M,3)
    sh_degree: int


def mean_knn_dist2(points: torch.Tensor) -> torch.Tensor:
    """Mean squared distance to the 3 nearest other points (host, exact; scipy KD-tree)."""
    from scipy.spatial import cKDTree

    pts = points.double().numpy()
    tree = cKDTree(pts)
    k = min(4, pts.shape[0])
    d, _ = tree.query(pts, k=k, workers=-1)
    d2 = (d[:, 1:] ** 2).sum(axis=1) / 3.0
    return torch.from_numpy(d2).float()


def make_scene(P: int, sh_degree: int, r0: float, seed: int) -> Scene:
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(P, generator=g)
    dirs = torch.randn(P, 3, generator=g)
    dirs = dirs / dirs.norm(dim=1, keepdim=True).clamp_min(1e-12)
    xyz = (r0 * u.pow(1.0 / 3.0))[:, None] * dirs
    s0 = torch.sqrt(mean_knn_dist2(xyz).clamp_min(1e-7))
    aniso = torch.exp((torch.rand(P, 3, generator=g) * 2 - 1) * 0.7)
    scales = s0[:, None] * aniso
    q = torch.randn(P, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True).clamp_min(1e-12)
    opac = torch.sigmoid(torch.randn(P, 1, generator=g) * 1.5)
    M = (sh_degree + 1) ** 2
    dc = (torch.rand(P, 1, 3, generator=g) - 0.5) / SH_C0
    rest = torch.randn(P, M - 1, 3, generator=g) * 0.05
    shs = torch.cat([dc, rest], dim=1).contiguous()
    return Scene(xyz.contiguous().float(), scales.contiguous().float(), q.contiguous().float(),
                 opac.contiguous().float(), shs.float(), sh_degree)


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """Same construction as utils/sugar_utils.py:809-829 (getProjectionMatrix)."""
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def cam_info_gaussian(c2w: torch.Tensor, fovx: float, fovy: float, znear: float = 0.1,
                      zfar: float = 100.0):
    """threestudio.utils.ops.get_cam_info_gaussian restated (call site
    renderer/gaussian_batch_renderer.py:24-26): flip the y/z camera axes of c2w, invert,
    transpose -> (world_view_transform, full_proj_transform, camera_center)."""
    c2w = c2w.clone().float()
    c2w[:3, 1:3] *= -1
    w2c = torch.inverse(c2w)
    wvt = w2c.transpose(0, 1).contiguous()
    proj = projection_matrix(znear, zfar, fovx, fovy).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = wvt.inverse()[3, :3].contiguous()
    return wvt, full, center


def look_at_c2w(pos: torch.Tensor, center: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """data/uncond.py:303-315."""
    lookat = torch.nn.functional.normalize(center - pos, dim=-1)
    right = torch.nn.functional.normalize(torch.linalg.cross(lookat, up), dim=-1)
    up2 = torch.nn.functional.normalize(torch.linalg.cross(right, lookat), dim=-1)
    c2w = torch.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up2, -lookat, pos
    return c2w


def _camera(pos, fovy, H, W) -> Camera:
    c2w = look_at_c2w(pos, torch.zeros(3), torch.tensor([0.0, 0.0, 1.0]))
    wvt, full, center = cam_info_gaussian(c2w, fovy, fovy)
    t = math.tan(fovy * 0.5)
    return Camera(H, W, t, t, wvt, full, center, fovy)


def sds_cameras(B: int, H: int, W: int, seed: int, camera_distance: float = 2.5,
                fovy_deg=(60.0, 70.0), elevation_deg=(-20.0, 90.0)):
    """Random cameras as data/uncond.py:150-352 with configs/gaussian_splatting.yaml:11-16
    (batch-uniform azimuth, 50/50 elevation sampling, no perturbations)."""
    g = torch.Generator().manual_seed(seed)
    if float(torch.rand(1, generator=g)) < 0.5:
        el = torch.rand(B, generator=g) * (elevation_deg[1] - elevation_deg[0]) + elevation_deg[0]
        el = el * math.pi / 180
    else:
        lo, hi = (math.sin(e / 180 * math.pi) for e in elevation_deg)
        el = torch.asin(torch.rand(B, generator=g) * (hi - lo) + lo)
    az = (torch.rand(B, generator=g) + torch.arange(B)) / B * 360.0 - 180.0
    az = az * math.pi / 180
    fovy = (torch.rand(B, generator=g) * (fovy_deg[1] - fovy_deg[0]) + fovy_deg[0]) * math.pi / 180
    cams = []
    for i in range(B):
        pos = camera_distance * torch.stack([torch.cos(el[i]) * torch.cos(az[i]),
                                             torch.cos(el[i]) * torch.sin(az[i]), torch.sin(el[i])])
        cams.append(_camera(pos, float(fovy[i]), H, W))
    return cams


def mvdream_cameras(B: int, H: int, W: int, seed: int, n_view: int = 4):
    """MVDream rig of configs/gaussian_splatting_mvdream.yaml:9-23: groups of n_view evenly spaced
    azimuths sharing one elevation U(0,30) deg and one fovy U(15,60) deg, distance
    U(0.8,1.0)/tan(fovy/2) ("relative"; formula of the un-vendored mvdream datamodule, restated)."""
    g = torch.Generator().manual_seed(seed)
    cams = []
    groups = (B + n_view - 1) // n_view
    for _ in range(groups):
        el = float(torch.rand(1, generator=g)) * 30.0 * math.pi / 180
        fovy = (float(torch.rand(1, generator=g)) * 45.0 + 15.0) * math.pi / 180
        dist = (float(torch.rand(1, generator=g)) * 0.2 + 0.8) / math.tan(fovy / 2)
        az0 = float(torch.rand(1, generator=g)) * 2 * math.pi
        for v in range(n_view):
            az = az0 + 2 * math.pi * v / n_view
            pos = dist * torch.tensor([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az),
                                       math.sin(el)])
            cams.append(_camera(pos, fovy, H, W))
    return cams[:B]


def pixel_grads(H: int, W: int, seed: int):
    """Upstream gradients dL/dcolor, dL/ddepth, dL/dalpha ~ N(0,1)/(HW)."""
    g = torch.Generator().manual_seed(seed)
    s = 1.0 / (H * W)
    return (torch.randn(3, H, W, generator=g) * s, torch.randn(1, H, W, generator=g) * s,
            torch.randn(1, H, W, generator=g) * s)


# name -> (P, sh_degree, r0, H, W, views, camera rig)
WORKLOADS = {
    "config1_16k_128_sh0": (16384, 0, 0.8, 128, 128, 1, "sds"),
    "config2_100k_512_sh0_b4": (100_000, 0, 0.8, 512, 512, 4, "sds"),
    "config3_300k_512_sh0_b4": (300_000, 0, 0.8, 512, 512, 4, "sds"),
    "config4_1m_256_sh3_b32": (1_000_000, 3, 0.5, 256, 256, 32, "mvdream"),
    "headline_1m_512_sh3": (1_000_000, 3, 0.5, 512, 512, 1, "mvdream"),
    "stress_4m_1024_sh3_b64": (4_000_000, 3, 0.5, 1024, 1024, 64, "mvdream"),
    # BASELINE.json configs[4] as written: the spacetime call shape (b200splat/spacetime.py) on the same scene size
    "stress_4m_1024_st_b64": (4_000_000, 0, 0.5, 1024, 1024, 64, "mvdream"),
}


def make_workload(name: str, views: int | None = None):
    P, deg, r0, H, W, B, rig = WORKLOADS[name]
    idx = list(WORKLOADS).index(name)
    B = views or B
    scene = make_scene(P, deg, r0, seed=1234 + idx)
    cams = (sds_cameras if rig == "sds" else mvdream_cameras)(B, H, W, seed=4321 + idx)
    return scene, cams
