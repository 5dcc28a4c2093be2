"""The spacetime renderer's call shape (BASELINE.json configs[4]; renderer/diff_gaussian_rasterizer_st.py:135-150,
geometry/spacetime_gaussian.py:291-334): every view of a step has its own timestamp, so its own NON-LEAF ``means3D``
(cubic B-spline of the per-Gaussian control knots, basis of geometry/spline_utils.py:109-119) and ``rotations``
(``normalize(q + dq[frame])``), with ``colors_precomp`` instead of SH, ``sigmoid`` opacities and ``exp`` scales; the
rasterizer is called once per view (the timed parameters differ per view) and autograd carries the gradients back to
the knots.

Host-side PyTorch only (parameter plumbing above the operator); the reference's rotation spline lives in pypose, which
is not in the image -- the per-frame quaternion offset of the non-spline branch (geometry/spacetime_gaussian.py:326)
stands in for it.  Used by bench.py's ``stress_4m_1024_st_b64`` workload and tests/test_spacetime_gpu.py.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn.functional as F


class SpacetimeParams(NamedTuple):
    knots: torch.Tensor       # (P, K, 3) control knots of the position spline
    rotation: torch.Tensor    # (P, 4) raw quaternion (r, x, y, z)
    omega: torch.Tensor       # (Fr, P, 4) per-frame quaternion offset
    colors: torch.Tensor      # (P, 3) precomputed RGB
    opacity: torch.Tensor     # (P, 1) raw (pre-sigmoid)
    scaling: torch.Tensor     # (P, 3) raw (pre-exp)


def bspline_position(knots: torch.Tensor, t: float) -> torch.Tensor:
    """Uniform cubic B-spline through K control knots at t in [0, 1]: segment i = floor(t (K - 3)), local u, blend of
    knots i..i+3 with the coefficients of geometry/spline_utils.py:109-119 (``coeffs_t``)."""
    K = knots.shape[1]
    x = min(max(float(t), 0.0), 1.0) * (K - 3)
    i = min(int(x), K - 4)
    u = x - i
    uu, uuu, oos = u * u, u * u * u, 1.0 / 6.0
    c = knots.new_tensor([oos - 0.5 * u + 0.5 * uu - oos * uuu, 4.0 * oos - uu + 0.5 * uuu,
                          oos + 0.5 * u + 0.5 * uu - 0.5 * uuu, oos * uuu])
    # one slice of the four knots (its backward is ONE dense gradient tensor, not four)
    return (knots[:, i:i + 4] * c[None, :, None]).sum(dim=1)


def timed_all(p: SpacetimeParams, t: float, frame: int):
    """-> (means3D, scales, rotations, opacity, colors_precomp) as ``get_timed_all`` returns them."""
    means3D = bspline_position(p.knots, t)
    rotations = F.normalize(p.rotation + p.omega[frame], dim=-1)
    return means3D, torch.exp(p.scaling), rotations, torch.sigmoid(p.opacity), p.colors


def make_params(scene, frames: int = 12, knots: int = 12, seed: int = 0, motion: float = 0.05) -> SpacetimeParams:
    """Spacetime parameters around a static synthetic scene (b200splat.scenes.make_scene): the knots wander around the
    static position, the raw activations invert the scene's activated values."""
    g = torch.Generator().manual_seed(seed)
    P = scene.means3D.shape[0]
    drift = torch.cumsum(torch.randn(P, knots, 3, generator=g) * (motion / knots ** 0.5), dim=1)
    kn = scene.means3D[:, None, :] + drift - drift.mean(dim=1, keepdim=True)
    omega = torch.randn(frames, P, 4, generator=g) * 0.05
    colors = torch.rand(P, 3, generator=g)
    return SpacetimeParams(kn.contiguous(), scene.rotations.clone(), omega.contiguous(), colors,
                           torch.logit(scene.opacities.clamp(1e-4, 1 - 1e-4)), torch.log(scene.scales))
