"""PyTorch-CPU float32 restatement of the Gaussian-splatting rasterizer hot path.

TEST INFRASTRUCTURE ONLY (see oracle/spec.py header): the checker for the CUDA
kernels and the timed CPU baseline.  The product package never imports it.

PARITY UNPINNED UPSTREAM.  The reference (/root/reference) only *calls* this path
(renderer/diff_gaussian_rasterizer.py:98-131, geometry/gaussian_base.py:815-851);
the arithmetic lives in the un-vendored, unpinned pip packages
ashawkey/diff-gaussian-rasterization (forward.cu, backward.cu, rasterizer_impl.cu,
auxiliary.h) and DSaurus/simple-knn (README.md:17-20).  Each stage below restates
the published algorithm of those files and is
self-pinned by oracle/dense_f64.py (float64 autograd dense renderer) and the analytic
known-answer tests in tests/test_oracle_*.py.

Operation order is spelled out (one IEEE float32 op per torch call, left-to-right
association as written, no fused multiply-add) for every quantity that feeds the
bit-exact outputs: radii, tile rectangles, tiles_touched, depth bits of the sort key.
csrc/preprocess.cu follows the same order and is compiled with -fmad=false.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import spec


class Settings(NamedTuple):
    """Mirror of upstream ``GaussianRasterizationSettings``
    (constructed at renderer/diff_gaussian_rasterizer.py:83-96)."""

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


def _f32(v) -> np.float32:
    return np.float32(v)


def derived_scalars(s: Settings):
    """Host-side float32 scalars exactly as the C-ABI computes them (csrc/api.cu)."""
    W, H = int(s.image_width), int(s.image_height)
    tanx, tany = _f32(s.tanfovx), _f32(s.tanfovy)
    fx = _f32(W) / (_f32(2.0) * tanx)
    fy = _f32(H) / (_f32(2.0) * tany)
    limx = _f32(spec.FOV_CLAMP) * tanx
    limy = _f32(spec.FOV_CLAMP) * tany
    gx = (W + spec.BLOCK_X - 1) // spec.BLOCK_X
    gy = (H + spec.BLOCK_Y - 1) // spec.BLOCK_Y
    return dict(W=W, H=H, fx=float(fx), fy=float(fy), limx=float(limx), limy=float(limy),
                grid_x=gx, grid_y=gy, mod=float(_f32(s.scale_modifier)))


def effective_sh_degree(sh_degree: int, M: int) -> int:
    """Clamp the active degree to what an (P, M, 3) tensor holds.

    renderer/diff_gaussian_rasterizer_shading.py:178-187 passes a (P,1,3) "SH" tensor
    with sh_degree possibly > 0; reading coefficients 1.. would be out of bounds.
    """
    cap = int(math.isqrt(max(M, 1))) - 1
    return max(0, min(int(sh_degree), cap, 3))


# ----------------------------------------------------------------------------------------
# stage a3: preprocess (upstream forward.cu preprocessCUDA + computeCov3D/computeCov2D/
#           computeColorFromSH/in_frustum/getRect)
# ----------------------------------------------------------------------------------------

def cov3d_from_scale_rot(scales, rotations, mod: float):
    """Sigma = R S^2 R^T with R built from the *un-normalised* quaternion (r,x,y,z).

    Same Sigma as geometry/gaussian_base.py:99-134,234-238 (build_scaling_rotation, L L^T).
    Returns the 6 upper-triangular entries (00,01,02,11,12,22).
    """
    sx, sy, sz = (mod * scales[:, 0], mod * scales[:, 1], mod * scales[:, 2])
    r, x, y, z = rotations.unbind(-1)
    R00 = 1.0 - 2.0 * (y * y + z * z)
    R01 = 2.0 * (x * y - r * z)
    R02 = 2.0 * (x * z + r * y)
    R10 = 2.0 * (x * y + r * z)
    R11 = 1.0 - 2.0 * (x * x + z * z)
    R12 = 2.0 * (y * z - r * x)
    R20 = 2.0 * (x * z - r * y)
    R21 = 2.0 * (y * z + r * x)
    R22 = 1.0 - 2.0 * (x * x + y * y)
    L00, L01, L02 = R00 * sx, R01 * sy, R02 * sz
    L10, L11, L12 = R10 * sx, R11 * sy, R12 * sz
    L20, L21, L22 = R20 * sx, R21 * sy, R22 * sz
    c0 = L00 * L00 + L01 * L01 + L02 * L02
    c1 = L00 * L10 + L01 * L11 + L02 * L12
    c2 = L00 * L20 + L01 * L21 + L02 * L22
    c3 = L10 * L10 + L11 * L11 + L12 * L12
    c4 = L10 * L20 + L11 * L21 + L12 * L22
    c5 = L20 * L20 + L21 * L21 + L22 * L22
    return c0, c1, c2, c3, c4, c5


def eval_sh_rgb(deg: int, shs, dx, dy, dz):
    """SH -> RGB before the +0.5 / clamp; basis as geometry/sugar.py:775-818.

    shs: (N, M, 3).  dx,dy,dz: unit direction columns.  Returns 3 columns.
    """
    out = []
    for c in range(3):
        sh = shs[:, :, c]
        res = spec.SH_C0 * sh[:, 0]
        if deg > 0:
            res = res - spec.SH_C1 * dy * sh[:, 1] + spec.SH_C1 * dz * sh[:, 2] - spec.SH_C1 * dx * sh[:, 3]
            if deg > 1:
                xx, yy, zz = dx * dx, dy * dy, dz * dz
                xy, yz, xz = dx * dy, dy * dz, dx * dz
                res = (res
                       + spec.SH_C2[0] * xy * sh[:, 4]
                       + spec.SH_C2[1] * yz * sh[:, 5]
                       + spec.SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6]
                       + spec.SH_C2[3] * xz * sh[:, 7]
                       + spec.SH_C2[4] * (xx - yy) * sh[:, 8])
                if deg > 2:
                    res = (res
                           + spec.SH_C3[0] * dy * (3.0 * xx - yy) * sh[:, 9]
                           + spec.SH_C3[1] * xy * dz * sh[:, 10]
                           + spec.SH_C3[2] * dy * (4.0 * zz - xx - yy) * sh[:, 11]
                           + spec.SH_C3[3] * dz * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[:, 12]
                           + spec.SH_C3[4] * dx * (4.0 * zz - xx - yy) * sh[:, 13]
                           + spec.SH_C3[5] * dz * (xx - yy) * sh[:, 14]
                           + spec.SH_C3[6] * dx * (xx - 3.0 * yy) * sh[:, 15])
        out.append(res)
    return out


def _trunc_to_int(f: torch.Tensor) -> torch.Tensor:
    """C ``(int)f`` with CUDA's saturating semantics for out-of-range values."""
    g = torch.nan_to_num(f.detach(), nan=0.0, posinf=2.0e9, neginf=-2.0e9)
    return torch.clamp(g, -2147483520.0, 2147483520.0).to(torch.int64).to(torch.int32)


def _project(means3D, s: Settings, means2D_dummy=None):
    """View/projection transform and the near cull of ``in_frustum``."""
    V = s.viewmatrix.reshape(-1).to(torch.float32)
    Pm = s.projmatrix.reshape(-1).to(torch.float32)
    x, y, z = means3D.unbind(-1)
    tvx = V[0] * x + V[4] * y + V[8] * z + V[12]
    tvy = V[1] * x + V[5] * y + V[9] * z + V[13]
    tvz = V[2] * x + V[6] * y + V[10] * z + V[14]
    hx = Pm[0] * x + Pm[4] * y + Pm[8] * z + Pm[12]
    hy = Pm[1] * x + Pm[5] * y + Pm[9] * z + Pm[13]
    hw = Pm[3] * x + Pm[7] * y + Pm[11] * z + Pm[15]
    pw = 1.0 / (hw + spec.PW_EPS)
    ndcx = hx * pw
    ndcy = hy * pw
    if means2D_dummy is not None:
        # dL/dmeans2D is the loss gradient w.r.t. the NDC position.
        ndcx = ndcx + means2D_dummy[:, 0]
        ndcy = ndcy + means2D_dummy[:, 1]
    return tvx, tvy, tvz, ndcx, ndcy


def _cov2d(tvx, tvy, tvz, cov3d, s: Settings, d):
    """EWA projection of Sigma3 -> (a, b, c) with the +0.3 dilation (computeCov2D)."""
    V = s.viewmatrix.reshape(-1).to(torch.float32)
    limx, limy, fx, fy = d["limx"], d["limy"], d["fx"], d["fy"]
    txtz = tvx / tvz
    tytz = tvy / tvz
    cx = torch.clamp(txtz.detach(), -limx, limx) * tvz.detach()
    cy = torch.clamp(tytz.detach(), -limy, limy) * tvz.detach()
    # backward treats the clamped t.x,t.y as independent of t.z and
    # multiplies dL/dt.x by 0 when the clamp is active.
    xmul = ((txtz.detach() >= -limx) & (txtz.detach() <= limx)).to(tvx.dtype)
    ymul = ((tytz.detach() >= -limy) & (tytz.detach() <= limy)).to(tvx.dtype)
    tx = cx + xmul * (tvx - tvx.detach())
    ty = cy + ymul * (tvy - tvy.detach())
    J00 = fx / tvz
    J02 = -(fx * tx) / (tvz * tvz)
    J11 = fy / tvz
    J12 = -(fy * ty) / (tvz * tvz)
    # Rw[i][j] = W2C[i][j] = viewmatrix_flat[4*j + i]
    M00 = J00 * V[0] + J02 * V[2]
    M01 = J00 * V[4] + J02 * V[6]
    M02 = J00 * V[8] + J02 * V[10]
    M10 = J11 * V[1] + J12 * V[2]
    M11 = J11 * V[5] + J12 * V[6]
    M12 = J11 * V[9] + J12 * V[10]
    c0, c1, c2, c3, c4, c5 = cov3d
    N00 = M00 * c0 + M01 * c1 + M02 * c2
    N01 = M00 * c1 + M01 * c3 + M02 * c4
    N02 = M00 * c2 + M01 * c4 + M02 * c5
    N10 = M10 * c0 + M11 * c1 + M12 * c2
    N11 = M10 * c1 + M11 * c3 + M12 * c4
    N12 = M10 * c2 + M11 * c4 + M12 * c5
    a = N00 * M00 + N01 * M01 + N02 * M02 + spec.DILATION
    b = N00 * M10 + N01 * M11 + N02 * M12
    c = N10 * M10 + N11 * M11 + N12 * M12 + spec.DILATION
    return a, b, c


def _colors(means3D, shs, colors_precomp, s: Settings):
    if colors_precomp is not None:
        r, g, b = colors_precomp.unbind(-1)
        z = torch.zeros_like(r, dtype=torch.bool)
        return (r, g, b), (z, z, z)
    M = shs.shape[1]
    deg = effective_sh_degree(s.sh_degree, M)
    cam = s.campos.reshape(-1).to(torch.float32)
    x, y, z = means3D.unbind(-1)
    dx, dy, dz = x - cam[0], y - cam[1], z - cam[2]
    n = torch.sqrt(dx * dx + dy * dy + dz * dz)
    dx, dy, dz = dx / n, dy / n, dz / n
    raw = eval_sh_rgb(deg, shs, dx, dy, dz)
    raw = [c + 0.5 for c in raw]
    clamped = tuple((c.detach() < 0) for c in raw)
    rgb = tuple(torch.clamp_min(c, 0.0) for c in raw)
    return rgb, clamped


def preprocess(means3D, opacities, scales, rotations, cov3D_precomp, shs, colors_precomp,
               s: Settings, means2D_dummy=None, subset: Optional[torch.Tensor] = None):
    """Per-Gaussian stage.  When ``subset`` (index tensor of visible Gaussians) is given,
    only those rows are computed -- that is the differentiable path used by backward,
    which makes culled Gaussians receive exactly-zero gradients (Appendix A.1 item 5)."""
    d = derived_scalars(s)
    if subset is not None:
        pick = lambda t: None if t is None else t[subset]
        means3D, opacities, scales, rotations = map(pick, (means3D, opacities, scales, rotations))
        cov3D_precomp, shs, colors_precomp, means2D_dummy = map(
            pick, (cov3D_precomp, shs, colors_precomp, means2D_dummy))
    tvx, tvy, tvz, ndcx, ndcy = _project(means3D, s, means2D_dummy)
    in_front = tvz.detach() > spec.NEAR_CULL
    if cov3D_precomp is not None:
        cov3d = cov3D_precomp.unbind(-1)
    else:
        cov3d = cov3d_from_scale_rot(scales, rotations, d["mod"])
    a, b, c = _cov2d(tvx, tvy, tvz, cov3d, s, d)
    det = a * c - b * b
    det_inv = 1.0 / det
    conic_a, conic_b, conic_c = c * det_inv, -b * det_inv, a * det_inv
    mid = 0.5 * (a + c)
    disc = torch.sqrt(torch.clamp_min(mid * mid - det, spec.LAMBDA_FLOOR))
    lam = torch.maximum(mid + disc, mid - disc)
    radius_f = torch.ceil(spec.RADIUS_SIGMAS * torch.sqrt(lam))
    W, H = d["W"], d["H"]
    px = ((ndcx + 1.0) * float(W) - 1.0) * 0.5
    py = ((ndcy + 1.0) * float(H) - 1.0) * 0.5
    gx, gy = d["grid_x"], d["grid_y"]
    rmin_x = torch.clamp(_trunc_to_int((px - radius_f) / float(spec.BLOCK_X)), 0, gx)
    rmin_y = torch.clamp(_trunc_to_int((py - radius_f) / float(spec.BLOCK_Y)), 0, gy)
    rmax_x = torch.clamp(_trunc_to_int((px + radius_f + float(spec.BLOCK_X - 1)) / float(spec.BLOCK_X)), 0, gx)
    rmax_y = torch.clamp(_trunc_to_int((py + radius_f + float(spec.BLOCK_Y - 1)) / float(spec.BLOCK_Y)), 0, gy)
    area = (rmax_x - rmin_x) * (rmax_y - rmin_y)
    visible = in_front & (det.detach() != 0) & (area > 0)
    rgb, clamped = _colors(means3D, shs, colors_precomp, s)
    radii = torch.where(visible, _trunc_to_int(radius_f), torch.zeros_like(area))
    tiles = torch.where(visible, area, torch.zeros_like(area))
    return dict(
        visible=visible, radii=radii, tiles_touched=tiles,
        rect_min=torch.stack([rmin_x, rmin_y], -1), rect_max=torch.stack([rmax_x, rmax_y], -1),
        depth=tvz, px=px, py=py, conic=(conic_a, conic_b, conic_c), opacity=opacities.reshape(-1),
        rgb=rgb, clamped=clamped, cov3d=cov3d, cov2d=(a, b, c),
    )


# ----------------------------------------------------------------------------------------
# stages a4-a7: InclusiveSum, duplicateWithKeys, SortPairs, identifyTileRanges  (integer, bit-exact)
# ----------------------------------------------------------------------------------------

def bin_and_sort(pre, s: Settings):
    d = derived_scalars(s)
    gx, gy = d["grid_x"], d["grid_y"]
    tiles = pre["tiles_touched"].to(torch.int64)
    point_offsets = torch.cumsum(tiles, 0)
    R = int(point_offsets[-1]) if tiles.numel() else 0
    idx = torch.repeat_interleave(torch.arange(tiles.numel()), tiles)
    start = point_offsets - tiles
    k = torch.arange(R) - start[idx]
    rmin = pre["rect_min"].to(torch.int64)
    rmax = pre["rect_max"].to(torch.int64)
    w = (rmax[:, 0] - rmin[:, 0])[idx]
    ty = rmin[idx, 1] + torch.div(k, torch.clamp_min(w, 1), rounding_mode="floor")
    tx = rmin[idx, 0] + k % torch.clamp_min(w, 1)
    depth_bits = pre["depth"].detach().to(torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys = ((ty * gx + tx) << 32) | depth_bits[idx]
    keys_sorted, perm = torch.sort(keys, stable=True)
    point_list = idx[perm].to(torch.int32)
    tile_of = keys_sorted >> 32
    T = gx * gy
    ranges = torch.zeros(T, 2, dtype=torch.int32)
    if R > 0:
        t_ids = torch.arange(T)
        lo = torch.searchsorted(tile_of, t_ids, right=False)
        hi = torch.searchsorted(tile_of, t_ids, right=True)
        touched = hi > lo
        ranges[touched, 0] = lo[touched].to(torch.int32)
        ranges[touched, 1] = hi[touched].to(torch.int32)
    return dict(point_offsets=point_offsets.to(torch.int32), num_rendered=R, keys_unsorted=keys,
                vals_unsorted=idx.to(torch.int32), keys_sorted=keys_sorted, point_list=point_list,
                ranges=ranges, sort_bits=32 + spec.higher_msb(T))


# ----------------------------------------------------------------------------------------
# stage a8: renderCUDA forward (front-to-back alpha blending, early termination)
# ----------------------------------------------------------------------------------------

class Cuts(NamedTuple):
    """The three hard cut-offs of the blend loop.  The nominal values are upstream's; the parity tests also run the
    backward with slightly more permissive / stricter values to learn which gradient elements depend on a decision
    that a 1-ulp difference between exp() implementations can flip (tests/util.py::check_grads_bounded)."""
    alpha_min: float = spec.ALPHA_MIN
    t_min: float = spec.T_MIN
    power_max: float = 0.0


NOMINAL_CUTS = Cuts()


def _blend_tile(pixx, pixy, xy_x, xy_y, con_a, con_b, con_c, opac, rgb, depth, T0=None, cuts: Cuts = NOMINAL_CUTS):
    """Blend the entries (already in list order) over the given pixels.

    pix*: (Np,) float pixel coordinates.  Per-entry tensors: (G,).  rgb: (G,3).
    Returns per-pixel C (Np,3), D, A, T_final, last (1-based index of last blended entry,
    0 if none) and ``blended`` mask.  Masks are constants for autograd (A.1 item 6);
    ``alpha = min(0.99, o G)`` is straight-through (A.1 item 1).
    """
    dx = xy_x[None, :] - pixx[:, None]
    dy = xy_y[None, :] - pixy[:, None]
    power = -0.5 * (con_a[None, :] * dx * dx + con_c[None, :] * dy * dy) - con_b[None, :] * dx * dy
    G = torch.exp(power)
    raw = opac[None, :] * G
    alpha = raw + (torch.clamp_max(raw.detach(), spec.ALPHA_MAX) - raw.detach())
    valid = (power.detach() <= cuts.power_max) & (alpha.detach() >= cuts.alpha_min)
    a_eff = torch.where(valid, alpha.detach(), torch.zeros_like(alpha))
    first = torch.ones_like(a_eff[:, :1]) if T0 is None else T0[:, None].to(a_eff.dtype)
    # sequential product T_k = T_{k-1} (1 - alpha_k), carry-in first
    T_incl = torch.cumprod(torch.cat([first, 1.0 - a_eff], dim=1), dim=1)[:, 1:]
    term = valid & (T_incl < cuts.t_min)
    stopped = torch.cummax(term.to(torch.int8), dim=1).values.bool()
    blended = valid & ~stopped
    a_use = torch.where(blended, alpha, torch.zeros_like(alpha))
    T_all = torch.cumprod(torch.cat([first, 1.0 - a_use], dim=1), dim=1)
    T_before = T_all[:, :-1]
    T_fin = T_all[:, -1]
    w = a_use * T_before
    C = w @ rgb
    D = w @ depth
    A = w.sum(dim=1)
    Gn = blended.shape[1]
    pos = torch.arange(1, Gn + 1)[None, :]
    last = torch.where(blended, pos, torch.zeros_like(pos)).max(dim=1).values
    any_stop = stopped[:, -1]
    n_trav = torch.where(any_stop, Gn - stopped.sum(dim=1) + 1, torch.full_like(last, Gn))
    return C, D, A, T_fin, last, any_stop, n_trav


def render_forward(pre, binned, s: Settings, chunk: int = 1024, tiles=None):
    """Per-tile blend.  Returns color (3,H,W), depth (1,H,W), alpha (1,H,W), n_contrib (H,W) int32
    and counters n_eval_fwd / n_eval_bwd."""
    d = derived_scalars(s)
    W, H, gx, gy = d["W"], d["H"], d["grid_x"], d["grid_y"]
    bg = s.bg.reshape(-1).to(torch.float32)
    color = torch.zeros(3, H, W)
    depth_o = torch.zeros(1, H, W)
    alpha_o = torch.zeros(1, H, W)
    ncontrib = torch.zeros(H, W, dtype=torch.int32)
    n_eval_fwd = 0
    px, py = pre["px"].detach(), pre["py"].detach()
    ca, cb, cc = (t.detach() for t in pre["conic"])
    op = pre["opacity"].detach()
    rgb = torch.stack([t.detach() for t in pre["rgb"]], -1)
    dep = pre["depth"].detach()
    pl = binned["point_list"].to(torch.int64)
    ranges = binned["ranges"]
    # ``tiles``: optional subset of tile indices (bounded-sample CPU baseline); other tiles stay zero
    for tile in (range(gx * gy) if tiles is None else tiles):
        ty, tx = divmod(int(tile), gx)
        if True:
            r0, r1 = int(ranges[ty * gx + tx, 0]), int(ranges[ty * gx + tx, 1])
            x0, y0 = tx * spec.BLOCK_X, ty * spec.BLOCK_Y
            x1, y1 = min(x0 + spec.BLOCK_X, W), min(y0 + spec.BLOCK_Y, H)
            ys, xs = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
            pixx = xs.reshape(-1).to(torch.float32)
            pixy = ys.reshape(-1).to(torch.float32)
            Np = pixx.numel()
            T = torch.ones(Np)
            C = torch.zeros(Np, 3)
            D = torch.zeros(Np)
            A = torch.zeros(Np)
            last = torch.zeros(Np, dtype=torch.int64)
            done = torch.zeros(Np, dtype=torch.bool)
            base = 0
            while r0 + base < r1 and not bool(done.all()):
                ids = pl[r0 + base: min(r0 + base + chunk, r1)]
                act = ~done
                Cc, Dc, Ac, Tf, lc, stop, ntr = _blend_tile(
                    pixx[act], pixy[act], px[ids], py[ids], ca[ids], cb[ids], cc[ids], op[ids],
                    rgb[ids], dep[ids], T0=T[act])
                # entries traversed: up to and including the terminating one
                C[act] += Cc
                D[act] += Dc
                A[act] += Ac
                T[act] = Tf
                last[act] = torch.where(lc > 0, lc + base, last[act])
                n_eval_fwd += int(ntr.sum())
                nd = done.clone()
                nd[act] = stop
                done = nd
                base += ids.numel()
            color[:, y0:y1, x0:x1] = (C + T[:, None] * bg[None, :]).t().reshape(3, y1 - y0, x1 - x0)
            depth_o[0, y0:y1, x0:x1] = D.reshape(y1 - y0, x1 - x0)
            alpha_o[0, y0:y1, x0:x1] = A.reshape(y1 - y0, x1 - x0)
            ncontrib[y0:y1, x0:x1] = last.reshape(y1 - y0, x1 - x0).to(torch.int32)
    return dict(color=color, depth=depth_o, alpha=alpha_o, n_contrib=ncontrib,
                n_eval_fwd=n_eval_fwd, n_eval_bwd=int(ncontrib.sum()))


def rasterize_forward(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                      cov3D_precomp, s: Settings):
    """Whole forward; returns (outputs dict, pre, binned)."""
    with torch.no_grad():
        pre = preprocess(means3D, opacities, scales, rotations, cov3D_precomp, shs, colors_precomp, s)
        binned = bin_and_sort(pre, s)
        out = render_forward(pre, binned, s)
    return out, pre, binned


# ----------------------------------------------------------------------------------------
# stages a9-a10: backward.  Autograd of the differentiable restatement with the Appendix A.1
# conventions; tile by tile so memory stays bounded.
# ----------------------------------------------------------------------------------------

def rasterize_backward(inputs, s: Settings, pre, binned, fwd, dL_dcolor, dL_ddepth, dL_dalpha, tiles=None,
                       cuts: Cuts = NOMINAL_CUTS):
    """Returns dict of gradients for means3D, means2D, shs, colors_precomp, opacities, scales,
    rotations, cov3D_precomp (None where the input was None).  Also returns the per-Gaussian
    2D-stage gradients (dL/dxy_pix, dL/dconic, dL/dopacity, dL/drgb, dL/ddepth) that the CUDA
    render-backward kernel emits, for stage-level debugging.
    ``cuts``: non-nominal cut-offs re-decide every blend decision inside this call (the forward's n_contrib only
    bounds how far each tile's list is read, with a margin)."""
    means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp = inputs
    d = derived_scalars(s)
    W, H, gx, gy = d["W"], d["H"], d["grid_x"], d["grid_y"]
    P = means3D.shape[0]
    vis_idx = pre["visible"].nonzero().reshape(-1)
    bg = s.bg.reshape(-1).to(torch.float32)

    # differentiable preprocess on the visible subset
    leaf = lambda t: None if t is None else t.detach().clone().requires_grad_(True)
    m3, op_, sc_, ro_, c3_, sh_, cp_ = map(leaf, (means3D, opacities, scales, rotations,
                                                cov3D_precomp, shs, colors_precomp))
    m2 = torch.zeros(P, 3, requires_grad=True)
    with torch.enable_grad():
        sub = preprocess(m3, op_, sc_, ro_, c3_, sh_, cp_, s, means2D_dummy=m2, subset=vis_idx)
        two_d = [sub["px"], sub["py"], sub["conic"][0], sub["conic"][1], sub["conic"][2],
                 sub["opacity"], sub["rgb"][0], sub["rgb"][1], sub["rgb"][2], sub["depth"]]
    # map global Gaussian index -> row of the visible subset
    row_of = torch.full((P,), -1, dtype=torch.int64)
    row_of[vis_idx] = torch.arange(vis_idx.numel())
    acc = [torch.zeros(vis_idx.numel()) for _ in two_d]

    pl = binned["point_list"].to(torch.int64)
    ranges = binned["ranges"]
    ncontrib = fwd["n_contrib"]
    det2d = [t.detach() for t in two_d]
    for tile in (range(gx * gy) if tiles is None else tiles):
        ty, tx = divmod(int(tile), gx)
        if True:
            r0 = int(ranges[ty * gx + tx, 0])
            x0, y0 = tx * spec.BLOCK_X, ty * spec.BLOCK_Y
            x1, y1 = min(x0 + spec.BLOCK_X, W), min(y0 + spec.BLOCK_Y, H)
            nmax = int(ncontrib[y0:y1, x0:x1].max())
            if cuts is not NOMINAL_CUTS:   # a later termination / an extra blended entry may lie behind n_contrib
                nmax = min(int(ranges[ty * gx + tx, 1]) - r0, nmax + 64)
            if nmax == 0:
                continue
            ids = row_of[pl[r0:r0 + nmax]]
            ys, xs = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
            pixx = xs.reshape(-1).to(torch.float32)
            pixy = ys.reshape(-1).to(torch.float32)
            loc = [t[ids].clone().requires_grad_(True) for t in det2d]
            with torch.enable_grad():
                rgb = torch.stack(loc[6:9], -1)
                C, D, A, Tf, _, _, _ = _blend_tile(pixx, pixy, loc[0], loc[1], loc[2], loc[3], loc[4],
                                                loc[5], rgb, loc[9], cuts=cuts)
                col = C + Tf[:, None] * bg[None, :]
                gC = dL_dcolor[:, y0:y1, x0:x1].reshape(3, -1).t()
                gD = dL_ddepth[0, y0:y1, x0:x1].reshape(-1)
                gA = dL_dalpha[0, y0:y1, x0:x1].reshape(-1)
                obj = (col * gC).sum() + (D * gD).sum() + (A * gA).sum()
            grads = torch.autograd.grad(obj, loc, allow_unused=True)
            for k, g in enumerate(grads):
                if g is not None:
                    acc[k].index_add_(0, ids, g)
    with torch.enable_grad():
        obj = sum((t * a).sum() for t, a in zip(two_d, acc))
    wrt = [t for t in (m3, m2, sh_, cp_, op_, sc_, ro_, c3_) if t is not None]
    if vis_idx.numel() > 0:
        g = list(torch.autograd.grad(obj, wrt, allow_unused=True))
    else:
        g = [None] * len(wrt)
    g = [torch.zeros_like(t) if gi is None else gi for gi, t in zip(g, wrt)]
    it = iter(g)
    res = {}
    for name, t in (("means3D", m3), ("means2D", m2), ("shs", sh_), ("colors_precomp", cp_),
                    ("opacities", op_), ("scales", sc_), ("rotations", ro_), ("cov3D_precomp", c3_)):
        res[name] = next(it) if t is not None else None
    stage = {}
    full = lambda a: torch.zeros(P).index_add_(0, vis_idx, a)
    stage["dL_dpx"], stage["dL_dpy"] = full(acc[0]), full(acc[1])
    stage["dL_dconic"] = torch.stack([full(acc[2]), full(acc[3]), full(acc[4])], -1)
    stage["dL_dopacity"] = full(acc[5])
    stage["dL_drgb"] = torch.stack([full(acc[6]), full(acc[7]), full(acc[8])], -1)
    stage["dL_ddepth"] = full(acc[9])
    res["stage"] = stage
    return res


def mark_visible(means3D, s: Settings):
    """Upstream ``markVisible``/``checkFrustum``: the near-plane test only."""
    _, _, tvz, _, _ = _project(means3D, s)
    return tvz > spec.NEAR_CULL
