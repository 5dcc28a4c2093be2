"""Float64 dense autograd renderer: the self-pin of oracle/torch_oracle.py.

TEST INFRASTRUCTURE ONLY.  Written independently of the float32 oracle on purpose:
matrix formulation (J W Sigma W^T J^T via matmul), no tile lists (one global, stable
depth ordering; every pixel walks every Gaussian sequentially), plain autograd.  It keeps
only the *semantic* cut-offs of the path (SURVEY.md 8c "Self-pinning the oracle"):
tile-rectangle membership of the 3-sigma box, power > 0, alpha < 1/255, T' < 1e-4,
alpha = min(0.99, o G) straight-through, the FOV-clamp gradient convention, and the
means2D NDC dummy.  For <= a few hundred Gaussians at <= 64x64.
"""
from __future__ import annotations

import math

import torch

from . import spec

_C2 = spec.SH_C2
_C3 = spec.SH_C3


def _sh_basis(deg, d):
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    one = torch.ones_like(x)
    b = [spec.SH_C0 * one]
    if deg > 0:
        b += [-spec.SH_C1 * y, spec.SH_C1 * z, -spec.SH_C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [_C2[0] * xy, _C2[1] * yz, _C2[2] * (2 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy)]
    if deg > 2:
        b += [_C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
              _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy),
              _C3[5] * z * (xx - yy), _C3[6] * x * (xx - 3 * yy)]
    return torch.stack(b, -1)


def render_dense(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                 s, tile_quantised: bool = True):
    """All tensor arguments float64 (requires_grad as the caller wishes).  ``s`` is a
    torch_oracle.Settings whose tensors are cast to float64 here.  Returns color (3,H,W),
    depth (1,H,W), alpha (1,H,W), radii (P,) int."""
    f64 = torch.float64
    H, W = int(s.image_height), int(s.image_width)
    Vt = s.viewmatrix.to(f64)      # row-vector convention: p_view = [p,1] @ Vt
    Pt = s.projmatrix.to(f64)
    cam = s.campos.to(f64).reshape(3)
    bg = s.bg.to(f64).reshape(3)
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=f64)
    hom = torch.cat([means3D, ones], 1)
    pv = hom @ Vt
    ph = hom @ Pt
    pw = 1.0 / (ph[:, 3] + spec.PW_EPS)
    ndc = ph[:, :2] * pw[:, None]
    if means2D is not None:
        ndc = ndc + means2D[:, :2]
    tz = pv[:, 2]
    front = tz.detach() > spec.NEAR_CULL
    # Sigma3
    if cov3D_precomp is not None:
        c = cov3D_precomp
        Sig = torch.stack([torch.stack([c[:, 0], c[:, 1], c[:, 2]], -1),
                           torch.stack([c[:, 1], c[:, 3], c[:, 4]], -1),
                           torch.stack([c[:, 2], c[:, 4], c[:, 5]], -1)], 1)
    else:
        r, x, y, z = rotations.unbind(-1)
        R = torch.stack([
            torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
            torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
            torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], 1)
        L = R * (float(s.scale_modifier) * scales)[:, None, :]
        Sig = L @ L.transpose(1, 2)
    # EWA
    tanx, tany = float(s.tanfovx), float(s.tanfovy)
    fx, fy = W / (2 * tanx), H / (2 * tany)
    limx, limy = spec.FOV_CLAMP * tanx, spec.FOV_CLAMP * tany
    tzs = torch.where(front, tz, torch.ones_like(tz))      # keep culled rows finite
    rx, ry = (pv[:, 0] / tzs), (pv[:, 1] / tzs)
    inx = ((rx.detach() >= -limx) & (rx.detach() <= limx)).to(f64)
    iny = ((ry.detach() >= -limy) & (ry.detach() <= limy)).to(f64)
    txc = (rx.detach().clamp(-limx, limx) * tzs.detach()) + inx * (pv[:, 0] - pv[:, 0].detach())
    tyc = (ry.detach().clamp(-limy, limy) * tzs.detach()) + iny * (pv[:, 1] - pv[:, 1].detach())
    zero = torch.zeros_like(tzs)
    J = torch.stack([torch.stack([fx / tzs, zero, -fx * txc / (tzs * tzs)], -1),
                     torch.stack([zero, fy / tzs, -fy * tyc / (tzs * tzs)], -1)], 1)   # (P,2,3)
    Rw = Vt[:3, :3].t()                                                                # W2C rotation
    Mx = J @ Rw
    cov = Mx @ Sig @ Mx.transpose(1, 2)
    a = cov[:, 0, 0] + spec.DILATION
    b = cov[:, 0, 1]
    c_ = cov[:, 1, 1] + spec.DILATION
    det = a * c_ - b * b
    dets = torch.where(det.detach() != 0, det, torch.ones_like(det))
    conA, conB, conC = c_ / dets, -b / dets, a / dets
    mid = 0.5 * (a + c_)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, spec.LAMBDA_FLOOR))
    radius = torch.ceil(spec.RADIUS_SIGMAS * torch.sqrt(lam)).detach()
    px = ((ndc[:, 0] + 1) * W - 1) * 0.5
    py = ((ndc[:, 1] + 1) * H - 1) * 0.5
    gx = (W + spec.BLOCK_X - 1) // spec.BLOCK_X
    gy = (H + spec.BLOCK_Y - 1) // spec.BLOCK_Y
    tr = lambda t: torch.trunc(t.detach())
    rminx = tr((px - radius) / spec.BLOCK_X).clamp(0, gx)
    rminy = tr((py - radius) / spec.BLOCK_Y).clamp(0, gy)
    rmaxx = tr((px + radius + spec.BLOCK_X - 1) / spec.BLOCK_X).clamp(0, gx)
    rmaxy = tr((py + radius + spec.BLOCK_Y - 1) / spec.BLOCK_Y).clamp(0, gy)
    vis = front & (det.detach() != 0) & (((rmaxx - rminx) * (rmaxy - rminy)) > 0)
    # colour
    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        M = shs.shape[1]
        deg = max(0, min(int(s.sh_degree), int(math.isqrt(M)) - 1, 3))
        dirs = means3D - cam[None, :]
        dirs = dirs / dirs.norm(dim=1, keepdim=True)
        B = _sh_basis(deg, dirs)                       # (P, K)
        rgb = torch.einsum("pk,pkc->pc", B, shs[:, : B.shape[1], :]) + 0.5
        rgb = torch.clamp_min(rgb, 0.0)
    # global stable ordering by float32 depth bits (== the sort key's low word), then index
    depth32 = tz.detach().to(torch.float32)
    order = torch.argsort(depth32, stable=True)
    order = order[vis[order]]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=f64), torch.arange(W, dtype=f64), indexing="ij")
    tyix = torch.div(ys, spec.BLOCK_Y, rounding_mode="floor")
    txix = torch.div(xs, spec.BLOCK_X, rounding_mode="floor")
    T = torch.ones(H, W, dtype=f64)
    C = torch.zeros(3, H, W, dtype=f64)
    D = torch.zeros(H, W, dtype=f64)
    A = torch.zeros(H, W, dtype=f64)
    done = torch.zeros(H, W, dtype=torch.bool)
    op = opacities.reshape(-1)
    for g in order.tolist():
        member = (txix >= rminx[g]) & (txix < rmaxx[g]) & (tyix >= rminy[g]) & (tyix < rmaxy[g])
        if not tile_quantised:
            member = torch.ones_like(member)
        dx = px[g] - xs
        dy = py[g] - ys
        power = -0.5 * (conA[g] * dx * dx + conC[g] * dy * dy) - conB[g] * dx * dy
        raw = op[g] * torch.exp(power)
        alpha = raw + (raw.detach().clamp(max=spec.ALPHA_MAX) - raw.detach())
        ok = member & ~done & (power.detach() <= 0) & (alpha.detach() >= spec.ALPHA_MIN)
        testT = T * (1 - alpha)
        stop = ok & (testT.detach() < spec.T_MIN)
        done = done | stop
        use = ok & ~stop
        wgt = torch.where(use, alpha * T, torch.zeros_like(T))
        C = C + wgt[None] * rgb[g][:, None, None]
        D = D + wgt * tz[g]
        A = A + wgt
        T = torch.where(use, testT, T)
    color = C + T[None] * bg[:, None, None]
    radii = torch.where(vis, radius, torch.zeros_like(radius)).to(torch.int64)
    return color, D[None], A[None], radii
