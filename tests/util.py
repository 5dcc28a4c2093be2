"""Shared helpers for the parity tests: one scene -> oracle inputs and CUDA inputs."""
from __future__ import annotations

import torch

from b200splat import scenes
from oracle import torch_oracle as O


def oracle_settings(cam: scenes.Camera, sh_degree: int, bg=(1.0, 1.0, 1.0), scale_modifier=1.0) -> O.Settings:
    return O.Settings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                      torch.tensor(bg, dtype=torch.float32), scale_modifier, cam.viewmatrix, cam.projmatrix,
                      sh_degree, cam.campos, False, False)


def cuda_settings(s: O.Settings, device="cuda"):
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=s.image_height, image_width=s.image_width, tanfovx=s.tanfovx, tanfovy=s.tanfovy,
        bg=s.bg.to(device), scale_modifier=s.scale_modifier, viewmatrix=s.viewmatrix.to(device),
        projmatrix=s.projmatrix.to(device), sh_degree=s.sh_degree, campos=s.campos.to(device),
        prefiltered=False, debug=False)


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    """||g - g*||_inf / max(||g*||_inf, eps)  (SURVEY.md 8c tolerance definition)."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-20))


def borderline_pixels(pre, binned, s: O.Settings, out, tol=2e-5):
    """Pixels where some evaluated pair sits within ``tol`` (relative) of a hard cut-off
    (alpha = 1/255, T' = 1e-4, power = 0): there a 1-ulp difference in exp() legitimately flips
    a blend decision, so they are excluded from the max-abs image check (and counted)."""
    from oracle import spec
    d = O.derived_scalars(s)
    W, H, gx, gy = d["W"], d["H"], d["grid_x"], d["grid_y"]
    mask = torch.zeros(H, W, dtype=torch.bool)
    px, py = pre["px"], pre["py"]
    ca, cb, cc = pre["conic"]
    op = pre["opacity"]
    pl = binned["point_list"].long()
    rg = binned["ranges"]
    for ty in range(gy):
        for tx in range(gx):
            r0, r1 = int(rg[ty * gx + tx, 0]), int(rg[ty * gx + tx, 1])
            x0, y0 = tx * 16, ty * 16
            x1, y1 = min(x0 + 16, W), min(y0 + 16, H)
            nmax = int(out["n_contrib"][y0:y1, x0:x1].max())
            # one past the last contributor can be the terminating entry: look a bit further
            ids = pl[r0:min(r1, r0 + nmax + 64)]
            if ids.numel() == 0:
                continue
            ys, xs = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
            fx, fy = xs.reshape(-1, 1).float(), ys.reshape(-1, 1).float()
            dx, dy = px[ids][None] - fx, py[ids][None] - fy
            power = -0.5 * (ca[ids][None] * dx * dx + cc[ids][None] * dy * dy) - cb[ids][None] * dx * dy
            alpha = torch.clamp_max(op[ids][None] * torch.exp(power), spec.ALPHA_MAX)
            near_a = ((alpha - spec.ALPHA_MIN).abs() < tol * spec.ALPHA_MIN) | (power.abs() < 1e-6)
            valid = (power <= 0) & (alpha >= spec.ALPHA_MIN)
            T = torch.cumprod(torch.where(valid, 1 - alpha, torch.ones_like(alpha)), 1)
            near_t = valid & ((T - spec.T_MIN).abs() < 50 * tol * spec.T_MIN)
            bad = (near_a | near_t).any(1)
            mask[y0:y1, x0:x1] |= bad.reshape(y1 - y0, x1 - x0)
    return mask
