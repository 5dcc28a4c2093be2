"""Checker-side helpers shared by tests/ and __graft_entry__.smoke(): which pixels sit on a hard cut-off of the
blend loop, how far a flipped decision may move them, and the image comparison that uses both.

TEST INFRASTRUCTURE ONLY (see oracle/spec.py header): never imported by the product package.
"""
from __future__ import annotations

import torch

from . import torch_oracle as O


def borderline_bounds(pre, binned, s: O.Settings, out, tol=2e-5):
    """Pixels where some evaluated pair sits within ``tol`` (relative) of a hard cut-off
    (alpha = 1/255, T' = 1e-4, power = 0): there a 1-ulp difference in exp() legitimately flips
    a blend decision.  Such a pixel is NOT exempt from the image check: its error must stay within what
    the flipped decision can move.  Flipping entry e (alpha_e, transmittance T_e before it, colour c_e)
    changes the pixel by alpha_e (T_e c_e - Rest), |Rest| <= T_e cmax (everything behind e incl. the
    background), so

        |delta colour| <= alpha_e T_e (|c_e| + cmax),  |delta depth| <= alpha_e T_e (|z_e| + zmax),
        |delta alpha|  <= alpha_e T_e

    and a flipped termination test (T' ~ 1e-4) additionally lets the tail through: + T' cmax.
    Returns dict(mask (H,W) bool, color / depth / alpha (H,W) fp32 bounds, zero outside the mask).

    A flipped alpha ~ 1/255 entry scales T of everything behind it by (1 - 1/255); a termination test
    further down the same pixel's list that sits within that factor of T_min may flip as a consequence,
    so behind a flagged entry the window of the termination test is widened to 1.1/255 (relative)."""
    from oracle import spec
    d = O.derived_scalars(s)
    W, H, gx, gy = d["W"], d["H"], d["grid_x"], d["grid_y"]
    mask = torch.zeros(H, W, dtype=torch.bool)
    b_col = torch.zeros(H, W)
    b_dep = torch.zeros(H, W)
    b_alp = torch.zeros(H, W)
    px, py = pre["px"].detach(), pre["py"].detach()
    ca, cb, cc = (t.detach() for t in pre["conic"])
    op = pre["opacity"].detach()
    cn = torch.stack([t.detach().abs() for t in pre["rgb"]], -1).max(-1).values
    zz = pre["depth"].detach().abs()
    bgmax = float(s.bg.abs().max())
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
            a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
            T = torch.cumprod(1 - a_eff, 1)                       # transmittance behind entry e
            T_before = torch.cat([torch.ones_like(T[:, :1]), T[:, :-1]], 1)
            # entries strictly behind the one that terminates the pixel are never evaluated
            term = valid & (T < spec.T_MIN)
            stopped_before = torch.cat([torch.zeros_like(term[:, :1]),
                                        torch.cummax(term.to(torch.int8), 1).values.bool()[:, :-1]], 1)
            near_a = near_a & ~stopped_before
            after_flag = torch.cat([torch.zeros_like(near_a[:, :1]),
                                    torch.cummax(near_a.to(torch.int8), 1).values.bool()[:, :-1]], 1)
            rel_t = (T - spec.T_MIN).abs() / spec.T_MIN
            near_t = valid & ~stopped_before & ((rel_t < 50 * tol) | (after_flag & (rel_t < 1.1 / 255.0)))
            flag = near_a | near_t
            bad = flag.any(1)
            if not bool(bad.any()):
                continue
            cmax = max(bgmax, float(cn[ids].max()))
            zmax = float(zz[ids].max())
            a_flip = torch.where(near_a & ~valid, torch.clamp_min(alpha, spec.ALPHA_MIN), alpha)   # alpha if it were blended
            w = a_flip * T_before
            tail = torch.where(near_t, T, torch.zeros_like(T))
            f = flag.float()
            bc = (f * (w * (cn[ids][None] + cmax) + tail * cmax)).sum(1)
            bd = (f * (w * (zz[ids][None] + zmax) + tail * zmax)).sum(1)
            ba = (f * (w + tail)).sum(1)
            sh = (y1 - y0, x1 - x0)
            mask[y0:y1, x0:x1] |= bad.reshape(sh)
            b_col[y0:y1, x0:x1] = torch.where(bad, bc, torch.zeros_like(bc)).reshape(sh)
            b_dep[y0:y1, x0:x1] = torch.where(bad, bd, torch.zeros_like(bd)).reshape(sh)
            b_alp[y0:y1, x0:x1] = torch.where(bad, ba, torch.zeros_like(ba)).reshape(sh)
    return dict(mask=mask, color=b_col, depth=b_dep, alpha=b_alp)


def borderline_pixels(pre, binned, s: O.Settings, out, tol=2e-5):
    """The mask of ``borderline_bounds`` (pixels whose error bar is wider than the plain tolerance)."""
    return borderline_bounds(pre, binned, s, out, tol)["mask"]


def check_images(cu, out, bounds, tol, max_frac=0.01):
    """Image / depth / alpha within ``tol`` everywhere except the borderline pixels, which must stay within
    ``tol`` + the bound of their flipped blend decision; n_contrib (when given) equal outside the mask.
    ``cu``: dict of CUDA outputs (color, depth, alpha[, n_contrib]); ``out``: the oracle's."""
    mask = bounds["mask"]
    frac = float(mask.float().mean())
    assert frac < max_frac, f"too many borderline pixels: {frac}"
    worst = {}
    for name in ("color", "depth", "alpha"):
        err = (cu[name].detach().cpu().float() - out[name]).abs()
        lim = tol + bounds[name][None]
        over = err - lim
        worst[name] = float(err.max())
        assert float(over.max()) <= 0.0, (f"{name}: error {float(err.flatten()[over.argmax()])} exceeds its bound "
                                          f"{float(lim.expand_as(err).flatten()[over.argmax()])} "
                                          f"(borderline pixel: {bool(mask.flatten()[over.argmax() % mask.numel()])})")
    if cu.get("n_contrib") is not None:
        nc = cu["n_contrib"].cpu().to(torch.int64).reshape(mask.shape)
        ref = out["n_contrib"].to(torch.int64)
        neq = (nc != ref) & ~mask
        assert int(neq.sum()) == 0, f"n_contrib differs on {int(neq.sum())} non-borderline pixels"
    return frac, worst


def cut_variants(tol=2e-5):
    """(permissive, strict) cut-offs around the nominal ones: every blend decision that lies within ``tol``
    (relative; 50 tol for the transmittance test, 1e-6 absolute for power <= 0 -- the windows of
    ``borderline_bounds``) of a cut-off is taken one way by the first and the other way by the second."""
    from . import spec
    perm = O.Cuts(spec.ALPHA_MIN * (1 - tol), spec.T_MIN * (1 - 50 * tol), 1e-6)
    strict = O.Cuts(spec.ALPHA_MIN * (1 + tol), spec.T_MIN * (1 + 50 * tol), -1e-6)
    return perm, strict


def check_grads_bounded(got, nominal, permissive, strict, tol):
    """Gradient parity that knows about borderline blend decisions.  For every gradient element

        |got - nominal| <= tol * ||nominal||_inf + |permissive - nominal| + |strict - nominal|

    where the three oracle results differ only in how the decisions that sit on a hard cut-off were taken
    (``cut_variants``).  An element no such decision touches has all three equal and gets the plain ``tol`` bar; an
    element a flipped pair feeds may differ by what the flips move (each flagged decision is in exactly one of the
    two differences, so any combination of flips stays inside the sum).  Returns {name: (plain rel err, number of
    elements that needed the wider bar)}."""
    rep = {}
    for k, ref in nominal.items():
        if k == "stage" or ref is None:
            continue
        g = got[k]
        assert g is not None, f"missing grad {k}"
        g = g.detach().cpu().double().reshape(ref.shape)
        ref64 = ref.double()
        scale = float(ref64.abs().max().clamp_min(1e-20))
        slack = (permissive[k].double() - ref64).abs() + (strict[k].double() - ref64).abs()
        err = (g - ref64).abs()
        over = err - (tol * scale + slack)
        assert float(over.max()) <= 0.0, (f"grad {k}: error {float(err.flatten()[over.argmax()]) / scale:.3e} (relative to "
                                          f"the tensor's max) exceeds {tol} + the borderline slack "
                                          f"{float(slack.flatten()[over.argmax()]) / scale:.3e}")
        rep[k] = (float(err.max()) / scale, int((err > tol * scale).sum()))
    return rep
