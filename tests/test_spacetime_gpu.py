"""BASELINE.json configs[4] call shape at scale (renderer/diff_gaussian_rasterizer_st.py:135-150): per-view non-leaf
``means3D`` (cubic B-spline of 12 knots) and ``rotations`` (normalize(q + dq[frame])), ``colors_precomp``, activated
opacity / scale; the gradient must reach the knots.  CUDA: drop-in ``GaussianRasterizer`` + autograd.  Oracle: the CPU
rasterizer oracle per view, chained through the same spline by autograd."""
import pytest
import torch

from b200splat import scenes, spacetime
from oracle import torch_oracle as O
from util import check_grads_bounded, cuda_settings, cut_variants, oracle_settings

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_gradient_reaches_the_spline_knots_at_100k_gaussians():
    from diff_gaussian_rasterization import GaussianRasterizer
    P, H, W, V = 100_000, 256, 256, 2
    scene = scenes.make_scene(P, 0, 0.5, seed=901)
    cams = scenes.mvdream_cameras(V, H, W, seed=902)
    params = spacetime.make_params(scene, seed=903)
    times = [(0.23, 3), (0.81, 9)]                      # (timestamp, frame) of the two views
    pgs = [scenes.pixel_grads(H, W, 904 + v) for v in range(V)]
    names = spacetime.SpacetimeParams._fields

    # ---- CUDA: the reference's loop -- one rasterizer call per view on that view's timed parameters ----------------
    dev = "cuda"
    leaves = spacetime.SpacetimeParams(*[t.to(dev).clone().requires_grad_(True) for t in params])
    loss = 0.0
    for v in range(V):
        m3, scl, rot, opa, col = spacetime.timed_all(leaves, *times[v])
        assert not m3.is_leaf and not rot.is_leaf
        m2 = torch.zeros_like(m3, requires_grad=True) + 0
        rs = cuda_settings(oracle_settings(cams[v], 0), dev)
        c, r, d, a = GaussianRasterizer(raster_settings=rs)(means3D=m3, means2D=m2, shs=None, colors_precomp=col,
                                                            opacities=opa, scales=scl, rotations=rot, cov3D_precomp=None)
        loss = loss + sum((x * g.to(dev)).sum() for x, g in zip((c, d, a), pgs[v]))
    loss.backward()
    got = {n: t.grad for n, t in zip(names, leaves)}
    assert float(got["knots"].abs().max()) > 0

    # ---- oracle: per view forward + backward (nominal / permissive / strict cut-offs), chained through the spline ----
    perm, strict = cut_variants()
    want = {}
    for tag, cuts in (("nominal", O.NOMINAL_CUTS), ("perm", perm), ("strict", strict)):
        cl = spacetime.SpacetimeParams(*[t.clone().requires_grad_(True) for t in params])
        outs, gs = [], []
        for v in range(V):
            m3, scl, rot, opa, col = spacetime.timed_all(cl, *times[v])
            s = oracle_settings(cams[v], 0)
            inputs = (m3.detach(), None, None, col.detach(), opa.detach(), scl.detach(), rot.detach(), None)
            out, pre, binned = O.rasterize_forward(*inputs, s)
            g = O.rasterize_backward(inputs, s, pre, binned, out, *pgs[v], cuts=cuts)
            outs += [m3, scl, rot, opa, col]
            gs += [g["means3D"], g["scales"], g["rotations"], g["opacities"], g["colors_precomp"]]
        torch.autograd.backward(outs, gs)
        want[tag] = {n: t.grad for n, t in zip(names, cl)}
    rep = check_grads_bounded(got, want["nominal"], want["perm"], want["strict"], 1e-3)
    assert set(rep) == set(names)
