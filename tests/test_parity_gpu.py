"""GPU parity: CUDA path (through the drop-in Python surface -> C ABI) vs the CPU oracle.

Bars (BASELINE.json north_star): radii, tiles_touched, sorted keys, tile ranges bit-exact;
image / depth / alpha max-abs <= 1e-4; parameter gradients <= 1e-3 relative.
"""
import pytest
import torch

from b200splat import scenes
from oracle import torch_oracle as O
from oracle.knn import dist2_oracle
from util import (borderline_bounds, check_grads_bounded, check_images, cuda_settings, cut_variants, oracle_settings,
                  rel_err)

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_TOL = 1e-3


def _run_cuda(sc, s, *, colors_precomp=None, cov3D=None, shs=None, grads=None, dev="cuda"):
    from b200splat import ops
    from diff_gaussian_rasterization import GaussianRasterizer
    rs = cuda_settings(s, dev)
    req = lambda t: None if t is None else t.to(dev).clone().requires_grad_(True)
    m3 = req(sc.means3D)
    m2 = torch.zeros_like(m3, requires_grad=True)
    m2n = m2 + 0
    m2n.retain_grad()
    op = req(sc.opacities)
    sh = req(shs if shs is not None else (None if colors_precomp is not None else sc.shs))
    cp = req(colors_precomp)
    scl, rot, c3 = (None, None, req(cov3D)) if cov3D is not None else (req(sc.scales), req(sc.rotations), None)
    color, radii, depth, alpha = GaussianRasterizer(raster_settings=rs)(
        means3D=m3, means2D=m2n, shs=sh, colors_precomp=cp, opacities=op, scales=scl, rotations=rot,
        cov3D_precomp=c3)
    res = dict(color=color, radii=radii, depth=depth, alpha=alpha)
    if grads is not None:
        gc, gd, ga = (g.to(dev) for g in grads)
        loss = (color * gc).sum() + (depth * gd).sum() + (alpha * ga).sum()
        loss.backward()
        res["grads"] = dict(means3D=m3.grad, means2D=m2n.grad, opacities=op.grad,
                            shs=None if sh is None else sh.grad, colors_precomp=None if cp is None else cp.grad,
                            scales=None if scl is None else scl.grad, rotations=None if rot is None else rot.grad,
                            cov3D_precomp=None if c3 is None else c3.grad)
    return res


def _run_oracle(sc, s, *, colors_precomp=None, cov3D=None, shs=None, grads=None):
    shs = shs if shs is not None else (None if colors_precomp is not None else sc.shs)
    scl, rot = (None, None) if cov3D is not None else (sc.scales, sc.rotations)
    out, pre, binned = O.rasterize_forward(sc.means3D, None, shs, colors_precomp, sc.opacities, scl, rot, cov3D, s)
    res = dict(out=out, pre=pre, binned=binned)
    if grads is not None:
        inputs = (sc.means3D, None, shs, colors_precomp, sc.opacities, scl, rot, cov3D)
        res["grads"] = O.rasterize_backward(inputs, s, pre, binned, out, *grads)
        # the same backward with every borderline blend decision taken the permissive / the strict way
        perm, strict = cut_variants()
        res["grads_perm"] = O.rasterize_backward(inputs, s, pre, binned, out, *grads, cuts=perm)
        res["grads_strict"] = O.rasterize_backward(inputs, s, pre, binned, out, *grads, cuts=strict)
    return res


def _check_forward(cu, orc, s, allow_borderline=True):
    """radii bit-exact; image / depth / alpha within IMG_TOL, except that a pixel with a blend decision on a hard
    cut-off may deviate by what flipping that one decision moves (tests/util.py::borderline_bounds) -- and no more;
    n_contrib (when the caller has it) equal outside those pixels."""
    out, pre, binned = orc["out"], orc["pre"], orc["binned"]
    assert torch.equal(cu["radii"].cpu(), pre["radii"]), "radii not bit-exact"
    bounds = borderline_bounds(pre, binned, s, out)
    if not allow_borderline:
        bounds = {k: torch.zeros_like(v) for k, v in bounds.items()}
    return check_images(cu, out, bounds, IMG_TOL)


def _check_grads(cu, orc, tol=GRAD_TOL):
    """<= 1e-3 relative for every gradient element no borderline blend decision feeds; an element such a decision
    feeds may differ by what flipping it moves (oracle/checks.py::check_grads_bounded)."""
    return check_grads_bounded(cu["grads"], orc["grads"], orc["grads_perm"], orc["grads_strict"], tol)


def _scene(P, deg, H, W, seed, r0=0.8, scale_mul=1.0):
    sc = scenes.make_scene(P, deg, r0, seed=seed)
    if scale_mul != 1.0:
        sc = sc._replace(scales=sc.scales * scale_mul)
    cam = scenes.sds_cameras(1, H, W, seed=seed + 100)[0]
    return sc, cam


def test_sort_pairs_matches_stable_sort():
    from b200splat import ops
    g = torch.Generator().manual_seed(0)
    for n, bits in ((1, 64), (37, 64), (4096, 43), (4097, 39), (100_003, 45), (1_000_000, 43), (50_000, 8)):
        keys = torch.randint(0, 2 ** 62, (n,), generator=g, dtype=torch.int64)
        if bits < 64:
            keys = keys & ((1 << bits) - 1)
        # many duplicates in the upper bits (tile ids) and in the depth exponent byte
        keys = keys & ~(0xF << 28)
        vals = torch.arange(n, dtype=torch.int32)
        ks, vs = ops.sort_pairs(keys.cuda(), vals.cuda(), end_bit=bits)
        rk, perm = torch.sort(keys, stable=True)
        assert torch.equal(ks.cpu(), rk), (n, bits)
        assert torch.equal(vs.cpu(), vals[perm]), (n, bits)


def test_inclusive_scan():
    from b200splat import ops
    g = torch.Generator().manual_seed(1)
    for n in (1, 7, 2048, 2049, 1_000_003):
        x = torch.randint(0, 50, (n,), generator=g, dtype=torch.int32)
        y = ops.inclusive_scan_u32(x.cuda())
        assert torch.equal(y.cpu(), torch.cumsum(x, 0).to(torch.int32)), n


@pytest.mark.parametrize("P,deg,H,W,seed", [
    (16384, 0, 128, 128, 1235),      # BASELINE configs[0] shape
    (4000, 3, 96, 80, 7),            # SH degree 3, non-square, W not a multiple of 16
    (3000, 1, 50, 70, 8),
    (3000, 2, 64, 64, 9),
])
@pytest.mark.parametrize("staging", ["ldgsts", "bulk"])
def test_forward_backward_parity(P, deg, H, W, seed, staging):
    """Both staging variants of the render kernels: per-entry cp.async (LDGSTS) and cp.async.bulk (UBLKCP,
    complete_tx on the stage's mbarrier)."""
    from b200splat import _lib
    sc, cam = _scene(P, deg, H, W, seed)
    _lib.set_staging(staging)
    try:
        _parity_case(sc, cam, deg, H, W, seed)
    finally:
        _lib.set_staging(None)


def _parity_case(sc, cam, deg, H, W, seed):
    s = oracle_settings(cam, deg)
    grads = scenes.pixel_grads(H, W, seed + 1)
    orc = _run_oracle(sc, s, grads=grads)
    cu = _run_cuda(sc, s, grads=grads)
    _check_forward(cu, orc, s)
    _check_grads(cu, orc)


@pytest.mark.parametrize("P,H,W", [(16384, 128, 128),     # 64 tiles: one wide partition pass (tile id <= 10 bits)
                                    (6000, 528, 528)])      # 1089 tiles: two 8-bit passes of the generic kernel
def test_intermediates_bit_exact(P, H, W):
    """tiles_touched, point_offsets, depth bits, sorted keys, point list and tile ranges."""
    from b200splat import ops
    sc, cam = _scene(P, 0, H, W, 1235)
    s = oracle_settings(cam, 0)
    orc = _run_oracle(sc, s)
    camc = ops.make_cam(cuda_settings(s), "cuda")
    d = lambda t: t.cuda().contiguous()
    color, radii, depth, alpha, st = ops.forward(camc, d(sc.means3D), d(sc.shs), None, d(sc.opacities),
                                                 d(sc.scales), d(sc.rotations), None)
    v = {k: t.cpu() for k, t in ops.forward_views(camc, st).items()}
    pre, binned = orc["pre"], orc["binned"]
    assert st.num_rendered == binned["num_rendered"]
    assert torch.equal(radii.cpu(), pre["radii"])
    assert torch.equal(v["tiles_touched"], pre["tiles_touched"])
    vis = pre["visible"]
    # the Gaussians are depth-sorted first (stable: ties keep index order) and the offsets are the scan in THAT order.
    # Culled Gaussians emit nothing, so the sort is free to leave them anywhere (it sorts key - min on 24 bits when the
    # depth range allows); the VISIBLE ones must come in exactly the oracle's (depth bits, index) order.
    dbits = pre["depth"].contiguous().view(torch.int32).long() & 0xFFFFFFFF
    order_cuda = (v["gaussian_order"] & 0xFFFFFFFF).long()
    assert torch.equal(torch.sort(order_cuda).values, torch.arange(P)), "the depth order must be a permutation"
    vis_idx = vis.nonzero().reshape(-1)
    want_vis = vis_idx[torch.sort(dbits[vis_idx], stable=True).indices]
    assert torch.equal(order_cuda[vis[order_cuda]], want_vis)
    assert torch.equal(v["point_offsets"].long(), torch.cumsum(pre["tiles_touched"].long()[order_cuda], 0))
    assert int(v["point_offsets"][-1]) == int(binned["point_offsets"][-1])
    assert torch.equal(v["depths"][vis].view(torch.int32), pre["depth"][vis].contiguous().view(torch.int32))
    assert torch.equal(v["keys_sorted"], binned["keys_sorted"])
    assert torch.equal(v["point_list"], binned["point_list"])
    assert torch.equal(v["ranges"], binned["ranges"])
    check_images(dict(color=color, depth=depth, alpha=alpha, n_contrib=v["n_contrib"]), orc["out"],
                 borderline_bounds(pre, binned, s, orc["out"]), IMG_TOL)


def test_colors_precomp_cov3d_precomp_and_bg():
    sc, cam = _scene(3000, 0, 64, 64, 21)
    s = oracle_settings(cam, 0, bg=(0.0, 0.0, 0.0))
    g = torch.Generator().manual_seed(5)
    cp = torch.rand(3000, 3, generator=g)
    c0, c1, c2, c3, c4, c5 = O.cov3d_from_scale_rot(sc.scales, sc.rotations, 1.0)
    cov = torch.stack([c0, c1, c2, c3, c4, c5], -1).contiguous()
    grads = scenes.pixel_grads(64, 64, 22)
    orc = _run_oracle(sc, s, colors_precomp=cp, cov3D=cov, grads=grads)
    cu = _run_cuda(sc, s, colors_precomp=cp, cov3D=cov, grads=grads)
    _check_forward(cu, orc, s)
    _check_grads(cu, orc)


def test_degenerate_sh_tensor_with_high_degree():
    """(P,1,3) 'SH' with sh_degree 3 (renderer/diff_gaussian_rasterizer_shading.py:178-187)."""
    sc, cam = _scene(2000, 0, 64, 64, 31)
    s = oracle_settings(cam, 3)
    grads = scenes.pixel_grads(64, 64, 32)
    orc = _run_oracle(sc, s, shs=sc.shs[:, :1].contiguous(), grads=grads)
    cu = _run_cuda(sc, s, shs=sc.shs[:, :1].contiguous(), grads=grads)
    _check_forward(cu, orc, s)
    _check_grads(cu, orc)


def test_offscreen_and_behind_camera():
    """FOV clamp gradient convention + near cull + off-screen rectangles."""
    sc, cam = _scene(3000, 1, 64, 64, 41, r0=3.0, scale_mul=1.0)
    s = oracle_settings(cam, 1)
    grads = scenes.pixel_grads(64, 64, 42)
    orc = _run_oracle(sc, s, grads=grads)
    assert int((orc["pre"]["radii"] == 0).sum()) > 0
    cu = _run_cuda(sc, s, grads=grads)
    _check_forward(cu, orc, s)
    _check_grads(cu, orc)
    culled = (orc["pre"]["radii"] == 0)
    for k, g in cu["grads"].items():
        if g is not None:
            assert float(g.cpu()[culled].abs().max()) == 0.0, f"culled Gaussians must get zero {k} grad"


def test_empty_and_double_backward():
    from diff_gaussian_rasterization import GaussianRasterizer
    sc, cam = _scene(1000, 0, 32, 32, 51)
    s = oracle_settings(cam, 0)
    rs = cuda_settings(s)
    z = lambda *sh: torch.zeros(*sh, device="cuda")
    color, radii, depth, alpha = GaussianRasterizer(raster_settings=rs)(
        means3D=z(0, 3), means2D=z(0, 3), shs=z(0, 1, 3), colors_precomp=None, opacities=z(0, 1),
        scales=z(0, 3), rotations=z(0, 4), cov3D_precomp=None)
    assert color.shape == (3, 32, 32) and float((color - 1).abs().max()) == 0 and radii.numel() == 0
    # two backward passes over one forward (system/gaussian_splatting.py:129,137-138)
    m3 = sc.means3D.cuda().requires_grad_(True)
    m2 = torch.zeros_like(m3, requires_grad=True)
    out = GaussianRasterizer(raster_settings=rs)(
        means3D=m3, means2D=m2, shs=sc.shs.cuda(), colors_precomp=None, opacities=sc.opacities.cuda(),
        scales=sc.scales.cuda(), rotations=sc.rotations.cuda(), cov3D_precomp=None)
    out[0].sum().backward(retain_graph=True)
    g1 = m3.grad.clone()
    out[0].sum().backward()
    assert rel_err(m3.grad, 2 * g1) < 1e-5
    # outputs are independent tensors the caller may write into
    out[2][out[3] < 0.5] = 0.0
    with pytest.raises(Exception):
        GaussianRasterizer(raster_settings=rs)(means3D=m3, means2D=m2, shs=None, colors_precomp=None,
                                               opacities=sc.opacities.cuda(), scales=sc.scales.cuda(),
                                               rotations=sc.rotations.cuda(), cov3D_precomp=None)


def test_dist2():
    from simple_knn._C import distCUDA2
    g = torch.Generator().manual_seed(3)
    for P in (4096, 5, 20000, 40000):
        pts = torch.randn(P, 3, generator=g) * torch.tensor([1.0, 0.5, 2.0])
        got = distCUDA2(pts.cuda()).cpu()
        ref = dist2_oracle(pts)
        assert rel_err(got, ref) < 1e-5, P
    z = distCUDA2(torch.zeros(4096, 3, device="cuda"))     # checkpoint-load path: identical points
    assert float(z.abs().max()) == 0.0
    z = distCUDA2(torch.zeros(20000, 3, device="cuda"))
    assert float(z.abs().max()) == 0.0


def test_cuda_matches_committed_golden_vectors():
    """CUDA path vs tests/golden/config1_oracle.npz (BASELINE.json configs[0]); nothing here reads
    /root/reference or runs the oracle."""
    import hashlib
    from pathlib import Path
    import numpy as np
    from b200splat import ops
    import importlib.util
    gdir = Path(__file__).parent / "golden"
    gold = np.load(gdir / "config1_oracle.npz")
    spec_ = importlib.util.spec_from_file_location("make_golden", gdir / "make_golden.py")
    mg = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mg)
    scene, cam, pix_grads = mg.load_inputs(gold)     # the fixture's own inputs: no host-dependent regeneration
    s = oracle_settings(cam, 0)
    camc = ops.make_cam(cuda_settings(s), "cuda")
    d = lambda t: t.cuda().contiguous()
    m3, sh, op, scl, rot = map(d, (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations))
    color, radii, depth, alpha, st = ops.forward(camc, m3, sh, None, op, scl, rot, None)
    v = {k: t.cpu() for k, t in ops.forward_views(camc, st).items()}
    sha = lambda t: hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()
    assert np.array_equal(radii.cpu().numpy(), gold["radii"])
    assert np.array_equal(v["tiles_touched"].numpy(), gold["tiles_touched"])
    assert st.num_rendered == int(gold["num_rendered"])
    assert sha(v["keys_sorted"]) == str(gold["keys_sorted_sha256"])
    assert sha(v["point_list"]) == str(gold["point_list_sha256"])
    assert np.array_equal(v["ranges"].numpy(), gold["ranges"])
    # the fixture carries the oracle's borderline pixels and the error a flipped blend decision may cause there
    bounds = dict(mask=torch.from_numpy(gold["borderline_mask"]), color=torch.from_numpy(gold["bound_color"]),
                  depth=torch.from_numpy(gold["bound_depth"]), alpha=torch.from_numpy(gold["bound_alpha"]))
    ref = {k: torch.from_numpy(gold[k]) for k in ("color", "depth", "alpha", "n_contrib")}
    check_images(dict(color=color, depth=depth, alpha=alpha, n_contrib=v["n_contrib"]), ref, bounds, IMG_TOL)
    gc, gd, ga = (g.cuda() for g in pix_grads)
    g = ops.backward(camc, st, m3, sh, None, op, scl, rot, None, radii, alpha, gc, gd, ga)
    for k, gk in (("means3D", "g_means3D"), ("means2D", "g_means2D"), ("shs", "g_shs"), ("opacities", "g_opacities"),
                  ("scales", "g_scales"), ("rotations", "g_rotations")):
        assert rel_err(g[k], torch.from_numpy(gold[gk])) <= GRAD_TOL, k


def test_fused_densification_stats_and_accumulate():
    """Fused statistics epilogue == the reference's per-view update (geometry/gaussian_base.py:815-851);
    accumulate mode sums parameter gradients over views."""
    from b200splat import batched, ops
    sc, cam0 = _scene(5000, 1, 64, 64, 61)
    cams_h = scenes.sds_cameras(3, 64, 64, seed=62)
    dev = "cuda"
    d = lambda t: t.to(dev).contiguous()
    m3, sh, op, scl, rot = map(d, (sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations))
    cams = [ops.make_cam(cuda_settings(oracle_settings(c, 1)), dev) for c in cams_h]
    pgs = [tuple(d(g) for g in scenes.pixel_grads(64, 64, 70 + i)) for i in range(3)]
    pk = batched.PackedGrads(5000, sh.shape[1], dev)
    batched.render_views_fwd_bwd(cams, m3, sh, None, op, scl, rot, pgs, pk)
    acc = torch.zeros(5000, device=dev)
    den = torch.zeros(5000, device=dev)
    mr = torch.zeros(5000, device=dev)
    gsum = None
    for cam, pg in zip(cams, pgs):
        color, radii, depth, alpha, st = ops.forward(cam, m3, sh, None, op, scl, rot, None)
        g = ops.backward(cam, st, m3, sh, None, op, scl, rot, None, radii, alpha, *pg)
        vis = radii > 0
        mr = torch.max(mr, radii.float())
        acc[vis] += g["means2D"][vis, :2].norm(dim=-1)
        den[vis] += 1
        gsum = {k: v.clone() for k, v in g.items()} if gsum is None else {k: gsum[k] + v for k, v in g.items()}
    assert torch.equal(pk.max_radii, mr) and torch.equal(pk.views["denom"], den)
    assert rel_err(pk.views["grad_accum"], acc) < 1e-5
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        assert rel_err(pk.views[k], gsum[k]) < 1e-4, k


def test_batched_entry_points_match_per_view_path():
    """b200splat_forward_batched / _backward_batched (one launch per phase for all views, device-side
    num_rendered, gradients summed in registers) == the per-view single-view calls, bit-exact on the
    integer outputs."""
    from b200splat import batched, ops
    P, deg, H, W, V = 6000, 3, 80, 96, 3
    sc, _ = _scene(P, deg, H, W, 71)
    cams_h = scenes.sds_cameras(V, H, W, seed=72)
    dev = "cuda"
    d = lambda t: t.to(dev).contiguous()
    m3, sh, op, scl, rot = map(d, (sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations))
    cams = [ops.make_cam(cuda_settings(oracle_settings(c, deg)), dev) for c in cams_h]
    pgs = [tuple(d(g) for g in scenes.pixel_grads(H, W, 80 + i)) for i in range(V)]
    # per-view path
    pk_ref = batched.PackedGrads(P, sh.shape[1], dev)
    ref_imgs = batched.render_views_fwd_bwd(cams, m3, sh, None, op, scl, rot, pgs, pk_ref, keep_images=True)
    # batched path (tiny initial capacity forces one overflow + regrow in calibrate)
    br = batched.BatchRenderer(P, sh.shape[1], H, W, dev, views=V)
    br.ws[0]._alloc_binning(1024)
    br.step(cams, m3, sh, None, op, scl, rot, pgs)
    assert not br.overflowed()
    ws = br.ws[0]
    for v in range(V):
        color, depth, alpha, radii = ref_imgs[v]
        assert torch.equal(ws.radii[v], radii)
        assert float((ws.color[v] - color).abs().max()) < 1e-6
        assert float((ws.depth[v] - depth).abs().max()) < 1e-5
        assert float((ws.alpha[v] - alpha).abs().max()) < 1e-6
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        assert rel_err(br.packed.views[k], pk_ref.views[k]) < 1e-4, k
    assert torch.equal(br.packed.max_radii, pk_ref.max_radii)
    assert torch.equal(br.packed.views["denom"], pk_ref.views["denom"])
    assert rel_err(br.packed.views["grad_accum"], pk_ref.views["grad_accum"]) < 1e-5
    # the backward scratch cleans itself (scratch_clean): a second and a third step give the same gradients,
    # and so does the CUDA-graph replay of the step
    first = br.packed.buffer.clone()
    for _ in range(2):
        br.step(cams, m3, sh, None, op, scl, rot, pgs)
        assert rel_err(br.packed.buffer, first) < 1e-5
    for t in ws.scratch:
        assert int(t.view(torch.int32).ne(0).sum()) == 0, "scratch must be all-zero between steps"
    graph = br.capture_step(cams, m3, sh, None, op, scl, rot, pgs)
    br.packed.buffer.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert rel_err(br.packed.buffer, first) < 1e-5
    # the multi-GPU split of the step (forward + render backward | preprocess backward in Gaussian ranges, each
    # followed by its exchange on a side stream) computes the same gradients
    ranges = []
    br.packed.buffer.zero_()
    br.step_head(cams, m3, sh, None, op, scl, rot, pgs)
    br.step_tail(cams, m3, sh, None, op, scl, rot, lambda g0, g1: ranges.append((g0, g1)), chunks=3)
    torch.cuda.synchronize()
    assert ranges[0][0] == 0 and ranges[-1][1] == P and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert rel_err(br.packed.buffer, first) < 1e-5
    segs = br.packed.segments(ranges[1][0], ranges[1][1])
    assert len(segs) == len(br.packed.fields) + 1 and all(sg[0] % 4 == 0 and sg[1] % 4 == 0 for sg in segs)
    # sorted keys / ranges of a batched view are those of the single-view call
    st_b = ws.states(sh.shape[1])[1]
    vb = ops.forward_views(cams[1], st_b)
    _, _, _, _, st_s = ops.forward(cams[1], m3, sh, None, op, scl, rot, None)
    vs = ops.forward_views(cams[1], st_s)
    R = st_s.num_rendered
    assert int(vb["point_offsets"][-1]) == R
    assert torch.equal(vb["keys_sorted"][:R], vs["keys_sorted"]) and torch.equal(vb["point_list"][:R], vs["point_list"])
    assert torch.equal(vb["ranges"], vs["ranges"]) and torch.equal(vb["n_contrib"], vs["n_contrib"])


def test_view_batch_rasterizer_autograd_matches_per_view_operator():
    """ViewBatchRasterizer (batched autograd op) == a loop of GaussianRasterizer calls: images, radii,
    parameter gradients and the per-view means2D gradients the densification statistics read."""
    from b200splat.batched import ViewBatchRasterizer
    from diff_gaussian_rasterization import GaussianRasterizer
    P, deg, H, W, V = 5000, 2, 64, 64, 4
    sc, _ = _scene(P, deg, H, W, 91)
    cams_h = scenes.sds_cameras(V, H, W, seed=92)
    dev = "cuda"
    rss = [cuda_settings(oracle_settings(c, deg), dev) for c in cams_h]
    pgs = [tuple(g.to(dev) for g in scenes.pixel_grads(H, W, 93 + i)) for i in range(V)]

    def params():
        return [t.to(dev).clone().requires_grad_(True) for t in (sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations)]

    # reference: per-view loop
    m3, sh, op, scl, rot = params()
    m2s, loss, imgs = [], 0.0, []
    for v in range(V):
        m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
        c, r, d, a = GaussianRasterizer(raster_settings=rss[v])(means3D=m3, means2D=m2, shs=sh, colors_precomp=None,
                                                                opacities=op, scales=scl, rotations=rot,
                                                                cov3D_precomp=None)
        loss = loss + (c * pgs[v][0]).sum() + (d * pgs[v][1]).sum() + (a * pgs[v][2]).sum()
        m2s.append(m2); imgs.append((c, r, d, a))
    loss.backward()
    ref = [t.grad.clone() for t in (m3, sh, op, scl, rot)]
    # batched
    m3b, shb, opb, sclb, rotb = params()
    m2b = torch.zeros(V, P, 3, device=dev, requires_grad=True)
    rast = ViewBatchRasterizer(V, P, H, W, dev)
    C, R, D, A = rast(rss, means3D=m3b, means2D=m2b, opacities=opb, shs=shb, scales=sclb, rotations=rotb)
    lb = sum((C[v] * pgs[v][0]).sum() + (D[v] * pgs[v][1]).sum() + (A[v] * pgs[v][2]).sum() for v in range(V))
    lb.backward()
    assert not rast.check_overflow()
    for v in range(V):
        assert torch.equal(R[v], imgs[v][1])
        assert float((C[v] - imgs[v][0]).abs().max()) < 1e-6
        assert rel_err(m2b.grad[v], m2s[v].grad) < 1e-4
    for got, want, name in zip((m3b, shb, opb, sclb, rotb), ref, ("means3D", "shs", "opacities", "scales", "rotations")):
        assert rel_err(got.grad, want) < 1e-4, name
    # second step reuses the calibrated workspace without a host sync
    m2c = torch.zeros(V, P, 3, device=dev, requires_grad=True)
    C2, _, _, _ = rast(rss, means3D=m3b, means2D=m2c, opacities=opb, shs=shb, scales=sclb, rotations=rotb)
    assert float((C2 - C).abs().max()) == 0.0


@pytest.mark.timeout(600)
def test_config2_shape_single_view_parity():
    """BASELINE.json configs[1] shape (100K Gaussians, 512x512, SH degree 0), one of its views, against the oracle."""
    scene, cams = scenes.make_workload("config2_100k_512_sh0_b4", views=1)
    cam = cams[0]
    s = oracle_settings(cam, 0)
    grads = scenes.pixel_grads(512, 512, 5)
    orc = _run_oracle(scene, s, grads=grads)
    cu = _run_cuda(scene, s, grads=grads)
    _check_forward(cu, orc, s)
    _check_grads(cu, orc)


@pytest.mark.timeout(600)
def test_full_size_properties_headline_workload():
    """BASELINE headline size (1M Gaussians, SH3, 512x512): size-independent properties of the CUDA path --
    sortedness, range partition, bounds, determinism of the integer outputs, linearity of backward in the
    upstream gradient, batched == single-view."""
    from b200splat import batched, ops
    scene, cams_h = scenes.make_workload("headline_1m_512_sh3", views=2)
    dev = "cuda"
    d = lambda t: t.to(dev).contiguous()
    m3, sh, op, scl, rot = map(d, (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations))
    P, H, W = m3.shape[0], 512, 512
    cams = [ops.make_cam(cuda_settings(oracle_settings(c, 3)), dev) for c in cams_h]
    color, radii, depth, alpha, st = ops.forward(cams[0], m3, sh, None, op, scl, rot, None)
    v = ops.forward_views(cams[0], st)
    R = st.num_rendered
    assert R == int(v["tiles_touched"].long().sum()) == int(v["point_offsets"][-1])
    ks = v["keys_sorted"]
    assert bool((ks[1:] >= ks[:-1]).all()), "keys not sorted"
    assert bool((torch.sort(v["point_list"].long()).values == torch.sort(
        torch.repeat_interleave(torch.arange(P, device=dev), v["tiles_touched"].long())).values).all()), \
        "point list is not a permutation of the duplicated indices"
    rg = v["ranges"].long()
    lens = rg[:, 1] - rg[:, 0]
    assert int(lens.sum()) == R and bool((lens >= 0).all())
    tiles_of = (ks >> 32)
    nz = lens > 0
    assert bool((tiles_of[rg[nz, 0]] == torch.nonzero(nz).reshape(-1)).all())
    # depth bits of the key == float bits of the Gaussian's depth
    assert torch.equal((ks & 0xFFFFFFFF).to(torch.int32), v["depths"][v["point_list"].long()].view(torch.int32))
    nc = v["n_contrib"].long().reshape(H // 16, 16, W // 16, 16).permute(0, 2, 1, 3).reshape(-1, 256)
    assert bool((nc.max(1).values <= lens).all())
    assert torch.isfinite(color).all() and float(alpha.min()) >= 0 and float(alpha.max()) <= 1 + 1e-5
    assert bool(((radii > 0) == (v["tiles_touched"] > 0)).all())
    # second forward: integer outputs identical (deterministic), images identical
    color2, radii2, _, alpha2, st2 = ops.forward(cams[0], m3, sh, None, op, scl, rot, None)
    v2 = ops.forward_views(cams[0], st2)
    assert torch.equal(radii, radii2) and torch.equal(v["keys_sorted"], v2["keys_sorted"])
    assert torch.equal(v["point_list"], v2["point_list"]) and torch.equal(color, color2)
    # backward is linear in the upstream gradient
    g1 = tuple(d(g) for g in scenes.pixel_grads(H, W, 3))
    ga = ops.backward(cams[0], st, m3, sh, None, op, scl, rot, None, radii, alpha, *g1)
    gb = ops.backward(cams[0], st, m3, sh, None, op, scl, rot, None, radii, alpha, *(2.5 * g for g in g1))
    for k in ("means3D", "means2D", "shs", "opacities", "scales", "rotations"):
        assert rel_err(gb[k], 2.5 * ga[k]) < 1e-4, k
        culled = radii == 0
        assert float(ga[k][culled].abs().max()) == 0.0 if bool(culled.any()) else True
    # batched path == single-view path at full size
    br = batched.BatchRenderer(P, sh.shape[1], H, W, dev, views=2)
    pgs = [g1, tuple(d(g) for g in scenes.pixel_grads(H, W, 4))]
    br.step(cams, m3, sh, None, op, scl, rot, pgs)
    assert not br.overflowed()
    assert torch.equal(br.ws[0].radii[0], radii) and float((br.ws[0].color[0] - color).abs().max()) < 1e-6
    # the batched pipeline (fused scan + duplicateWithKeys, look-back-free partition from the count matrix) produces the
    # single-view pipeline's sorted list bit for bit
    vb = ops.forward_views(cams[0], br.ws[0].states(sh.shape[1])[0])
    assert int(vb["point_offsets"][-1]) == R
    assert torch.equal(vb["keys_sorted"][:R], v["keys_sorted"]) and torch.equal(vb["point_list"][:R], v["point_list"])
    assert torch.equal(vb["ranges"], v["ranges"]) and torch.equal(vb["n_contrib"], v["n_contrib"])
    _, radii_b, _, alpha_b, st_b = ops.forward(cams[1], m3, sh, None, op, scl, rot, None)
    gsum = {k: ga[k] + ops.backward(cams[1], st_b, m3, sh, None, op, scl, rot, None, radii_b, alpha_b, *pgs[1])[k]
            for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    for k, t in gsum.items():
        assert rel_err(br.packed.views[k], t) < 1e-4, k


@pytest.mark.parametrize("P,H,W,scale_mul", [(20000, 256, 256, 6.0),     # huge splats: > 8 partition tiles per CTA
                                              (50000, 512, 512, 1.0),     # 1024 tiles: every bin of the wide pass
                                              (700, 64, 48, 1.0),         # less than one partition tile
                                              (60000, 1024, 1024, 1.5)])  # 4096 tiles: fused kernel + two 8-bit passes
def test_batched_binning_matches_single_view_bit_for_bit(P, H, W, scale_mul):
    """The batched forward's binning (depth sort -> fused scan + duplicateWithKeys with the per-(partition tile, image
    tile) count matrix -> partition offsets -> look-back-free partition) against the single-view pipeline (stand-alone
    scan, duplicateWithKeys, look-back partition), which the oracle tests pin: sorted words, list, ranges, n_contrib."""
    from b200splat import batched, ops
    sc, _ = _scene(P, 0, H, W, 905)
    cams_h = scenes.sds_cameras(2, H, W, seed=906)
    dev = "cuda"
    d = lambda t: t.to(dev).contiguous()
    m3, sh, op, rot = map(d, (sc.means3D, sc.shs, sc.opacities, sc.rotations))
    scl = d(sc.scales * scale_mul)
    cams = [ops.make_cam(cuda_settings(oracle_settings(c, 0)), dev) for c in cams_h]
    pgs = [tuple(d(g) for g in scenes.pixel_grads(H, W, 910 + i)) for i in range(2)]
    br = batched.BatchRenderer(P, sh.shape[1], H, W, dev, views=2)
    br.step(cams, m3, sh, None, op, scl, rot, pgs)
    assert not br.overflowed()
    for i in range(2):
        color, radii, depth, alpha, st = ops.forward(cams[i], m3, sh, None, op, scl, rot, None)
        vs = ops.forward_views(cams[i], st)
        vb = ops.forward_views(cams[i], br.ws[0].states(sh.shape[1])[i])
        R = st.num_rendered
        assert int(vb["point_offsets"][-1]) == R
        assert torch.equal(vb["keys_sorted"][:R], vs["keys_sorted"]), "sorted pair words"
        assert torch.equal(vb["point_list"][:R], vs["point_list"]), "point list"
        assert torch.equal(vb["ranges"], vs["ranges"]) and torch.equal(vb["n_contrib"], vs["n_contrib"])
        assert torch.equal(br.ws[0].color[i], color)


def test_stress_shape_pair_mode_and_large_tile_count():
    """600K Gaussians at 1024x1024 (T = 4096 tiles, the stress config's tile count: the pair words go through two
    8-bit passes of the generic onesweep kernel instead of the single wide partition): batched == per-view, sortedness."""
    from b200splat import batched, ops
    P, H, W = 600_000, 1024, 1024
    sc = scenes.make_scene(P, 0, 0.5, seed=5)
    cams_h = scenes.mvdream_cameras(2, H, W, seed=6)
    dev = "cuda"
    d = lambda t: t.to(dev).contiguous()
    m3, sh, op, scl, rot = map(d, (sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations))
    cams = [ops.make_cam(cuda_settings(oracle_settings(c, 0)), dev) for c in cams_h]
    color, radii, depth, alpha, st = ops.forward(cams[0], m3, sh, None, op, scl, rot, None)
    v = ops.forward_views(cams[0], st)
    ks = v["keys_sorted"]
    assert bool((ks[1:] >= ks[:-1]).all())
    assert int((v["ranges"][:, 1] - v["ranges"][:, 0]).sum()) == st.num_rendered
    pgs = [tuple(d(g) for g in scenes.pixel_grads(H, W, 7 + i)) for i in range(2)]
    pk = batched.PackedGrads(P, 1, dev)
    batched.render_views_fwd_bwd(cams, m3, sh, None, op, scl, rot, pgs, pk)
    br = batched.BatchRenderer(P, 1, H, W, dev, views=2)
    br.step(cams, m3, sh, None, op, scl, rot, pgs)
    assert not br.overflowed()
    assert torch.equal(br.ws[0].radii[0], radii) and float((br.ws[0].color[0] - color).abs().max()) < 1e-6
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        assert rel_err(br.packed.views[k], pk.views[k]) < 1e-4, k


# ----------------------------------------------------------------------------------------------------------------
# Oracle parity AT THE BENCHMARKED SHAPES (BASELINE.json configs / the bench workloads): one view each, the whole
# operator -- integer stages bit-exact, images 1e-4 (bounded on borderline pixels), n_contrib, gradients 1e-3.
# The CPU oracle needs a few seconds per view at these sizes.
# ----------------------------------------------------------------------------------------------------------------
def _full_scale_parity(scene, cam, bg=(1.0, 1.0, 1.0), seed=5, colors_precomp=None):
    from b200splat import ops
    deg = scene.sh_degree
    H, W = cam.image_height, cam.image_width
    s = oracle_settings(cam, deg, bg=bg)
    grads = scenes.pixel_grads(H, W, seed)
    orc = _run_oracle(scene, s, grads=grads, colors_precomp=colors_precomp)
    pre, binned, out = orc["pre"], orc["binned"], orc["out"]
    camc = ops.make_cam(cuda_settings(s), "cuda")
    d = lambda t: None if t is None else t.cuda().contiguous()
    m3, op, scl, rot = map(d, (scene.means3D, scene.opacities, scene.scales, scene.rotations))
    sh = None if colors_precomp is not None else d(scene.shs)
    cp = d(colors_precomp)
    color, radii, depth, alpha, st = ops.forward(camc, m3, sh, cp, op, scl, rot, None)
    v = {k: t.cpu() for k, t in ops.forward_views(camc, st).items()}
    assert st.num_rendered == binned["num_rendered"]
    assert torch.equal(radii.cpu(), pre["radii"]), "radii"
    assert torch.equal(v["tiles_touched"], pre["tiles_touched"]), "tiles_touched"
    assert torch.equal(v["keys_sorted"], binned["keys_sorted"]), "sorted keys"
    assert torch.equal(v["point_list"], binned["point_list"]), "point_list"
    assert torch.equal(v["ranges"], binned["ranges"]), "ranges"
    bounds = borderline_bounds(pre, binned, s, out)
    frac, worst = check_images(dict(color=color, depth=depth, alpha=alpha, n_contrib=v["n_contrib"]), out, bounds,
                               IMG_TOL)
    g = ops.backward(camc, st, m3, sh, cp, op, scl, rot, None, radii, alpha, *(t.cuda() for t in grads))
    errs = _check_grads(dict(grads=g), orc)
    wide = sum(n for _, n in errs.values())
    total = sum(t.numel() for k, t in orc["grads"].items() if k != "stage" and t is not None)
    assert wide <= 1e-4 * total, f"{wide} of {total} gradient elements needed the borderline slack"
    return dict(borderline_frac=frac, worst=worst, grad_err=errs, R=st.num_rendered)


@pytest.mark.timeout(900)
def test_headline_1m_512_sh3_parity():
    """The bench workload (BASELINE.json metric: 1 M Gaussians, SH degree 3, 512 x 512)."""
    scene, cams = scenes.make_workload("headline_1m_512_sh3", views=2)
    for cam in cams:          # two of the rig's views (opposite azimuths)
        _full_scale_parity(scene, cam)


@pytest.mark.timeout(900)
def test_config3_300k_512_raster_and_shading_postops_parity():
    """BASELINE.json configs[2]: 300 K Gaussians, 512 x 512, the shading renderer -- the raster pass over a black
    background (renderer/diff_gaussian_rasterizer_shading.py:94) against the oracle, then the fused post-ops at this
    size against oracle/postops.py on the same raster images (forward 1e-4, pixel gradients 1e-3)."""
    from b200splat import ops
    from b200splat.postops import postprocess_views
    from oracle import postops as PO
    scene, cams = scenes.make_workload("config3_300k_512_sh0_b4", views=1)
    cam = cams[0]
    _full_scale_parity(scene, cam, bg=(0.0, 0.0, 0.0))
    H, W = cam.image_height, cam.image_width
    s = oracle_settings(cam, 0, bg=(0.0, 0.0, 0.0))
    camc = ops.make_cam(cuda_settings(s), "cuda")
    d = lambda t: t.cuda().contiguous()
    color, radii, depth, alpha, st = ops.forward(camc, d(scene.means3D), d(scene.shs), None, d(scene.opacities),
                                                 d(scene.scales), d(scene.rotations), None)
    g = torch.Generator().manual_seed(77)
    rays_d = torch.nn.functional.normalize(torch.randn(H, W, 3, generator=g) * 0.3 + torch.tensor([0.0, 1.0, 0.0]), dim=-1)
    rays_o = cam.campos[None, None, :].expand(H, W, 3).contiguous()
    bgm, light = torch.rand(H, W, 3, generator=g), torch.randn(3, generator=g) * 2
    w = {k: torch.randn(c, H, W, generator=g) / (H * W) for k, c in (("render", 3), ("normal", 3), ("depth", 1))}
    for shading in ("diffuse", "textureless", "albedo"):
        lv = {k: t.detach().clone().requires_grad_(True) for k, t in (("image", color), ("depth", depth), ("alpha", alpha))}
        bgc = bgm.cuda().requires_grad_(True)
        post = postprocess_views("shading", lv["image"][None], lv["depth"][None], lv["alpha"][None], bg=bgc[None],
                                 rays_o=rays_o.cuda()[None], rays_d=rays_d.cuda()[None],
                                 light_positions=light.cuda()[None], shading=shading)
        sum((post[k] * w[k].cuda()[None]).sum() for k in w).backward()
        lo = {k: t.detach().cpu().clone().requires_grad_(True) for k, t in (("image", color), ("depth", depth), ("alpha", alpha))}
        bgo = bgm.clone().requires_grad_(True)
        ref = PO.postprocess_view(PO.MODE_SHADING, lo["image"], lo["depth"], lo["alpha"], rays_o, rays_d, bgo, light,
                                  torch.tensor([0.1] * 3), torch.tensor([0.9] * 3), shading, None)
        sum((ref[k] * w[k]).sum() for k in w).backward()
        for k in w:
            assert float((post[k][0].detach().cpu() - ref[k].detach()).abs().max()) <= IMG_TOL, (shading, k)
        for k in lv:
            if lo[k].grad is None:
                assert lv[k].grad is None or float(lv[k].grad.abs().max()) == 0.0, (shading, k)
            else:
                assert rel_err(lv[k].grad, lo[k].grad) <= GRAD_TOL, (shading, k)
        if bgo.grad is not None:
            assert rel_err(bgc.grad, bgo.grad) <= GRAD_TOL, (shading, "bg")


@pytest.mark.timeout(900)
def test_config4_1m_256_sh3_parity():
    """BASELINE.json configs[3]: 1 M Gaussians, SH degree 3, 256 x 256 (256 tiles), MVDream rig."""
    scene, cams = scenes.make_workload("config4_1m_256_sh3_b32", views=5)
    for cam in (cams[0], cams[4]):   # one view of each of two camera groups (different fovy / distance)
        _full_scale_parity(scene, cam)


@pytest.mark.timeout(900)
def test_4096_tile_grid_sh3_parity():
    """The stress configuration's image size (1024 x 1024: 4096 tiles, the pair words take the wide-grid path) with
    250 K SH-degree-3 Gaussians."""
    scene = scenes.make_scene(250_000, 3, 0.5, seed=4242)
    cam = scenes.mvdream_cameras(1, 1024, 1024, seed=4243)[0]
    _full_scale_parity(scene, cam)


@pytest.mark.timeout(900)
def test_spacetime_call_shape_parity_at_scale():
    """BASELINE.json configs[4] call shape (renderer/diff_gaussian_rasterizer_st.py:135-150): precomputed colours
    instead of SH, 200 K Gaussians at 512 x 512."""
    scene = scenes.make_scene(200_000, 0, 0.5, seed=777)
    cam = scenes.mvdream_cameras(1, 512, 512, seed=778)[0]
    cp = torch.rand(200_000, 3, generator=torch.Generator().manual_seed(779))
    _full_scale_parity(scene, cam, colors_precomp=cp)


def test_binning_capacity_heals_itself():
    """a view that outgrows the calibrated binning capacity must not come back as background.
    The capacity is shrunk below what a closer, narrower camera needs; the forward notices (pair-count notice from
    the scan kernel), regrows and renders again -- images, radii and gradients equal the per-view operator's."""
    from b200splat.batched import ViewBatchRasterizer
    from diff_gaussian_rasterization import GaussianRasterizer
    P, deg, H, W, V = 20000, 1, 96, 96, 2
    sc, _ = _scene(P, deg, H, W, 301)
    far = scenes.sds_cameras(V, H, W, seed=302, camera_distance=4.0, fovy_deg=(69.0, 70.0))
    near = scenes.sds_cameras(V, H, W, seed=302, camera_distance=1.6, fovy_deg=(30.0, 31.0))
    dev = "cuda"
    mk = lambda: [t.to(dev).clone().requires_grad_(True) for t in (sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations)]
    rast = ViewBatchRasterizer(V, P, H, W, dev)
    m3, sh, op, scl, rot = mk()
    rss = [cuda_settings(oracle_settings(c, deg), dev) for c in far]
    rast(rss, means3D=m3, means2D=torch.zeros(V, P, 3, device=dev), opacities=op, shs=sh, scales=scl, rotations=rot)
    small = max(rast.last_pairs) + 64
    rast.ws._alloc_binning(small)                       # what a tight calibration on the far cameras would give
    cap0, regrown0 = rast.ws.capacity, rast.regrown
    rss = [cuda_settings(oracle_settings(c, deg), dev) for c in near]
    pgs = [tuple(g.to(dev) for g in scenes.pixel_grads(H, W, 303 + i)) for i in range(V)]
    m2 = torch.zeros(V, P, 3, device=dev, requires_grad=True)
    C, R, D, A = rast(rss, means3D=m3, means2D=m2, opacities=op, shs=sh, scales=scl, rotations=rot)
    assert max(rast.last_pairs) > cap0, "the near cameras were meant to overflow the shrunk capacity"
    assert rast.regrown == regrown0 + 1 and rast.ws.capacity >= max(rast.last_pairs)
    assert not rast.check_overflow()
    sum((C[v] * pgs[v][0]).sum() + (D[v] * pgs[v][1]).sum() + (A[v] * pgs[v][2]).sum() for v in range(V)).backward()
    n3, nsh, nop, nscl, nrot = mk()
    loss = 0.0
    for v in range(V):
        n2 = torch.zeros(P, 3, device=dev, requires_grad=True)
        c, r, d, a = GaussianRasterizer(raster_settings=rss[v])(means3D=n3, means2D=n2, shs=nsh, colors_precomp=None,
                                                                opacities=nop, scales=nscl, rotations=nrot,
                                                                cov3D_precomp=None)
        assert torch.equal(R[v], r) and float((C[v] - c).abs().max()) < 1e-6 and float((A[v] - a).abs().max()) < 1e-6
        assert float(a.max()) > 0.5, "the view must actually show the object"
        loss = loss + (c * pgs[v][0]).sum() + (d * pgs[v][1]).sum() + (a * pgs[v][2]).sum()
    loss.backward()
    for got, want, name in zip((m3, sh, op, scl, rot), (n3, nsh, nop, nscl, nrot),
                               ("means3D", "shs", "opacities", "scales", "rotations")):
        assert rel_err(got.grad, want.grad) < 1e-4, name


@pytest.mark.parametrize("P", [4999, 5001, 5002])
def test_odd_gaussian_counts_through_the_packed_paths(P):
    """P is arbitrary after densify / prune.  Every field of the packed gradient buffer stays
    16-byte aligned (rotations are stored as float4), through BatchRenderer.step, the per-view loop into the packed
    buffer, the step_head / step_tail split and FusedGaussianAdam's view of the same buffer."""
    from b200splat import batched, ops
    deg, H, W, V = 1, 64, 64, 2
    sc, _ = _scene(P, deg, H, W, 311)
    cams_h = scenes.sds_cameras(V, H, W, seed=312)
    dev = "cuda"
    d = lambda t: t.to(dev).contiguous()
    m3, sh, op, scl, rot = map(d, (sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations))
    cams = [ops.make_cam(cuda_settings(oracle_settings(c, deg)), dev) for c in cams_h]
    pgs = [tuple(d(g) for g in scenes.pixel_grads(H, W, 313 + i)) for i in range(V)]
    pk_ref = batched.PackedGrads(P, sh.shape[1], dev)
    assert all(off % 4 == 0 for _, off, _ in pk_ref.fields) and pk_ref.views["rotations"].data_ptr() % 16 == 0
    batched.render_views_fwd_bwd(cams, m3, sh, None, op, scl, rot, pgs, pk_ref)
    # independent reference: the drop-in operator + autograd
    leaves = [t.clone().requires_grad_(True) for t in (m3, sh, op, scl, rot)]
    from diff_gaussian_rasterization import GaussianRasterizer
    loss = 0.0
    for v in range(V):
        rs = cuda_settings(oracle_settings(cams_h[v], deg), dev)
        c, r, dd, a = GaussianRasterizer(raster_settings=rs)(means3D=leaves[0], means2D=torch.zeros(P, 3, device=dev),
                                                             shs=leaves[1], colors_precomp=None, opacities=leaves[2],
                                                             scales=leaves[3], rotations=leaves[4], cov3D_precomp=None)
        loss = loss + (c * pgs[v][0]).sum() + (dd * pgs[v][1]).sum() + (a * pgs[v][2]).sum()
    loss.backward()
    for k, t in zip(("means3D", "shs", "opacities", "scales", "rotations"), leaves):
        assert rel_err(pk_ref.views[k], t.grad) < 1e-4, k
    br = batched.BatchRenderer(P, sh.shape[1], H, W, dev, views=V)
    br.step(cams, m3, sh, None, op, scl, rot, pgs)
    assert not br.overflowed()
    for k in ("means3D", "shs", "opacities", "scales", "rotations", "grad_accum", "denom"):
        assert rel_err(br.packed.views[k], pk_ref.views[k]) < 1e-4, k
    first = br.packed.buffer.clone()
    br.packed.buffer.zero_()
    br.step_head(cams, m3, sh, None, op, scl, rot, pgs)
    seen = []
    br.step_tail(cams, m3, sh, None, op, scl, rot, lambda g0, g1: seen.append(br.packed.segments(g0, g1)), chunks=3)
    torch.cuda.synchronize()
    assert rel_err(br.packed.buffer, first) < 1e-5
    assert all(sg[0] % 4 == 0 and sg[1] % 4 == 0 for segs in seen for sg in segs)
    # the padding Gaussians of every field stay zero (they are summed by the exchange)
    for name, off, w in br.packed.fields:
        assert float(br.packed.buffer[off + w * P: off + w * br.packed.P4].abs().sum()) == 0.0, name


def test_view_batch_rasterizer_survives_densify_and_prune():
    """The number of Gaussians changes between steps (geometry/gaussian_base.py:853-869): the same
    ViewBatchRasterizer keeps working (grow-only workspace), no rebuild, results equal the per-view operator's."""
    from b200splat.batched import ViewBatchRasterizer
    from diff_gaussian_rasterization import GaussianRasterizer
    deg, H, W, V = 0, 64, 64, 2
    dev = "cuda"
    cams_h = scenes.sds_cameras(V, H, W, seed=322)
    rss = [cuda_settings(oracle_settings(c, deg), dev) for c in cams_h]
    rast = ViewBatchRasterizer(V, 3000, H, W, dev)
    geoms = []
    for P in (3000, 4501, 2000, 6000):
        sc, _ = _scene(P, deg, H, W, 320 + P)
        leaves = [t.to(dev).clone().requires_grad_(True) for t in (sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations)]
        m2 = torch.zeros(V, P, 3, device=dev, requires_grad=True)
        C, R, D, A = rast(rss, means3D=leaves[0], means2D=m2, opacities=leaves[2], shs=leaves[1], scales=leaves[3],
                          rotations=leaves[4])
        (C.sum() + D.sum()).backward()
        geoms.append(rast.ws.geom[0].data_ptr())
        ref = [t.detach().clone().requires_grad_(True) for t in leaves]
        loss = 0.0
        for v in range(V):
            n2 = torch.zeros(P, 3, device=dev, requires_grad=True)
            c, r, dd, a = GaussianRasterizer(raster_settings=rss[v])(means3D=ref[0], means2D=n2, shs=ref[1],
                                                                     colors_precomp=None, opacities=ref[2], scales=ref[3],
                                                                     rotations=ref[4], cov3D_precomp=None)
            assert torch.equal(R[v], r) and float((C[v] - c).abs().max()) < 1e-6
            loss = loss + c.sum() + dd.sum()
        loss.backward()
        for a_, b_ in zip(leaves, ref):
            assert rel_err(a_.grad, b_.grad) < 1e-4
    assert geoms[1] == geoms[2], "shrinking P must not reallocate the workspace"
