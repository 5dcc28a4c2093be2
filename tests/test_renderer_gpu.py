"""GPU: ``B200GaussianBatchRenderer.batch_forward`` (one batched raster call + one fused post-op launch) against the
composition the reference runs per view -- the single-view operator (parity-checked against the oracle in
test_parity_gpu.py) followed by the reference's post-ops (oracle/postops.py, pinned to the reference files) -- on the
same ``batch`` dict, for every renderer variant it mirrors."""
import random
import types

import pytest
import torch

from b200splat import scenes
from oracle import postops as PO
from util import rel_err

pytestmark = pytest.mark.gpu

OUT_TOL, GRAD_TOL = 1e-4, 1e-3
SH_C0 = 0.28209479177387814


class _Geometry:
    def __init__(self, sc, pred_normal, seed):
        mk = lambda t: t.cuda().clone().requires_grad_(True)
        self.get_xyz, self.get_opacity, self.get_scaling = mk(sc.means3D), mk(sc.opacities), mk(sc.scales)
        self.get_rotation, self.get_features = mk(sc.rotations), mk(sc.shs)
        g = torch.Generator().manual_seed(seed)
        self.get_normal = mk(torch.nn.functional.normalize(torch.randn(sc.means3D.shape[0], 3, generator=g), dim=-1))
        self.cfg = types.SimpleNamespace(pred_normal=pred_normal)
        self.active_sh_degree = sc.sh_degree

    def leaves(self):
        return dict(xyz=self.get_xyz, opacity=self.get_opacity, scaling=self.get_scaling, rotation=self.get_rotation,
                    features=self.get_features, normal=self.get_normal)


def _batch(V, H, W, seed):
    cams = scenes.sds_cameras(V, H, W, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    c2ws = []
    for cam in cams:
        c2w = torch.inverse(cam.viewmatrix.t())
        c2w[:3, 1:3] *= -1
        c2ws.append(c2w)
    rays_d = torch.nn.functional.normalize(torch.randn(V, H, W, 3, generator=g) * 0.3 + torch.tensor([0.0, 1.0, 0.0]),
                                           dim=-1)
    rays_o = torch.stack([cam.campos for cam in cams])[:, None, None, :].expand(V, H, W, 3).contiguous()
    batch = dict(c2w=torch.stack(c2ws).cuda(), fovy=torch.tensor([c.fovy for c in cams]).cuda(), height=H, width=W,
                 rays_o=rays_o.cuda(), rays_d=rays_d.cuda(), light_positions=(torch.randn(V, 3, generator=g) * 2).cuda())
    return batch, cams


def _weights(V, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(V, H, W, c, generator=g) / (H * W) for k, c in
            (("comp_rgb", 3), ("comp_normal", 3), ("comp_pred_normal", 3), ("comp_depth", 1), ("comp_mask", 1))}


def _loss(outputs, wts, dev):
    return sum((outputs[k] * w.to(dev)).sum() for k, w in wts.items() if k in outputs)


@pytest.mark.parametrize("variant,pred_normal,train,soft", [
    ("plain", False, False, False), ("advanced", False, True, False), ("background", False, True, False),
    ("normal", False, True, False), ("normal", True, True, False), ("shading", False, True, False),
    ("shading", True, False, False), ("shading", False, True, True)])
def test_batch_forward_matches_the_per_view_composition(variant, pred_normal, train, soft):
    from b200splat.renderer import B200GaussianBatchRenderer, _settings
    from diff_gaussian_rasterization import GaussianRasterizer
    P, H, W, V = 5000, 64, 80, 3
    sc = scenes.make_scene(P, 1, 0.8, seed=201)
    batch, cams = _batch(V, H, W, 202)
    wts = _weights(V, H, W, 203)
    bg_map = torch.rand(V, H, W, 3, generator=torch.Generator().manual_seed(204)).cuda().requires_grad_(True)
    bgt = torch.tensor([1.0, 1.0, 1.0], device="cuda")

    class Ren(B200GaussianBatchRenderer):
        pass

    ren = Ren()
    ren.variant, ren.training, ren.background_tensor = variant, train, bgt
    ren.geometry = _Geometry(sc, pred_normal, 205)
    ren.background = lambda dirs: bg_map
    mat = types.SimpleNamespace(ambient_light_color=torch.tensor([0.1, 0.1, 0.1]),
                                diffuse_light_color=torch.tensor([0.9, 0.9, 0.9]), ambient_only=False, training=train,
                                cfg=types.SimpleNamespace(diffuse_prob=0.75, textureless_prob=0.5, soft_shading=soft))
    ren.material = mat
    random.seed(7)
    out = ren.batch_forward(batch)
    _loss(out, wts, "cuda").backward()
    got_grads = {k: t.grad.clone() for k, t in ren.geometry.leaves().items() if t.grad is not None}
    got_bg = None if bg_map.grad is None else bg_map.grad.clone()
    bg_map.grad = None

    # ---- expected: the reference's per-view loop (same random call sequence for the material) -----------------------
    random.seed(7)
    geo = _Geometry(sc, pred_normal, 205)
    exp = {k: [] for k in ("comp_rgb", "comp_normal", "comp_pred_normal", "comp_depth", "comp_mask")}
    m2_grads, radii_all = [], []
    m2s = []
    for v, cam in enumerate(cams):
        # invert_bg_prob = 1: never inverted in training, always in eval (renderer/diff_gaussian_rasterizer.py:59-64)
        raster_bg = (bgt if train else 1.0 - bgt) if variant in ("plain", "advanced", "normal") else bgt * 0
        # the camera is rebuilt from batch["c2w"] / batch["fovy"] exactly as the reference's loop does
        # (renderer/gaussian_batch_renderer.py:23-49 -> get_cam_info_gaussian)
        rs = _settings(batch, v, raster_bg, 1, "cuda")
        m2 = torch.zeros(P, 3, device="cuda", requires_grad=True)
        m2s.append(m2)
        kw = dict(means3D=geo.get_xyz, opacities=geo.get_opacity, scales=geo.get_scaling, rotations=geo.get_rotation,
                  cov3D_precomp=None)
        image, radii, depth, alpha = GaussianRasterizer(raster_settings=rs)(means2D=m2, shs=geo.get_features,
                                                                            colors_precomp=None, **kw)
        radii_all.append(radii)
        pred = None
        if pred_normal:
            pred = GaussianRasterizer(raster_settings=rs)(means2D=torch.zeros_like(m2), shs=geo.get_normal.unsqueeze(1),
                                                          colors_precomp=None, **kw)[0]
        shading = "diffuse"
        ambient_c, diffuse_c = mat.ambient_light_color, mat.diffuse_light_color
        if variant == "shading":
            if train and soft:   # material/gaussian_material.py:58-63: a new ambient ratio per view, drawn first
                r = random.random()
                diffuse_c = torch.full_like(mat.diffuse_light_color, r)
                ambient_c = 1.0 - diffuse_c
            if train:
                shading = "albedo" if random.random() > 0.75 else ("textureless" if random.random() < 0.5 else "diffuse")
        mode = {"plain": PO.MODE_PLAIN, "advanced": PO.MODE_PLAIN, "background": PO.MODE_BACKGROUND,
                "normal": PO.MODE_NORMAL, "shading": PO.MODE_SHADING}[variant]
        post = PO.postprocess_view(mode, image.cpu(), depth.cpu(), alpha.cpu(), batch["rays_o"][v].cpu(),
                                   batch["rays_d"][v].cpu(), bg_map[v].cpu(), batch["light_positions"][v].cpu(),
                                   ambient_c, diffuse_c, shading,
                                   None if pred is None else pred.cpu())
        exp["comp_rgb"].append(post["render"].permute(1, 2, 0))
        if variant in ("normal", "shading"):
            exp["comp_normal"].append(post["normal"].permute(1, 2, 0))
            if pred is not None:
                exp["comp_pred_normal"].append(pred.cpu().permute(1, 2, 0))
        if variant not in ("plain", "background"):
            exp["comp_depth"].append(post["depth"].permute(1, 2, 0))
            exp["comp_mask"].append(alpha.cpu().permute(1, 2, 0))
    exp = {k: torch.stack(v) for k, v in exp.items() if v}
    _loss(exp, wts, "cpu").backward()

    assert set(exp) <= set(out), (sorted(exp), sorted(out))
    for k, ref in exp.items():
        assert out[k].shape == ref.shape, (k, out[k].shape, ref.shape)
        err = float((out[k].detach().cpu() - ref.detach()).abs().max())
        assert err <= OUT_TOL, f"{variant}.{k}: {err}"
    for k, t in geo.leaves().items():
        if t.grad is None:
            assert k not in got_grads or float(got_grads[k].abs().max()) == 0.0, k
            continue
        assert rel_err(got_grads[k], t.grad) <= GRAD_TOL, f"{variant} grad {k}: {rel_err(got_grads[k], t.grad)}"
    if bg_map.grad is not None:
        assert rel_err(got_bg, bg_map.grad) <= GRAD_TOL
    for v in range(V):
        assert torch.equal(out["radii"][v], radii_all[v])
        assert torch.equal(out["visibility_filter"][v], radii_all[v] > 0)
        if not pred_normal:
            assert rel_err(out["viewspace_points"][v].grad, m2s[v].grad) <= GRAD_TOL
    # the unchanged update_states statistics (geometry/gaussian_base.py:815-819, :845-851) run on the outputs
    accum, denom = torch.zeros(P, 1, device="cuda"), torch.zeros(P, 1, device="cuda")
    for v in range(V):
        vis = out["visibility_filter"][v]
        accum[vis] += torch.norm(out["viewspace_points"][v].grad[vis, :2], dim=-1, keepdim=True)
        denom[vis] += 1
    assert float(denom.sum()) == float(sum(int((r > 0).sum()) for r in radii_all))


def test_two_renders_before_backward_do_not_share_a_workspace():
    """system/gaussian_zero123.py:212-235 renders twice per step and calls backward afterwards."""
    from b200splat.renderer import B200GaussianBatchRenderer
    P, H, W, V = 3000, 48, 48, 2
    sc = scenes.make_scene(P, 0, 0.8, seed=221)
    batch_a, _ = _batch(V, H, W, 222)
    batch_b, _ = _batch(V, H, W, 223)

    class Ren(B200GaussianBatchRenderer):
        variant = "advanced"

    ren = Ren()
    ren.training, ren.background_tensor = True, torch.ones(3, device="cuda")
    ren.geometry = _Geometry(sc, False, 224)
    out_a = ren.batch_forward(batch_a)
    out_b = ren.batch_forward(batch_b)               # same shape, first graph still alive
    assert len(ren._b200_rasterizers[(V, H, W, "cuda:0")]) == 2
    (out_a["comp_rgb"].sum() + out_b["comp_rgb"].sum()).backward()
    g_two = ren.geometry.get_xyz.grad.clone()
    # the same two renders one after the other
    ren.geometry.get_xyz.grad = None
    ren.batch_forward(batch_a)["comp_rgb"].sum().backward()
    ren.batch_forward(batch_b)["comp_rgb"].sum().backward()
    assert rel_err(g_two, ren.geometry.get_xyz.grad) <= 1e-4
    assert all(not r.pending for r in ren._b200_rasterizers[(V, H, W, "cuda:0")])
    with torch.no_grad():                             # inference never holds a workspace
        for _ in range(6):
            ren.batch_forward(batch_a)
    assert len(ren._b200_rasterizers[(V, H, W, "cuda:0")]) == 2


def test_device_cameras_match_the_reference_construction_and_batch_forward_never_synchronises():
    """(1) ``_camera_batch`` (batched device ops) == get_cam_info_gaussian per view on the host (restated in
    b200splat.scenes.cam_info_gaussian; call site renderer/gaussian_batch_renderer.py:23-49).  (2) After a warm-up
    call, ``batch_forward`` + backward run with torch's synchronisation debug mode set to "error": no .item() / .cpu() /
    float(tensor) anywhere on the path (the reference's loop has one D2H read of fovy per view,
    renderer/diff_gaussian_rasterizer.py:80-81)."""
    import math
    from b200splat.renderer import B200GaussianBatchRenderer, _camera_batch
    P, H, W, V = 4000, 64, 64, 4
    sc = scenes.make_scene(P, 0, 0.8, seed=231)
    batch, cams = _batch(V, H, W, 232)
    wvt, full, centre, tan = _camera_batch(batch)
    for v in range(V):
        fovy = float(batch["fovy"][v])
        rw, rf, rc = scenes.cam_info_gaussian(batch["c2w"][v].cpu(), fovy, fovy, 0.1, 100.0)
        assert float((wvt[v].cpu() - rw).abs().max()) < 2e-6 and float((full[v].cpu() - rf).abs().max()) < 1e-5
        assert float((centre[v].cpu() - rc).abs().max()) < 1e-5
        assert float(tan[v]) == float(torch.tensor(math.tan(fovy * 0.5), dtype=torch.float32))

    class Ren(B200GaussianBatchRenderer):
        variant = "shading"

    ren = Ren()
    ren.training, ren.background_tensor = True, torch.ones(3, device="cuda")
    ren.geometry = _Geometry(sc, True, 233)
    bg_map = torch.rand(V, H, W, 3, device="cuda")
    ren.background = lambda dirs: bg_map
    ren.material = types.SimpleNamespace(ambient_light_color=torch.tensor([0.1, 0.1, 0.1], device="cuda"),
                                         diffuse_light_color=torch.tensor([0.9, 0.9, 0.9], device="cuda"),
                                         ambient_only=False, training=True,
                                         cfg=types.SimpleNamespace(diffuse_prob=0.75, textureless_prob=0.5, soft_shading=True))
    out = ren.batch_forward(batch)                   # warm-up: workspaces, pinned notice words, light-colour cache
    (out["comp_rgb"].sum() + out["comp_depth"].sum()).backward()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        out = ren.batch_forward(batch)
        (out["comp_rgb"].sum() + out["comp_depth"].sum()).backward()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    assert torch.isfinite(out["comp_rgb"]).all()
