"""The reference's UNCHANGED files -- renderer/diff_gaussian_rasterizer.py (``DiffGaussian.forward``),
renderer/gaussian_batch_renderer.py (``batch_forward``: the per-view loop) and geometry/gaussian_base.py
(``GaussianBaseModel``: activations, ``update_states``, optimizer) -- running on the CUDA backend: the product packages
``diff_gaussian_rasterization`` / ``simple_knn`` under the reference's own import names, on a GPU.

The GPU box has no /root/reference; scripts/stage_reference_for_gpu.py stages the imported files (unmodified, git-ignored)
under tests/_refcopy/ before the snapshot is taken.  Skipped when neither exists."""
import warnings

import pytest
import torch

import ref_harness as H
from b200splat import scenes
from util import borderline_bounds, check_images, oracle_settings, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not H.available(), reason="reference files not staged")]


def _inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def _geometry(geo_mod, scene):
    geo = geo_mod.GaussianBaseModel({"init_num_pts": 4096, "sh_degree": 0, "pc_init_radius": 0.8})
    assert geo._xyz.is_cuda and geo._xyz.shape == (4096, 3) and torch.isfinite(geo._scaling).all()   # distCUDA2 init
    P = scene.means3D.shape[0]
    dev = "cuda"
    geo._xyz = torch.nn.Parameter(scene.means3D.to(dev).clone())
    geo._scaling = torch.nn.Parameter(torch.log(scene.scales).to(dev))
    geo._rotation = torch.nn.Parameter(scene.rotations.to(dev).clone())
    geo._opacity = torch.nn.Parameter(_inverse_sigmoid(scene.opacities).to(dev))
    geo._features_dc = torch.nn.Parameter(scene.shs[:, :1].to(dev).clone())
    geo._features_rest = torch.nn.Parameter(scene.shs[:, 1:].to(dev).clone())
    geo.max_radii2D = torch.zeros(P, device=dev)
    geo.training_setup()
    geo.update_learning_rate(0)
    return geo


@pytest.mark.timeout(600)
def test_reference_batch_forward_and_update_states_on_the_cuda_backend():
    from oracle import torch_oracle as O
    warnings.filterwarnings("ignore")
    ren_mod, geo_mod = H.load("cuda")
    import diff_gaussian_rasterization
    assert "b200" in diff_gaussian_rasterization.__file__ or "threestudio-3dgs_b200" in diff_gaussian_rasterization.__file__
    V = 4
    scene, cams = scenes.make_workload("config2_100k_512_sh0_b4", views=V)     # BASELINE.json configs[1]
    P = scene.means3D.shape[0]
    geo = _geometry(geo_mod, scene)
    ren = ren_mod.DiffGaussian({}, geometry=geo, material=None, background=None)
    ren.train()
    c2ws = []
    for cam in cams:
        c2w = torch.inverse(cam.viewmatrix.t())
        c2w[:3, 1:3] *= -1
        c2ws.append(c2w)
    H_, W_ = cams[0].image_height, cams[0].image_width
    batch = {"c2w": torch.stack(c2ws).cuda(), "fovy": torch.tensor([c.fovy for c in cams]).cuda(), "width": W_,
             "height": H_}
    out = ren.batch_forward(batch)                      # the reference's unchanged per-view loop
    assert out["comp_rgb"].shape == (V, H_, W_, 3) and out["comp_rgb"].is_cuda
    gcs = [scenes.pixel_grads(H_, W_, 700 + v)[0] for v in range(V)]
    loss = sum((out["comp_rgb"][v].permute(2, 0, 1) * gcs[v].cuda()).sum() for v in range(V))
    loss.backward()
    # every view against the oracle: the image (clamped like the renderer does), radii, and the view-space gradient
    import math
    from threestudio.utils.ops import get_cam_info_gaussian          # the stub the reference files import (restated)
    for v, cam in enumerate(cams):
        # the camera exactly as the reference's loop derives it from the batch (renderer/gaussian_batch_renderer.py:
        # 23-26, renderer/diff_gaussian_rasterizer.py:80-96): c2w -> (world_view, full_proj, centre), tan(fovy / 2)
        fovy = batch["fovy"][v].cpu()
        wvt, full, centre = get_cam_info_gaussian(c2w=batch["c2w"][v].cpu(), fovx=fovy, fovy=fovy, znear=0.1, zfar=100)
        tanf = math.tan(fovy * 0.5)
        s = O.Settings(H_, W_, tanf, tanf, torch.ones(3), 1.0, wvt, full, 0, centre, False, False)
        o, pre, binned = O.rasterize_forward(scene.means3D, None, scene.shs, None, scene.opacities, scene.scales,
                                             scene.rotations, None, s)
        assert torch.equal(out["radii"][v].cpu(), pre["radii"])
        assert torch.equal(out["visibility_filter"][v].cpu(), pre["radii"] > 0)
        b = borderline_bounds(pre, binned, s, o)
        comp = out["comp_rgb"][v].permute(2, 0, 1).detach().cpu()
        err = (comp - o["color"].clamp(0, 1)).abs()
        assert float((err - (1e-4 + b["color"][None])).max()) <= 0.0
        vsp = out["viewspace_points"][v]
        assert vsp.grad is not None and vsp.grad.shape == (P, 3) and float(vsp.grad[:, 2].abs().max()) == 0.0
    assert float(geo._xyz.grad.abs().max()) > 0 and torch.isfinite(geo._xyz.grad).all()
    # the unchanged densification statistics (geometry/gaussian_base.py:815-819, :845-851)
    geo.update_states(1, out["visibility_filter"], out["radii"], out["viewspace_points"])
    den = sum((r > 0).float() for r in out["radii"])
    assert torch.equal(geo.denom[:, 0], den)
    acc = sum(torch.where(r > 0, vp.grad[:, :2].norm(dim=-1), torch.zeros_like(den))
              for r, vp in zip(out["radii"], out["viewspace_points"]))
    assert rel_err(geo.xyz_gradient_accum[:, 0], acc) < 1e-6
    assert torch.equal(geo.max_radii2D, torch.stack([r.float() for r in out["radii"]]).max(0).values)
    # one optimizer step of the reference's Adam groups on the gradients the operator produced
    before = geo._xyz.detach().clone()
    geo.optimizer.step()
    assert float((geo._xyz.detach() - before).abs().max()) > 0
    # two backward passes over one graph (system/gaussian_splatting.py:129,137-138); eval mode (inverted background)
    out2 = ren.batch_forward(batch)
    out2["comp_rgb"].sum().backward(retain_graph=True)
    out2["comp_rgb"].mean().backward()
    ren.eval()
    with torch.no_grad():
        out3 = ren.batch_forward(batch)
    assert torch.isfinite(out3["comp_rgb"]).all()
