"""Generate tests/golden/config1_oracle.npz: outputs of the CPU oracle on BASELINE.json configs[0]
(16K random Gaussians, 128x128, SH degree 0, seed 1234) -- the golden vectors the CPU tests re-check the
oracle against and the GPU tests check the CUDA path against.

    python tests/golden/make_golden.py

The reference itself holds no golden vectors for this path, and its CUDA rasterizer is not
vendored, so these are oracle outputs (self-pinned by oracle/dense_f64.py; tests/test_oracle_cpu.py).
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "threestudio-3dgs_b200")); sys.path.insert(0, str(ROOT / "tests"))
from b200splat import scenes  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402
from util import borderline_bounds, oracle_settings  # noqa: E402


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


INPUT_KEYS = ("in_means3D", "in_scales", "in_rotations", "in_opacities", "in_shs", "in_viewmatrix", "in_projmatrix",
              "in_campos", "in_cam_scalars", "in_grad_color", "in_grad_depth", "in_grad_alpha")


def make_inputs():
    """The synthetic inputs of BASELINE.json configs[0].  torch's CPU randn / exp / pow differ in the last ulp between
    hosts (vector ISA), so the fixture CARRIES the inputs it was made from and every consumer reads them back
    (``load_inputs``) instead of regenerating them."""
    scene, cams = scenes.make_workload("config1_16k_128_sh0", views=1)
    cam = cams[0]
    gc, gd, ga = scenes.pixel_grads(cam.image_height, cam.image_width, 2024)
    return dict(in_means3D=scene.means3D.numpy(), in_scales=scene.scales.numpy(), in_rotations=scene.rotations.numpy(),
                in_opacities=scene.opacities.numpy(), in_shs=scene.shs.numpy(), in_viewmatrix=cam.viewmatrix.numpy(),
                in_projmatrix=cam.projmatrix.numpy(), in_campos=cam.campos.numpy(),
                in_cam_scalars=np.array([cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, cam.fovy],
                                        dtype=np.float64),
                in_grad_color=gc.numpy(), in_grad_depth=gd.numpy(), in_grad_alpha=ga.numpy())


def load_inputs(blob):
    """(scene, camera, (dL/dcolor, dL/ddepth, dL/dalpha)) from a fixture (np.load result or dict)."""
    t = lambda k: torch.from_numpy(np.array(blob[k]))
    scene = scenes.Scene(t("in_means3D"), t("in_scales"), t("in_rotations"), t("in_opacities"), t("in_shs"), 0)
    h, w, tx, ty, fovy = (float(x) for x in blob["in_cam_scalars"])
    cam = scenes.Camera(int(h), int(w), tx, ty, t("in_viewmatrix"), t("in_projmatrix"), t("in_campos"), fovy)
    return scene, cam, (t("in_grad_color"), t("in_grad_depth"), t("in_grad_alpha"))


def build(inputs=None):
    """Oracle outputs for ``inputs`` (a fixture's own inputs when re-checking it; fresh ones when writing it)."""
    inputs = make_inputs() if inputs is None else {k: np.array(inputs[k]) for k in INPUT_KEYS}
    scene, cam, grads = load_inputs(inputs)
    s = oracle_settings(cam, 0)
    out, pre, binned = O.rasterize_forward(scene.means3D, None, scene.shs, None, scene.opacities, scene.scales,
                                           scene.rotations, None, s)
    g = O.rasterize_backward((scene.means3D, None, scene.shs, None, scene.opacities, scene.scales, scene.rotations,
                              None), s, pre, binned, out, *grads)
    # pixels whose blend decisions sit on a hard cut-off, with the error a flipped decision may cause (oracle/checks.py):
    # the CUDA path must stay within IMG_TOL + bound there and within IMG_TOL everywhere else
    bb = borderline_bounds(pre, binned, s, out)
    return dict(
        inputs,
        borderline_mask=bb["mask"].numpy(), bound_color=bb["color"].numpy(), bound_depth=bb["depth"].numpy(),
        bound_alpha=bb["alpha"].numpy(),
        color=out["color"].numpy(), depth=out["depth"].numpy(), alpha=out["alpha"].numpy(),
        n_contrib=out["n_contrib"].numpy().astype(np.int32), radii=pre["radii"].numpy().astype(np.int32),
        tiles_touched=pre["tiles_touched"].numpy().astype(np.int32),
        num_rendered=np.array(binned["num_rendered"]), ranges=binned["ranges"].numpy().astype(np.int32),
        keys_sorted_sha256=np.array(digest(binned["keys_sorted"])),
        point_list_sha256=np.array(digest(binned["point_list"])),
        g_means3D=g["means3D"].numpy(), g_means2D=g["means2D"].numpy(), g_shs=g["shs"].numpy(),
        g_opacities=g["opacities"].numpy(), g_scales=g["scales"].numpy(), g_rotations=g["rotations"].numpy(),
    )


if __name__ == "__main__":
    d = build()
    p = Path(__file__).with_name("config1_oracle.npz")
    np.savez_compressed(p, **d)
    print("wrote", p, p.stat().st_size, "bytes; R =", int(d["num_rendered"]))
