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


def build():
    scene, cams = scenes.make_workload("config1_16k_128_sh0", views=1)
    cam = cams[0]
    s = oracle_settings(cam, 0)
    H, W = cam.image_height, cam.image_width
    grads = scenes.pixel_grads(H, W, 2024)
    out, pre, binned = O.rasterize_forward(scene.means3D, None, scene.shs, None, scene.opacities, scene.scales,
                                           scene.rotations, None, s)
    g = O.rasterize_backward((scene.means3D, None, scene.shs, None, scene.opacities, scene.scales, scene.rotations,
                              None), s, pre, binned, out, *grads)
    # pixels whose blend decisions sit on a hard cut-off, with the error a flipped decision may cause (tests/util.py):
    # the CUDA path must stay within IMG_TOL + bound there and within IMG_TOL everywhere else
    bb = borderline_bounds(pre, binned, s, out)
    return dict(
        borderline_mask=bb["mask"].numpy(), bound_color=bb["color"].numpy(), bound_depth=bb["depth"].numpy(),
        bound_alpha=bb["alpha"].numpy(),
        inputs_sha256=np.array(digest(torch.cat([scene.means3D.reshape(-1), scene.scales.reshape(-1),
                                                 scene.rotations.reshape(-1), scene.opacities.reshape(-1),
                                                 scene.shs.reshape(-1), cam.viewmatrix.reshape(-1),
                                                 cam.projmatrix.reshape(-1)]))),
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
