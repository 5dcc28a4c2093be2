"""CPU: the post-op oracle (oracle/postops.py) against golden vectors produced by the reference's own renderer /
material files (tests/golden/make_postops_golden.py), and -- where /root/reference exists -- against those files
run live."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import postops as PO

GOLDEN = Path(__file__).parent / "golden" / "postops_ref.npz"
sys.path.insert(0, str(GOLDEN.parent))
import make_postops_golden as MK  # noqa: E402

MODES = {"shading": PO.MODE_SHADING, "normal": PO.MODE_NORMAL, "background": PO.MODE_BACKGROUND}


def oracle_case(inp, variant, shading, pred):
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("image", "depth", "alpha", "bg")}
    out = PO.postprocess_view(MODES[variant], leaves["image"], leaves["depth"], leaves["alpha"], inp["rays_o"],
                              inp["rays_d"], leaves["bg"], inp["light"], torch.tensor([0.1, 0.1, 0.1]),
                              torch.tensor([0.9, 0.9, 0.9]), shading, inp["pred"] if pred else None)
    loss = (out["render"] * inp["g_render"]).sum()
    if out["normal"] is not None:
        loss = loss + (out["normal"] * inp["g_normal"]).sum() + (out["depth"] * inp["g_depth"]).sum()
    loss.backward()
    res = {k: v.detach() for k, v in out.items() if v is not None}
    for k, t in leaves.items():
        res["d_" + k] = t.grad if t.grad is not None else torch.zeros_like(t)
    return res


def _close(a, b, what, tol=2e-5):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = float((a - b).abs().max() / b.abs().max().clamp_min(1.0))
    assert err <= tol, f"{what}: {err}"


@pytest.mark.parametrize("variant,shading,pred", MK.CASES)
def test_postops_oracle_matches_reference_golden(variant, shading, pred):
    blob = np.load(GOLDEN)
    inp = {k[3:]: torch.from_numpy(blob[k]) for k in blob.files if k.startswith("in_")}
    res = oracle_case(inp, variant, shading, pred)
    name = MK.case_name(variant, shading, pred)
    keys = [k.split("__")[1] for k in blob.files if k.startswith(name + "__")]
    assert "render" in keys and "d_image" in keys
    for k in keys:
        if variant == "background" and k in ("depth", "normal"):
            continue
        _close(res[k], blob[f"{name}__{k}"], f"{name}.{k}")


@pytest.mark.skipif(not Path("/root/reference/renderer/diff_gaussian_rasterizer_shading.py").exists(),
                    reason="/root/reference not present")
def test_golden_is_what_the_reference_files_produce_now():
    blob = np.load(GOLDEN)
    inp = MK.make_inputs(4321)
    for k, v in inp.items():
        assert np.array_equal(blob["in_" + k], v.numpy())
    for variant, shading, pred in (("shading", "diffuse", False), ("normal", "diffuse", False)):
        res = MK.run_case(variant, inp, shading, pred)
        for k, v in res.items():
            _close(v, blob[f"{MK.case_name(variant, shading, pred)}__{k}"], k, tol=1e-6)
