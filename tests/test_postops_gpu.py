"""GPU parity of the fused post-op kernels (csrc/postops.cu through b200splat.postops -> C ABI) against
(1) golden vectors produced by the reference's own renderer / material files and (2) the CPU oracle on larger
multi-view inputs.  Bars: outputs max-abs <= 1e-4 (values are O(1)); gradients <= 1e-3 relative."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import postops as PO
from util import rel_err

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).parent / "golden" / "postops_ref.npz"
sys.path.insert(0, str(GOLDEN.parent))
import make_postops_golden as MK  # noqa: E402

OUT_TOL, GRAD_TOL = 1e-4, 1e-3


def _cuda_case(inps, variant, shading, pred):
    """inps: list of per-view input dicts (CPU).  Returns stacked outputs and gradients (CPU)."""
    from b200splat.postops import postprocess_views
    st = lambda k: torch.stack([i[k] for i in inps]).cuda()
    leaves = {k: st(k).requires_grad_(True) for k in ("image", "depth", "alpha", "bg")}
    out = postprocess_views(variant, leaves["image"], leaves["depth"], leaves["alpha"], bg=leaves["bg"],
                            rays_o=st("rays_o"), rays_d=st("rays_d"), light_positions=st("light"),
                            pred_normal=st("pred") if pred else None, shading=shading)
    loss = (out["render"] * st("g_render")).sum()
    if out["normal"] is not None:
        loss = loss + (out["normal"] * st("g_normal")).sum() + (out["depth"] * st("g_depth")).sum()
    loss.backward()
    res = {k: v.detach().cpu() for k, v in out.items() if v is not None}
    for k, t in leaves.items():
        res["d_" + k] = (t.grad if t.grad is not None else torch.zeros_like(t)).cpu()
    return res


@pytest.mark.parametrize("variant,shading,pred", MK.CASES)
def test_postops_cuda_matches_reference_golden(variant, shading, pred):
    blob = np.load(GOLDEN)
    inp = {k[3:]: torch.from_numpy(blob[k]) for k in blob.files if k.startswith("in_")}
    res = _cuda_case([inp], variant, shading, pred)
    name = MK.case_name(variant, shading, pred)
    for key in [k.split("__")[1] for k in blob.files if k.startswith(name + "__")]:
        ref = torch.from_numpy(blob[f"{name}__{key}"])
        got = res[key][0]
        if key.startswith("d_"):
            assert rel_err(got, ref) <= GRAD_TOL, f"{name}.{key}: {rel_err(got, ref)}"
        else:
            assert float((got - ref).abs().max()) <= OUT_TOL, f"{name}.{key}"


@pytest.mark.parametrize("variant,shading,pred", MK.CASES + [("plain", "diffuse", False)])
def test_postops_cuda_matches_oracle_multi_view(variant, shading, pred):
    V, H, W = 3, 70, 93                      # not multiples of the 32x8 block
    inps = [MK.make_inputs(100 + v, H, W) for v in range(V)]
    res = _cuda_case(inps, variant, shading, pred)
    mode = {"plain": PO.MODE_PLAIN, "background": PO.MODE_BACKGROUND, "normal": PO.MODE_NORMAL,
            "shading": PO.MODE_SHADING}[variant]
    for v, inp in enumerate(inps):
        leaves = {k: inp[k].clone().requires_grad_(True) for k in ("image", "depth", "alpha", "bg")}
        out = PO.postprocess_view(mode, leaves["image"], leaves["depth"], leaves["alpha"], inp["rays_o"], inp["rays_d"],
                                  leaves["bg"], inp["light"], torch.tensor([0.1] * 3), torch.tensor([0.9] * 3), shading,
                                  inp["pred"] if pred else None)
        loss = (out["render"] * inp["g_render"]).sum()
        if out["normal"] is not None:
            loss = loss + (out["normal"] * inp["g_normal"]).sum() + (out["depth"] * inp["g_depth"]).sum()
        loss.backward()
        for key in ("render", "normal", "depth"):
            if out[key] is not None and key in res:
                assert float((res[key][v] - out[key].detach()).abs().max()) <= OUT_TOL, (variant, key, v)
        for key in ("image", "depth", "alpha", "bg"):
            ref = leaves[key].grad
            if ref is None:
                assert float(res["d_" + key][v].abs().max()) == 0.0
            else:
                assert rel_err(res["d_" + key][v], ref) <= GRAD_TOL, (variant, key, v, rel_err(res["d_" + key][v], ref))


def test_postops_refuses_cpu_tensors():
    from b200splat.postops import postprocess_views
    with pytest.raises(RuntimeError):
        postprocess_views("plain", torch.zeros(1, 3, 4, 4), torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))
