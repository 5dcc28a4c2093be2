"""GPU: fused activation-backward + Adam (csrc/adam.cu through b200splat.optim -> C ABI) against the CPU oracle
(oracle/optim.py: autograd through the reference's activations + torch.optim.Adam with the reference's groups).
Tolerance: 1e-5 relative on parameters and optimizer states after several steps with changing learning rates."""
import pytest
import torch

from oracle.optim import GROUPS, OracleGaussianAdam
from util import rel_err

pytestmark = pytest.mark.gpu


def _raw(P, M, seed):
    g = torch.Generator().manual_seed(seed)
    return dict(xyz=torch.randn(P, 3, generator=g), f_dc=torch.randn(P, 1, 3, generator=g) * 1.5,
                f_rest=torch.randn(P, M - 1, 3, generator=g) * 0.05, opacity=torch.randn(P, 1, generator=g) * 2,
                scaling=torch.randn(P, 3, generator=g) - 3.0, rotation=torch.randn(P, 4, generator=g))


@pytest.mark.parametrize("P,M,clip", [(5000, 16, 2.0), (777, 1, float("inf")), (1, 4, 0.5)])
def test_fused_adam_matches_oracle(P, M, clip):
    from b200splat.optim import FusedGaussianAdam
    raw = _raw(P, M, 31)
    lrs = dict(xyz=1.6e-4, f_dc=2.5e-3, f_rest=2.5e-3 / 20, opacity=5e-2, scaling=5e-3, rotation=1e-3)
    orc = OracleGaussianAdam(raw, lrs, eps=1e-15, color_clip=clip)
    dev = {k: v.clone().cuda() for k, v in raw.items()}
    opt = FusedGaussianAdam(dev, lrs, eps=1e-15, color_clip=clip)
    g = torch.Generator().manual_seed(32)
    for it in range(5):
        lrs = {k: v * (0.9 ** it) for k, v in lrs.items()}
        orc.set_lrs(lrs)
        opt.lrs = dict(lrs)
        grads = dict(means3D=torch.randn(P, 3, generator=g), shs=torch.randn(P, M, 3, generator=g),
                     opacities=torch.randn(P, 1, generator=g), scales=torch.randn(P, 3, generator=g),
                     rotations=torch.randn(P, 4, generator=g))
        if it == 2:
            grads["means3D"].zero_()          # zero gradients keep decaying the moments
        orc.step(grads)
        opt.step({k: v.cuda() for k, v in grads.items()})
    for k in GROUPS:
        if raw[k].numel() == 0:
            continue
        m, v = orc.state(k)
        assert rel_err(dev[k], orc.p[k]) <= 1e-5, f"param {k}: {rel_err(dev[k], orc.p[k])}"
        assert rel_err(opt.exp_avg[k], m) <= 1e-5, f"exp_avg {k}"
        assert rel_err(opt.exp_avg_sq[k], v) <= 1e-5, f"exp_avg_sq {k}"


def test_fused_adam_consumes_the_packed_gradient_buffer():
    """The step reads the rasterizer's packed buffer in place (the all-reduce payload), no repacking."""
    from b200splat.batched import PackedGrads
    from b200splat.optim import FusedGaussianAdam
    P, M = 3000, 4
    raw = _raw(P, M, 41)
    lrs = dict(xyz=1e-3, f_dc=1e-3, f_rest=1e-4, opacity=1e-2, scaling=1e-3, rotation=1e-3)
    packed = PackedGrads(P, M, "cuda")
    g = torch.Generator().manual_seed(42)
    packed.buffer.copy_(torch.randn(packed.buffer.numel(), generator=g).cuda())
    orc = OracleGaussianAdam(raw, lrs)
    orc.step({k: v.cpu() for k, v in packed.grads().items()})
    dev = {k: v.clone().cuda() for k, v in raw.items()}
    FusedGaussianAdam(dev, lrs).step(packed.grads())
    for k in GROUPS:
        assert rel_err(dev[k], orc.p[k]) <= 1e-5, k


def test_fused_adam_refuses_cpu_tensors():
    from b200splat.optim import FusedGaussianAdam
    with pytest.raises(RuntimeError):
        FusedGaussianAdam(_raw(4, 1, 1), dict.fromkeys(GROUPS, 1e-3))
