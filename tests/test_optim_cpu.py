"""CPU: the optimizer oracle (oracle/optim.py) against the reference's own GaussianBaseModel -- its getters, its
``training_setup`` Adam groups and its ``update_learning_rate`` schedule, imported unchanged from /root/reference."""
import warnings

import pytest
import torch

import ref_harness as H
from oracle.optim import GROUPS, OracleGaussianAdam

pytestmark = pytest.mark.skipif(not H.available(), reason="/root/reference not present")


def test_oracle_adam_matches_reference_geometry_optimizer():
    warnings.filterwarnings("ignore")
    _, geo_mod = H.load("oracle")
    g = torch.Generator().manual_seed(11)
    with H.CudaToCpu():
        geo = geo_mod.GaussianBaseModel({"init_num_pts": 256, "sh_degree": 2, "pc_init_radius": 0.8})
        geo._features_rest.data.normal_(0, 0.05, generator=g)
        geo._rotation.data.normal_(0, 1, generator=g)
        geo._features_dc.data.mul_(3.0)           # some DC values beyond the colour clip
        geo.training_setup()
        raw = dict(xyz=geo._xyz, f_dc=geo._features_dc, f_rest=geo._features_rest, opacity=geo._opacity,
                   scaling=geo._scaling, rotation=geo._rotation)
        orc = None
        for it in range(1, 4):
            geo.update_learning_rate(it)
            lrs = {grp["name"]: grp["lr"] for grp in geo.optimizer.param_groups}
            if orc is None:
                orc = OracleGaussianAdam(raw, lrs, eps=1e-15, color_clip=geo.color_clip)
            orc.set_lrs(lrs)
            orc.color_clip = geo.color_clip
            act = dict(means3D=geo.get_xyz, shs=geo.get_features, opacities=geo.get_opacity, scales=geo.get_scaling,
                       rotations=geo.get_rotation)
            grads = {k: torch.randn(v.shape, generator=g) for k, v in act.items()}
            geo.optimizer.zero_grad(set_to_none=True)
            torch.autograd.backward(list(act.values()), [grads[k] for k in act])
            geo.optimizer.step()
            orc.step(grads)
            for k in GROUPS:
                assert torch.allclose(orc.p[k].detach(), raw[k].detach(), rtol=1e-6, atol=1e-7), (it, k)
