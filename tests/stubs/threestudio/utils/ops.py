import math

import torch


def _projection(znear, zfar, fovx, fovy):
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    P = torch.zeros(4, 4)
    P[0, 0] = 1.0 / tx
    P[1, 1] = 1.0 / ty
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def get_cam_info_gaussian(c2w, fovx, fovy, znear, zfar):
    """threestudio.utils.ops.get_cam_info_gaussian restated (call site
    renderer/gaussian_batch_renderer.py:24-26): flip the camera y/z axes, invert, transpose."""
    dev = c2w.device
    c2w = c2w.detach().clone().float().cpu()
    c2w[:3, 1:3] *= -1
    w2c = torch.inverse(c2w)
    wvt = w2c.transpose(0, 1).contiguous()
    proj = _projection(znear, zfar, float(fovx), float(fovy)).transpose(0, 1)
    full = wvt.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0).contiguous()
    center = wvt.inverse()[3, :3].contiguous()
    return wvt.to(dev), full.to(dev), center.to(dev)


def dot(x, y):
    return torch.sum(x * y, -1, keepdim=True)


def get_activation(name):
    """threestudio.utils.ops.get_activation restated for the names the reference's material uses
    (material/gaussian_material.py:9, :115): only needed so the file imports."""
    if name is None or str(name).lower() == "none":
        return lambda x: x
    name = str(name).lower()
    table = {"sigmoid": torch.sigmoid, "tanh": torch.tanh, "exp": torch.exp, "relu": torch.relu,
             "softplus": torch.nn.functional.softplus}
    if name not in table:
        raise ValueError(f"unknown activation {name}")
    return table[name]
