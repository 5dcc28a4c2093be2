"""Localise the gradient error of one Gaussian: per tile, then per pixel (diagnostic, GPU)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "threestudio-3dgs_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch
from b200splat import scenes, ops
from oracle import torch_oracle as O
from util import oracle_settings, cuda_settings

P, res, seed, gid = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
scene = scenes.make_scene(P, 3, 0.5, seed=seed)
cam = scenes.mvdream_cameras(1, res, res, seed=seed + 1)[0]
s = oracle_settings(cam, 3)
grads = scenes.pixel_grads(res, res, 5)
inputs = (scene.means3D, None, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None)
out, pre, binned = O.rasterize_forward(*inputs, s)
camc = ops.make_cam(cuda_settings(s), "cuda")
d = lambda t: t.cuda().contiguous()
m3, sh, op, scl, rot = map(d, (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations))
color, radii, depth, alpha, st = ops.forward(camc, m3, sh, None, op, scl, rot, None)
gx = res // 16
x0, y0 = pre["rect_min"][gid].tolist(); x1, y1 = pre["rect_max"][gid].tolist()
print("rect", x0, y0, x1, y1, "conic", [float(c[gid]) for c in pre["conic"]], "px", float(pre["px"][gid]), float(pre["py"][gid]),
      "depth", float(pre["depth"][gid]))
def masked(mask):
    return tuple(g * mask for g in grads)
def both(mask, tiles):
    mg = masked(mask)
    ref = O.rasterize_backward(inputs, s, pre, binned, out, *mg, tiles=tiles)
    g = ops.backward(camc, st, m3, sh, None, op, scl, rot, None, radii, alpha, *(t.cuda() for t in mg))
    r = torch.cat([ref["means2D"][gid, :2], ref["opacities"][gid].reshape(-1), ref["stage"]["dL_dconic"][gid]])
    c = torch.cat([g["means2D"][gid, :2].cpu(), g["opacities"][gid].reshape(-1).cpu()])
    return r, c
worst = None
for ty in range(y0, y1):
    for tx in range(x0, x1):
        mask = torch.zeros(1, res, res); mask[:, ty*16:ty*16+16, tx*16:tx*16+16] = 1
        r, c = both(mask, [ty * gx + tx])
        e = (r[:3] - c).abs().max().item()
        if worst is None or e > worst[0]:
            worst = (e, tx, ty, r, c)
        if e > 1e-9:
            print("tile", tx, ty, "err", e, "ref", r[:3].tolist(), "cuda", c.tolist())
e, tx, ty, r, c = worst
print("worst tile", tx, ty, e)
t = ty * gx + tx
r0, r1 = binned["ranges"][t].tolist()
pos = (binned["point_list"][r0:r1].long() == gid).nonzero().item()
print("list length", r1 - r0, "position of gid", pos, "n_contrib max in tile", int(out["n_contrib"][ty*16:ty*16+16, tx*16:tx*16+16].max()))
for py in range(16):
    for pxl in range(16):
        mask = torch.zeros(1, res, res); mask[:, ty*16+py, tx*16+pxl] = 1
        r, c = both(mask, [t])
        e = (r[:3] - c).abs().max().item()
        if e > 1e-10:
            Y, X = ty*16+py, tx*16+pxl
            dx, dy = float(pre["px"][gid]) - X, float(pre["py"][gid]) - Y
            ca, cb, cc = [float(cx[gid]) for cx in pre["conic"]]
            power = -0.5*(ca*dx*dx + cc*dy*dy) - cb*dx*dy
            import math
            print("pixel", X, Y, "err", e, "ref", r[:3].tolist(), "cuda", c.tolist(), "power", power,
                  "alpha", float(scene.opacities[gid]) * math.exp(power), "n_contrib", int(out["n_contrib"][Y, X]))
