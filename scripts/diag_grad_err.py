"""Where does the gradient error of a parity case come from?  (diagnostic, GPU)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "threestudio-3dgs_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch
from b200splat import scenes, ops
from oracle import torch_oracle as O
from oracle.checks import borderline_bounds
from util import oracle_settings, cuda_settings

P, res, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
scene = scenes.make_scene(P, 3, 0.5, seed=seed)
cam = scenes.mvdream_cameras(1, res, res, seed=seed + 1)[0]
s = oracle_settings(cam, 3)
grads = scenes.pixel_grads(res, res, 5)
out, pre, binned = O.rasterize_forward(scene.means3D, None, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None, s)
ref = O.rasterize_backward((scene.means3D, None, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None), s, pre, binned, out, *grads)
camc = ops.make_cam(cuda_settings(s), "cuda")
d = lambda t: t.cuda().contiguous()
m3, sh, op, scl, rot = map(d, (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations))
color, radii, depth, alpha, st = ops.forward(camc, m3, sh, None, op, scl, rot, None)
g = ops.backward(camc, st, m3, sh, None, op, scl, rot, None, radii, alpha, *(t.cuda() for t in grads))
bb = borderline_bounds(pre, binned, s, out)
print("borderline frac", float(bb["mask"].float().mean()))
v = ops.forward_views(camc, st)
ncd = (v["n_contrib"].cpu().long() != out["n_contrib"].long())
print("n_contrib mismatches", int(ncd.sum()), "outside mask", int((ncd & ~bb["mask"]).sum()))
for k in ("means3D", "means2D", "shs", "opacities", "scales", "rotations"):
    a, b = g[k].cpu().double(), ref[k].double()
    err = (a - b).abs()
    flat = err.reshape(err.shape[0], -1).max(1).values
    top = torch.topk(flat, 5)
    print(k, "rel", float(err.max() / b.abs().max()), "ref max", float(b.abs().max()))
    for e, i in zip(top.values.tolist(), top.indices.tolist()):
        print("   gauss", i, "err", e, "ref", b[i].reshape(-1).abs().max().item(), "radius", int(pre["radii"][i]),
              "opac", float(scene.opacities[i]), "scales", scene.scales[i].tolist(), "px,py", float(pre["px"][i]), float(pre["py"][i]))
# the 2D-stage gradients of the worst Gaussian
