import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/threestudio-3dgs_b200"); sys.path.insert(0, "/root/repo/tests")
from b200splat import scenes, ops
from oracle import torch_oracle as O
from util import *
import test_parity_gpu as T
P, deg, H, W, seed = 16384, 0, 128, 128, 1235
sc, cam = T._scene(P, deg, H, W, seed)
s = oracle_settings(cam, deg)
grads = scenes.pixel_grads(H, W, seed + 1)
orc = T._run_oracle(sc, s, grads=grads)
bad = borderline_pixels(orc["pre"], orc["binned"], s, orc["out"])
print("borderline px", int(bad.sum()))
cu = T._run_cuda(sc, s, grads=grads)
for k, ref in orc["grads"].items():
    if k == "stage" or ref is None: continue
    got = cu["grads"][k].cpu()
    d = (got - ref).abs()
    i = int(d.reshape(d.shape[0], -1).max(1).values.argmax())
    print(k, "rel", rel_err(got, ref), "max ref", float(ref.abs().max()), "worst gaussian", i, "err", float(d.reshape(d.shape[0],-1)[i].max()), "ref there", float(ref.reshape(ref.shape[0],-1)[i].abs().max()))
camc = ops.make_cam(cuda_settings(s), "cuda")
d_ = lambda t: t.cuda().contiguous()
color, radii, depth, alpha, st = ops.forward(camc, d_(sc.means3D), d_(sc.shs), None, d_(sc.opacities), d_(sc.scales), d_(sc.rotations), None)
v = ops.forward_views(camc, st)
nc = v["n_contrib"].cpu(); onc = orc["out"]["n_contrib"]
mm = (nc != onc)
print("n_contrib mismatches", int(mm.sum()), "of which borderline", int((mm & bad).sum()))
# masked-grad experiment
keep = (~bad).float()
g2 = (grads[0]*keep, grads[1]*keep, grads[2]*keep)
orc2 = T._run_oracle(sc, s, grads=g2); cu2 = T._run_cuda(sc, s, grads=g2)
for k, ref in orc2["grads"].items():
    if k == "stage" or ref is None: continue
    print("masked", k, rel_err(cu2["grads"][k], ref))
