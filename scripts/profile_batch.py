"""A few batched steps (V views fwd+bwd) of a workload: the short command profiled under ncu."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "threestudio-3dgs_b200"))
import torch
from b200splat import scenes, ops, batched

name = sys.argv[1] if len(sys.argv) > 1 else "headline_1m_512_sh3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
V = int(sys.argv[3]) if len(sys.argv) > 3 else 4
scene, cams_h = scenes.make_workload(name, views=V)
dev = torch.device("cuda", 0)
to = lambda t: t.to(dev).contiguous()
m3, sh, op, sc, ro = map(to, (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations))
class S: pass
cams = []
for c in cams_h:
    s = S()
    s.image_height, s.image_width, s.tanfovx, s.tanfovy = c.image_height, c.image_width, c.tanfovx, c.tanfovy
    s.bg, s.scale_modifier, s.viewmatrix, s.projmatrix = torch.ones(3, device=dev), 1.0, c.viewmatrix, c.projmatrix
    s.sh_degree, s.campos, s.prefiltered, s.debug = scene.sh_degree, c.campos, False, False
    cams.append(ops.make_cam(s, dev))
H, W = cams_h[0].image_height, cams_h[0].image_width
pg = [tuple(to(g) for g in scenes.pixel_grads(H, W, 99 + v)) for v in range(V)]
br = batched.BatchRenderer(m3.shape[0], sh.shape[1], H, W, dev, views=V)
br.calibrate(cams, m3, sh, None, op, sc, ro)
for i in range(iters):
    if i == iters - 1:          # ncu --profile-from-start off captures only the last step
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    br.step(cams, m3, sh, None, op, sc, ro, pg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(br.packed.buffer.abs().sum()), br.overflowed())
