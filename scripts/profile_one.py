"""One-view fwd+bwd of a workload, a few iterations: the short command profiled under ncu."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "threestudio-3dgs_b200"))
import torch
from b200splat import scenes, ops, batched

name = sys.argv[1] if len(sys.argv) > 1 else "headline_1m_512_sh3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene, cams = scenes.make_workload(name, views=1)
dev = torch.device("cuda", 0)
to = lambda t: t.to(dev).contiguous()
m3, sh, op, sc, ro = map(to, (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations))
c = cams[0]
class S: pass
s = S()
s.image_height, s.image_width, s.tanfovx, s.tanfovy = c.image_height, c.image_width, c.tanfovx, c.tanfovy
s.bg, s.scale_modifier, s.viewmatrix, s.projmatrix = torch.ones(3, device=dev), 1.0, c.viewmatrix, c.projmatrix
s.sh_degree, s.campos, s.prefiltered, s.debug = scene.sh_degree, c.campos, False, False
cam = ops.make_cam(s, dev)
pg = [tuple(to(g) for g in scenes.pixel_grads(c.image_height, c.image_width, 99))]
packed = batched.PackedGrads(m3.shape[0], sh.shape[1], dev)
for _ in range(iters):
    batched.render_views_fwd_bwd([cam], m3, sh, None, op, sc, ro, pg, packed)
torch.cuda.synchronize()
print("ok", float(packed.buffer.abs().sum()))
