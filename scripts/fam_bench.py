"""Per-kernel-family device times of the batched step (CUDA events inside the library): quick A/B of kernel variants.
    python scripts/fam_bench.py [workload] [views] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "threestudio-3dgs_b200"))
import torch
from b200splat import _lib, batched, ops, scenes

name = sys.argv[1] if len(sys.argv) > 1 else "headline_1m_512_sh3"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
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
for _ in range(5):
    br.step(cams, m3, sh, None, op, sc, ro, pg)
g = br.capture_step(cams, m3, sh, None, op, sc, ro, pg)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(steps):
    g.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
_lib.profile_enable(True)
for _ in range(steps):
    br.step(cams, m3, sh, None, op, sc, ro, pg)
torch.cuda.synchronize()
prof = _lib.profile_read(); _lib.profile_enable(False)
tag = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("B200SPLAT_"))
print(f"[{name} V={V} {tag}] step {ms*1e3:.0f} us = {V/ms*1e3:.0f} renders/s | " +
      " ".join(f"{k} {t/max(n,1)/V*1e3:.1f}" for k, (t, n) in prof.items() if n), "| overflow", br.overflowed())
