"""The kernels either side of the rasterizer on the headline shapes, inside a cudaProfilerStart/Stop window:
one batched step with 3 extra feature channels, the fused post-ops (shading mode, forward + backward) and one fused
Adam step.   ncu --set full --clock-control none --import-source on --profile-from-start off -o <rep> python scripts/profile_aux.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "threestudio-3dgs_b200"))
import torch
from b200splat import scenes, ops, batched
from b200splat.optim import FusedGaussianAdam
from b200splat.postops import postprocess_views

V = 4
scene, cams_h = scenes.make_workload("headline_1m_512_sh3", views=V)
dev = torch.device("cuda", 0)
to = lambda t: t.to(dev).contiguous()
m3, sh, op, sc, ro = map(to, (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations))
P, M = m3.shape[0], sh.shape[1]
class S: pass
cams = []
for c in cams_h:
    s = S()
    s.image_height, s.image_width, s.tanfovx, s.tanfovy = c.image_height, c.image_width, c.tanfovx, c.tanfovy
    s.bg, s.scale_modifier, s.viewmatrix, s.projmatrix = torch.zeros(3, device=dev), 1.0, c.viewmatrix, c.projmatrix
    s.sh_degree, s.campos, s.prefiltered, s.debug = scene.sh_degree, c.campos, False, False
    cams.append(ops.make_cam(s, dev))
H, W = cams_h[0].image_height, cams_h[0].image_width
g = torch.Generator().manual_seed(1)
ws = batched.BatchWorkspace(V, P, H, W, dev)
nr, ov = batched.forward_batched(ws, cams, m3, sh, None, op, sc, ro, sync=True)
ws._alloc_binning(int(max(nr) * 1.25) + 4096)
extra = torch.rand(P, 3, generator=g).to(dev)
eo = [torch.empty(3, H, W, device=dev) for _ in range(V)]
pg = [tuple(to(t) for t in scenes.pixel_grads(H, W, 99 + v)) for v in range(V)]
eg = [torch.randn(3, H, W, generator=g).to(dev) / (H * W) for _ in range(V)]
new = lambda *s: torch.empty(*s, device=dev)
outg = {"means3D": new(P, 3), "opacities": new(P, 1), "scales": new(P, 3), "rotations": new(P, 4), "shs": new(P, M, 3),
        "extra_features": new(P, 3)}

def extra_step():
    batched.forward_batched(ws, cams, m3, sh, None, op, sc, ro, extra_features=extra, extra_out=eo)
    batched.backward_batched(ws, cams, m3, sh, None, op, sc, ro, pg, outg, extra_features=extra, extra_grads=eg)

rays_d = torch.nn.functional.normalize(torch.randn(V, H, W, 3, generator=g), dim=-1).to(dev)
rays_o = torch.stack([c.campos for c in cams_h])[:, None, None, :].expand(V, H, W, 3).contiguous().to(dev)
bgm, light = torch.rand(V, H, W, 3, generator=g).to(dev), (torch.randn(V, 3, generator=g) * 3).to(dev)

def post_step():
    img, dep, alp = (torch.stack(t).detach().requires_grad_(True) for t in (ws.color, ws.depth, ws.alpha))
    r = postprocess_views("shading", img, dep, alp, bg=bgm, rays_o=rays_o, rays_d=rays_d, light_positions=light)
    (r["render"].sum() + r["normal"].sum() + r["depth"].sum()).backward()

raw = dict(xyz=m3.clone(), f_dc=sh[:, :1].contiguous(), f_rest=sh[:, 1:].contiguous(),
           opacity=torch.logit(op.clamp(1e-4, 1 - 1e-4)), scaling=torch.log(sc), rotation=ro.clone())
opt = FusedGaussianAdam(raw, dict.fromkeys(("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"), 1e-4))

for _ in range(2):
    extra_step(); post_step(); opt.step(outg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
extra_step(); post_step(); opt.step(outg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
