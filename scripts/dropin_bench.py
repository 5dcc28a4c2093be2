"""e2e renders/s of the unchanged per-view loop (GaussianRasterizer called once per view, autograd, one backward per
step) on the headline workload: pooled fast path vs the allocating entry point (B200SPLAT_DROPIN_POOL=0)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "threestudio-3dgs_b200"))
import torch
from b200splat import scenes
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
V, steps = 4, 20
scene, cams = scenes.make_workload("headline_1m_512_sh3", views=V)
dev = torch.device("cuda", 0)
params = [t.to(dev).clone().requires_grad_(True) for t in (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations)]
H = W = 512
pgs = [tuple(g.to(dev) for g in scenes.pixel_grads(H, W, 99 + v)) for v in range(V)]
rss = [GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, torch.ones(3, device=dev), 1.0, c.viewmatrix.to(dev),
                                     c.projmatrix.to(dev), 3, c.campos.to(dev), False, False) for c in cams]
def step():
    for p in params: p.grad = None
    loss = 0.0
    for v in range(V):
        m2 = torch.zeros_like(params[0], requires_grad=True)
        c, r, d, a = GaussianRasterizer(raster_settings=rss[v])(means3D=params[0], means2D=m2, shs=params[1], colors_precomp=None,
                                                                opacities=params[2], scales=params[3], rotations=params[4], cov3D_precomp=None)
        loss = loss + (c * pgs[v][0]).sum() + (d * pgs[v][1]).sum() + (a * pgs[v][2]).sum()
    loss.backward()
    return loss
for _ in range(3): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(steps): l = step()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"[pool={os.environ.get('B200SPLAT_DROPIN_POOL','1')}] per-view loop: {V*steps/dt:.0f} renders/s ({dt/steps/V*1e3:.3f} ms/view) loss {float(l):.6f}")
