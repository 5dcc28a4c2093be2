"""Where the per-view drop-in path (GaussianRasterizer called once per view, the reference's unchanged loop) spends
its time on the headline workload: wall clock vs device time vs host-side hot spots (cProfile)."""
import cProfile, io, pstats, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "threestudio-3dgs_b200"))
import torch
from b200splat import scenes, _lib
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

dev = torch.device("cuda", 0)
V = 4
scene, cams = scenes.make_workload("headline_1m_512_sh3", views=V)
H, W = cams[0].image_height, cams[0].image_width
params = [t.to(dev).contiguous().requires_grad_(True) for t in (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations)]
bg = torch.ones(3, device=dev)
rss = [GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, c.viewmatrix.to(dev), c.projmatrix.to(dev), scene.sh_degree, c.campos.to(dev), False, False) for c in cams]
pgs = [tuple(g.to(dev) for g in scenes.pixel_grads(H, W, 99 + v)) for v in range(V)]

def step():
    for p in params:
        p.grad = None
    loss = None
    for v in range(V):
        m2 = torch.zeros_like(params[0], requires_grad=True)
        color, radii, depth, alpha = GaussianRasterizer(raster_settings=rss[v])(
            means3D=params[0], means2D=m2, shs=params[1], colors_precomp=None, opacities=params[2], scales=params[3],
            rotations=params[4], cov3D_precomp=None)
        l = (color * pgs[v][0]).sum() + (depth * pgs[v][1]).sum() + (alpha * pgs[v][2]).sum()
        loss = l if loss is None else loss + l
    loss.backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N
_lib.profile_enable(True)
for _ in range(5):
    step()
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
dev_ms = sum(ms for ms, n in prof.values()) / 5
print(f"per step of {V} views: wall {wall*1e3:.3f} ms ({V/wall:.0f} renders/s); libb200splat kernels {dev_ms:.3f} ms")
print({k: round(ms / 5 / V * 1e3, 1) for k, (ms, n) in prof.items() if n})
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:4000])
