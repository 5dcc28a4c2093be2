"""Where does the e2e step (ViewBatchRasterizer + autograd, host inputs in, images + loss out) spend its time?
Variants: full | no image read-back | no pixel-gradient upload | neither; plus the host-only enqueue time of a step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "threestudio-3dgs_b200"))
import torch
from b200splat import scenes
from b200splat.batched import ViewBatchRasterizer
from diff_gaussian_rasterization import GaussianRasterizationSettings
V, steps = 4, 30
scene, cams_host = scenes.make_workload("headline_1m_512_sh3", views=V)
dev = torch.device("cuda", 0)
H = W = 512
P = scene.means3D.shape[0]
params = [t.to(dev).clone().requires_grad_(True) for t in (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations)]
pin = lambda t: t.contiguous().pin_memory()
pgrads_host = [scenes.pixel_grads(H, W, 99 + v) for v in range(V)]
cam_block = pin(torch.stack([torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1), c.campos.reshape(-1), torch.ones(3)]) for c in cams_host]))
pg_block = pin(torch.stack([torch.cat([g.reshape(-1, H, W) for g in pg]) for pg in pgrads_host]))
pg_dev = pg_block.to(dev)
img_host = torch.empty(V, 3, H, W).pin_memory()
loss_host = torch.empty(1).pin_memory()
vbr = ViewBatchRasterizer(V, P, H, W, dev)
copy_stream = torch.cuda.Stream(device=dev)

def step(upload=True, readback=True, sync=True):
    for p in params: p.grad = None
    main = torch.cuda.current_stream()
    cb = cam_block.to(dev, non_blocking=True)
    rss = [GaussianRasterizationSettings(H, W, cams_host[v].tanfovx, cams_host[v].tanfovy, cb[v, 35:38], 1.0, cb[v, 0:16].view(4, 4),
                                         cb[v, 16:32].view(4, 4), 3, cb[v, 32:35], False, False) for v in range(V)]
    if upload:
        with torch.cuda.stream(copy_stream):
            pgd = pg_block.to(dev, non_blocking=True)
    else:
        pgd = pg_dev
    m2 = torch.zeros(V, P, 3, device=dev, requires_grad=True)
    C, R, D, A = vbr(rss, means3D=params[0], means2D=m2, opacities=params[2], shs=params[1], scales=params[3], rotations=params[4])
    if upload:
        main.wait_stream(copy_stream); pgd.record_stream(main)
    loss = (C * pgd[:, 0:3]).sum() + (D * pgd[:, 3:4]).sum() + (A * pgd[:, 4:5]).sum()
    if readback:
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            img_host.copy_(C.detach(), non_blocking=True)
        C.record_stream(copy_stream)
    loss.backward()
    loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
    if sync:
        torch.cuda.synchronize()

def timed(**kw):
    for _ in range(3): step(**kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): step(**kw)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3

print("full, sync each step      %.3f ms" % timed())
print("full, no per-step sync    %.3f ms" % timed(sync=False))
print("no image read-back        %.3f ms" % timed(readback=False, sync=False))
print("no pixel-grad upload      %.3f ms" % timed(upload=False, sync=False))
print("neither                   %.3f ms" % timed(upload=False, readback=False, sync=False))
# host-only time of a step: enqueue with an idle GPU queue measured by syncing BEFORE and timing until the call returns
torch.cuda.synchronize(); ts = []
for _ in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(upload=False, readback=False, sync=False); ts.append(time.perf_counter() - t0)
print("host time of one step (returns before the GPU is done): %.3f ms" % (sorted(ts)[len(ts)//2] * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step(upload=False, readback=False, sync=False)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
