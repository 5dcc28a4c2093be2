import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "threestudio-3dgs_b200"))
import torch
from b200splat import scenes, ops
name = sys.argv[1] if len(sys.argv) > 1 else "headline_1m_512_sh3"
scene, cams = scenes.make_workload(name, views=4)
dev = torch.device("cuda", 0)
to = lambda t: t.to(dev).contiguous()
m3, sh, op, sc, ro = map(to, (scene.means3D, scene.shs, scene.opacities, scene.scales, scene.rotations))
class S: pass
for c in cams:
    s = S()
    s.image_height, s.image_width, s.tanfovx, s.tanfovy = c.image_height, c.image_width, c.tanfovx, c.tanfovy
    s.bg, s.scale_modifier, s.viewmatrix, s.projmatrix = torch.ones(3, device=dev), 1.0, c.viewmatrix, c.projmatrix
    s.sh_degree, s.campos, s.prefiltered, s.debug = scene.sh_degree, c.campos, False, False
    cam = ops.make_cam(s, dev)
    color, radii, depth, alpha, st = ops.forward(cam, m3, sh, None, op, sc, ro, None)
    v = ops.forward_views(cam, st)
    H, W = c.image_height, c.image_width
    rg = v["ranges"].long(); ln = (rg[:,1]-rg[:,0])
    nv = v["n_visited"].long().reshape(H//16,16,W//16,16).permute(0,2,1,3).reshape(-1,256)
    nc = v["n_contrib"].long().reshape(H//16,16,W//16,16).permute(0,2,1,3).reshape(-1,256)
    trav = nv.max(1).values
    print(f"fovy={c.fovy*57.3:.1f} R={st.num_rendered} radii mean={radii[radii>0].float().mean():.1f} max={int(radii.max())} "
          f"list len mean={ln.float().mean():.0f} max={int(ln.max())}; traversed per tile mean={trav.float().mean():.0f} max={int(trav.max())} "
          f"sum_trav={int(trav.sum())} ; n_contrib max per tile mean={nc.max(1).values.float().mean():.0f} max={int(nc.max())}; "
          f"alpha mean={alpha.mean():.3f}; tiles with trav>2000: {int((trav>2000).sum())}, >5000: {int((trav>5000).sum())}")
    qs = torch.quantile(trav.float(), torch.tensor([0.5,0.9,0.99],device=dev))
    print("   trav quantiles 50/90/99:", qs.tolist())
