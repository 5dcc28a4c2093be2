"""GPU parity of the extra feature channels (``extra_features=``): one pass must equal the reference's TWO passes.

The reference's normal / shading variants call the rasterizer a second time with the per-Gaussian normals in
place of the colour and a zero background (renderer/diff_gaussian_rasterizer_shading.py:94, :177-187;
renderer/diff_gaussian_rasterizer_normal.py:175-185).  The expected values here are therefore built from two
ORACLE passes -- pass 1 with the colours, pass 2 with ``colors_precomp = extra`` and ``bg = 0`` -- whose gradients
add (both passes feed the same Gaussian parameters).  Bars as everywhere: images max-abs <= 1e-4, gradients
<= 1e-3 relative.
"""
import pytest
import torch

from b200splat import scenes
from oracle import torch_oracle as O
from util import borderline_pixels, cuda_settings, oracle_settings, rel_err

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_TOL = 1e-3


def _scene(P, deg, H, W, seed):
    sc = scenes.make_scene(P, deg, 0.8, seed=seed)
    cam = scenes.sds_cameras(1, H, W, seed=seed + 100)[0]
    return sc, cam


def _leaf(t, dev="cuda"):
    return t.to(dev).clone().requires_grad_(True)


def _cuda_one_pass(sc, s, extra, grads, g_extra):
    from diff_gaussian_rasterization import GaussianRasterizer
    m3, op, sh, scl, rot, ex = map(_leaf, (sc.means3D, sc.opacities, sc.shs, sc.scales, sc.rotations, extra))
    m2 = torch.zeros_like(m3, requires_grad=True)
    out = GaussianRasterizer(raster_settings=cuda_settings(s))(
        means3D=m3, means2D=m2, shs=sh, colors_precomp=None, opacities=op, scales=scl, rotations=rot,
        cov3D_precomp=None, extra_features=ex)
    assert len(out) == 5
    color, radii, depth, alpha, eimg = out
    gc, gd, ga = (g.cuda() for g in grads)
    ((color * gc).sum() + (depth * gd).sum() + (alpha * ga).sum() + (eimg * g_extra.cuda()).sum()).backward()
    return dict(color=color, radii=radii, depth=depth, alpha=alpha, extra=eimg,
                grads=dict(means3D=m3.grad, means2D=m2.grad, opacities=op.grad, shs=sh.grad, scales=scl.grad,
                           rotations=rot.grad, extra=ex.grad))


def _oracle_two_passes(sc, s, extra, grads, g_extra):
    out1, pre, binned = O.rasterize_forward(sc.means3D, None, sc.shs, None, sc.opacities, sc.scales, sc.rotations,
                                            None, s)
    g1 = O.rasterize_backward((sc.means3D, None, sc.shs, None, sc.opacities, sc.scales, sc.rotations, None), s, pre,
                              binned, out1, *grads)
    C = extra.shape[1]
    ex3 = torch.cat([extra, torch.zeros(extra.shape[0], 3 - C)], 1) if C < 3 else extra[:, :3].contiguous()
    ge3 = torch.cat([g_extra, torch.zeros(3 - C, *g_extra.shape[1:])], 0) if C < 3 else g_extra[:3].contiguous()
    s0 = s._replace(bg=torch.zeros(3))
    z1 = torch.zeros(1, s.image_height, s.image_width)
    out2, pre2, binned2 = O.rasterize_forward(sc.means3D, None, None, ex3, sc.opacities, sc.scales, sc.rotations,
                                              None, s0)
    g2 = O.rasterize_backward((sc.means3D, None, None, ex3, sc.opacities, sc.scales, sc.rotations, None), s0, pre2,
                              binned2, out2, ge3, z1, z1)
    total = {k: g1[k] + g2[k] for k in ("means3D", "opacities", "scales", "rotations")}
    # the reference's second call gets a gradient-free zeros tensor as means2D
    # (renderer/diff_gaussian_rasterizer_shading.py:177-187): viewspace gradients come from the colour pass alone
    total["means2D"] = g1["means2D"]
    total["shs"] = g1["shs"]
    total["extra"] = g2["colors_precomp"][:, :C]
    return dict(out=out1, pre=pre, binned=binned, extra=out2["color"][:C], grads=total)


@pytest.mark.parametrize("C", [3, 1])
def test_extra_channels_equal_the_reference_second_pass(C):
    P, H, W = 4000, 64, 64
    sc, cam = _scene(P, 1, H, W, 71)
    s = oracle_settings(cam, 1)
    g = torch.Generator().manual_seed(72)
    extra = torch.rand(P, C, generator=g)
    grads = scenes.pixel_grads(H, W, 73)
    g_extra = torch.randn(C, H, W, generator=g) / (H * W)
    orc = _oracle_two_passes(sc, s, extra, grads, g_extra)
    cu = _cuda_one_pass(sc, s, extra, grads, g_extra)
    assert torch.equal(cu["radii"].cpu(), orc["pre"]["radii"])
    keep = ~borderline_pixels(orc["pre"], orc["binned"], s, orc["out"])
    assert float((~keep).float().mean()) < 0.01
    for name, ref in (("color", orc["out"]["color"]), ("depth", orc["out"]["depth"]), ("alpha", orc["out"]["alpha"]),
                      ("extra", orc["extra"])):
        err = ((cu[name].cpu() - ref).abs() * keep[None]).max().item()
        assert err <= IMG_TOL, f"{name} max-abs {err}"
    for k, ref in orc["grads"].items():
        e = rel_err(cu["grads"][k], ref)
        assert e <= GRAD_TOL, f"grad {k}: rel err {e}"


def test_extra_image_is_bit_identical_to_a_second_cuda_pass_and_four_channels():
    """Same kernels, same alphas, same accumulation order: the fused extra image equals the colour image of a second
    pass with colors_precomp = extra and bg = 0 bit for bit; 4 channels = two such passes."""
    from diff_gaussian_rasterization import GaussianRasterizer
    P, H, W = 20000, 128, 96
    sc, cam = _scene(P, 0, H, W, 81)
    s = oracle_settings(cam, 0)
    extra = torch.rand(P, 4, generator=torch.Generator().manual_seed(82)).cuda()
    a = dict(means3D=sc.means3D.cuda(), means2D=torch.zeros(P, 3, device="cuda"), opacities=sc.opacities.cuda(),
             scales=sc.scales.cuda(), rotations=sc.rotations.cuda(), cov3D_precomp=None)
    with torch.no_grad():
        color, radii, depth, alpha, eimg = GaussianRasterizer(raster_settings=cuda_settings(s))(
            shs=sc.shs.cuda(), colors_precomp=None, extra_features=extra, **a)
        plain = GaussianRasterizer(raster_settings=cuda_settings(s))(shs=sc.shs.cuda(), colors_precomp=None, **a)
        s0 = cuda_settings(s._replace(bg=torch.zeros(3)))
        p2 = GaussianRasterizer(raster_settings=s0)(shs=None, colors_precomp=extra[:, :3].contiguous(), **a)
        p3 = GaussianRasterizer(raster_settings=s0)(shs=None, colors_precomp=extra[:, 1:].contiguous(), **a)
    assert eimg.shape == (4, H, W)
    for got, ref in zip((color, radii, depth, alpha), plain):
        assert torch.equal(got, ref), "the extra channels must not change the regular outputs"
    assert torch.equal(eimg[:3], p2[0])
    assert torch.equal(eimg[1:], p3[0])


def test_extra_channels_through_the_batched_operator():
    from b200splat.batched import ViewBatchRasterizer
    from diff_gaussian_rasterization import GaussianRasterizer
    P, H, W, V = 6000, 64, 64, 3
    sc = scenes.make_scene(P, 1, 0.8, seed=91)
    cams = scenes.sds_cameras(V, H, W, seed=92)
    g = torch.Generator().manual_seed(93)
    extra = torch.rand(P, 3, generator=g)
    gws = [torch.randn(8, H, W, generator=g).cuda() / (H * W) for _ in range(V)]   # colour 3 | depth | alpha | extra 3

    def loss_of(out, v):
        color, _, depth, alpha, eimg = out
        gw = gws[v]
        return (color * gw[:3]).sum() + (depth * gw[3:4]).sum() + (alpha * gw[4:5]).sum() + (eimg * gw[5:]).sum()

    settings = [cuda_settings(oracle_settings(c, 1)) for c in cams]
    # per-view operator
    leaves = [_leaf(t) for t in (sc.means3D, sc.opacities, sc.shs, sc.scales, sc.rotations, extra)]
    m3, op, sh, scl, rot, ex = leaves
    ref_imgs, ref_m2 = [], []
    for v in range(V):
        m2 = torch.zeros_like(m3, requires_grad=True)
        out = GaussianRasterizer(raster_settings=settings[v])(
            means3D=m3, means2D=m2, shs=sh, colors_precomp=None, opacities=op, scales=scl, rotations=rot,
            cov3D_precomp=None, extra_features=ex)
        loss_of(out, v).backward()
        ref_imgs.append(out[4].detach())
        ref_m2.append(m2.grad)
    ref = [t.grad.clone() for t in leaves]
    # batched operator
    leaves_b = [_leaf(t) for t in (sc.means3D, sc.opacities, sc.shs, sc.scales, sc.rotations, extra)]
    m3b, opb, shb, sclb, rotb, exb = leaves_b
    m2b = torch.zeros(V, P, 3, device="cuda", requires_grad=True)
    rast = ViewBatchRasterizer(V, P, H, W)
    out = rast(settings, m3b, m2b, opb, shs=shb, scales=sclb, rotations=rotb, extra_features=exb)
    assert len(out) == 5 and out[4].shape == (V, 3, H, W)
    sum(loss_of(tuple(o[v] for o in out), v) for v in range(V)).backward()
    for v in range(V):
        assert torch.equal(out[4][v], ref_imgs[v])
        assert rel_err(m2b.grad[v], ref_m2[v]) <= 1e-5
    for name, got, want in zip(("means3D", "opacities", "shs", "scales", "rotations", "extra"), leaves_b, ref):
        assert rel_err(got.grad, want) <= 1e-4, name
    # a second step on the same workspace: the self-cleaning scratch must have left the extra records zero
    for t in leaves_b:
        t.grad = None
    out = rast(settings, m3b, m2b, opb, shs=shb, scales=sclb, rotations=rotb, extra_features=exb)
    sum(loss_of(tuple(o[v] for o in out), v) for v in range(V)).backward()
    assert rel_err(exb.grad, ref[5]) <= 1e-4
