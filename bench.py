#!/usr/bin/env python
"""bench.py -- fwd+bwd splat renders/sec on the BASELINE.json headline workload.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU oracle (reference arm)

A *step* is one pass of the hot path over one batch of views on every GPU: for each of the rank's
views forward + backward (gradients summed into the packed buffer, densification statistics fused),
then -- for N > 1 -- one SUM all-reduce of the packed buffer and one MAX all-reduce of max_radii.
Weak scaling: every rank renders ``--views-per-gpu`` views of the same replicated 1M-Gaussian scene.
Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "threestudio-3dgs_b200"))

METRIC = "fwd+bwd splat renders/sec"
UNIT = "renders/s"
DEFAULT_WORKLOAD = "headline_1m_512_sh3"


# ------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--views-per-gpu", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "mc", "p2p", "nccl"],
                    help="multi-GPU gradient exchange: mc = own kernel through NVSwitch multicast (multimem.ld_reduce / "
                         "multimem.st, in-switch reduction), p2p = own kernel over NVLink peer memory, nccl; auto = p2p "
                         "at 2 GPUs, mc at 3..8 (where the box supports multicast; else p2p up to 4 GPUs, NCCL above)")
    ap.add_argument("--exchange-chunks", type=int, default=1,
                    help="multi-GPU: Gaussian ranges of the preprocess backward whose exchange overlaps the next range "
                         "(1 = one exchange after the step, the default.  Measured with the whole step incl. the exchange "
                         "kernels in ONE CUDA graph, 1/2/4 ranges: 2 GPUs p2p 4312/-/4167, multicast 3769/3719/3645; "
                         "8 GPUs NCCL 13952/12587/12252 renders/s -- preprocess backward is HBM-bound and so is the "
                         "exchange's local side: overlapped they slow each other down by more than the overlap returns)")
    ap.add_argument("--eager", action="store_true", help="launch the step from Python every time (no CUDA graph)")
    ap.add_argument("--dense-exchange", action="store_true",
                    help="multi-GPU: exchange every row of the packed gradients (default: rows of the rotation / SH "
                         "fields travel only when some rank has a gradient for the Gaussian)")
    ap.add_argument("--quick", action="store_true",
                    help="only the resident timed region (no e2e, no per-kernel table, no aux / cpu / configs legs): "
                         "for exchange experiments at N > 1")
    ap.add_argument("--views-total", type=int, default=0,
                    help="strong scaling: this many views per step in total, split evenly over the GPUs "
                         "(BASELINE.json configs[3]: 32 views of config4_1m_256_sh3_b32); 0 = weak scaling with "
                         "--views-per-gpu")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the short runs of the other BASELINE.json configurations (the `configs` array, N = 1 only)")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_oracle_render_time(scene, cam, threads: int):
    """Wall-clock seconds of ONE complete fwd+bwd render of the workload on the CPU oracle: every Gaussian, every
    tile, forward and backward -- nothing sampled, nothing scaled ("run once and say so")."""
    import torch
    from oracle import torch_oracle as O
    from b200splat import scenes
    torch.set_num_threads(threads)
    s = O.Settings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, torch.ones(3), 1.0,
                   cam.viewmatrix, cam.projmatrix, scene.sh_degree, cam.campos, False, False)
    H, W = cam.image_height, cam.image_width
    gc, gd, ga = scenes.pixel_grads(H, W, 7)
    inputs = (scene.means3D, None, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None)
    t0 = time.perf_counter()
    out, pre, binned = O.rasterize_forward(*inputs, s)
    t_fwd = time.perf_counter() - t0
    O.rasterize_backward(inputs, s, pre, binned, out, gc, gd, ga)
    total = time.perf_counter() - t0
    return total, dict(t_fwd=t_fwd, t_bwd=total - t_fwd, tiles=int(binned["ranges"].shape[0]),
                       num_rendered=int(binned["num_rendered"]))


def run_reference(args):
    """Reference arm: the CPU oracle (kind "port" -- the reference's CUDA rasterizer is un-vendored and
    cannot be built here) on the host cores, same workload / metric / unit.  Each step is one COMPLETE
    fwd+bwd render of one view (no tile sampling); `ms_per_step` is the measured mean wall-clock of the timed steps.
    The run is bounded to about four minutes: if the first timed step shows that K steps would not fit, fewer steps
    are timed and `steps` reports how many were."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from b200splat import scenes
    threads = os.cpu_count() or 1
    t_start = time.perf_counter()
    scene, cams = scenes.make_workload(args.workload, views=1)
    budget_s = 200.0
    warm = min(args.warmup, 1)
    for _ in range(warm):
        cpu_oracle_render_time(scene, cams[0], threads)
    times, info = [], {}
    for i in range(args.steps):
        t, info = cpu_oracle_render_time(scene, cams[0], threads)
        times.append(t)
        elapsed = time.perf_counter() - t_start
        if i + 1 < args.steps and elapsed + t > budget_s:
            break
    sec = sum(times) / len(times)
    val = 1.0 / sec
    sample = (f"the whole workload per step: 1 view, all {scene.means3D.shape[0]} Gaussians, all {info['tiles']} tiles, "
              f"forward + backward of the CPU oracle (no sampling, no scaling); {len(times)} timed steps after {warm} "
              f"warm-up, {sum(times):.1f} s timed, {time.perf_counter() - t_start:.1f} s in total")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "views_per_step": 1, "device": "cpu",
                       "steps_requested": args.steps, "warmup_requested": args.warmup},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measure_config(name, V, steps, warmup, dev):
    """Short resident run of another BASELINE.json configuration on this GPU (the `configs` array of the bench line):
    the same batched step, CUDA-graph replay, CUDA events on the launch stream."""
    import torch
    from b200splat import batched, ops, scenes
    scene, cams_host = scenes.make_workload(name, views=V)
    H, W = cams_host[0].image_height, cams_host[0].image_width
    P, M = scene.means3D.shape[0], scene.shs.shape[1]
    to = lambda t: t.to(dev).contiguous()
    means3D, shs, opac, scales, rots = map(to, (scene.means3D, scene.shs, scene.opacities, scene.scales,
                                                scene.rotations))
    bg = torch.ones(3, device=dev)
    cams = [ops.make_cam(_settings_like(c, bg, scene.sh_degree), dev) for c in cams_host]
    pgrads = [tuple(to(g) for g in scenes.pixel_grads(H, W, 99 + v)) for v in range(V)]
    renderer = batched.BatchRenderer(P, M, H, W, dev, views=V)
    renderer.calibrate(cams, means3D, shs, None, opac, scales, rots)
    graph = renderer.capture_step(cams, means3D, shs, None, opac, scales, rots, pgrads)
    for _ in range(max(warmup, 3)):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ok = not renderer.overflowed()
    nr = [int(x) for ws in renderer.ws for x in ws.num_rendered]
    del graph, renderer
    torch.cuda.empty_cache()
    return {"workload": name, "value": V / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "views_per_step": V,
            "steps": steps, "gaussians": P, "sh_degree": scene.sh_degree, "image": [H, W],
            "pairs_per_view_at_calibration": int(sum(nr) / max(len(nr), 1)), "valid": ok}


def measure_spacetime(name, V, steps, warmup, dev, world=1, rank=0):
    """BASELINE.json configs[4] as written: the spacetime call shape (b200splat/spacetime.py; reference
    renderer/diff_gaussian_rasterizer_st.py:135-150).  Every view has its own timestamp, hence its own non-leaf
    ``means3D`` (cubic B-spline of 12 knots per Gaussian) and ``rotations``; the rasterizer is called once per view
    through the drop-in ``GaussianRasterizer`` (as the reference's loop does) and autograd carries the gradients to the
    knots.  A step = V views forward + one backward (+ an NCCL all-reduce of the parameter gradients at N > 1).  Timed
    end to end on the device (CUDA events), parameters resident."""
    import torch
    import torch.distributed as dist
    from b200splat import scenes, spacetime
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    scene, cams_all = scenes.make_workload(name, views=V * world)
    cams_host = cams_all[rank * V:(rank + 1) * V]
    H, W = cams_host[0].image_height, cams_host[0].image_width
    P = scene.means3D.shape[0]
    params = spacetime.SpacetimeParams(*[t.to(dev).contiguous().requires_grad_(True)
                                         for t in spacetime.make_params(scene, seed=77)])
    bg = torch.ones(3, device=dev)
    rss = [GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, c.viewmatrix.to(dev), c.projmatrix.to(dev),
                                         0, c.campos.to(dev), False, False) for c in cams_host]
    pgs = [tuple(g.to(dev) for g in scenes.pixel_grads(H, W, 99 + rank * V + v)) for v in range(V)]
    frames = params.omega.shape[0]
    times = [((rank * V + v + 0.5) / (V * world), int((rank * V + v + 0.5) / (V * world) * frames)) for v in range(V)]

    def step():
        for t in params:
            t.grad = None
        loss = None
        for v in range(V):
            m3, scl, rot, opa, col = spacetime.timed_all(params, *times[v])
            m2 = torch.zeros_like(m3, requires_grad=True)
            c, r, d, a = GaussianRasterizer(raster_settings=rss[v])(means3D=m3, means2D=m2, shs=None, colors_precomp=col,
                                                                    opacities=opa, scales=scl, rotations=rot,
                                                                    cov3D_precomp=None)
            l = (c * pgs[v][0]).sum() + (d * pgs[v][1]).sum() + (a * pgs[v][2]).sum()
            loss = l if loss is None else loss + l
        loss.backward()
        if world > 1:
            for t in params:
                dist.all_reduce(t.grad)

    from b200splat import _lib
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    knots_grad = float(params.knots.grad.abs().max())
    return {"workload": name, "value": V * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "views_per_step": V * world, "views_per_gpu_per_step": V, "steps": steps, "gaussians": P, "image": [H, W],
            "call_shape": "spacetime: per-view spline means3D + normalize(q + dq[frame]) (non-leaf), colors_precomp; one "
                          "GaussianRasterizer call per view, autograd to the 12 knots per Gaussian",
            "gpu_launches": int(_lib.launch_count() - l0), "knots_grad_max": knots_grad, "valid": knots_grad > 0}


class _S:
    pass


def _settings_like(c, bg, sh_degree):
    s = _S()
    s.image_height, s.image_width, s.tanfovx, s.tanfovy = c.image_height, c.image_width, c.tanfovx, c.tanfovy
    s.bg, s.scale_modifier, s.viewmatrix, s.projmatrix = bg, 1.0, c.viewmatrix, c.projmatrix
    s.sh_degree, s.campos, s.prefiltered, s.debug = sh_degree, c.campos, False, False
    return s


def aux_kernels(dev, P, M, H, W, V, cams, params, hbm_gbs):
    """The kernels either side of the rasterizer, timed alone with CUDA events on the launch stream on
    the bench workload's shapes: fused post-ops (shading mode, V views), fused activation-backward + Adam, and the
    cost of 3 extra feature channels riding the raster pass.  Algorithmic bytes per launch as stated in DESIGN.md."""
    import torch
    from b200splat import batched
    from b200splat.optim import FusedGaussianAdam
    means3D, shs, opac, scales, rots = params

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    def row(name, ms, nbytes, note):
        ach = nbytes / (ms * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "avg_ms": ms, "achieved": ach, "peak": hbm_gbs, "unit": "GB/s",
                "frac": ach / hbm_gbs, "work_per_launch": nbytes, "note": note}

    out = []
    try:
        g = torch.Generator(device="cpu").manual_seed(5)
        rnd = lambda *s: torch.rand(*s, generator=g).to(dev)
        img, dep, alp = rnd(V, 3, H, W), rnd(V, 1, H, W) * 3, rnd(V, 1, H, W)
        rays_o, rays_d, bgm, light = rnd(V, H, W, 3), torch.nn.functional.normalize(rnd(V, H, W, 3), dim=-1), \
            rnd(V, H, W, 3), rnd(V, 3) * 3
        # timed through the C ABI directly (preallocated outputs): the autograd wrapper's host work would hide
        # kernels this short
        import ctypes as C
        from b200splat import postops as PO
        from b200splat._lib import lib, check
        render, normal, dout = torch.empty(V, 3, H, W, device=dev), torch.empty(V, 3, H, W, device=dev), \
            torch.empty(V, 1, H, W, device=dev)
        gr, gn, gd = torch.randn_like(render), torch.randn_like(normal), torch.randn_like(dout)
        d_img, d_dep, d_alp, d_bg = torch.empty_like(img), torch.empty_like(dep), torch.empty_like(alp), \
            torch.empty_like(bgm)
        scratch = torch.empty(int(lib.b200splat_postprocess_scratch_bytes(V, H, W)), dtype=torch.uint8, device=dev)
        pa = PO._args(3, 2, img.detach(), dep.detach(), alp.detach(), rays_o, rays_d, bgm, light, None,
                      [0.1] * 3, [0.9] * 3)
        pa.render, pa.normal, pa.depth_out = render.data_ptr(), normal.data_ptr(), dout.data_ptr()
        pa.g_render, pa.g_normal, pa.g_depth = gr.data_ptr(), gn.data_ptr(), gd.data_ptr()
        pa.d_image, pa.d_depth, pa.d_alpha, pa.d_bg = d_img.data_ptr(), d_dep.data_ptr(), d_alp.data_ptr(), \
            d_bg.data_ptr()
        pa.scratch, pa.scratch_bytes = scratch.data_ptr(), scratch.numel()
        t_f = timed(lambda: check(lib.b200splat_postprocess_forward(C.byref(pa)), "postprocess_forward"), 20)
        t_b = timed(lambda: check(lib.b200splat_postprocess_backward(C.byref(pa)), "postprocess_backward"), 20)
        px = V * H * W
        out.append(row("postprocess_fwd", t_f, 84 * px, f"shading mode, {V} views {H}x{W}: 14 floats in, 7 out per pixel"))
        out.append(row("postprocess_bwd", t_b, 116 * px,
                       "two kernels (local + stencil gather); 21 floats in, 8 out per pixel; the 9-float scratch "
                       "round trip (72 B/pixel) is not counted as algorithmic"))
        raw = dict(xyz=means3D.clone(), f_dc=shs[:, :1].contiguous(), f_rest=shs[:, 1:].contiguous(),
                   opacity=torch.logit(opac.clamp(1e-4, 1 - 1e-4)), scaling=torch.log(scales), rotation=rots.clone())
        opt = FusedGaussianAdam(raw, dict.fromkeys(("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"), 1e-4))
        grads = dict(means3D=torch.randn_like(means3D), shs=torch.randn_like(shs), opacities=torch.randn_like(opac),
                     scales=torch.randn_like(scales), rotations=torch.randn_like(rots))
        t_a = timed(lambda: opt.step(grads))
        out.append(row("adam_step", t_a, 7 * 4 * (11 + 3 * M) * P,
                       "activation backward + Adam, 2 kernels (flat elementwise + quaternion): gradients read, "
                       "parameter / exp_avg / exp_avg_sq read+written"))
        del opt, raw, grads
        # 3 extra feature channels in the same raster pass: step time with and without them
        ws = batched.BatchWorkspace(V, P, H, W, dev)
        extra = torch.rand(P, 3, generator=g).to(dev)
        eo = [torch.empty(3, H, W, device=dev) for _ in range(V)]
        pg = [(torch.randn(3, H, W, device=dev) / (H * W), None, None) for _ in range(V)]
        eg = [torch.randn(3, H, W, device=dev) / (H * W) for _ in range(V)]
        new = lambda *s: torch.empty(*s, device=dev)
        outg = {"means3D": new(P, 3), "opacities": new(P, 1), "scales": new(P, 3), "rotations": new(P, 4),
                "shs": new(P, M, 3), "extra_features": new(P, 3)}
        nr, ov = batched.forward_batched(ws, cams, means3D, shs, None, opac, scales, rots, sync=True)
        ws._alloc_binning(int(max(nr) * 1.25) + 4096)

        def plain():
            batched.forward_batched(ws, cams, means3D, shs, None, opac, scales, rots)
            batched.backward_batched(ws, cams, means3D, shs, None, opac, scales, rots, pg, outg)

        def with_extra():
            batched.forward_batched(ws, cams, means3D, shs, None, opac, scales, rots, extra_features=extra, extra_out=eo)
            batched.backward_batched(ws, cams, means3D, shs, None, opac, scales, rots, pg, outg, extra_features=extra,
                                     extra_grads=eg)
        t_p, t_e = timed(plain, 5), timed(with_extra, 5)
        out.append({"kernel": "extra_channels_step", "avg_ms": t_e, "plain_step_ms": t_p, "overhead_frac": t_e / t_p - 1.0,
                    "note": f"eager fwd+bwd of {V} views with 3 extra channels in the same pass vs without; the "
                            "reference pays a second full rasterizer pass (+100 %) for them"})
    except Exception as exc:   # reported in the JSON line, never hidden
        out.append({"error": repr(exc)})
    return out


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from b200splat import _lib, batched, ops, scenes
    from b200splat import dist as bdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback); "
                         "use --impl reference for the CPU oracle")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the caller's NCCL_DEBUG stands (the driver reads the communicator lines); NCCL's log goes to stderr so that
        # stdout stays the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    N = world
    if args.workload == "stress_4m_1024_st_b64":
        # BASELINE.json configs[4] as written (spacetime call shape): its own short bench line
        V = args.views_per_gpu if args.views_per_gpu != 4 else 8
        res = measure_spacetime(args.workload, V, args.steps, args.warmup, dev, world, rank)
        if rank == 0:
            res.update({"metric": METRIC, "n_gpus": N, "warmup": max(args.warmup, 3), "higher_is_better": True,
                        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                        "config": {"workload": args.workload, "views_per_gpu_per_step": V}})
            print(json.dumps(res), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    V = args.views_per_gpu
    scaling = "weak"
    if args.views_total:
        if args.views_total % N:
            raise SystemExit("--views-total must be a multiple of the number of GPUs")
        V, scaling = args.views_total // N, "strong"

    # ---- workload: replicated scene, per-rank views ------------------------------------------------
    scene, cams_all = scenes.make_workload(args.workload, views=V * N)
    cams_host = cams_all[rank * V:(rank + 1) * V]
    H, W = cams_host[0].image_height, cams_host[0].image_width
    P = scene.means3D.shape[0]
    M = scene.shs.shape[1]
    to = lambda t: t.to(dev).contiguous()
    means3D, shs, opac, scales, rots = map(to, (scene.means3D, scene.shs, scene.opacities, scene.scales,
                                                scene.rotations))
    bg = torch.ones(3, device=dev)

    def dev_cam(c):
        return ops.make_cam(_settings_like(c, bg, scene.sh_degree), dev)

    cams = [dev_cam(c) for c in cams_host]
    pgrads_host = [scenes.pixel_grads(H, W, 99 + rank * V + v) for v in range(V)]
    pgrads = [tuple(to(g) for g in pg) for pg in pgrads_host]
    # multi-GPU: the packed gradient buffer lives in CUDA-IPC memory mapped by all ranks and is all-reduced in
    # place by one kernel per rank over NVLink peer memory (csrc/p2p.cu); --allreduce nccl uses NCCL instead
    p2p, allreduce_mode = None, ("none" if world == 1 else "nccl")
    # row-sparse exchange (own kernels): room for the live map behind max_radii; --dense-exchange leaves it out
    n_tail = 0 if args.dense_exchange else batched.PackedGrads.live_floats(P)
    # auto (measured on this pool, profiles/r2_scaling.md): 2 GPUs -> peer-memory kernel (a rank's own half never
    # crosses NVLink), 3..8 GPUs -> multicast kernel (in-switch reduction), NCCL only as the fall-back
    if world > 1 and (args.allreduce == "mc" or (args.allreduce == "auto" and world > 2)):
        try:
            p2p = bdist.MulticastAllReduce(batched.PackedGrads.floats(P, M), batched.PackedGrads.padded(P), dev,
                                           device_epoch=True, n_tail=n_tail)
            allreduce_mode = "nvswitch_multicast_kernel"
        except Exception as exc:
            print(f"bench: multicast all-reduce unavailable ({exc!r})", file=sys.stderr)
            p2p = None
    use_p2p = p2p is None and (args.allreduce == "p2p" or (args.allreduce in ("auto", "mc") and world <= 4))
    if world > 1 and use_p2p:
        try:
            p2p = bdist.P2PAllReduce(batched.PackedGrads.floats(P, M), batched.PackedGrads.padded(P), dev, device_epoch=True,
                                     n_tail=n_tail)
            allreduce_mode = "p2p_nvlink_kernel"
        except Exception as exc:
            print(f"bench: P2P all-reduce unavailable ({exc!r}); using NCCL", file=sys.stderr)
            p2p = None
    renderer = batched.BatchRenderer(P, M, H, W, dev, views=V, packed_storage=None if p2p is None else p2p.buffer)
    packed = renderer.packed
    if p2p is not None and packed.live_map is not None:
        p2p.live_offset = packed.live_offset_bytes
        allreduce_mode += "+row_sparse"
    renderer.calibrate(cams, means3D, shs, None, opac, scales, rots)

    # the step's launches are recorded once into a CUDA graph (the C ABI never synchronises in the batched path);
    # --eager launches them from Python every step instead.  With an own exchange kernel (multicast / peer memory) the
    # exchange is part of the graph: after the step, or range by range on a forked stream (--exchange-chunks).
    graph, launch_mode = None, "eager"
    chunks = args.exchange_chunks if world > 1 else 1
    l_a = _lib.launch_count()
    renderer.step(cams, means3D, shs, None, opac, scales, rots, pgrads)
    launches_per_step = _lib.launch_count() - l_a

    def exchange(g0, g1):
        if p2p is not None:
            p2p(packed.segments(g0, g1))
        else:   # one coalesced SUM launch per range; the (small) MAX of max_radii once, with the last range
            bdist.allreduce_packed_range(packed, g0, g1, with_max=False)
            if g1 == P:
                dist.all_reduce(packed.max_radii, op=dist.ReduceOp.MAX)

    graph_has_exchange = False
    if not args.eager:
        try:
            if p2p is not None:
                l_b = _lib.launch_count()
                graph = renderer.capture_step(cams, means3D, shs, None, opac, scales, rots, pgrads, exchange=exchange,
                                              chunks=chunks)
                graph_has_exchange = True
                launches_per_step = (_lib.launch_count() - l_b) // 2      # the eager pass + the captured pass
            else:
                graph = renderer.capture_step(cams, means3D, shs, None, opac, scales, rots, pgrads, head_only=chunks > 1)
            launch_mode = "cuda_graph"
        except Exception as exc:   # report, do not hide: the run continues on the eager path
            print(f"bench: CUDA graph capture failed ({exc!r}); eager launches", file=sys.stderr)
            graph, graph_has_exchange = None, False
    if world > 1:
        dist.barrier()

    def step():
        if graph_has_exchange:
            graph.replay()
            return
        if chunks > 1:
            # multi-GPU: forward + render backward, then preprocess backward in Gaussian ranges whose gradients are
            # exchanged on a side stream while the next range is computed
            if graph is not None:
                graph.replay()
            else:
                renderer.step_head(cams, means3D, shs, None, opac, scales, rots, pgrads)
            renderer.step_tail(cams, means3D, shs, None, opac, scales, rots, exchange, chunks=chunks)
            return
        if graph is not None:
            graph.replay()
        else:
            renderer.step(cams, means3D, shs, None, opac, scales, rots, pgrads)
        if p2p is not None:
            p2p()
        else:
            bdist.allreduce_packed(packed.buffer, packed.max_radii)

    def verify_exchange():
        """After the timed region (N > 1): one more step WITHOUT the exchange, a copy of the local sums, then the
        exchange as timed -- its result must be bit-identical on all ranks and equal to an independent NCCL SUM / MAX
        of the copies (tolerance: the summation order over the ranks may differ, fp32)."""
        if graph is not None and chunks == 1 and not graph_has_exchange:
            graph.replay()
        else:
            renderer.step(cams, means3D, shs, None, opac, scales, rots, pgrads)
        torch.cuda.synchronize()
        ref_sum, ref_max = packed.buffer.clone(), packed.max_radii.clone()
        if p2p is not None:
            p2p()
        else:
            bdist.allreduce_packed(packed.buffer, packed.max_radii)
        dist.all_reduce(ref_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(ref_max, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
        scale = float(ref_sum.abs().max().clamp_min(1e-30))
        err = float((packed.buffer - ref_sum).abs().max()) / scale
        max_equal = bool(torch.equal(packed.max_radii, ref_max))
        # bit-identical on every rank: compare two integer checksums of the raw words
        words = packed.buffer.view(torch.int32).to(torch.int64)
        chk = torch.stack([words.sum(), (words * (torch.arange(words.numel(), device=dev) % 8191 + 1)).sum(),
                           packed.max_radii.view(torch.int32).to(torch.int64).sum()])
        allchk = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        identical = all(bool(torch.equal(c, allchk[0])) for c in allchk)
        nonzero = bool(ref_sum.abs().sum() > 0)
        return {"ok": bool(err <= 2e-6 and max_equal and identical and nonzero), "max_rel_err_vs_nccl_sum": err,
                "max_radii_equal": max_equal, "bit_identical_across_ranks": identical}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0          # launched from the host inside the timed region (incl. p2p kernels)
    if graph_has_exchange:                       # everything is a graph node: the step and its exchange kernels
        launches += launches_per_step * args.steps
    elif graph is not None:                      # + the graph's nodes: the whole step, or the step minus the
        launches += (launches_per_step - (1 if chunks > 1 else 0)) * args.steps   # chunked preprocess backward
    if renderer.overflowed():
        raise SystemExit("bench: a view exceeded its binning capacity during the timed region (invalid run)")
    if p2p is not None and p2p.failed():
        raise SystemExit("bench: the P2P all-reduce timed out waiting for a peer (invalid run)")
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    exchange_check = None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        exchange_check = verify_exchange()
    total_ms = float(ms.item())
    ms_per_step = total_ms / args.steps
    value = N * V * args.steps / (total_ms / 1e3)

    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps,
                              "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                              "scaling": scaling, "quick": True,
                              "config": {"workload": args.workload, "views_per_gpu_per_step": V, "launch": launch_mode,
                                         "allreduce": allreduce_mode, "exchange_chunks": chunks},
                              "exchange_verified": None if exchange_check is None else exchange_check["ok"],
                              "exchange_check": exchange_check, "gpu_launches": launches}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: public API (GaussianRasterizer + autograd), host buffers for the step's inputs ----------
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    pin = lambda t: t.contiguous().pin_memory()
    cam_host_t = [(pin(c.viewmatrix), pin(c.projmatrix), pin(c.campos)) for c in cams_host]
    bg_host = pin(torch.ones(3))
    pg_pinned = [tuple(pin(g) for g in pg) for pg in pgrads_host]
    img_host = torch.empty(V, 3, H, W).pin_memory()
    loss_host = torch.empty(1).pin_memory()
    params = [t.clone().requires_grad_(True) for t in (means3D, shs, opac, scales, rots)]
    h2d = sum(sum(t.numel() * 4 for t in ct) for ct in cam_host_t) + V * 12 + \
        sum(sum(g.numel() * 4 for g in pg) for pg in pg_pinned)
    d2h = img_host.numel() * 4 + 4

    def e2e_step():
        for p in params:
            p.grad = None
        loss = None
        for v in range(V):
            vm, pm, cp = (t.to(dev, non_blocking=True) for t in cam_host_t[v])
            bgd = bg_host.to(dev, non_blocking=True)
            gc, gd, ga = (t.to(dev, non_blocking=True) for t in pg_pinned[v])
            rs = GaussianRasterizationSettings(H, W, cams_host[v].tanfovx, cams_host[v].tanfovy, bgd, 1.0, vm, pm,
                                               scene.sh_degree, cp, False, False)
            m2 = torch.zeros_like(params[0], requires_grad=True)
            color, radii, depth, alpha = GaussianRasterizer(raster_settings=rs)(
                means3D=params[0], means2D=m2, shs=params[1], colors_precomp=None, opacities=params[2],
                scales=params[3], rotations=params[4], cov3D_precomp=None)
            l = (color * gc).sum() + (depth * gd).sum() + (alpha * ga).sum()
            loss = l if loss is None else loss + l
            img_host[v].copy_(color.detach(), non_blocking=True)
        loss.backward()
        if world > 1:
            for p in params:
                dist.all_reduce(p.grad)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_host[0])

    from b200splat.batched import ViewBatchRasterizer
    vbr = ViewBatchRasterizer(V, P, H, W, dev)
    copy_stream = torch.cuda.Stream(device=dev)      # device -> host (images, loss)
    upload_stream = torch.cuda.Stream(device=dev)    # host -> device: never waits for the main stream, so the pixel
    #                                                  gradients of step k+1 cross PCIe while step k's backward runs

    # host side of the batched step: ONE pinned block per kind of input, so a step costs two uploads
    cam_block = pin(torch.stack([torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1),
                                            c.campos.reshape(-1), torch.ones(3)]) for c in cams_host]))   # (V, 38)
    pg_block = pin(torch.stack([torch.cat([g.reshape(-1, H, W) for g in pg]) for pg in pgrads_host]))      # (V, 5, H, W)

    def e2e_step_batched():
        """Public batched operator (ViewBatchRasterizer + autograd).  Host inputs of the step (cameras, bg,
        upstream pixel gradients) come from pinned memory; the pixel-gradient upload runs on its own stream
        (under the previous step's backward and this step's forward); images and the loss are read back."""
        for p in params:
            p.grad = None
        main = torch.cuda.current_stream()
        cb = cam_block.to(dev, non_blocking=True)
        rss = [GaussianRasterizationSettings(H, W, cams_host[v].tanfovx, cams_host[v].tanfovy, cb[v, 35:38], 1.0,
                                             cb[v, 0:16].view(4, 4), cb[v, 16:32].view(4, 4), scene.sh_degree,
                                             cb[v, 32:35], False, False) for v in range(V)]
        with torch.cuda.stream(upload_stream):
            pgd = pg_block.to(dev, non_blocking=True)
        m2 = torch.zeros(V, P, 3, device=dev, requires_grad=True)
        C, R, D, A = vbr(rss, means3D=params[0], means2D=m2, opacities=params[2], shs=params[1], scales=params[3],
                         rotations=params[4])
        main.wait_stream(upload_stream)
        pgd.record_stream(main)
        loss = (C * pgd[:, 0:3]).sum() + (D * pgd[:, 3:4]).sum() + (A * pgd[:, 4:5]).sum()
        copy_stream.wait_stream(main)            # images leave over PCIe while the backward runs
        with torch.cuda.stream(copy_stream):
            img_host.copy_(C.detach(), non_blocking=True)
        C.record_stream(copy_stream)
        loss.backward()
        if world > 1:
            for p in params:
                dist.all_reduce(p.grad)
        # the loss of step k is read on the host while step k+1 is already queued (one step of lag, as a training
        # loop that logs asynchronously does); every step's loss and images are read inside the timed region
        slot = pending["k"] & 1
        loss_slots[slot].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = done_events[slot]
        copy_stream.wait_stream(main)
        ev.record(copy_stream)
        prev = pending.get("ev")
        pending["ev"], pending["slot"] = ev, slot
        pending["k"] += 1
        if prev is not None:
            prev.synchronize()
            return float(loss_slots[slot ^ 1][0])
        return None

    def e2e_drain():
        ev = pending.pop("ev", None)
        if ev is not None:
            ev.synchronize()
            return float(loss_slots[pending["slot"]][0])
        return None

    pending = {"k": 0}
    loss_slots = [torch.empty(1).pin_memory() for _ in range(2)]
    done_events = [torch.cuda.Event() for _ in range(2)]

    def time_e2e(fn, drain=None):
        for _ in range(3):
            fn()
        if drain:
            drain()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        if drain:
            drain()
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        return N * V * args.steps / float(e2e_s.item())

    e2e_dropin = time_e2e(e2e_step)
    e2e_value = time_e2e(e2e_step_batched, e2e_drain)
    # clocks / throttle reasons were sampled (20 ms period) from the start of the resident timed region to here
    clocks = sampler.stop() if rank == 0 else None
    if vbr.check_overflow():
        raise SystemExit("bench: binning capacity overflow in the e2e region (invalid run)")

    # ---- per-kernel roofline (rank 0): CUDA events on the launch stream, live ----------------------
    roof, kernels = None, None
    cpu_base = None
    if rank == 0:
        pk = peaks()
        prof_steps = max(2, min(5, args.steps))
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        for _ in range(prof_steps):
            renderer.step(cams, means3D, shs, None, opac, scales, rots, pgrads)
        torch.cuda.synchronize()
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        # one launch group covers the whole view batch: times and work below are PER LAUNCH (V views)
        # work counters of the rank's views (summed over the batch)
        cnt = dict(V=0, R=0, n_eval_fwd=0, n_eval_bwd=0, staged=0)
        for cam in cams:
            color, radii, depth, alpha, st = ops.forward(cam, means3D, shs, None, opac, scales, rots, None)
            vw = ops.forward_views(cam, st)
            cnt["V"] += int((radii > 0).sum())
            cnt["R"] += st.num_rendered
            cnt["n_eval_fwd"] += int(vw["n_visited"].sum())
            cnt["n_eval_bwd"] += int(vw["n_contrib"].sum())
        Tn = ((W + 15) // 16) * ((H + 15) // 16)
        tile_bits = max(1, (Tn - 1).bit_length())
        pair_passes = 1 if tile_bits <= 10 else (tile_bits + 7) // 8
        Mc, Rs = M, cnt["R"]
        alg = {  # algorithmic bytes / flops per launch of the view batch
            # parameters read once for the batch; per view: record 48 + depth 4 + radius 4 + tiles 4 + rect 8 +
            # depth-sort word 8 + clamp bits 1
            "preprocess": ("hbm", P * (44 + 12 * Mc) + V * P * 77),
            # per view: depth-order word 8 + gathered tile count 4 + offset 4
            "scan": ("hbm", V * 16 * P),
            # per view: order word 8 + rect 8 + offset 4 per Gaussian, one 8-byte pair word per pair
            "duplicate": ("hbm", V * 20 * P + 8 * Rs),
            # Gaussian depth sort: histogram read 8 P + 4 passes x 16 P; pair partition: passes x 16 R
            "sort": ("hbm", V * 72 * P + pair_passes * 16 * Rs),
            "ranges": ("hbm", V * 16 * Tn),
            "render_fwd": ("fp32", 30 * cnt["n_eval_fwd"]),
            "render_bwd": ("fp32", 90 * cnt["n_eval_bwd"]),
            # parameters read once, gradients written once; per view: 48-byte gradient record + radius 4 + clamp 1
            "preprocess_bwd": ("hbm", P * (44 + 12 * Mc) + P * (44 + 12 * Mc) + V * P * 53),
        }
        fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        traffic = {}
        tpath = ROOT / "profiles" / "ncu_traffic.json"
        if tpath.exists():
            try:
                traffic = json.loads(tpath.read_text()).get(args.workload, {})
            except Exception:
                traffic = {}
        kernels = []
        for name, (bound, work) in alg.items():
            tot, n = prof.get(name, (0.0, 0))
            if n == 0:
                continue
            avg_ms = tot / n
            if bound == "hbm":
                ach, peak, unit = work / (avg_ms * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s"
            else:
                ach, peak, unit = work / (avg_ms * 1e-3) / 1e12, fp32_peak, "TFLOP/s"
            kernels.append({"kernel": name, "bound": bound, "avg_ms": avg_ms, "launches": n, "achieved": ach,
                            "peak": peak, "unit": unit, "frac": ach / peak, "work_per_launch": work,
                            "views_per_launch": V, "us_per_view": 1e3 * avg_ms / V, "traffic": traffic.get(name)})
        step_ms = sum(k["avg_ms"] for k in kernels)
        for k in kernels:
            k["share_of_step"] = k["avg_ms"] / step_ms if step_ms else None
        dom = max(kernels, key=lambda k: k["avg_ms"])
        roof = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"],
                "unit": dom["unit"], "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": pk["source"] +
                (" HBM copy" if dom["bound"] == "hbm" else " sm_max_mhz x 148 SM x 128 lanes x 2 (non-tensor FP32)"),
                "avg_ms": dom["avg_ms"], "counters": cnt}
        if N == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            t_cpu = [cpu_oracle_render_time(scene, cams_host[0], threads)[0] for _ in range(3)]
            sec = statistics.median(t_cpu)
            cpu_base = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": (f"3 complete fwd+bwd renders of one view of this workload on the CPU oracle (all {P} "
                                   f"Gaussians, all tiles, nothing sampled or scaled), median; {sum(t_cpu):.1f} s of "
                                   "CPU work")}

    aux = None
    configs = None
    if rank == 0 and N == 1:
        aux = aux_kernels(dev, P, M, H, W, V, cams, (means3D, shs, opac, scales, rots), peaks()["hbm_gbs"])
        if not args.no_configs and args.workload == DEFAULT_WORKLOAD:
            # driver-run numbers for the other BASELINE.json configurations (short resident runs on this GPU)
            del renderer, graph, vbr
            torch.cuda.empty_cache()
            configs = []
            for name, v in (("config2_100k_512_sh0_b4", 4), ("config3_300k_512_sh0_b4", 4),
                            ("config4_1m_256_sh3_b32", 32), ("stress_4m_1024_sh3_b64", 8)):
                try:
                    configs.append(measure_config(name, v, 3, 3, dev))
                except Exception as exc:   # reported, never hidden
                    configs.append({"workload": name, "error": repr(exc)})
            try:    # configs[4] as written: the spacetime call shape, one GPU's share (8 of the 64 views)
                configs.append(measure_spacetime("stress_4m_1024_st_b64", 8, 3, 3, dev))
            except Exception as exc:
                configs.append({"workload": "stress_4m_1024_st_b64", "error": repr(exc)})
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "gaussians": P, "sh_degree": scene.sh_degree, "image": [H, W],
                       "views_per_gpu_per_step": V, "global_views_per_step": V * N,
                       "path": "b200splat_forward_batched/_backward_batched: one launch per phase for the V views",
                       "launch": launch_mode, "allreduce": allreduce_mode, "exchange_chunks": chunks,
                       "parallelism": f"view-dp{N}" if N > 1 else "single",
                       "allreduce_bytes_per_step": (packed.nbytes + 4 * P) if N > 1 else 0,
                       "l2": "inputs larger than L2: %.0f MB parameters + per-view key/value buffers > 126 MB"
                             % ((means3D.numel() + shs.numel() + opac.numel() + scales.numel() + rots.numel()) * 4 / 1e6)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "b200splat.batched.ViewBatchRasterizer + autograd (one call for the step's views); cameras, "
                           "bg and pixel gradients from pinned host memory each step; images + loss read back (the loss of step k is read while step k+1 is queued)",
                    "per_view_dropin_value": e2e_dropin,
                    "per_view_dropin_api": "diff_gaussian_rasterization.GaussianRasterizer called once per view "
                                           "(the reference's unchanged loop), same host traffic"},
            "gpu_launches": launches,
            "roofline": roof, "kernels": kernels, "cpu_baseline": cpu_base,
            "aux_kernels": aux, "configs": configs,
        }
        if exchange_check is not None:
            line["exchange_verified"] = exchange_check["ok"]
            line["exchange_check"] = exchange_check
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
