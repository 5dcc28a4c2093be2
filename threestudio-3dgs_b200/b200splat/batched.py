"""One training step's view batch on one GPU: the unit the reference runs as a Python loop of
single-view rasterizer calls (renderer/gaussian_batch_renderer.py:21-54) followed by per-view
densification statistics (geometry/gaussian_base.py:815-819, 846-851).

Here the whole batch goes through the C ABI's view-batched entry points (b200splat_forward_batched /
b200splat_backward_batched): the Gaussian parameters are read once per phase for all views, every
phase is one launch for the batch, num_rendered stays on the device (persistent binning buffers with a
capacity), the parameter gradients of all views are summed in registers and written once into one
packed buffer, and the densification statistics are a fused epilogue.  A multi-GPU step is then:
local views -> ONE sum all-reduce of the packed buffer + ONE max all-reduce of max_radii
(b200splat/dist.py).  Note the order the reference's statistics impose:
||means2D.grad|| is taken per view *before* any summation, so it is reduced per view on the rank and
only the accumulators are all-reduced, never means2D.grad itself.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, ops
from ._lib import check, lib

MAX_VIEWS = _lib.MAX_VIEWS


class PackedGrads:
    """[dL/dmeans3D 3 | dL/dscales 3 | dL/drotations 4 | dL/dopacity 1 | dL/dshs 3M (or colours 3) |
    grad_accum 1 | denom 1] x P floats in ONE contiguous fp32 buffer (the all-reduce payload,
    4*P*(13+3M) bytes), plus max_radii (P) reduced with MAX.

    Every field starts on a 16-byte boundary for ANY P (densify / prune make P arbitrary): a field of width w
    occupies w * P4 floats, P4 = P rounded up to a multiple of 4; the up to 3 padding Gaussians stay zero.  The
    kernels access rotations (and, when aligned, the SH rows) as float4, and the peer-memory all-reduce
    exchanges whole float4s."""

    @staticmethod
    def padded(P: int) -> int:
        return (P + 3) // 4 * 4

    @staticmethod
    def live_floats(P: int) -> int:
        """Floats behind max_radii that hold the live map of the row-sparse exchange (one byte per Gaussian)."""
        return PackedGrads.padded(P) // 4

    @staticmethod
    def floats(P: int, M: int, color_mode: str = "shs") -> int:
        """Floats of the SUM segment (the MAX segment, max_radii, is padded(P) more)."""
        return (13 + (3 * M if color_mode == "shs" else 3)) * PackedGrads.padded(P)

    def __init__(self, P: int, M: int, device, color_mode: str = "shs", storage: Optional[torch.Tensor] = None):
        """``storage``: optional flat fp32 tensor of at least floats() + padded(P) elements that backs the buffer
        and max_radii -- e.g. the exchange buffer of ``dist.P2PAllReduce``, so the all-reduce runs in place."""
        self.P, self.M, self.color_mode = P, M, color_mode
        P4 = self.P4 = self.padded(P)
        ncol = 3 * M if color_mode == "shs" else 3
        widths = [("means3D", 3), ("scales", 3), ("rotations", 4), ("opacities", 1),
                  (color_mode if color_mode == "shs" else "colors_precomp", ncol),
                  ("grad_accum", 1), ("denom", 1)]
        total = sum(w for _, w in widths) * P4
        if storage is not None:
            assert storage.dtype == torch.float32 and storage.is_contiguous() and storage.numel() >= total + P4
            assert storage.data_ptr() % 16 == 0
            storage[:total + P4].zero_()
        self.buffer = storage[:total] if storage is not None else torch.zeros(total, dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = {}
        self.fields: List[tuple] = []   # (name, offset in floats, floats per Gaussian)
        off = 0
        for name, w in widths:
            self.fields.append((name, off, w))
            self.views[name] = self.buffer[off:off + w * P].view(P, w) if w > 1 else self.buffer[off:off + P]
            off += w * P4
        self.views["opacities"] = self.views["opacities"].view(P, 1)
        if color_mode == "shs":
            self.views["shs"] = self.views["shs"].view(P, M, 3)
        self._max_all = storage[total:total + P4] if storage is not None else \
            torch.zeros(P4, dtype=torch.float32, device=device)
        self.max_radii = self._max_all[:P]
        # live map (shared exchange storage with room behind max_radii only): preprocess backward's scan writes one
        # byte per Gaussian, 1 = it received a gradient on this rank; the exchange kernels move only the rows of the
        # float4-granular per-Gaussian fields (rotations, SH) that are live on some rank (include/b200splat.h ABI 7)
        self.live_map, self.live_offset_bytes = None, 0
        if storage is not None and storage.numel() >= total + P4 + P4 // 4:
            lm = storage[total + P4:total + P4 + P4 // 4]
            lm.zero_()
            self.live_map = lm.view(torch.uint8)
            self.live_offset_bytes = (total + P4) * 4
        self.means2D_scratch = torch.empty(P, 3, dtype=torch.float32, device=device)

    @property
    def nbytes(self) -> int:
        return self.buffer.numel() * 4

    def segments(self, g0: int = 0, g1: Optional[int] = None):
        """(offset, count, op, row_floats, row0) float ranges of the exchange storage that hold Gaussians [g0, g1): one
        SUM range per field of the packed buffer plus the MAX range of max_radii (which sits right behind the buffer
        when the storage is shared).  op: 0 = sum, 1 = max.  row_floats > 0: the range is row-sparse (rows of that many
        floats, the first one is Gaussian row0; only rows live on some rank are exchanged).  g0 must be a multiple of 4; a range that ends at P also covers
        the field's (zero) padding, so offsets and counts are always multiples of 4 floats."""
        g1 = self.P if g1 is None else g1
        assert g0 % 4 == 0 and (g1 % 4 == 0 or g1 == self.P)
        e1 = self.P4 if g1 == self.P else g1
        sparse = lambda name, w: w if (self.live_map is not None and w % 4 == 0 and
                                       name not in ("grad_accum", "denom")) else 0
        segs = [(off + w * g0, w * (e1 - g0), 0, sparse(name, w), g0) for name, off, w in self.fields]
        segs.append((self.buffer.numel() + g0, e1 - g0, 1, 0, 0))
        return segs

    def zero_stats_(self):
        self.views["grad_accum"].zero_()
        self.views["denom"].zero_()
        self.max_radii.zero_()

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in self.views.items() if k not in ("grad_accum", "denom")}


def _parr(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


class BatchWorkspace:
    """Persistent device buffers of a view batch (torch owns them): per view the geometry / image /
    binning buffers, the backward scratch and the output images.  The binning buffers have a capacity in
    (tile, Gaussian) pairs.  The per-Gaussian buffers are grow-only in P: densify / prune
    (geometry/gaussian_base.py:853-869) change P every few hundred steps, ``resize`` keeps the allocation when the
    new P fits (layouts are computed per call from P; the scratch is all-zero between steps whatever P was)."""

    def __init__(self, V: int, P: int, H: int, W: int, device, capacity_pairs: Optional[int] = None):
        assert 1 <= V <= MAX_VIEWS
        self.V, self.H, self.W, self.device = V, H, W, torch.device(device)
        u8 = lambda n: torch.empty(int(n), dtype=torch.uint8, device=device)
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
        self.image = [u8(lib.b200splat_image_bytes(H, W)) for _ in range(V)]
        self.color = [f32(3, H, W) for _ in range(V)]
        self.depth = [f32(1, H, W) for _ in range(V)]
        self.alpha = [f32(1, H, W) for _ in range(V)]
        self.P = self.P_alloc = 0
        self.resize(P)
        self.binning: List[torch.Tensor] = []
        self.binning_bytes = 0
        self.num_rendered = [0] * V
        self._alloc_binning(capacity_pairs or max(4 * P, 1 << 16))
        # early notice of every view's pair count (b200splat_batch_forward_args.pairs_notify): pinned host words the
        # scan kernel writes, polled by the host while the rest of the forward is still queued
        self.notify = None
        self.epoch = 0
        if self.device.type == "cuda":
            self.notify = torch.zeros(V, dtype=torch.int64).pin_memory()
            self._notify_np = self.notify.numpy()

    def resize(self, P: int):
        if P > self.P_alloc:
            alloc = P if self.P_alloc == 0 else max(P, int(self.P_alloc * 1.25))
            dev = self.device
            self.geom = [torch.empty(int(lib.b200splat_geom_bytes(alloc)), dtype=torch.uint8, device=dev)
                         for _ in range(self.V)]
            # all-zero on entry to every backward and left all-zero by it (b200splat_batch_backward_args.scratch_clean)
            self.scratch = [torch.zeros(int(lib.b200splat_backward_scratch_bytes(alloc)), dtype=torch.uint8, device=dev)
                            for _ in range(self.V)]
            self._radii_all = [torch.empty(alloc, dtype=torch.int32, device=dev) for _ in range(self.V)]
            self.P_alloc = alloc
        self.P = P
        self.radii = [r[:P] for r in self._radii_all]

    def _alloc_binning(self, pairs: int):
        self.binning_bytes = int(lib.b200splat_binning_bytes(int(pairs)))
        self.binning = [torch.empty(self.binning_bytes, dtype=torch.uint8, device=self.device) for _ in range(self.V)]
        self.capacity = int(lib.b200splat_binning_capacity(self.binning_bytes))

    def wait_pair_counts(self, epoch: int, timeout_s: float = 60.0) -> List[int]:
        """Spin until every view's notice of forward ``epoch`` has arrived; returns the V pair counts."""
        import time
        t0 = None
        while True:
            w = self._notify_np.copy()
            if bool(((w >> 32) == epoch).all()):
                return [int(x) for x in (w & 0xFFFFFFFF)]
            if t0 is None:
                t0 = time.perf_counter()
            elif time.perf_counter() - t0 > timeout_s:
                torch.cuda.synchronize(self.device)   # surfaces a sticky CUDA error, if that is why nothing arrived
                raise RuntimeError("b200splat: the forward's pair-count notice never arrived")

    def states(self, M: int):
        """Per-view ops.ForwardState (for ops.forward_views / the single-view backward)."""
        return [ops.ForwardState(self.P, M, self.capacity, self.geom[v], self.binning[v], self.image[v])
                for v in range(self.V)]


def forward_batched(ws: BatchWorkspace, cams: Sequence[ops.Cam], means3D, shs, colors_precomp, opacities, scales,
                    rotations, sync: bool = False, extra_features=None, extra_out=None, notify: bool = False):
    """One launch set for all views.  With sync=True returns (num_rendered list, overflow list).
    extra_features (P,C') + extra_out (list of V (C',H,W) tensors): extra channels blended by the same pass.
    notify=True: returns (epoch, None); ``ws.wait_pair_counts(epoch)`` then yields the views' pair counts as soon as
    the scan kernel has run, without synchronising the stream."""
    V = len(cams)
    assert V == ws.V
    M = 0 if shs is None else int(shs.shape[1])
    cam_arr = (_lib.Camera * V)(*[c.c_struct() for c in cams])
    a = _lib.BatchForwardArgs()
    a.V, a.cams, a.P, a.M = V, cam_arr, ws.P, M
    a.means3D, a.shs, a.colors_precomp = ops._ptr(means3D), ops._ptr(shs), ops._ptr(colors_precomp)
    a.opacities, a.scales, a.rotations = ops._ptr(opacities), ops._ptr(scales), ops._ptr(rotations)
    keep = [_parr([t.data_ptr() for t in lst]) for lst in (ws.color, ws.depth, ws.alpha, ws.radii, ws.geom,
                                                            ws.image, ws.binning)]
    a.out_color, a.out_depth, a.out_alpha, a.radii, a.geom_buffer, a.image_buffer, a.binning_buffer = keep
    a.binning_bytes = ws.binning_bytes
    a.stream = ops._stream(ws.device)
    a.sync = int(sync)
    if notify:
        ws.epoch = (ws.epoch % 0x7FFFFFFF) + 1
        a.pairs_notify, a.notify_epoch = ws.notify.data_ptr(), ws.epoch
    nr = (C.c_int64 * V)()
    ov = (C.c_int32 * V)()
    a.num_rendered_out, a.overflow_out = nr, ov
    n_extra = ops._check_extra(extra_features, ws.P)
    if n_extra:
        eo = _parr([t.data_ptr() for t in extra_out])
        a.extra_features, a.n_extra, a.out_extra = extra_features.data_ptr(), n_extra, eo
    with torch.cuda.device(ws.device):
        check(lib.b200splat_forward_batched(C.byref(a)), "b200splat_forward_batched")
    if sync:
        ws.num_rendered = [int(x) for x in nr]
        return ws.num_rendered, [int(x) for x in ov]
    return (ws.epoch if notify else None), None


def backward_batched(ws: BatchWorkspace, cams: Sequence[ops.Cam], means3D, shs, colors_precomp, opacities, scales,
                     rotations, pixel_grads, out: Dict[str, torch.Tensor], accumulate: bool = False,
                     stats=None, means2D_out: Optional[Sequence[Optional[torch.Tensor]]] = None,
                     phase: int = 0, g_range: Optional[tuple] = None, extra_features=None, extra_grads=None,
                     live_map: Optional[torch.Tensor] = None):
    """phase 0: whole backward; 1: render backward only; 2: preprocess backward only, Gaussians g_range=(g0, g1).
    extra_features (P,C') + extra_grads (list of V (C',H,W) tensors or None): out["extra_features"] receives
    dL/dextra_features summed over the views."""
    V = len(cams)
    M = 0 if shs is None else int(shs.shape[1])
    cam_arr = (_lib.Camera * V)(*[c.c_struct() for c in cams])
    a = _lib.BatchBackwardArgs()
    a.V, a.cams, a.P, a.M = V, cam_arr, ws.P, M
    a.means3D, a.shs, a.colors_precomp = ops._ptr(means3D), ops._ptr(shs), ops._ptr(colors_precomp)
    a.opacities, a.scales, a.rotations = ops._ptr(opacities), ops._ptr(scales), ops._ptr(rotations)
    gp = lambda i: _parr([ops._ptr(pg[i]) if pg is not None and pg[i] is not None else None for pg in pixel_grads])
    keep = [_parr([t.data_ptr() for t in lst]) for lst in (ws.radii, ws.geom, ws.image, ws.binning, ws.scratch)]
    a.radii, a.geom_buffer, a.image_buffer, a.binning_buffer, a.scratch = keep
    a.binning_bytes = ws.binning_bytes
    gc, gd, ga = gp(0), gp(1), gp(2)
    a.dL_dout_color, a.dL_dout_depth, a.dL_dout_alpha = gc, gd, ga
    m2 = None
    if means2D_out is not None:
        m2 = _parr([ops._ptr(t) for t in means2D_out])
        a.dL_dmeans2D = m2
    a.dL_dmeans3D, a.dL_dshs, a.dL_dcolors = ops._ptr(out["means3D"]), ops._ptr(out.get("shs")), \
        ops._ptr(out.get("colors_precomp"))
    a.dL_dopacity, a.dL_dscales, a.dL_drotations = ops._ptr(out["opacities"]), ops._ptr(out["scales"]), \
        ops._ptr(out["rotations"])
    a.accumulate = int(bool(accumulate))
    if stats is not None:
        a.stat_grad_accum, a.stat_denom, a.stat_max_radii = (ops._ptr(t) for t in stats)
    a.stream = ops._stream(ws.device)
    a.phase = int(phase)
    if g_range is not None:
        a.g_begin, a.g_end = int(g_range[0]), int(g_range[1])
    n_extra = ops._check_extra(extra_features, ws.P)
    if n_extra:
        eg = _parr([ops._ptr(t) for t in (extra_grads or [None] * V)])
        a.extra_features, a.n_extra = extra_features.data_ptr(), n_extra
        a.dL_dout_extra, a.dL_dextra = eg, out["extra_features"].data_ptr()
    if live_map is not None:
        a.live_map = live_map.data_ptr()
    a.scratch_clean = 1   # ws.scratch is zero on entry; the library leaves it zero (it re-zeroes what it consumed)
    try:
        with torch.cuda.device(ws.device):
            check(lib.b200splat_backward_batched(C.byref(a)), "b200splat_backward_batched")
    except Exception:
        for t in ws.scratch:   # a failed launch may leave partial sums behind: restore the invariant
            t.zero_()
        raise


class BatchRenderer:
    """fwd+bwd of a step's views with persistent buffers.  ``step`` returns nothing: gradients and
    statistics are in ``packed``; rendered images stay in ``ws.color/depth/alpha`` until the next step."""

    def __init__(self, P: int, M: int, H: int, W: int, device, views: int, color_mode: str = "shs",
                 packed_storage: Optional[torch.Tensor] = None):
        self.P, self.M, self.H, self.W, self.device = P, M, H, W, device
        self.chunks = [min(MAX_VIEWS, views - i) for i in range(0, views, MAX_VIEWS)]
        self.ws = [BatchWorkspace(v, P, H, W, device) for v in self.chunks]
        self.packed = PackedGrads(P, M, device, color_mode, storage=packed_storage)
        self.calibrated = False

    def calibrate(self, cams, means3D, shs, colors_precomp, opacities, scales, rotations, headroom: float = 1.25):
        """Size the binning buffers from one synchronous forward (num_rendered changes slowly between steps)."""
        i = 0
        for ws, n in zip(self.ws, self.chunks):
            while True:
                nr, ov = forward_batched(ws, cams[i:i + n], means3D, shs, colors_precomp, opacities, scales, rotations,
                                         sync=True)
                need = int(max(nr) * headroom) + 4096
                if any(ov) or ws.capacity < max(nr):
                    ws._alloc_binning(max(need, 2 * ws.capacity))
                    continue
                if ws.capacity > 2 * need or ws.capacity < need:
                    ws._alloc_binning(need)
                    continue
                break
            i += n
        self.calibrated = True

    def step(self, cams, means3D, shs, colors_precomp, opacities, scales, rotations, pixel_grads):
        """pixel_grads[v] = (dL/dcolor, dL/ddepth|None, dL/dalpha|None) or a callable(color, depth, alpha)."""
        if not self.calibrated:
            self.calibrate(cams, means3D, shs, colors_precomp, opacities, scales, rotations)
        pk = self.packed
        pk.zero_stats_()
        out = pk.grads()
        stats = (pk.views["grad_accum"], pk.views["denom"], pk.max_radii)
        i = 0
        for ci, (ws, n) in enumerate(zip(self.ws, self.chunks)):
            cs = cams[i:i + n]
            forward_batched(ws, cs, means3D, shs, colors_precomp, opacities, scales, rotations, sync=False)
            pgs = []
            for v in range(n):
                pg = pixel_grads[i + v]
                pgs.append(pg(ws.color[v], ws.depth[v], ws.alpha[v]) if callable(pg) else pg)
            backward_batched(ws, cs, means3D, shs, colors_precomp, opacities, scales, rotations, pgs, out,
                             accumulate=ci > 0, stats=stats, live_map=pk.live_map)
            i += n

    def step_head(self, cams, means3D, shs, colors_precomp, opacities, scales, rotations, pixel_grads):
        """First part of a step whose gradients are exchanged chunk by chunk (multi-GPU): forward + render
        backward of the view batch.  ``step_tail`` finishes it.  One view chunk (<= MAX_VIEWS views) only."""
        assert len(self.ws) == 1, "chunked exchange needs the step's views in one batch"
        if not self.calibrated:
            self.calibrate(cams, means3D, shs, colors_precomp, opacities, scales, rotations)
        pk, ws = self.packed, self.ws[0]
        pk.zero_stats_()
        forward_batched(ws, cams, means3D, shs, colors_precomp, opacities, scales, rotations, sync=False)
        pgs = [pg(ws.color[v], ws.depth[v], ws.alpha[v]) if callable(pg) else pg for v, pg in enumerate(pixel_grads)]
        backward_batched(ws, cams, means3D, shs, colors_precomp, opacities, scales, rotations, pgs, pk.grads(),
                         stats=(pk.views["grad_accum"], pk.views["denom"], pk.max_radii), phase=1)

    def step_tail(self, cams, means3D, shs, colors_precomp, opacities, scales, rotations, exchange, chunks: int = 4):
        """Preprocess backward in ``chunks`` Gaussian ranges; after each range ``exchange(g0, g1)`` runs on a side
        stream (all-reduce of that range of the packed buffer across the ranks), so the exchange of range c is
        under way while range c+1 is computed.  Returns with the side stream joined."""
        pk, ws, P = self.packed, self.ws[0], self.P
        main = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_side"):
            # high priority: the exchange kernel's CTAs take SM slots as the running compute CTAs retire, instead of
            # queueing behind the whole next preprocess-backward range (which fills every SM's registers)
            self._side = torch.cuda.Stream(device=self.device, priority=-1)
            self._events = [torch.cuda.Event() for _ in range(16)]
        side = self._side
        step = ((P + chunks - 1) // chunks + 127) // 128 * 128   # whole CTAs of the kernel, multiples of 4 Gaussians
        none_pg = [None] * ws.V
        for c, g0 in enumerate(range(0, P, step)):
            g1 = min(P, g0 + step)
            backward_batched(ws, cams, means3D, shs, colors_precomp, opacities, scales, rotations, none_pg,
                             pk.grads(), stats=(pk.views["grad_accum"], pk.views["denom"], pk.max_radii),
                             phase=2, g_range=(g0, g1), live_map=pk.live_map)
            ev = self._events[c % len(self._events)]
            ev.record(main)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                exchange(g0, g1)
        main.wait_stream(side)

    def capture_step(self, cams, means3D, shs, colors_precomp, opacities, scales, rotations, pixel_grads,
                     head_only: bool = False, exchange=None, chunks: int = 1):
        """Record one step (every launch of the view batch's forward + backward) into a CUDA graph and return it;
        ``graph.replay()`` then re-runs the step on the same buffers without per-launch host work.  The inputs
        must keep their addresses (parameters updated in place, cameras / pixel gradients written into the same
        tensors); pixel_grads must be tensors, not callables.  head_only: record ``step_head`` (the part before
        the chunked preprocess backward / exchange of a multi-GPU step).
        exchange(g0, g1): the multi-GPU gradient exchange of Gaussians [g0, g1) -- recorded INTO the graph (own
        kernels with a device-side call counter only: ``P2PAllReduce`` / ``MulticastAllReduce`` with
        ``device_epoch=True``): after the whole step when ``chunks`` <= 1, else range by range on a forked stream while
        the next range's preprocess backward runs (``step_tail``)."""
        assert self.calibrated, "calibrate() first: the binning capacity is baked into the graph"
        assert not any(callable(pg) for pg in pixel_grads)
        args = (cams, means3D, shs, colors_precomp, opacities, scales, rotations, pixel_grads)
        params = (cams, means3D, shs, colors_precomp, opacities, scales, rotations)

        def fn():
            if exchange is None:
                (self.step_head if head_only else self.step)(*args)
            elif chunks <= 1:
                self.step(*args)
                exchange(0, self.P)
            else:
                self.step_head(*args)
                self.step_tail(*params, exchange, chunks=chunks)

        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            fn()   # eager once on the side stream: one-time attribute / allocation work happens here
            if head_only and exchange is None:   # leave the scratch clean (the eager head's records are consumed by a full tail)
                backward_batched(self.ws[0], cams, means3D, shs, colors_precomp, opacities, scales, rotations,
                                 [None] * self.ws[0].V, self.packed.grads(), phase=2)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        return graph

    def overflowed(self) -> bool:
        """Lazy overflow check (one small D2H): True if any view of the last step exceeded its capacity."""
        flags = []
        for ws in self.ws:
            for v in range(ws.V):
                st = ops.ForwardState(self.P, self.M, ws.capacity, ws.geom[v], ws.binning[v], ws.image[v])
                flags.append(ops.status_tensor(self.H, self.W, st))
        bad = bool(torch.stack(flags)[:, 0].any().item())
        if bad:
            self.calibrated = False
        return bad


def render_views_fwd_bwd(cams: Sequence[ops.Cam], means3D, shs, colors_precomp, opacities, scales, rotations,
                         pixel_grads, packed: PackedGrads, keep_images: bool = False):
    """Per-view reference loop over the single-view entry points (the shape of the reference's own
    batch_forward): forward + backward of every view, gradients accumulated into ``packed``.  Kept as the
    baseline the batched path is tested against."""
    packed.zero_stats_()
    out = dict(packed.grads())
    out["means2D"] = packed.means2D_scratch
    stats = (packed.views["grad_accum"], packed.views["denom"], packed.max_radii)
    images = []
    for v, cam in enumerate(cams):
        color, radii, depth, alpha, st = ops.forward(cam, means3D, shs, colors_precomp, opacities, scales,
                                                     rotations, None)
        pg = pixel_grads[v]
        if callable(pg):
            pg = pg(color, depth, alpha)
        ops.backward(cam, st, means3D, shs, colors_precomp, opacities, scales, rotations, None, radii, alpha,
                     pg[0], pg[1], pg[2], out=out, accumulate=v > 0, stats=stats)
        if keep_images:
            images.append((color, depth, alpha, radii))
    return images


# ------------------------------------------------------------------------------------------------------
# autograd surface of the batched path: the operator a GaussianBatchRenderer-style caller uses instead of
# a Python loop over single-view GaussianRasterizer calls (renderer/gaussian_batch_renderer.py:21-54)
# ------------------------------------------------------------------------------------------------------
class _RasterizeViews(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, cams, means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                extra_features=None):
        ws = owner.ws
        V, P, H, W = ws.V, ws.P, ws.H, ws.W
        dev = means3D.device
        f = ops._f32c
        m3, sh_, cp_, op_ = f(means3D, "means3D"), f(shs, "shs"), f(colors_precomp, "colors_precomp"), \
            f(opacities, "opacities")
        sc_, ro_ = f(scales, "scales"), f(rotations, "rotations")
        # fresh outputs every call (callers write into them in place); the workspace only lends scratch.
        # owner.single: the one-view operator (GaussianRasterizer) -- outputs without the view dimension
        lead = () if owner.single else (V,)
        unb = (lambda t: [t]) if owner.single else (lambda t: list(t.unbind(0)))
        color = torch.empty(*lead, 3, H, W, dtype=torch.float32, device=dev)
        depth = torch.empty(*lead, 1, H, W, dtype=torch.float32, device=dev)
        alpha = torch.empty(*lead, 1, H, W, dtype=torch.float32, device=dev)
        radii = torch.empty(*lead, P, dtype=torch.int32, device=dev)
        ws.color, ws.depth, ws.alpha = unb(color), unb(depth), unb(alpha)
        ws.radii = unb(radii)
        n_extra = ops._check_extra(extra_features, P)
        ex_ = f(extra_features, "extra_features") if n_extra else None
        extra = torch.empty(*lead, n_extra, H, W, dtype=torch.float32, device=dev)
        ekw = dict(extra_features=ex_, extra_out=unb(extra)) if n_extra else {}
        owner.render_checked(cams, m3, sh_, cp_, op_, sc_, ro_, **ekw)
        owner.generation += 1
        ctx.owner, ctx.cams, ctx.generation = owner, cams, owner.generation
        ctx.has = (sh_ is not None, cp_ is not None)
        e = lambda t: t if t is not None else torch.empty(0, device=dev)
        ctx.n_extra = n_extra
        ctx.save_for_backward(m3, e(sh_), e(cp_), op_, sc_, ro_, e(ex_))
        ctx.mark_non_differentiable(radii)
        owner._ctx = weakref.ref(ctx)   # the live graph (if any) that owns the workspace: see ViewBatchRasterizer.pending
        return color, radii, depth, alpha, extra

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha, g_extra):
        owner = ctx.owner
        if ctx.generation != owner.generation:
            raise RuntimeError("ViewBatchRasterizer: the workspace was reused by a later forward before this "
                               "backward ran; use one ViewBatchRasterizer per live graph")
        ws = owner.ws
        m3, sh_, cp_, op_, sc_, ro_, ex_ = ctx.saved_tensors
        has_sh, has_cp = ctx.has
        n_extra = ctx.n_extra
        sh_ = sh_ if has_sh else None
        cp_ = cp_ if has_cp else None
        V, P = ws.V, ws.P
        dev = m3.device
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        out = {"means3D": new(P, 3), "opacities": new(P, 1), "scales": new(P, 3), "rotations": new(P, 4)}
        if has_sh:
            out["shs"] = new(P, sh_.shape[1], 3)
        if has_cp:
            out["colors_precomp"] = new(P, 3)
        single = owner.single
        m2 = new(P, 3) if single else new(V, P, 3)
        g = lambda t: None if t is None else ops._f32c(t, "grad")
        gc, gd, ga = g(g_color), g(g_depth), g(g_alpha)
        pick = (lambda t, v: t) if single else (lambda t, v: t[v])
        pgs = [(None if gc is None else pick(gc, v), None if gd is None else pick(gd, v),
                None if ga is None else pick(ga, v)) for v in range(V)]
        ekw = {}
        if n_extra:
            out["extra_features"] = new(P, n_extra)
            ge = g(g_extra)
            ekw = dict(extra_features=ex_, extra_grads=None if ge is None else ([ge] if single else list(ge.unbind(0))))
        backward_batched(ws, ctx.cams, m3, sh_, cp_, op_, sc_, ro_, pgs, out, accumulate=False,
                         means2D_out=[m2] if single else list(m2.unbind(0)), **ekw)
        return (None, None, out["means3D"], m2, out.get("shs"), out.get("colors_precomp"), out["opacities"],
                out["scales"], out["rotations"], out.get("extra_features"))


class ViewBatchRasterizer(torch.nn.Module):
    """Batched counterpart of ``GaussianRasterizer``: ``forward(raster_settings_list, means3D, means2D (V,P,3),
    opacities, shs=None, colors_precomp=None, scales=, rotations=)`` -> ``(color (V,3,H,W), radii (V,P) int32,
    depth (V,1,H,W), alpha (V,1,H,W))`` (+ ``extra (V,C',H,W)`` when ``extra_features (P,C')`` is given) with autograd; ``means2D.grad[v]`` is view v's NDC gradient, exactly what
    the per-view call would have produced (geometry/gaussian_base.py:815-819 reads it per view).  Holds the
    persistent device workspace of the batch; one instance per concurrently live autograd graph.

    The number of Gaussians may change between calls (densify / prune): the workspace grows when it has to.
    The binning capacity heals itself: every forward learns its views' pair counts from the scan kernel through
    pinned host memory (no stream synchronisation -- the rest of the forward is already queued behind the scan) and,
    if a view outgrew the capacity (closer camera, narrower fov, grown scales), enlarges the buffers and renders
    again before anything is returned; a view can never silently come back as background."""

    HEADROOM = 1.5

    def __init__(self, views: int, P: int, H: int, W: int, device="cuda", single: bool = False):
        super().__init__()
        assert views <= MAX_VIEWS, f"at most {MAX_VIEWS} views per batch call"
        assert not single or views == 1
        self.single = single      # one-view operator: tensors without the leading view dimension (see one_view_forward)
        self.ws = BatchWorkspace(views, P, H, W, torch.device(device))
        self.generation = 0
        self.regrown = 0          # how many times a forward had to be repeated with larger binning buffers
        self.last_pairs: List[int] = []
        self._ctx = None

    @property
    def pending(self) -> bool:
        """True while an autograd graph recorded by this rasterizer can still run its backward: the graph's node is
        alive and its saved tensors have not been released (they are released by a backward without
        ``retain_graph`` -- the reference runs two backward passes over one graph, system/gaussian_splatting.py:
        129-138 -- or when the outputs are dropped)."""
        ctx = self._ctx() if self._ctx is not None else None
        if ctx is None:
            return False
        try:
            ctx.saved_tensors
        except RuntimeError:
            return False
        return True

    def render_checked(self, cams, means3D, shs, colors_precomp, opacities, scales, rotations, **extra):
        """The batch's forward with the capacity check described in the class docstring."""
        ws = self.ws
        for attempt in range(4):
            epoch, _ = forward_batched(ws, cams, means3D, shs, colors_precomp, opacities, scales, rotations, sync=False,
                                       notify=True, **extra)
            pairs = ws.wait_pair_counts(epoch)
            self.last_pairs = pairs
            if max(pairs) <= ws.capacity:
                return
            self.regrown += 1
            ws._alloc_binning(int(max(pairs) * self.HEADROOM) + 4096)
        raise RuntimeError("ViewBatchRasterizer: the binning buffers still overflow after regrowing them")

    def calibrate(self, cams, means3D, shs, colors_precomp, opacities, scales, rotations, headroom: float = 1.5,
                  **extra):
        """Optional: size the binning buffers ahead of the first step (a forward does it on demand)."""
        ws = self.ws
        nr, ov = forward_batched(ws, cams, means3D, shs, colors_precomp, opacities, scales, rotations, sync=True, **extra)
        need = int(max(nr) * headroom) + 4096
        if any(ov) or ws.capacity < need:
            ws._alloc_binning(need)

    def check_overflow(self) -> bool:
        """One small device read of the views' status words; True if a view of the LAST forward exceeded the
        capacity.  ``forward`` repairs that itself, so this is a cross-check (tests, bench)."""
        ws = self.ws
        flags = [ops.status_tensor(ws.H, ws.W, st) for st in ws.states(0)]
        return bool(torch.stack(flags)[:, 0].any().item())

    def forward(self, raster_settings, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, extra_features=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if scales is None or rotations is None:
            raise Exception("The batched rasterizer needs the scale/rotation pair")
        ws = self.ws
        if means3D.shape[0] != ws.P:
            if self.pending:
                raise RuntimeError("ViewBatchRasterizer: the number of Gaussians changed while a graph recorded with "
                                   "the old one is still waiting for its backward")
            ws.resize(int(means3D.shape[0]))
        cams = [ops.make_cam(rs, means3D.device) for rs in raster_settings]
        out = _RasterizeViews.apply(self, cams, means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                    extra_features)
        if not (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (
                means3D, means2D, opacities, shs, colors_precomp, scales, rotations, extra_features))):
            self._ctx = None
        return out if extra_features is not None else out[:4]


# ------------------------------------------------------------------------------------------------------
# The one-view operator on pooled persistent workspaces: what ``diff_gaussian_rasterization.GaussianRasterizer``
# (the reference's unchanged per-view loop, renderer/gaussian_batch_renderer.py:21-54) runs on.  Compared with the
# allocate-per-call single-view entry point it has no host round trip in the middle of the forward (the pair count is
# learnt from the scan kernel's notice while the rest of the forward is already queued), no per-call allocation of the
# geometry / binning / scratch buffers and no 64 MB scratch memset per backward (self-cleaning scratch).
# ------------------------------------------------------------------------------------------------------
_ONE_VIEW_POOL: Dict[tuple, List["ViewBatchRasterizer"]] = {}
ONE_VIEW_POOL_MAX = 16   # live graphs per (device, H, W): 4 views per step, two batches per step in the zero123 systems


def one_view_rasterizer(P: int, H: int, W: int, device) -> Optional["ViewBatchRasterizer"]:
    """A pooled one-view rasterizer whose last graph has been consumed, or None when ONE_VIEW_POOL_MAX graphs of this
    shape are still alive (the caller then takes the allocating path)."""
    pool = _ONE_VIEW_POOL.setdefault((str(device), H, W), [])
    for rast in pool:
        if not rast.pending:
            return rast
    if len(pool) >= ONE_VIEW_POOL_MAX:
        return None
    pool.append(ViewBatchRasterizer(1, P, H, W, device, single=True))
    return pool[-1]


def one_view_forward(rast: "ViewBatchRasterizer", cam: ops.Cam, means3D, means2D, shs, colors_precomp, opacities, scales,
                     rotations, extra_features=None):
    """-> (color (3,H,W), radii (P,), depth (1,H,W), alpha (1,H,W), extra (C',H,W)) with autograd."""
    if means3D.shape[0] != rast.ws.P:
        rast.ws.resize(int(means3D.shape[0]))
    return _RasterizeViews.apply(rast, [cam], means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                 extra_features)
