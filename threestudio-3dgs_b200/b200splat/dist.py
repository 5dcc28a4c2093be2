"""View-data-parallel plumbing: views of a step are split across the GPUs of one box, Gaussians are
replicated, and the packed gradient buffer + densification accumulators are all-reduced (NCCL over
NVLink 5 / NVSwitch on GPUs, gloo in the CPU tests).  The reference has no distributed code at all
(SURVEY.md 2.1); this is new capability behind the same per-view operator.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the step's views owned by ``rank`` (ragged allowed)."""
    base, rem = divmod(num_views, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_packed(buffer: torch.Tensor, max_radii: torch.Tensor, group=None):
    """SUM over ranks of the packed gradient/statistics buffer, MAX of max_radii.  No-op at world 1."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(buffer, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(max_radii, op=dist.ReduceOp.MAX, group=group)


def allreduce_packed_range(packed, g0: int, g1: int, group=None, with_max: bool = True):
    """NCCL/gloo version of the chunked exchange: the fields of the packed buffer (and max_radii), Gaussians [g0, g1).
    The per-field SUM all-reduces of a range are coalesced into ONE collective launch (ncclGroupStart/End) -- seven
    separate launches per range cost more than the overlap of a range's exchange with the next range's compute
    returns."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    e1 = packed.P4 if g1 == packed.P else g1
    slices = [packed.buffer[off + w * g0:off + w * e1] for _, off, w in packed.fields]
    if dist.get_backend(group) == "nccl":   # gloo has no coalescing (and a failed attempt leaves its queue open)
        from torch.distributed.distributed_c10d import _coalescing_manager
        with _coalescing_manager(group=group, device=packed.buffer.device, async_ops=False):
            for t in slices:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    else:
        for t in slices:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if with_max:
        dist.all_reduce(packed.max_radii[g0:g1], op=dist.ReduceOp.MAX, group=group)


class _RawCudaArray:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class P2PAllReduce:
    """In-place SUM (+ MAX tail) all-reduce of one fp32 buffer per rank over NVLink peer memory
    (include/b200splat.h b200splat_p2p_*, csrc/p2p.cu): every rank's buffer is cudaMalloc memory exported with
    CUDA IPC and mapped by all peers; one kernel per rank reduces its slice from all ranks and stores the result to
    all ranks.  ``buffer`` is the rank's exchange tensor (n_sum + n_max floats): build the step's PackedGrads on
    it (``PackedGrads(..., storage=ar.buffer)``) and call ``ar()`` after the backward.  One process per GPU of
    ONE box; the handles travel through ``torch.distributed`` (any backend)."""

    def __init__(self, n_sum: int, n_max: int, device, group=None, device_epoch: bool = False, n_tail: int = 0):
        """device_epoch: the kernel keeps the call counter in the rank's signal words (epoch argument 0), so a call can
        be captured in a CUDA graph and replayed; all ranks must choose the same mode.  n_tail: floats allocated behind
        the MAX part that are never exchanged (``PackedGrads`` keeps the row-sparse exchange's live map there;
        ``live_offset`` then tells the kernel where it is)."""
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        assert n_sum % 4 == 0 and n_max % 4 == 0, "segment lengths must be multiples of 4 floats"
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world <= 8, "one NVSwitch box: at most 8 ranks"
        self.n_sum, self.n_max, self.device, self.epoch = n_sum, n_max, torch.device(device), 0
        self.device_epoch = device_epoch
        self.n_tail, self.live_offset = int(n_tail), 0
        lib = _lib.lib
        # every step below is collective-safe: a rank that fails still takes part in the exchange of the status, so
        # all ranks raise together instead of some waiting in a collective for a peer that already gave up
        self._own, self._mapped = None, []
        self._bufs, self._sigs = [0] * self.world, [0] * self.world
        err = None
        with torch.cuda.device(self.device):
            buf, sig = C.c_void_p(), C.c_void_p()
            hb, hs = C.create_string_buffer(_lib.P2P_HANDLE_BYTES), C.create_string_buffer(_lib.P2P_HANDLE_BYTES)
            try:
                _lib.check(lib.b200splat_p2p_alloc((n_sum + n_max + n_tail) * 4, C.byref(buf), hb), "b200splat_p2p_alloc")
                _lib.check(lib.b200splat_p2p_alloc(_lib.P2P_SIGNAL_BYTES, C.byref(sig), hs), "b200splat_p2p_alloc")
                self._own = (buf.value, sig.value)
            except Exception as exc:
                err = repr(exc)
            handles: list = [None] * self.world
            dist.all_gather_object(handles, (err, hb.raw, hs.raw), group=group)
            if not any(h[0] for h in handles):
                try:
                    for k, (_, b_h, s_h) in enumerate(handles):
                        if k == self.rank:
                            self._bufs[k], self._sigs[k] = self._own
                            continue
                        pb, ps = C.c_void_p(), C.c_void_p()
                        _lib.check(lib.b200splat_p2p_open(b_h, C.byref(pb)), "b200splat_p2p_open")
                        self._mapped.append(pb.value)
                        _lib.check(lib.b200splat_p2p_open(s_h, C.byref(ps)), "b200splat_p2p_open")
                        self._mapped.append(ps.value)
                        self._bufs[k], self._sigs[k] = pb.value, ps.value
                except Exception as exc:
                    err = repr(exc)
            status: list = [None] * self.world
            dist.all_gather_object(status, err or next((h[0] for h in handles if h[0]), None), group=group)
        bad = [f"rank {k}: {e}" for k, e in enumerate(status) if e]
        if bad:
            self.buffer = None
            self._release()
            raise RuntimeError("P2PAllReduce setup failed (" + "; ".join(bad) + ")")
        self.buffer = torch.as_tensor(_RawCudaArray(self._own[0], n_sum + n_max + n_tail, "<f4"), device=self.device)
        dist.barrier(group=group)   # every rank has mapped every peer before the first exchange

    def __call__(self, segments=None):
        """Stream-ordered on the current stream: when the kernel completes the listed (offset, count, op) float
        ranges of the buffer (default: the whole SUM part and the MAX tail) hold the reduced values.  All ranks must
        make the same sequence of calls."""
        C, _lib = self._C, self._lib
        if segments is None:
            segments = [(0, self.n_sum, 0)] + ([(self.n_sum, self.n_max, 1)] if self.n_max else [])
        segments = [sg for sg in segments if sg[1] > 0]
        assert 1 <= len(segments) <= 16
        self.epoch += 1
        a = _lib.P2PArgs()
        a.rank, a.world, a.epoch, a.n_segments = self.rank, self.world, 0 if self.device_epoch else self.epoch, len(segments)
        for i, sg in enumerate(segments):
            a.seg_offset[i], a.seg_count[i], a.seg_op[i] = int(sg[0]), int(sg[1]), int(sg[2])
            if len(sg) > 3 and sg[3] and self.live_offset:    # row-sparse range (PackedGrads.segments)
                a.seg_row_floats[i], a.seg_row0[i] = int(sg[3]), int(sg[4])
        a.live_offset = int(self.live_offset)
        for k in range(self.world):
            a.bufs[k], a.signals[k] = self._bufs[k], self._sigs[k]
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.b200splat_p2p_allreduce(C.byref(a)), "b200splat_p2p_allreduce")

    def failed(self) -> bool:
        """True if a peer never showed up within the kernel's spin bound (synchronises the device)."""
        C = self._C
        torch.cuda.synchronize(self.device)
        flag = C.c_int32(0)
        self._lib.check(self._lib.lib.b200splat_p2p_error(self._own[1], C.byref(flag)), "b200splat_p2p_error")
        return flag.value != 0

    def _release(self):
        lib = self._lib.lib
        for p in self._mapped:
            lib.b200splat_p2p_close(p)
        self._mapped = []
        if self._own:
            lib.b200splat_p2p_free(self._own[0])
            lib.b200splat_p2p_free(self._own[1])
            self._own = None

    def close(self):
        torch.cuda.synchronize(self.device)
        self.buffer = None
        self._release()


class MulticastAllReduce:
    """In-place SUM (+ MAX tail) all-reduce of one fp32 buffer per rank through NVSwitch multicast
    (b200splat_mc_allreduce, csrc/p2p.cu: multimem.ld_reduce / multimem.st -- the switch adds the ranks' copies on the
    way in and replicates the result on the way out, N/N of the buffer per NVLink direction instead of 2 (N-1)/N).
    Same use as ``P2PAllReduce``: build the step's PackedGrads on ``buffer`` and call the object after the backward.
    The symmetric allocation, the multicast binding and the peer-mapped signal pads are torch's
    (``torch.distributed._symmetric_memory``: plumbing); the data path is this library's kernel.  Raises at
    construction -- on all ranks together -- when the box has no multicast support."""

    SIGNAL_OFFSET = 4096   # bytes into torch's signal pad (its own barrier uses the first words)

    def __init__(self, n_sum: int, n_max: int, device, group=None, device_epoch: bool = False, n_tail: int = 0):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib, self._C = _lib, C
        assert n_sum % 4 == 0 and n_max % 4 == 0, "segment lengths must be multiples of 4 floats"
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world <= 8, "one NVSwitch box: at most 8 ranks"
        self.n_sum, self.n_max, self.device, self.epoch = n_sum, n_max, torch.device(device), 0
        self.device_epoch = device_epoch
        self.n_tail, self.live_offset = int(n_tail), 0
        err = None
        try:
            with torch.cuda.device(self.device):
                self.buffer = symm_mem.empty(n_sum + n_max + n_tail, dtype=torch.float32, device=self.device)
                self._hdl = symm_mem.rendezvous(self.buffer, group)
            self._mc = int(self._hdl.multicast_ptr)
            if not self._mc:
                raise RuntimeError("no multicast address (NVSwitch multicast unsupported here)")
            if self._hdl.signal_pad_size < self.SIGNAL_OFFSET + _lib.P2P_SIGNAL_BYTES:
                raise RuntimeError("signal pad too small")
            self._sigs = [int(p) + self.SIGNAL_OFFSET for p in self._hdl.signal_pad_ptrs]
            self.buffer.zero_()
        except Exception as exc:
            err = repr(exc)
        status: list = [None] * self.world
        dist.all_gather_object(status, err, group=group)
        bad = [f"rank {k}: {e}" for k, e in enumerate(status) if e]
        if bad:
            self.buffer = None
            raise RuntimeError("MulticastAllReduce setup failed (" + "; ".join(bad) + ")")
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    def __call__(self, segments=None):
        """Stream-ordered on the current stream; see ``P2PAllReduce.__call__``."""
        C, _lib = self._C, self._lib
        if segments is None:
            segments = [(0, self.n_sum, 0)] + ([(self.n_sum, self.n_max, 1)] if self.n_max else [])
        segments = [sg for sg in segments if sg[1] > 0]
        assert 1 <= len(segments) <= 16
        self.epoch += 1
        a = _lib.MCArgs()
        a.rank, a.world, a.epoch, a.n_segments = self.rank, self.world, 0 if self.device_epoch else self.epoch, len(segments)
        a.mc_buffer = self._mc
        for i, sg in enumerate(segments):
            a.seg_offset[i], a.seg_count[i], a.seg_op[i] = int(sg[0]), int(sg[1]), int(sg[2])
            if len(sg) > 3 and sg[3] and self.live_offset:    # row-sparse range (PackedGrads.segments)
                a.seg_row_floats[i], a.seg_row0[i] = int(sg[3]), int(sg[4])
        a.live_offset = int(self.live_offset)
        for k in range(self.world):
            a.signals[k] = self._sigs[k]
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.b200splat_mc_allreduce(C.byref(a)), "b200splat_mc_allreduce")

    def failed(self) -> bool:
        C = self._C
        torch.cuda.synchronize(self.device)
        flag = C.c_int32(0)
        self._lib.check(self._lib.lib.b200splat_p2p_error(self._sigs[self.rank], C.byref(flag)), "b200splat_p2p_error")
        return flag.value != 0

    def close(self):
        torch.cuda.synchronize(self.device)
        self.buffer, self._hdl = None, None


def reference_update_states(xyz_gradient_accum, denom, max_radii2D, grad_accum_step, denom_step, max_radii_step):
    """Apply one step's (all-reduced) statistics to the persistent accumulators exactly as the
    reference's per-view loop does (geometry/gaussian_base.py:815-819, 846-851)."""
    xyz_gradient_accum += grad_accum_step.reshape(xyz_gradient_accum.shape)
    denom += denom_step.reshape(denom.shape)
    torch.maximum(max_radii2D, max_radii_step.reshape(max_radii2D.shape), out=max_radii2D)
