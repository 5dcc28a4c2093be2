"""Thin torch <-> C-ABI glue: torch owns device memory and streams, libb200splat.so does the work.

Every function here requires CUDA tensors and raises if handed anything else: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch

from . import _lib
from ._lib import lib, check


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"b200splat: {name} must be a CUDA tensor (no CPU fallback); got {t.device}")
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def _stream(dev=None) -> int:
    """Handle of torch's current stream ON THE TENSORS' DEVICE (not on whatever device is current)."""
    return torch.cuda.current_stream(dev).cuda_stream


class Cam(NamedTuple):
    """Host mirror of GaussianRasterizationSettings with tensors normalised to contiguous fp32 CUDA."""
    H: int
    W: int
    tanfovx: float
    tanfovy: float
    scale_modifier: float
    sh_degree: int
    prefiltered: bool
    debug: bool
    bg: torch.Tensor
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    campos: torch.Tensor
    scalars: Optional[torch.Tensor] = None   # device (focal_x, focal_y, limx, limy) when tanfov is a device tensor

    def c_struct(self) -> _lib.Camera:
        return _lib.Camera(self.H, self.W, float(self.tanfovx), float(self.tanfovy), float(self.scale_modifier),
                           int(self.sh_degree), int(bool(self.prefiltered)), int(bool(self.debug)),
                           self.bg.data_ptr(), self.viewmatrix.data_ptr(), self.projmatrix.data_ptr(),
                           self.campos.data_ptr(), None if self.scalars is None else self.scalars.data_ptr())


def camera_scalars(tanfovx: torch.Tensor, tanfovy: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """(..., 4) fp32 (focal_x, focal_y, limx, limy) from device tangents, in the operation order of the C side
    (csrc/api.cu fill_camera): W / (2 tan), H / (2 tan), 1.3 tan -- one IEEE fp32 operation each, so the values equal
    what the host computes from the same fp32 tangents bit for bit."""
    tx, ty = tanfovx.float(), tanfovy.float()
    return torch.stack([float(W) / (2.0 * tx), float(H) / (2.0 * ty), tx * 1.3, ty * 1.3], dim=-1).contiguous()


def make_cam(settings, device) -> Cam:
    t = lambda x, n: _f32c(torch.as_tensor(x, device=device) if not torch.is_tensor(x) else x.to(device), n)
    H, W = int(settings.image_height), int(settings.image_width)
    tx, ty, scalars = settings.tanfovx, settings.tanfovy, None
    if torch.is_tensor(tx) and tx.is_cuda:   # the field of view lives on the device: keep it there (no .item() sync)
        scalars = camera_scalars(tx.reshape(()), (ty if torch.is_tensor(ty) else tx.new_tensor(ty)).reshape(()), H, W)
        tx = ty = 0.0
    return Cam(H, W, float(tx), float(ty), float(settings.scale_modifier), int(settings.sh_degree),
               bool(settings.prefiltered), bool(settings.debug), t(settings.bg, "bg").reshape(-1),
               t(settings.viewmatrix, "viewmatrix"), t(settings.projmatrix, "projmatrix"),
               t(settings.campos, "campos").reshape(-1), scalars)


class ForwardState(NamedTuple):
    P: int
    M: int
    num_rendered: int
    geom: torch.Tensor
    binning: Optional[torch.Tensor]
    image: torch.Tensor


MAX_EXTRA = 4


def _check_extra(extra_features, P):
    if extra_features is None:
        return 0
    if extra_features.dim() != 2 or extra_features.shape[0] != P or not 1 <= extra_features.shape[1] <= MAX_EXTRA:
        raise ValueError(f"extra_features must be (P, 1..{MAX_EXTRA}); got {tuple(extra_features.shape)}")
    return int(extra_features.shape[1])


def forward(cam: Cam, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
            extra_features=None):
    """Returns color (3,H,W), radii (P,) int32, depth (1,H,W), alpha (1,H,W), ForwardState; with
    ``extra_features`` (P,C'), C' <= 4, additionally the (C',H,W) image of those channels blended by the same
    pass (inserted before the ForwardState)."""
    dev = means3D.device
    if not means3D.is_cuda:
        raise RuntimeError("b200splat.forward: tensors must be on a CUDA device (no CPU fallback)")
    P = means3D.shape[0]
    M = 0 if shs is None else int(shs.shape[1])
    H, W = cam.H, cam.W
    color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
    depth = torch.empty(1, H, W, dtype=torch.float32, device=dev)
    alpha = torch.empty(1, H, W, dtype=torch.float32, device=dev)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    n_extra = _check_extra(extra_features, P)
    extra_img = torch.empty(n_extra, H, W, dtype=torch.float32, device=dev) if n_extra else None
    geom_bytes = lib.b200splat_geom_bytes(P)
    image_bytes = lib.b200splat_image_bytes(H, W)
    geom = torch.empty(geom_bytes, dtype=torch.uint8, device=dev)
    image = torch.empty(image_bytes, dtype=torch.uint8, device=dev)
    held = []

    def _alloc(_user, nbytes):
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        held.append(t)
        return t.data_ptr()

    cb = _lib.ALLOC_FN(_alloc)
    nr = C.c_int64(0)
    bout = C.c_void_p(0)
    a = _lib.ForwardArgs()
    a.cam = cam.c_struct()
    a.P, a.M = P, M
    a.means3D, a.shs, a.colors_precomp = _ptr(means3D), _ptr(shs), _ptr(colors_precomp)
    a.opacities, a.scales, a.rotations = _ptr(opacities), _ptr(scales), _ptr(rotations)
    a.cov3D_precomp = _ptr(cov3D_precomp)
    a.out_color, a.out_depth, a.out_alpha, a.radii = color.data_ptr(), depth.data_ptr(), alpha.data_ptr(), _ptr(radii)
    a.geom_buffer, a.geom_bytes = geom.data_ptr(), geom_bytes
    a.image_buffer, a.image_bytes = image.data_ptr(), image_bytes
    a.binning_buffer, a.binning_bytes = None, 0
    a.binning_alloc, a.alloc_user = cb, None
    a.stream = _stream(dev)
    a.num_rendered_out = C.pointer(nr)
    a.binning_out = C.pointer(bout)
    if n_extra:
        a.extra_features, a.n_extra, a.out_extra = extra_features.data_ptr(), n_extra, extra_img.data_ptr()
    with torch.cuda.device(dev):
        check(lib.b200splat_forward(C.byref(a)), "b200splat_forward")
    binning = held[0] if held else None
    st = ForwardState(P, M, int(nr.value), geom, binning, image)
    if n_extra:
        return color, radii, depth, alpha, extra_img, st
    return color, radii, depth, alpha, st


def backward(cam: Cam, st: ForwardState, means3D, shs, colors_precomp, opacities, scales, rotations,
             cov3D_precomp, radii, out_alpha, g_color, g_depth, g_alpha, out=None, accumulate=False, stats=None,
             extra_features=None, g_extra=None):
    """Returns dict of dense gradients (tensors allocated here unless ``out`` supplies them)."""
    dev = means3D.device
    P, M = st.P, st.M
    new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
    out = dict(out or {})
    out.setdefault("means3D", new(P, 3))
    out.setdefault("means2D", new(P, 3))
    out.setdefault("opacities", new(P, 1))
    if shs is not None:
        out.setdefault("shs", new(P, M, 3))
    if colors_precomp is not None:
        out.setdefault("colors_precomp", new(P, 3))
    if scales is not None:
        out.setdefault("scales", new(P, 3))
        out.setdefault("rotations", new(P, 4))
    if cov3D_precomp is not None:
        out.setdefault("cov3D_precomp", new(P, 6))
    n_extra = _check_extra(extra_features, P)
    if n_extra:
        out.setdefault("extra_features", new(P, n_extra))
    scratch_bytes = lib.b200splat_backward_scratch_bytes(P)
    scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
    a = _lib.BackwardArgs()
    a.cam = cam.c_struct()
    a.P, a.M, a.num_rendered = P, M, st.num_rendered
    a.means3D, a.shs, a.colors_precomp = _ptr(means3D), _ptr(shs), _ptr(colors_precomp)
    a.opacities, a.scales, a.rotations = _ptr(opacities), _ptr(scales), _ptr(rotations)
    a.cov3D_precomp, a.radii, a.out_alpha = _ptr(cov3D_precomp), _ptr(radii), _ptr(out_alpha)
    a.geom_buffer, a.binning_buffer, a.image_buffer = _ptr(st.geom), _ptr(st.binning), _ptr(st.image)
    a.dL_dout_color, a.dL_dout_depth, a.dL_dout_alpha = _ptr(g_color), _ptr(g_depth), _ptr(g_alpha)
    a.dL_dmeans3D, a.dL_dmeans2D = _ptr(out["means3D"]), _ptr(out["means2D"])
    a.dL_dshs, a.dL_dcolors = _ptr(out.get("shs")), _ptr(out.get("colors_precomp"))
    a.dL_dopacity, a.dL_dscales = _ptr(out["opacities"]), _ptr(out.get("scales"))
    a.dL_drotations, a.dL_dcov3D = _ptr(out.get("rotations")), _ptr(out.get("cov3D_precomp"))
    a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch_bytes
    a.accumulate = int(bool(accumulate))
    if stats is not None:   # (grad_accum, denom, max_radii) each (P,) fp32, updated in place
        a.stat_grad_accum, a.stat_denom, a.stat_max_radii = (_ptr(t) for t in stats)
    a.stream = _stream(dev)
    if n_extra:
        a.extra_features, a.n_extra = extra_features.data_ptr(), n_extra
        a.dL_dout_extra, a.dL_dextra = _ptr(g_extra), out["extra_features"].data_ptr()
    with torch.cuda.device(dev):
        check(lib.b200splat_backward(C.byref(a)), "b200splat_backward")
    return out


def mark_visible(positions: torch.Tensor, viewmatrix: torch.Tensor, projmatrix: torch.Tensor) -> torch.Tensor:
    positions = _f32c(positions, "positions")
    P = 0 if positions is None else positions.shape[0]
    dev = viewmatrix.device
    present = torch.zeros(P, dtype=torch.uint8, device=dev)
    if P:
        v, p = _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix")
        with torch.cuda.device(dev):
            check(lib.b200splat_mark_visible(P, positions.data_ptr(), v.data_ptr(), p.data_ptr(),
                                             present.data_ptr(), _stream(dev)), "b200splat_mark_visible")
    return present.bool()


def dist2(points: torch.Tensor) -> torch.Tensor:
    if not points.is_cuda:
        raise RuntimeError("distCUDA2: points must be a CUDA tensor (no CPU fallback)")
    pts = _f32c(points, "points")
    P = 0 if pts is None else pts.shape[0]
    out = torch.empty(P, dtype=torch.float32, device=points.device)
    if P == 0:
        return out
    nbytes = lib.b200splat_dist2_workspace_bytes(P)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
    with torch.cuda.device(points.device):
        check(lib.b200splat_dist2(P, pts.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, _stream(points.device)),
              "b200splat_dist2")
    return out


def sort_pairs(keys: torch.Tensor, vals: torch.Tensor, end_bit: int = 64):
    """Stable ascending sort of (int64 keys viewed as u64, int32 values viewed as u32) on bits [0,end_bit)."""
    assert keys.is_cuda and keys.dtype == torch.int64 and vals.dtype == torch.int32
    n = keys.numel()
    k0, v0 = keys.clone(), vals.clone()
    k1, v1 = torch.empty_like(k0), torch.empty_like(v0)
    if n == 0:
        return k0, v0
    nbytes = lib.b200splat_sort_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=keys.device)
    sel = C.c_int32(0)
    with torch.cuda.device(keys.device):
        check(lib.b200splat_sort_pairs(n, end_bit, k0.data_ptr(), v0.data_ptr(), k1.data_ptr(), v1.data_ptr(),
                                       ws.data_ptr(), nbytes, C.byref(sel), _stream(keys.device)), "b200splat_sort_pairs")
    return (k1, v1) if sel.value else (k0, v0)


def inclusive_scan_u32(x: torch.Tensor) -> torch.Tensor:
    assert x.is_cuda and x.dtype == torch.int32
    n = x.numel()
    out = torch.empty_like(x)
    if n == 0:
        return out
    nbytes = lib.b200splat_scan_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.b200splat_inclusive_scan_u32(n, x.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, _stream(x.device)),
              "b200splat_inclusive_scan_u32")
    return out


def status_tensor(H: int, W: int, st: ForwardState) -> torch.Tensor:
    """(4,) int32 device tensor aliasing the view's status words ([0] != 0: binning capacity overflow)."""
    v = _lib.ForwardViews()
    check(lib.b200splat_forward_views_get(st.P, H, W, 0, _ptr(st.geom), None, _ptr(st.image), C.byref(v)),
          "b200splat_forward_views_get")
    off = v.status - st.image.data_ptr()
    return st.image[off:off + 16].view(torch.int32)


def forward_views(cam: Cam, st: ForwardState):
    """Copies of the forward's intermediate buffers (for the bit-exact parity tests)."""
    v = _lib.ForwardViews()
    check(lib.b200splat_forward_views_get(st.P, cam.H, cam.W, st.num_rendered, _ptr(st.geom), _ptr(st.binning),
                                          _ptr(st.image), C.byref(v)), "b200splat_forward_views_get")
    dev = st.geom.device
    T = ((cam.W + 15) // 16) * ((cam.H + 15) // 16)

    def view(buf, ptr, count, dtype):
        if not ptr or count == 0:
            return torch.empty(0, dtype=dtype, device=dev)
        off = ptr - buf.data_ptr()
        nb = count * torch.empty(0, dtype=dtype).element_size()
        return buf[off:off + nb].view(dtype).clone()

    R, P = st.num_rendered, st.P
    if R and P:
        # a batched view's state carries the binning CAPACITY as num_rendered: only the first point_offsets[P-1]
        # pair words are live, the rest of the buffer is uninitialised
        R = min(R, int(view(st.geom, v.point_offsets, P, torch.int32)[-1].item()))
    words = view(st.binning, v.keys_sorted, R, torch.int64) if R else torch.empty(0, dtype=torch.int64, device=dev)
    depths = view(st.geom, v.depths, P, torch.float32)
    # sorted pair words are (tile << 32 | gaussian); upstream's key = (tile << 32) | float_bits(depth[gaussian])
    plist = (words & 0xFFFFFFFF).to(torch.int32)
    dbits = depths.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys = ((words >> 32) << 32) | (dbits[plist.long()] if R else words)
    return dict(
        tiles_touched=view(st.geom, v.tiles_touched, P, torch.int32),
        point_offsets=view(st.geom, v.point_offsets, P, torch.int32),
        depths=view(st.geom, v.depths, P, torch.float32),
        gauss2d=view(st.geom, v.gauss2d, P * 12, torch.float32).reshape(P, 12),
        keys_sorted=keys,
        point_list=plist,
        gaussian_order=view(st.geom, v.gaussian_order, P, torch.int64),
        ranges=view(st.image, v.ranges, T * 2, torch.int32).reshape(T, 2),
        n_contrib=view(st.image, v.n_contrib, cam.H * cam.W, torch.int32).reshape(cam.H, cam.W),
        n_visited=view(st.image, v.n_visited, cam.H * cam.W, torch.int32).reshape(cam.H, cam.W),
        status=view(st.image, v.status, 4, torch.int32),
    )
