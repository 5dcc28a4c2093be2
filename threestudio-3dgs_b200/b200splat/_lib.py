"""ctypes binding of libb200splat.so (C ABI in include/b200splat.h).

The product path has NO fallback: if the shared library is missing or does not load, importing this
module raises -- build it with ``python threestudio-3dgs_b200/b200splat/build.py`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("B200SPLAT_LIB", _HERE / "libb200splat.so"))

ABI_VERSION = 7
MAX_VIEWS = 8

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class Camera(C.Structure):
    _fields_ = [
        ("image_height", C.c_int32), ("image_width", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("sh_degree", C.c_int32), ("prefiltered", C.c_int32), ("debug", C.c_int32),
        ("bg", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
        ("scalars_dev", C.c_void_p),
    ]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("cam", Camera), ("P", C.c_int32), ("M", C.c_int32),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
        ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
        ("cov3D_precomp", C.c_void_p),
        ("out_color", C.c_void_p), ("out_depth", C.c_void_p), ("out_alpha", C.c_void_p), ("radii", C.c_void_p),
        ("geom_buffer", C.c_void_p), ("geom_bytes", C.c_size_t),
        ("image_buffer", C.c_void_p), ("image_bytes", C.c_size_t),
        ("binning_buffer", C.c_void_p), ("binning_bytes", C.c_size_t),
        ("binning_alloc", ALLOC_FN), ("alloc_user", C.c_void_p),
        ("stream", C.c_void_p),
        ("num_rendered_out", C.POINTER(C.c_int64)), ("binning_out", C.POINTER(C.c_void_p)),
        ("extra_features", C.c_void_p), ("n_extra", C.c_int32), ("out_extra", C.c_void_p),
    ]


class BackwardArgs(C.Structure):
    _fields_ = [
        ("cam", Camera), ("P", C.c_int32), ("M", C.c_int32), ("num_rendered", C.c_int64),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
        ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
        ("cov3D_precomp", C.c_void_p), ("radii", C.c_void_p), ("out_alpha", C.c_void_p),
        ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p), ("image_buffer", C.c_void_p),
        ("dL_dout_color", C.c_void_p), ("dL_dout_depth", C.c_void_p), ("dL_dout_alpha", C.c_void_p),
        ("dL_dmeans3D", C.c_void_p), ("dL_dmeans2D", C.c_void_p), ("dL_dshs", C.c_void_p),
        ("dL_dcolors", C.c_void_p), ("dL_dopacity", C.c_void_p), ("dL_dscales", C.c_void_p),
        ("dL_drotations", C.c_void_p), ("dL_dcov3D", C.c_void_p),
        ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t),
        ("accumulate", C.c_int32),
        ("stat_grad_accum", C.c_void_p), ("stat_denom", C.c_void_p), ("stat_max_radii", C.c_void_p),
        ("stream", C.c_void_p),
        ("extra_features", C.c_void_p), ("n_extra", C.c_int32), ("dL_dout_extra", C.c_void_p),
        ("dL_dextra", C.c_void_p),
    ]


class ForwardViews(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "tiles_touched", "point_offsets", "depths", "gauss2d", "cov3D", "keys_sorted", "point_list",
        "ranges", "n_contrib", "n_visited", "status")] + [("packed_idx_bits", C.c_int64),
                                                            ("gaussian_order", C.c_void_p)]


PP = C.POINTER(C.c_void_p)   # host array of device pointers


class BatchForwardArgs(C.Structure):
    _fields_ = [
        ("V", C.c_int32), ("cams", C.POINTER(Camera)), ("P", C.c_int32), ("M", C.c_int32),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
        ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
        ("out_color", PP), ("out_depth", PP), ("out_alpha", PP), ("radii", PP),
        ("geom_buffer", PP), ("image_buffer", PP), ("binning_buffer", PP), ("binning_bytes", C.c_size_t),
        ("stream", C.c_void_p), ("sync", C.c_int32),
        ("num_rendered_out", C.POINTER(C.c_int64)), ("overflow_out", C.POINTER(C.c_int32)),
        ("extra_features", C.c_void_p), ("n_extra", C.c_int32), ("out_extra", PP),
        ("pairs_notify", C.c_void_p), ("notify_epoch", C.c_uint32),
    ]


class BatchBackwardArgs(C.Structure):
    _fields_ = [
        ("V", C.c_int32), ("cams", C.POINTER(Camera)), ("P", C.c_int32), ("M", C.c_int32),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
        ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
        ("radii", PP), ("geom_buffer", PP), ("image_buffer", PP), ("binning_buffer", PP),
        ("binning_bytes", C.c_size_t),
        ("dL_dout_color", PP), ("dL_dout_depth", PP), ("dL_dout_alpha", PP), ("dL_dmeans2D", PP),
        ("dL_dmeans3D", C.c_void_p), ("dL_dshs", C.c_void_p), ("dL_dcolors", C.c_void_p),
        ("dL_dopacity", C.c_void_p), ("dL_dscales", C.c_void_p), ("dL_drotations", C.c_void_p),
        ("scratch", PP), ("accumulate", C.c_int32), ("scratch_clean", C.c_int32),
        ("phase", C.c_int32), ("g_begin", C.c_int32), ("g_end", C.c_int32),
        ("stat_grad_accum", C.c_void_p), ("stat_denom", C.c_void_p), ("stat_max_radii", C.c_void_p),
        ("stream", C.c_void_p),
        ("extra_features", C.c_void_p), ("n_extra", C.c_int32), ("dL_dout_extra", PP), ("dL_dextra", C.c_void_p),
        ("live_map", C.c_void_p),
    ]


class PostprocessArgs(C.Structure):
    _fields_ = [
        ("V", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("mode", C.c_int32), ("shading", C.c_int32),
        ("image", C.c_void_p), ("depth", C.c_void_p), ("alpha", C.c_void_p), ("rays_o", C.c_void_p),
        ("rays_d", C.c_void_p), ("bg", C.c_void_p), ("light", C.c_void_p), ("pred_normal", C.c_void_p),
        ("ambient", C.c_float * 3), ("diffuse", C.c_float * 3),
        ("render", C.c_void_p), ("normal", C.c_void_p), ("depth_out", C.c_void_p),
        ("g_render", C.c_void_p), ("g_normal", C.c_void_p), ("g_depth", C.c_void_p),
        ("d_image", C.c_void_p), ("d_depth", C.c_void_p), ("d_alpha", C.c_void_p), ("d_bg", C.c_void_p),
        ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t), ("stream", C.c_void_p),
        ("shading_per_view", C.POINTER(C.c_int32)), ("lights_per_view", C.POINTER(C.c_float)),
    ]


_ADAM_TENSORS = ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")


class AdamArgs(C.Structure):
    _fields_ = ([("P", C.c_int32), ("M", C.c_int32)] + [(n, C.c_void_p) for n in _ADAM_TENSORS] +
                [("m_" + n, C.c_void_p) for n in _ADAM_TENSORS] + [("v_" + n, C.c_void_p) for n in _ADAM_TENSORS] +
                [(n, C.c_void_p) for n in ("g_means3D", "g_shs", "g_opacities", "g_scales", "g_rotations")] +
                [("lr", C.c_double * 6), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_float),
                 ("color_clip", C.c_float), ("step", C.c_int32), ("stream", C.c_void_p)])


class P2PArgs(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("bufs", C.c_void_p * 8), ("signals", C.c_void_p * 8),
                ("n_segments", C.c_int32), ("seg_offset", C.c_int64 * 16), ("seg_count", C.c_int64 * 16),
                ("seg_op", C.c_int32 * 16), ("epoch", C.c_uint32), ("stream", C.c_void_p),
                ("seg_row_floats", C.c_int32 * 16), ("seg_row0", C.c_int64 * 16), ("live_offset", C.c_int64)]


class MCArgs(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("mc_buffer", C.c_void_p), ("signals", C.c_void_p * 8),
                ("n_segments", C.c_int32), ("seg_offset", C.c_int64 * 16), ("seg_count", C.c_int64 * 16),
                ("seg_op", C.c_int32 * 16), ("epoch", C.c_uint32), ("stream", C.c_void_p),
                ("seg_row_floats", C.c_int32 * 16), ("seg_row0", C.c_int64 * 16), ("live_offset", C.c_int64)]


P2P_HANDLE_BYTES, P2P_SIGNAL_BYTES = 64, 256

# every symbol include/b200splat.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b200splat_abi_version": (C.c_int, []),
    "b200splat_set_staging": (C.c_int, [C.c_int32]),
    "b200splat_last_error": (C.c_char_p, []),
    "b200splat_launch_count": (C.c_uint64, []),
    "b200splat_p2p_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "b200splat_p2p_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200splat_p2p_close": (C.c_int, [C.c_void_p]),
    "b200splat_p2p_free": (C.c_int, [C.c_void_p]),
    "b200splat_p2p_allreduce": (C.c_int, [C.POINTER(P2PArgs)]),
    "b200splat_p2p_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "b200splat_mc_allreduce": (C.c_int, [C.POINTER(MCArgs)]),
    "b200splat_geom_bytes": (C.c_size_t, [C.c_int32]),
    "b200splat_image_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "b200splat_binning_bytes": (C.c_size_t, [C.c_int64]),
    "b200splat_backward_scratch_bytes": (C.c_size_t, [C.c_int32]),
    "b200splat_forward": (C.c_int, [C.POINTER(ForwardArgs)]),
    "b200splat_backward": (C.c_int, [C.POINTER(BackwardArgs)]),
    "b200splat_binning_capacity": (C.c_int64, [C.c_size_t]),
    "b200splat_forward_batched": (C.c_int, [C.POINTER(BatchForwardArgs)]),
    "b200splat_backward_batched": (C.c_int, [C.POINTER(BatchBackwardArgs)]),
    "b200splat_postprocess_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "b200splat_postprocess_forward": (C.c_int, [C.POINTER(PostprocessArgs)]),
    "b200splat_postprocess_backward": (C.c_int, [C.POINTER(PostprocessArgs)]),
    "b200splat_adam_step": (C.c_int, [C.POINTER(AdamArgs)]),
    "b200splat_mark_visible": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200splat_dist2_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "b200splat_dist2": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200splat_sort_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "b200splat_sort_pairs": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.c_void_p]),
    "b200splat_scan_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "b200splat_inclusive_scan_u32": (C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                               C.c_void_p]),
    "b200splat_forward_views_get": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.POINTER(ForwardViews)]),
    "b200splat_profile_enable": (C.c_int, [C.c_int32]),
    "b200splat_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
}

FAMILIES = ("preprocess", "scan", "duplicate", "sort", "ranges", "render_fwd", "render_bwd", "preprocess_bwd",
            "dist2")


def _load():
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            f"`python {_HERE / 'build.py'}` (needs nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    got = lib.b200splat_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"libb200splat ABI {got} != binding ABI {ABI_VERSION}; rebuild")
    return lib


lib = _load()


class B200SplatError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.b200splat_last_error().decode("utf-8", "replace")
        raise B200SplatError(f"{what} failed (code {rc}): {msg}")


def profile_enable(on: bool) -> None:
    check(lib.b200splat_profile_enable(int(on)), "b200splat_profile_enable")


def profile_read():
    """{family: (total_ms, launch_groups)} since profile_enable(True); clears the record."""
    n = len(FAMILIES)
    ms = (C.c_float * n)()
    cnt = (C.c_int64 * n)()
    check(lib.b200splat_profile_read(ms, cnt), "b200splat_profile_read")
    return {FAMILIES[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}


def launch_count() -> int:
    return int(lib.b200splat_launch_count())


STAGING = {None: -1, "default": -1, "ldgsts": 0, "bulk": 1}


def set_staging(mode) -> int:
    """Select how the render kernels stage Gaussian records (None = default); returns the previous setting."""
    return int(lib.b200splat_set_staging(STAGING[mode]))
