"""The C-ABI shared library loads without a GPU and exports every symbol include/b200splat.h declares;
host-only entry points behave; argument errors are reported through the error-code convention."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "b200splat.h"


def _declared():
    txt = HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200splat_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from b200splat import _lib
    names = _declared()
    assert len(names) >= 18
    raw = C.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/b200splat.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature in b200splat/_lib.py"
    assert _lib.lib.b200splat_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header_field_order():
    from b200splat import _lib
    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    for cname, cls in (("b200splat_camera", _lib.Camera), ("b200splat_forward_args", _lib.ForwardArgs),
                       ("b200splat_backward_args", _lib.BackwardArgs),
                       ("b200splat_batch_forward_args", _lib.BatchForwardArgs),
                       ("b200splat_batch_backward_args", _lib.BatchBackwardArgs),
                       ("b200splat_forward_views", _lib.ForwardViews),
                       ("b200splat_postprocess_args", _lib.PostprocessArgs),
                       ("b200splat_adam_args", _lib.AdamArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), txt, flags=re.S).group(1)
        fields = [re.sub(r"\[.*\]", "", re.split(r"[\s\*]+", d.strip())[-1]) for d in body.split(";") if d.strip()]
        assert fields == [f[0] for f in cls._fields_], cname


def test_host_only_entry_points():
    from b200splat import _lib
    lib = _lib.lib
    assert lib.b200splat_geom_bytes(1000) > 1000 * (48 + 4 + 1 + 4 + 4)
    assert lib.b200splat_geom_bytes(2000) > lib.b200splat_geom_bytes(1000)
    assert lib.b200splat_image_bytes(512, 512) >= 1024 * 8 + 512 * 512 * 12
    assert lib.b200splat_binning_bytes(1_000_000) >= 1_000_000 * 24
    cap = lib.b200splat_binning_capacity(lib.b200splat_binning_bytes(1_000_000))
    assert cap >= 1_000_000 and lib.b200splat_binning_bytes(cap) <= lib.b200splat_binning_bytes(1_000_000)
    assert lib.b200splat_backward_scratch_bytes(1000) >= 48_000
    assert lib.b200splat_sort_workspace_bytes(1 << 20) > 0 and lib.b200splat_scan_workspace_bytes(1 << 20) > 0
    assert lib.b200splat_dist2_workspace_bytes(4096) > 0
    assert _lib.launch_count() == 0


def test_argument_errors_use_error_codes_not_exceptions():
    from b200splat import _lib
    lib = _lib.lib
    assert lib.b200splat_forward(None) == -1
    assert b"null" in lib.b200splat_last_error()
    a = _lib.ForwardArgs()
    a.P = 10                                   # neither shs nor colors_precomp
    assert lib.b200splat_forward(C.byref(a)) == -1
    assert b"exactly one" in lib.b200splat_last_error()
    with pytest.raises(_lib.B200SplatError):
        _lib.check(-1, "x")


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: the drop-in operator raises on CPU tensors instead of routing anywhere else."""
    import torch
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from simple_knn._C import distCUDA2
    rs = GaussianRasterizationSettings(16, 16, 0.5, 0.5, torch.ones(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                       torch.zeros(3), False, False)
    z = torch.zeros
    with pytest.raises(RuntimeError, match="CUDA"):
        GaussianRasterizer(raster_settings=rs)(means3D=z(4, 3), means2D=z(4, 3), shs=z(4, 1, 3),
                                               colors_precomp=None, opacities=z(4, 1), scales=z(4, 3),
                                               rotations=z(4, 4), cov3D_precomp=None)
    with pytest.raises(RuntimeError, match="CUDA"):
        distCUDA2(z(8, 3))
    import inspect
    import b200splat.ops as ops
    import diff_gaussian_rasterization as dgr
    for mod in (ops, dgr):
        assert "oracle" not in inspect.getsource(mod), "product code must not reference the oracle"
