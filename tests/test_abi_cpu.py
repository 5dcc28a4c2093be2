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


def test_new_entry_points_validate_arguments_before_launching():
    from b200splat import _lib
    lib = _lib.lib
    assert lib.b200splat_postprocess_forward(None) == -1
    p = _lib.PostprocessArgs()
    assert lib.b200splat_postprocess_forward(C.byref(p)) == -1 and b"positive" in lib.b200splat_last_error()
    p.V, p.H, p.W, p.mode = 1, 4, 4, 9
    assert lib.b200splat_postprocess_forward(C.byref(p)) == -1 and b"mode" in lib.b200splat_last_error()
    p.mode, p.image, p.depth, p.alpha = 3, 16, 16, 16            # shading mode without rays / bg / light
    assert lib.b200splat_postprocess_backward(C.byref(p)) == -1 and b"rays" in lib.b200splat_last_error()
    assert lib.b200splat_postprocess_scratch_bytes(4, 512, 512) == 4 * 512 * 512 * 36
    a = _lib.AdamArgs()
    a.P, a.M, a.step = 8, 1, 0
    assert lib.b200splat_adam_step(C.byref(a)) == -1 and b"1-based" in lib.b200splat_last_error()
    a.step = 1
    assert lib.b200splat_adam_step(C.byref(a)) == -1 and b"null" in lib.b200splat_last_error()
    a.P = 0
    assert lib.b200splat_adam_step(C.byref(a)) == 0               # nothing to do, nothing launched
    f = _lib.ForwardArgs()
    f.P, f.M, f.shs, f.n_extra = 4, 1, 16, 7
    f.scales = f.rotations = f.out_color = f.out_depth = f.out_alpha = f.image_buffer = 16
    f.image_bytes = 1 << 30
    f.cam.image_height = f.cam.image_width = 16
    f.cam.bg = f.cam.viewmatrix = f.cam.projmatrix = f.cam.campos = 16
    assert lib.b200splat_forward(C.byref(f)) == -1 and b"n_extra" in lib.b200splat_last_error()
    assert _lib.launch_count() == 0


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/b200splat.h compiles as C99 (no C++ in the boundary) and a C program links the library and calls its
    host-only entry points -- the binding a non-Python host would write."""
    import shutil
    import subprocess
    from b200splat import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "b200splat.h"
int main(void) {
    b200splat_forward_args f;
    b200splat_adam_args a;
    b200splat_postprocess_args p;
    memset(&f, 0, sizeof f); memset(&a, 0, sizeof a); memset(&p, 0, sizeof p);
    if (b200splat_abi_version() != B200SPLAT_ABI_VERSION) return 1;
    if (b200splat_geom_bytes(1000) == 0 || b200splat_backward_scratch_bytes(1000) < 64000) return 2;
    if (b200splat_forward(&f) != B200SPLAT_ERR_INVALID || strlen(b200splat_last_error()) == 0) return 3;
    if (b200splat_postprocess_forward(&p) != B200SPLAT_ERR_INVALID) return 4;
    a.step = 1; a.M = 1;
    if (b200splat_adam_step(&a) != B200SPLAT_OK) return 5;          /* P == 0: nothing to do */
    if (b200splat_launch_count() != 0) return 6;
    printf("abi %d ok\n", b200splat_abi_version());
    return 0;
}
''')
    exe = tmp_path / "abi"
    lib_dir = _lib.LIB_PATH.parent
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
           f"-L{lib_dir}", "-l:libb200splat.so", f"-Wl,-rpath,{lib_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
