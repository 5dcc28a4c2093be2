"""Import the reference's renderer / geometry files UNCHANGED (from /root/reference) with stub
dependencies, and -- when no GPU is present -- run their hard-coded ``device="cuda"`` code on CPU.

backend="oracle": ``diff_gaussian_rasterization`` / ``simple_knn._C`` resolve to oracle/api.py
                  (BASELINE.json configs[0]: the CPU-runnable plumbing case).
backend="cuda":   they resolve to the product packages in threestudio-3dgs_b200/.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

import torch
from torch.overrides import TorchFunctionMode

STUBS = Path(__file__).resolve().parent / "stubs"
# The reference tree.  It does not exist on the GPU box; scripts/stage_reference_for_gpu.py copies the handful of
# files the harness imports into tests/_refcopy/ (git-ignored, never committed; it travels with the gpurun snapshot
# like the built .so), so that the GPU tests can run the reference's UNCHANGED files on the CUDA backend.
_CANDIDATES = (Path("/root/reference"), Path(__file__).resolve().parent / "_refcopy")
NEEDED = ("renderer/diff_gaussian_rasterizer.py", "renderer/gaussian_batch_renderer.py", "geometry/gaussian_base.py",
          "geometry/gaussian_io.py", "geometry/mesh_utils.py")
REFERENCE = next((p for p in _CANDIDATES if (p / "renderer" / "diff_gaussian_rasterizer.py").exists()), _CANDIDATES[0])


def available() -> bool:
    return all((REFERENCE / f).exists() for f in NEEDED)


class CudaToCpu(TorchFunctionMode):
    """Rewrites device='cuda' -> 'cpu' in factory calls and makes .cuda() a no-op (no-GPU hosts)."""

    def __torch_function__(self, func, types_, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        d = kwargs.get("device")
        if d is not None and "cuda" in str(d):
            kwargs["device"] = "cpu"
        name = getattr(func, "__name__", "")
        if name == "cuda" and args and torch.is_tensor(args[0]):
            return args[0]
        if name == "to" and len(args) >= 2 and "cuda" in str(args[1]):
            args = (args[0], "cpu") + tuple(args[2:])
        return func(*args, **kwargs)


def load(backend: str = "oracle"):
    """Returns (renderer_module, geometry_module) of the reference, freshly imported."""
    for k in [k for k in sys.modules if k.split(".")[0] in ("ref3dgs", "threestudio", "plyfile", "mcubes",
                                                            "diff_gaussian_rasterization", "simple_knn")]:
        del sys.modules[k]
    if str(STUBS) not in sys.path:
        sys.path.insert(0, str(STUBS))
    if backend == "oracle":
        from oracle import api
        dgr = types.ModuleType("diff_gaussian_rasterization")
        dgr.GaussianRasterizationSettings = api.GaussianRasterizationSettings
        dgr.GaussianRasterizer = api.GaussianRasterizer
        sys.modules["diff_gaussian_rasterization"] = dgr
        sk = types.ModuleType("simple_knn")
        skc = types.ModuleType("simple_knn._C")
        skc.distCUDA2 = api.distCUDA2
        sk._C = skc
        sys.modules["simple_knn"] = sk
        sys.modules["simple_knn._C"] = skc
    else:
        import diff_gaussian_rasterization  # noqa: F401  (product package, CUDA only)
        import simple_knn._C  # noqa: F401
    pkg = types.ModuleType("ref3dgs")
    pkg.__path__ = [str(REFERENCE)]
    sys.modules["ref3dgs"] = pkg
    for sub in ("renderer", "geometry", "material", "background", "utils"):
        m = types.ModuleType(f"ref3dgs.{sub}")
        m.__path__ = [str(REFERENCE / sub)]
        sys.modules[f"ref3dgs.{sub}"] = m
    if not torch.cuda.is_available():
        torch.cuda.set_device = lambda *a, **k: None
        torch.cuda.empty_cache = lambda *a, **k: None
    geometry = importlib.import_module("ref3dgs.geometry.gaussian_base")
    renderer = importlib.import_module("ref3dgs.renderer.diff_gaussian_rasterizer")
    return renderer, geometry
