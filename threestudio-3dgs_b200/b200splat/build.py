"""In-tree build of libb200splat.so for sm_100a (nvcc cross-compiles without a GPU).

    python threestudio-3dgs_b200/b200splat/build.py [--force] [--verbose]

preprocess.cu is compiled with -fmad=false (bit-exact radii / tiles / depth keys, see the file header);
everything else with nvcc defaults (IEEE div/sqrt, FMA contraction on, no fast-math).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE.parent / "csrc"
INCLUDE = HERE.parent.parent / "include"
OBJ = HERE / "_build"
LIB = HERE / "libb200splat.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
SOURCES = {
    "preprocess.cu": ["-fmad=false"],
    "scan_sort.cu": [],
    "render.cu": [],
    "extra.cu": [],
    "postops.cu": [],
    "adam.cu": [],
    "preprocess_bwd.cu": [],
    "knn.cu": [],
    "p2p.cu": [],
    "api.cu": [],
}


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(out: Path, deps) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = [CSRC / "common.cuh", INCLUDE / "b200splat.h", Path(__file__)]
    jobs = []
    for src, extra in SOURCES.items():
        o = OBJ / (src[:-3] + ".o")
        if force or _stale(o, [CSRC / src, *headers]):
            jobs.append(([_nvcc(), *ARCH, *COMMON, *extra, "-c", str(CSRC / src), "-o", str(o)], src))

    def run(job):
        cmd, src = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(f"--- {src}\n{r.stdout}{r.stderr}\n")
        (OBJ / (src[:-3] + ".ptxas.log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")

    with ThreadPoolExecutor(max_workers=6) as ex:
        list(ex.map(run, jobs))
    objs = [str(OBJ / (s[:-3] + ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [_nvcc(), *ARCH, "-shared", "-o", str(LIB), *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
