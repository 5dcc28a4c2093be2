"""Copy the few reference files tests/ref_harness.py imports into tests/_refcopy/ so that the GPU box (which has no
/root/reference) can run them UNCHANGED on the CUDA backend (tests/test_reference_cuda_gpu.py).  tests/_refcopy/ is
git-ignored: reference sources are never committed; the directory travels with the gpurun snapshot like the built .so.

    python scripts/stage_reference_for_gpu.py
"""
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
SRC = Path("/root/reference")
DST = ROOT / "tests" / "_refcopy"

if __name__ == "__main__":
    import ref_harness
    if not (SRC / "renderer").exists():
        raise SystemExit("/root/reference not present here")
    for rel in ref_harness.NEEDED:
        (DST / rel).parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(SRC / rel, DST / rel)
        print("staged", rel)
