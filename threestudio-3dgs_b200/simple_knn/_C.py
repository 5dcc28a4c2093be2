"""Drop-in ``simple_knn._C``: ``distCUDA2(points (P,3) float32 cuda) -> (P,) float32``.

Mean squared distance to the 3 nearest other points, exact (call sites
geometry/gaussian_base.py:434-437, geometry/spacetime_gaussian.py:429-432).  Backed by
libb200splat.so; CUDA only, no CPU fallback.
"""
from b200splat.ops import dist2 as _dist2


def distCUDA2(points):
    return _dist2(points)
