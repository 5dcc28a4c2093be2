"""Oracle for ``simple_knn._C.distCUDA2`` (call sites: geometry/gaussian_base.py:434-437,
geometry/spacetime_gaussian.py:429-432).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED UPSTREAM (DSaurus/simple-knn is un-vendored).
Upstream's Morton-ordered box search is *exact* (boxes only prune), so brute force is a
valid restatement: mean of the squared Euclidean distances to the 3 nearest *other* points;
self is excluded by index, so duplicates count with distance 0.
With fewer than 4 points the missing neighbours contribute 0 (upstream initialises best[] to
FLT_MAX only inside the search; here we define the degenerate case as mean over what exists / 3).
"""
from __future__ import annotations

import torch

from . import spec


def dist2_oracle(points: torch.Tensor, chunk: int = 2048) -> torch.Tensor:
    P = points.shape[0]
    pts = points.to(torch.float32)
    out = torch.zeros(P, dtype=torch.float32)
    if P <= 1:
        return out
    k = min(spec.KNN_K, P - 1)
    for s0 in range(0, P, chunk):
        q = pts[s0:s0 + chunk]
        dx = q[:, None, 0] - pts[None, :, 0]
        dy = q[:, None, 1] - pts[None, :, 1]
        dz = q[:, None, 2] - pts[None, :, 2]
        d2 = dx * dx + dy * dy + dz * dz
        rows = torch.arange(q.shape[0])
        d2[rows, rows + s0] = float("inf")
        best = torch.topk(d2, k, dim=1, largest=False).values
        out[s0:s0 + chunk] = best.sum(dim=1) / float(spec.KNN_K)
    return out
