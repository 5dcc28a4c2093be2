"""One training step's view batch on one GPU: the unit the reference runs as a Python loop of
single-view rasterizer calls (renderer/gaussian_batch_renderer.py:21-54) followed by per-view
densification statistics (geometry/gaussian_base.py:815-819, 846-851).

Here the per-view backward writes straight into one packed gradient buffer (first view overwrites,
later views accumulate) and the statistics are fused into the preprocess-backward kernel, so a
multi-GPU step is: local views -> ONE sum all-reduce of the packed buffer + ONE max all-reduce of
max_radii (b200splat/dist.py).  Note the order the reference's statistics impose
(SURVEY.md 8e): ||means2D.grad|| is taken per view *before* any summation, so it is reduced locally
per view and only the accumulators are all-reduced, never means2D.grad itself.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops


class PackedGrads:
    """[dL/dmeans3D 3 | dL/dscales 3 | dL/drotations 4 | dL/dopacity 1 | dL/dshs 3M (or colours 3) |
    grad_accum 1 | denom 1] x P floats in ONE contiguous fp32 buffer (the all-reduce payload,
    4*P*(13+3M) bytes), plus max_radii (P) reduced with MAX."""

    def __init__(self, P: int, M: int, device, color_mode: str = "shs"):
        self.P, self.M, self.color_mode = P, M, color_mode
        ncol = 3 * M if color_mode == "shs" else 3
        widths = [("means3D", 3), ("scales", 3), ("rotations", 4), ("opacities", 1),
                  (color_mode if color_mode == "shs" else "colors_precomp", ncol),
                  ("grad_accum", 1), ("denom", 1)]
        total = sum(w for _, w in widths) * P
        self.buffer = torch.zeros(total, dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = {}
        off = 0
        for name, w in widths:
            self.views[name] = self.buffer[off:off + w * P].view(P, w) if w > 1 else self.buffer[off:off + P]
            off += w * P
        self.views["opacities"] = self.views["opacities"].view(P, 1)
        if color_mode == "shs":
            self.views["shs"] = self.views["shs"].view(P, M, 3)
        self.max_radii = torch.zeros(P, dtype=torch.float32, device=device)
        self.means2D_scratch = torch.empty(P, 3, dtype=torch.float32, device=device)

    @property
    def nbytes(self) -> int:
        return self.buffer.numel() * 4

    def zero_stats_(self):
        self.views["grad_accum"].zero_()
        self.views["denom"].zero_()
        self.max_radii.zero_()

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in self.views.items() if k not in ("grad_accum", "denom")}


def render_views_fwd_bwd(cams: Sequence[ops.Cam], means3D, shs, colors_precomp, opacities, scales, rotations,
                         pixel_grads, packed: PackedGrads, keep_images: bool = False):
    """Forward + backward of every view in ``cams``; parameter gradients summed over the views and the
    densification statistics of the views land in ``packed``.  ``pixel_grads[v]`` = (dL/dcolor (3,H,W),
    dL/ddepth (1,H,W) or None, dL/dalpha (1,H,W) or None) or a callable(color, depth, alpha) -> that tuple
    (the loss).  Returns the list of (color, depth, alpha, radii) when keep_images."""
    packed.zero_stats_()
    out = dict(packed.grads())
    out["means2D"] = packed.means2D_scratch
    stats = (packed.views["grad_accum"], packed.views["denom"], packed.max_radii)
    images = []
    for v, cam in enumerate(cams):
        color, radii, depth, alpha, st = ops.forward(cam, means3D, shs, colors_precomp, opacities, scales,
                                                     rotations, None)
        pg = pixel_grads[v]
        if callable(pg):
            pg = pg(color, depth, alpha)
        ops.backward(cam, st, means3D, shs, colors_precomp, opacities, scales, rotations, None, radii, alpha,
                     pg[0], pg[1], pg[2], out=out, accumulate=v > 0, stats=stats)
        if keep_images:
            images.append((color, depth, alpha, radii))
    return images
