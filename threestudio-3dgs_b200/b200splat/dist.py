"""View-data-parallel plumbing: views of a step are split across the GPUs of one box, Gaussians are
replicated, and the packed gradient buffer + densification accumulators are all-reduced (NCCL over
NVLink 5 / NVSwitch on GPUs, gloo in the CPU tests).  The reference has no distributed code at all
(SURVEY.md 2.1); this is new capability behind the same per-view operator.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the step's views owned by ``rank`` (ragged allowed)."""
    base, rem = divmod(num_views, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_packed(buffer: torch.Tensor, max_radii: torch.Tensor, group=None):
    """SUM over ranks of the packed gradient/statistics buffer, MAX of max_radii.  No-op at world 1."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(buffer, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(max_radii, op=dist.ReduceOp.MAX, group=group)


def reference_update_states(xyz_gradient_accum, denom, max_radii2D, grad_accum_step, denom_step, max_radii_step):
    """Apply one step's (all-reduced) statistics to the persistent accumulators exactly as the
    reference's per-view loop does (geometry/gaussian_base.py:815-819, 846-851)."""
    xyz_gradient_accum += grad_accum_step.reshape(xyz_gradient_accum.shape)
    denom += denom_step.reshape(denom.shape)
    torch.maximum(max_radii2D, max_radii_step.reshape(max_radii2D.shape), out=max_radii2D)
