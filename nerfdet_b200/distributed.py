"""View-sharded lifting across the GPUs of one box.

Rank r owns a contiguous slice of the views (the backbone would have produced them
there).  Every rank accumulates per-voxel ``S1 = sum f``, ``S2 = sum f^2`` and the valid-view
count over ITS views into one flat fp32 buffer, the buffers are summed with ONE
all-reduce (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests), and every
rank finalises mean / all-view variance with the GLOBAL number of views -- the
reference's variance runs over all views, invalid ones included (nerfdet.py:179), so
``n_views_total`` is part of the formula, not just a normaliser.

The reference has no collective on this path (it lifts one scene per GPU); this is the
multi-GPU extension BASELINE.json's north_star asks for.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def view_shard(n_views: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the views owned by ``rank``: contiguous, sizes differ by at most 1,
    earlier ranks take the remainder."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, rem = divmod(n_views, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _cuda_accumulate(features, points, projection):
    from . import ops
    return ops.lift_accumulate(features, points, projection, 0)


def _cuda_finalize(acc, n_views_total, channels, n_voxels, alpha, want_cov):
    from . import ops
    return ops.lift_finalize(acc, n_views_total, channels, n_voxels, alpha, want_cov)


def lift_mean_var_view_sharded(features_local: torch.Tensor, points: torch.Tensor,
                               projection_local: torch.Tensor, n_views_total: Optional[int] = None,
                               group=None, alpha: Optional[torch.Tensor] = None, want_cov: bool = True,
                               accumulate_fn: Callable = _cuda_accumulate,
                               finalize_fn: Callable = _cuda_finalize):
    """Each rank passes ITS views (``features_local [nv_r, C, H, W]``, ``projection_local
    [nv_r, 3, 4]``); returns the same ``(volume_mean [C,X,Y,Z], volume_cov, valid [1,X,Y,Z])``
    on every rank as the single-GPU ``lifting.lift_mean_var`` over all views would.

    ``accumulate_fn`` / ``finalize_fn`` default to the CUDA ops; the CPU (gloo) tests inject
    oracle-backed stand-ins to exercise the sharding / collective logic without a GPU."""
    c = features_local.shape[1]
    gx, gy, gz = points.shape[-3:]
    n = gx * gy * gz
    acc = accumulate_fn(features_local, points, projection_local)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if n_views_total is None:
        nv = torch.tensor([features_local.shape[0]], dtype=torch.int64, device=acc.device)
        if world > 1:
            dist.all_reduce(nv, group=group)
        n_views_total = int(nv.item())
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    mean, cov, count = finalize_fn(acc, n_views_total, c, n,
                                   alpha.reshape(-1) if alpha is not None else None, want_cov)
    return (mean.view(c, gx, gy, gz), cov.view(c, gx, gy, gz) if want_cov else None,
            count.view(1, gx, gy, gz))
