"""Multi-GPU sharding of a batch of scans.

Scans are independent, so the batch is cut into ``world_size`` contiguous blocks (balanced by point
count) and every rank builds its own block with no data-path collective.  ``gather_grids`` is the
optional final gather for a single consumer; it is the only place a collective appears.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_bounds(point_counts: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) scan ranges per rank, balancing the number of points.
    Deterministic, covers every scan exactly once, keeps order; ranks may get empty ranges when
    there are fewer scans than ranks."""
    counts = np.asarray(point_counts, dtype=np.int64)
    n = counts.size
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    csum = np.concatenate([[0], np.cumsum(counts + 1)])       # +1 so empty scans still cost something
    total = csum[-1]
    bounds, start = [], 0
    for r in range(world_size):
        target = total * (r + 1) / world_size
        end = int(np.searchsorted(csum, target, side="left"))
        end = min(max(end, start), n)
        if r == world_size - 1:
            end = n
        bounds.append((start, end))
        start = end
    return bounds


def my_shard(point_counts: Sequence[int], rank: int, world_size: int) -> Tuple[int, int]:
    return shard_bounds(point_counts, world_size)[rank]


def gather_grids(local_grids, n_total: int, bounds: Sequence[Tuple[int, int]], group=None):
    """Optional final gather: every rank ends up with the (n_total, R, S) tensor.  Works on the
    backend of the default process group (NCCL on GPUs, gloo in CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    assert len(bounds) == world
    tail = tuple(local_grids.shape[1:])
    out = torch.empty((n_total,) + tail, dtype=local_grids.dtype, device=local_grids.device)
    pieces = [out[a:b] for a, b in bounds]
    if all(b - a == bounds[0][1] - bounds[0][0] for a, b in bounds):
        dist.all_gather(pieces, local_grids.contiguous(), group=group)
    else:
        # ragged shards: one broadcast per rank (pieces are views of `out`)
        for r, piece in enumerate(pieces):
            if piece.shape[0] == 0:
                continue
            if dist.get_rank(group) == r:
                piece.copy_(local_grids)
            dist.broadcast(piece, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return out
