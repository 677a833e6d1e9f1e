"""Angle sharding for multi-GPU runs: one process per GPU, the angles of
every OS-SART subset interleaved over the ranks, volume replicated.  Host logic only."""
from __future__ import annotations

import numpy as np


def shard_blocks(blocks, rank, world):
    """Rank-local angle list and per-subset (first, count) ranges into it.

    Rank r owns ``block[r::world]`` of every subset, concatenated in subset order, so the
    union over ranks is every angle exactly once and the per-rank load differs by at most
    one angle per subset.
    """
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    local, ranges, pos = [], [], 0
    for blk in blocks:
        mine = np.asarray(blk)[rank::world]
        ranges.append((pos, len(mine)))
        local.append(mine)
        pos += len(mine)
    local = np.concatenate(local) if local else np.zeros((0,), dtype=np.int64)
    return local.astype(np.int64), ranges


def shard_angles(n_angles, rank, world):
    """Contiguous-interleaved split of a plain angle list (simulation Ax: no collective)."""
    return np.arange(n_angles, dtype=np.int64)[rank::world]


def subset_geometry(geo, idx, n):
    """Geometry whose per-angle attributes (if any) are restricted to ``idx``."""
    gs = geo.copy()
    for name, width in (("DSD", 1), ("DSO", 1), ("offOrigin", 3), ("offDetector", 2)):
        arr = np.asarray(getattr(geo, name), dtype=np.float64)
        if (width == 1 and arr.shape == (n,)) or (width > 1 and arr.shape == (n, width)):
            setattr(gs, name, arr[idx])
    return gs
