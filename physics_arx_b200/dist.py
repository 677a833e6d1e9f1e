"""Sharding for multi-GPU runs: one process per GPU, volume replicated, one
all-reduce of the partial volume update per OS-SART subset.  Host logic only.

Two ways to cut a subset's rays over the ranks, both linear in the projections so the partial
updates (and the partial V maps) add up exactly:
  'angles'  the prescribed scheme: the angles of every subset interleaved over the ranks;
  'rows'    every rank takes ALL angles of the subset but only a band of detector rows.  A row
            band of a detector is itself a (shorter, V-shifted) detector, so nothing below the
            geometry bag changes; bilinear interpolation across a band boundary splits into two
            ramps that sum to one.  Balanced for any rank count (20 angles over 8 ranks are not).
  'grid'    both at once: the ranks form an A x R grid (choose_shard); 'auto' picks the grid that gives
            every rank the same number of angles per subset."""
from __future__ import annotations

import numpy as np


def shard_blocks(blocks, rank, world, contiguous=False):
    """Rank-local angle list and per-subset (first, count) ranges into it.

    Rank r owns ``block[r::world]`` of every subset (or, with ``contiguous``, the r-th of `world` consecutive
    stretches of the block), concatenated in subset order, so the union over ranks is every angle exactly once
    and the per-rank load differs by at most one angle per subset.
    """
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    local, ranges, pos = [], [], 0
    for blk in blocks:
        blk = np.asarray(blk)
        mine = blk[len(blk) * rank // world:len(blk) * (rank + 1) // world] if contiguous else blk[rank::world]
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


def shard_rows(n_rows, rank, world, align=32):
    """Row band [v0, v1) of rank `rank`: contiguous, aligned to `align` rows where possible."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    units = -(-n_rows // align)
    if units >= world:
        lo = (units * rank) // world * align
        hi = (units * (rank + 1)) // world * align
    else:                                   # fewer aligned bands than ranks: split single rows
        lo = (n_rows * rank) // world
        hi = (n_rows * (rank + 1)) // world
    return min(lo, n_rows), min(hi, n_rows)


def row_window_geometry(geo, v0, v1, n_angles=None):
    """The detector rows [v0, v1) of `geo` as a geometry of their own (A.2: a pixel's v coordinate
    is (v - nV/2 + 1/2) dV + offV, so the band is a detector of v1-v0 rows shifted in V)."""
    gs = geo.copy()
    nv = int(np.asarray(geo.nDetector)[0])
    dv = float(np.asarray(geo.dDetector, dtype=np.float64)[0])
    shift = (0.5 * (v0 + v1) - 0.5 * nv) * dv
    gs.nDetector = np.array((v1 - v0, int(np.asarray(geo.nDetector)[1])))
    gs.sDetector = gs.nDetector * np.asarray(geo.dDetector, dtype=np.float64)
    off = np.array(np.asarray(geo.offDetector, dtype=np.float64), copy=True)
    off[..., 0] += shift
    gs.offDetector = off
    return gs


def col_window_geometry(geo, u0, u1):
    """The detector columns [u0, u1) of `geo` as a geometry of their own (a pixel's u coordinate is
    (u - nU/2 + 1/2) dU + offU, so the band is a detector of u1-u0 columns shifted in U).  Bilinear interpolation
    across a band edge splits into two ramps that sum to one, exactly as for a band of rows."""
    gs = geo.copy()
    nu = int(np.asarray(geo.nDetector)[1])
    du = float(np.asarray(geo.dDetector, dtype=np.float64)[1])
    shift = (0.5 * (u0 + u1) - 0.5 * nu) * du
    gs.nDetector = np.array((int(np.asarray(geo.nDetector)[0]), u1 - u0))
    gs.sDetector = gs.nDetector * np.asarray(geo.dDetector, dtype=np.float64)
    off = np.array(np.asarray(geo.offDetector, dtype=np.float64), copy=True)
    off[..., 1] += shift
    gs.offDetector = off
    return gs


def band_axis(shard):
    """Which detector axis the bands of a grid cut: 'cols' (default: the fan-plane projector's work is per
    detector column, so a band of columns costs its share and a band of rows costs the whole) or 'rows'."""
    return "rows" if shard == "rows" else "cols"


def choose_shard(blocks, world, shard="auto"):
    """(A, R): the ranks form an A x R grid -- the angles of every subset are interleaved over A groups and
    every group splits the detector rows into R bands; rank r sits at (r % A, r // A).

      'angles' -> (world, 1)   the prescribed scheme
      'rows'   -> (1, world)   bands of detector rows
      'cols'   -> (1, world)   bands of detector columns
      'grid' / 'auto' -> A = gcd(world, every subset's size), R = world / A: every rank gets the same
                number of angles in every subset; the bands cut the detector COLUMNS (band_axis).
                Blocksize 20 with a tail of 16 over 8 ranks gives 4 x 2: 5 (4) angles x half the columns per
                rank, instead of 3,3,3,3,2,2,2,2 whole projections.  See DESIGN.md section 6.
    """
    if isinstance(shard, (tuple, list)):                    # an explicit grid
        a, r = (int(v) for v in shard)
        if a < 1 or r < 1 or a * r != world:
            raise ValueError(f"grid {shard} does not have {world} ranks")
        return a, r
    if shard == "angles":
        return world, 1
    if shard in ("rows", "cols"):
        return 1, world
    if shard not in ("auto", "grid"):
        raise ValueError(f"unknown shard mode {shard!r}")
    a = world
    for b in blocks:
        a = int(np.gcd(a, len(b)))
    return a, world // a
