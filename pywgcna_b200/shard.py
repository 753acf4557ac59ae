"""Row-block sharding of the hot path across GPUs -- the host-side arithmetic and the one
exchange primitive, free of any CUDA dependency so that the N > 1 logic is testable with gloo on CPU.

Rank r of G owns the contiguous rows ``[r * per, min(n, (r + 1) * per))`` of the adjacency / TOM and the
same block of sweep columns; ``per = ceil(n / G)`` rounded up to ``align`` rows (the operand tile
height), so every block starts on a tile boundary and all blocks but the last have equal size.  Buffers
that are exchanged hold ``per * G`` rows, which lets one ``all_gather_into_tensor`` move every block.
"""
from __future__ import annotations

import torch


def rows_per_rank(n: int, world: int, align: int = 128) -> int:
    per = -(-n // world)
    return -(-per // align) * align


def row_block(n: int, world: int, rank: int, align: int = 128):
    """(row0, nrows) of ``rank``; nrows may be 0 for trailing ranks of a small problem."""
    per = rows_per_rank(n, world, align)
    r0 = min(n, rank * per)
    r1 = min(n, r0 + per)
    return r0, r1 - r0


def padded_rows(n: int, world: int, align: int = 128) -> int:
    return rows_per_rank(n, world, align) * world


def symmetric_tile_count(n: int, world: int = 1, rank: int = 0, per: int | None = None, tile: int = 256) -> int:
    """Number of 256 x 256 tiles the symmetric int8 TOM schedule gives to ``rank`` (same deal as
    b200w_tom_compute_i8: diagonal tiles to their owner, tile {a, b} to the owner of a if a + b is even,
    else to the owner of b).  world == 1: the whole upper triangle, T (T + 1) / 2."""
    T = -(-n // tile)
    if world == 1:
        return T * (T + 1) // 2
    ptiles = per // tile
    cnt = 0
    for a in range(T):
        oa = a // ptiles
        cnt += oa == rank
        for b in range(a + 1, T):
            ob = b // ptiles
            o = oa if oa == ob else (oa if (a + b) % 2 == 0 else ob)
            cnt += o == rank
    return cnt


def all_gather_rows(full: torch.Tensor, per: int, rank: int, world: int, group=None) -> None:
    """In-place all-gather: ``full`` is [per * world, ...] contiguous and rank r has filled rows
    ``[r * per, (r + 1) * per)``; afterwards every rank holds every block."""
    if world == 1:
        return
    import torch.distributed as dist

    assert full.is_contiguous() and full.shape[0] == per * world
    mine = full[rank * per:(rank + 1) * per]
    dist.all_gather_into_tensor(full, mine, group=group)
