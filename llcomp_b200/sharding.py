"""Multi-GPU sharding of the hot path: independent slices, one process per GPU, no data-path collective.

Slices (whole images of a batch, or the tiles of one large image) share nothing -- own payload, own
63 KB of adaptive state, own line buffers -- so rank r simply codes its contiguous block of slices on
its own GPU.  The only exchange is off the hot path: the per-rank payload byte counts are all-gathered
(a few bytes per rank) so that every rank knows where its block starts in the job-wide stream index.
"""
from __future__ import annotations

from typing import Sequence


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; blocks differ in size by at most one item."""
    if world < 1 or not 0 <= rank < world or n_items < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def tile_rows_for_rank(height: int, tile_h: int, rank: int, world: int) -> tuple[int, int]:
    """Single large image: rank r takes a contiguous range of tile rows -> pixel rows [y0, y1)."""
    n_tile_rows = (height + tile_h - 1) // tile_h
    lo, hi = shard_range(n_tile_rows, rank, world)
    return min(lo * tile_h, height), min(hi * tile_h, height)


def global_stream_index(local_sizes: Sequence[int], group=None):
    """All-gather of per-rank stream sizes -> (byte offset of this rank's block, per-rank totals).

    `local_sizes` are the byte sizes of the streams this rank produced, in slice order.  Works with any
    torch.distributed backend (nccl on the GPUs, gloo in the CPU tests); a few bytes per rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([int(sum(local_sizes)), len(local_sizes)], dtype=torch.int64, device=dev)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    totals = [int(t[0]) for t in allv]
    counts = [int(t[1]) for t in allv]
    return sum(totals[:rank]), totals, counts
