"""Host-side logic of the N>1 path on CPU: contiguous sharding of slices over ranks and the (off-hot-path)
all-gather of stream sizes, exercised with the gloo backend and world_size 2/3."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llcomp_b200.sharding import global_stream_index, shard_range, tile_rows_for_rank


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_tile_rows_cover_image():
    for h, th, world in ((4096, 512, 8), (600, 128, 2), (100, 512, 4), (16384, 256, 8)):
        spans = [tile_rows_for_rank(h, th, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == h
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert all(y0 % th == 0 for y0, _ in spans if y0 < h)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_items, rank, world)
        sizes = [1000 + 7 * k for k in range(lo, hi)]            # pretend stream sizes of my slices
        off, totals, counts = global_stream_index(sizes)
        out[rank] = (lo, hi, off, totals, counts)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_global_index_over_gloo(world):
    n_items = 11
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_items, out), nprocs=world, join=True)
        res = dict(out)
    all_sizes = [1000 + 7 * k for k in range(n_items)]
    pos = 0
    for r in range(world):
        lo, hi, off, totals, counts = res[r]
        assert off == pos == sum(all_sizes[:lo])
        assert counts[r] == hi - lo and totals[r] == sum(all_sizes[lo:hi])
        pos += totals[r]
    assert pos == sum(all_sizes)
