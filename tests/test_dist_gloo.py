"""CPU-only, world_size 2 over gloo: the host-side sharding logic of the multi-GPU build (row blocks, link
slices, the per-hop all-gather layout).  Kernels are not involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from subgraph_sketching_b200.dist import allgather_rows, link_slice, shard_bounds


def test_shard_bounds_cover_all_rows():
    for n in (0, 1, 7, 8, 9, 1000, 16_777_216):
        for g in (1, 2, 4, 8):
            seen = 0
            for r in range(g):
                per, lo, hi = shard_bounds(n, g, r)
                assert lo == min(r * per, n) and lo <= hi <= n and hi - lo <= per
                seen += hi - lo
            assert seen == n and per * g >= n
            assert sum(b - a for a, b in (link_slice(n, g, r) for r in range(g))) == n


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, width):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        per, lo, hi = shard_bounds(n, world, rank)
        full = torch.full((per * world, width), 255, dtype=torch.uint8)
        # each rank fills its own block with a rank/row specific pattern
        rows = torch.arange(lo, hi).view(-1, 1)
        full[lo:hi] = ((rows * 7 + torch.arange(width).view(1, -1) + rank) % 251).to(torch.uint8)
        allgather_rows(full, per, rank, world)
        for r in range(world):
            _, a, b = shard_bounds(n, world, r)
            rr = torch.arange(a, b).view(-1, 1)
            want = ((rr * 7 + torch.arange(width).view(1, -1) + r) % 251).to(torch.uint8)
            assert torch.equal(full[a:b], want), f'rank {rank}: block of rank {r} wrong'
        cards = torch.zeros((per * world, 3))
        cards[lo:hi] = rank + 1.0
        allgather_rows(cards, per, rank, world)
        assert float(cards[:n].sum()) == sum((shard_bounds(n, world, r)[2] - shard_bounds(n, world, r)[1]) * (r + 1.0) * 3
                                             for r in range(world))
    finally:
        dist.destroy_process_group()


def test_allgather_layout_world2():
    mp.spawn(_worker, args=(2, _free_port(), 37, 48), nprocs=2, join=True)
