"""CPU-only, world_size 2 over gloo: the host-side sharding logic of the multi-GPU build (row blocks, link
slices, the per-hop all-gather layout).  Kernels are not involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from subgraph_sketching_b200.dist import balanced_bounds, exchange_blocks, link_slice, shard_bounds


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


def test_balanced_bounds_split_neighbours_evenly():
    g = torch.Generator().manual_seed(0)
    # power-law in-degrees: the low ids are hubs, like R-MAT
    deg = (1000.0 / (1.0 + torch.arange(5000.0)) + 1).long() + torch.randint(0, 3, (5000,), generator=g)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(deg, 0)])
    nnz = int(rowptr[-1])
    for G in (1, 2, 4, 8):
        b = balanced_bounds(rowptr, G)
        assert b[0] == 0 and b[-1] == 5000 and len(b) == G + 1 and all(x <= y for x, y in zip(b, b[1:]))
        shares = [int(rowptr[b[r + 1]] - rowptr[b[r]]) for r in range(G)]
        assert sum(shares) == nnz
        assert max(shares) <= nnz / G + int(deg.max()), (G, shares)
        bw = balanced_bounds(rowptr, G, row_weight=50.0)  # row-dominated cost -> nearly equal row blocks
        assert bw[0] == 0 and bw[-1] == 5000 and all(x <= y for x, y in zip(bw, bw[1:]))
        if G > 1:
            assert bw[1] > b[1]
    assert balanced_bounds(torch.zeros(1, dtype=torch.long), 4) == [0, 0, 0, 0, 0]
    assert balanced_bounds(torch.zeros(6, dtype=torch.long), 2)[-1] == 5


def _worker(rank, world, port, n, width):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        bounds = [0, 9, n]  # deliberately unequal blocks
        lo, hi = bounds[rank], bounds[rank + 1]
        full = torch.full((n, width), 255, dtype=torch.uint8)
        rows = torch.arange(lo, hi).view(-1, 1)
        full[lo:hi] = ((rows * 7 + torch.arange(width).view(1, -1) + rank) % 251).to(torch.uint8)
        exchange_blocks(full, bounds)
        for r in range(world):
            a, b = bounds[r], bounds[r + 1]
            rr = torch.arange(a, b).view(-1, 1)
            want = ((rr * 7 + torch.arange(width).view(1, -1) + r) % 251).to(torch.uint8)
            assert torch.equal(full[a:b], want), f'rank {rank}: block of rank {r} wrong'
        cards = torch.zeros((n, 3))
        cards[lo:hi] = rank + 1.0
        exchange_blocks(cards, bounds)
        assert float(cards.sum()) == sum((bounds[r + 1] - bounds[r]) * (r + 1.0) * 3 for r in range(world))
    finally:
        dist.destroy_process_group()


def test_exchange_layout_world2():
    mp.spawn(_worker, args=(2, _free_port(), 37, 48), nprocs=2, join=True)


def test_halo_masks_and_cumulative_shares():
    """host logic of the halo push: who reads which of my rows, which rows my own copy holds"""
    from subgraph_sketching_b200.dist import cumulative_shares, halo_masks
    assert cumulative_shares(1) == []
    assert cumulative_shares(4) == [0.25, 0.5, 0.75]
    c = cumulative_shares(3, [2.0, 1.0, 1.0])
    assert abs(c[0] - 0.5) < 1e-12 and abs(c[1] - 0.75) < 1e-12
    n, world = 12, 3
    bounds = [0, 3, 8, 12]
    marks = torch.zeros((world, n), dtype=torch.uint8)
    marks[0, [0, 1, 5, 9]] = 1      # rank 0 reads rows 0, 1 (own) and 5, 9 (halo)
    marks[1, [2, 4, 5, 11]] = 1
    marks[2, [0, 4, 9, 10]] = 1
    m1, l1 = halo_masks(marks, bounds, 1)   # rank 1 owns rows 3..7; its peers in order: rank 0 (bit 0), rank 2 (bit 1)
    assert m1.tolist() == [0, 2, 1, 0, 0]   # row 4 -> rank 2, row 5 -> rank 0
    assert l1.tolist() == [0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1]
    m0, l0 = halo_masks(marks, bounds, 0)   # peers: rank 1 (bit 0), rank 2 (bit 1)
    assert m0.tolist() == [2, 0, 1] and l0.tolist() == [1, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 0]
    m2, _ = halo_masks(marks, bounds, 2)    # rows 8..11; peers: rank 0 (bit 0), rank 1 (bit 1)
    assert m2.tolist() == [0, 1, 0, 2]
