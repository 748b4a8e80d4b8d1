"""GPU tests of the node-sharded engine over NCCL: world_size 1 always, world_size 2 when two GPUs are
visible.  The sharded tables must be bit-identical to the single-GPU tables."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_args, rmat_edges

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, scale, K, exchange='auto'):
    import subgraph_sketching_b200 as ssb
    from subgraph_sketching_b200.dist import ShardedElphHashes, link_slice
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        n = 1 << scale
        ei = rmat_edges(scale, 16, 3).to(dev)
        g = torch.Generator().manual_seed(4)
        links = torch.randint(0, n, (20001, 2), generator=g).to(dev)
        sh = ShardedElphHashes(make_args(K), exchange=exchange)
        tables, cards = sh.build_hash_tables(n, ei)
        feats = sh.get_subgraph_features(links, tables, cards)
        print(f'rank {rank}/{world}: exchange={sh.exchange} ({sh.exchange_error})', flush=True)
        if exchange != 'auto':
            assert sh.exchange == exchange
        # a second build reuses the symmetric buffers and must give the same tables
        tables, cards = sh.build_hash_tables(n, ei)
        one = ssb.ElphHashes(make_args(K))
        t1, c1 = one.build_hash_tables(n, ei)
        f1 = one.get_subgraph_features(links, t1, c1)
        for k in range(K + 1):
            assert torch.equal(tables.records(k), t1.records(k)), f'rank {rank}: hop {k} records differ'
        assert torch.equal(cards, c1)
        lo, hi = link_slice(links.shape[0], world, rank)
        assert torch.equal(feats, f1[lo:hi])
        # a pinned HOST edge list: every rank pulls only its slice over PCIe, the rest arrives over NVLink
        th, ch = sh.build_hash_tables(n, ei.cpu().pin_memory())
        for k in range(K + 1):
            assert torch.equal(th.records(k), t1.records(k)), f'rank {rank}: host-fed hop {k} records differ'
        assert torch.equal(ch, c1)
        shares = [int(x) for x in sh.bounds]
        assert shares[0] == 0 and shares[-1] == n
        if world > 1:  # blocks are balanced by neighbour count: the hub block is much shorter in rows
            assert shares[1] < n // world
    finally:
        dist.destroy_process_group()


def test_sharded_world1_matches_single_gpu():
    mp.spawn(_worker, args=(1, _free_port(), 12, 2), nprocs=1, join=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('exchange', ['nccl', 'auto', 'mc'])
def test_sharded_world2_matches_single_gpu(exchange):
    mp.spawn(_worker, args=(2, _free_port(), 13, 3, exchange), nprocs=2, join=True)
