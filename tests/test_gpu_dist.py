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
        one = ssb.ElphHashes(make_args(K))
        t1, c1 = one.build_hash_tables(n, ei)
        f1 = one.get_subgraph_features(links, t1, c1)
        lo_l, hi_l = link_slice(links.shape[0], world, rank)

        def check(sh, tables, cards, what):
            complete = tables.sharding['complete']
            lo, hi = sh.bounds[rank], sh.bounds[rank + 1]
            for k in range(K + 1):
                if complete or k == 0:
                    assert torch.equal(tables.records(k), t1.records(k)), f'rank {rank} {what}: hop {k} records differ'
                else:  # halo build: the own block is authoritative, and so is every row this rank's lists read
                    assert torch.equal(tables.records(k)[lo:hi], t1.records(k)[lo:hi]), f'rank {rank} {what}: hop {k} own block'
                    if k < K:
                        have = sh._shard[1].bool()
                        assert torch.equal(tables.records(k)[have], t1.records(k)[have]), f'rank {rank} {what}: hop {k} halo'
            assert torch.equal(cards, c1), f'rank {rank} {what}: cards differ'
            feats = sh.get_subgraph_features(links, tables, cards)
            assert torch.equal(feats, f1[lo_l:hi_l]), f'rank {rank} {what}: features differ'

        sh = ShardedElphHashes(make_args(K), exchange=exchange)
        tables, cards = sh.build_hash_tables(n, ei)
        print(f'rank {rank}/{world}: exchange={sh.exchange} ({sh.exchange_error}) csr={sh.csr_path} '
              f'halo fraction={sh.halo_fraction}', flush=True)
        if exchange != 'auto':
            assert sh.exchange == exchange
        if world > 1:
            assert sh.csr_path == 'streaming'      # R-MAT lists are ordered by source and symmetric
        check(sh, tables, cards, 'first build')
        # a second build reuses the symmetric buffers: same result, and the first build's tables are now invalid
        tables2, cards2 = sh.build_hash_tables(n, ei)
        check(sh, tables2, cards2, 'second build')
        if world > 1 and sh.exchange != 'nccl':
            with pytest.raises(RuntimeError):
                tables.records(1)
        # a pinned HOST edge list: every rank pulls only its own slice over PCIe
        th, ch = sh.build_hash_tables(n, ei.cpu().pin_memory())
        check(sh, th, ch, 'host-fed build')
        # a list that is not ordered takes the histogram build (and, host-fed, the NVLink re-assembly)
        perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(9)).to(dev)
        ts, cs = sh.build_hash_tables(n, ei[:, perm])
        assert world == 1 or sh.csr_path == 'histogram'
        check(sh, ts, cs, 'shuffled build')
        ts, cs = sh.build_hash_tables(n, ei[:, perm].cpu().pin_memory())
        check(sh, ts, cs, 'shuffled host-fed build')
        shares = [int(x) for x in sh.bounds]
        assert shares[0] == 0 and shares[-1] == n
        if world > 1:  # blocks are balanced by neighbour count: the hub block is much shorter in rows
            assert shares[1] < n // world
        # fresh buffers per build keep earlier tables alive
        if world > 1 and exchange in ('auto', 'halo'):
            sh2 = ShardedElphHashes(make_args(K), exchange=exchange, reuse_buffers=False)
            ta, ca = sh2.build_hash_tables(n, ei)
            tb, cb = sh2.build_hash_tables(n, ei)
            lo, hi = sh2.bounds[rank], sh2.bounds[rank + 1]
            assert torch.equal(ta.records(K)[lo:hi], t1.records(K)[lo:hi]) and torch.equal(tb.records(K)[lo:hi], t1.records(K)[lo:hi])
    finally:
        dist.destroy_process_group()


def test_sharded_world1_matches_single_gpu():
    mp.spawn(_worker, args=(1, _free_port(), 12, 2), nprocs=1, join=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('exchange', ['nccl', 'halo', 'p2p', 'mc'])
def test_sharded_world2_matches_single_gpu(exchange):
    mp.spawn(_worker, args=(2, _free_port(), 13, 3, exchange), nprocs=2, join=True)
