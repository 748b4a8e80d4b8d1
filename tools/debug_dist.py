"""2-rank smoke of the sharded build (debug aid): torchrun --nproc-per-node 2 tools/debug_dist.py [exchange]"""
import os
import sys
from argparse import Namespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402
from subgraph_sketching_b200.dist import ShardedElphHashes, link_slice  # noqa: E402
from subgraph_sketching_b200.graphs import rmat_edges  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
exchange = sys.argv[1] if len(sys.argv) > 1 else 'halo'
scale, K = 13, 3
n = 1 << scale
ei = rmat_edges(scale, 16, 3).to(dev)
args = Namespace(max_hash_hops=K, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False)
links = torch.randint(0, n, (20001, 2), generator=torch.Generator().manual_seed(4)).to(dev)
one = ssb.ElphHashes(args)
t1, c1 = one.build_hash_tables(n, ei)
f1 = one.get_subgraph_features(links, t1, c1)
sh = ShardedElphHashes(args, exchange=exchange)
tables, cards = sh.build_hash_tables(n, ei)
lo, hi = sh.bounds[rank], sh.bounds[rank + 1]
print(f'[rank {rank}] exchange={sh.exchange} csr={sh.csr_path} bounds={sh.bounds} halo={sh.halo_fraction}', flush=True)
for k in range(K + 1):
    print(f'[rank {rank}] hop {k} own block equal: {torch.equal(tables.records(k)[lo:hi], t1.records(k)[lo:hi])}', flush=True)
print(f'[rank {rank}] cards equal: {torch.equal(cards, c1)}', flush=True)
feats = sh.get_subgraph_features(links, tables, cards)
a, b = link_slice(links.shape[0], world, rank)
print(f'[rank {rank}] features equal: {torch.equal(feats, f1[a:b])}', flush=True)
dist.destroy_process_group()
