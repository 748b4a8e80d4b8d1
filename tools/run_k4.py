"""one get_subgraph_features call on an R-MAT graph (profiling target): python tools/run_k4.py K [scale] [links] [grouped]"""
import os
import sys
from argparse import Namespace

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402
from subgraph_sketching_b200.graphs import rmat_edges, sample_links  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 22
L = int(sys.argv[3]) if len(sys.argv) > 3 else 5_000_000
dev = torch.device('cuda', 0)
n = 1 << scale
ei = rmat_edges(scale, 16, 0, dev)
eh = ssb.ElphHashes(Namespace(max_hash_hops=K, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
tables, cards = eh.build_hash_tables(n, ei)
links = sample_links(n, ei, L // 2, L - L // 2, 0, dev)
if len(sys.argv) > 4:
    links[:, 0] = links[torch.arange(L, device=dev) // 1001, 0]
for _ in range(2):
    f = eh.get_subgraph_features(links, tables, cards)
torch.cuda.synchronize()
print(f.shape)
