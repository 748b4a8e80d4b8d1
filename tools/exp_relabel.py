"""Experiment (GPU only): does the ORDER of the node ids matter for the k-hop merge?
R-MAT hubs sit at ids with few set bits (0, 1, 2, 4, ... 2^k): with a 1024-byte record pitch their rows are
spaced by large powers of two.  This times one hop (hop 1 table -> hop 2 table, production kernel and layout,
NO persisting-L2 set-aside) on the same graph under different relabellings:
    identity | random permutation | descending degree | log2-degree buckets (stable counting sort, what a
    cheap on-device relabel could afford)
usage: python tools/exp_relabel.py [scale]      -> stdout + gpurun_out/exp_relabel.json"""
import json
import os
import sys
from argparse import Namespace

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402
from subgraph_sketching_b200.graphs import rmat_edges  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device('cuda', 0)
n = 1 << scale
eh = ssb.ElphHashes(Namespace(max_hash_hops=2, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
ei = rmat_edges(scale, 16, 0, dev)
deg = torch.bincount(ei[1], minlength=n)


def perm_for(kind):
    if kind == 'identity':
        return None
    if kind == 'random':
        order = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    elif kind == 'degree-desc':
        order = torch.argsort(deg, descending=True, stable=True)
    elif kind == 'log2-buckets':
        bucket = 63 - torch.log2(deg.double() + 1).floor().long()  # high degree first
        order = torch.argsort(bucket, stable=True)
    perm = torch.empty(n, dtype=torch.int64, device=dev)
    perm[order] = torch.arange(n, device=dev)
    return perm


rows = []
for kind in ('identity', 'random', 'degree-desc', 'log2-buckets'):
    perm = perm_for(kind)
    e2 = ei if perm is None else perm[ei]
    rowptr, colidx, nnz, _ = ssb.build_csr(e2, dev, num_rows=n, add_loops=True)
    del e2
    rec = eh._alloc_hop_tables(n, 768, dev)
    eh._init_records(n, dev, out=rec[0])
    cards = torch.zeros((n, 2), device=dev)
    ws = eh._merge(rowptr, colidx, nnz, rec[0], rec[1], cards[:, 0], dev)
    times = []
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eh._merge(rowptr, colidx, nnz, rec[1], rec[2], cards[:, 1], dev, ws)
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    row = dict(relabel=kind, stride=int(rec[1].stride(0)), hop2_ms=min(times[1:]), all_ms=times)
    rows.append(row)
    print(f'{kind:14s} stride {row["stride"]}: hop 2 = {row["hop2_ms"]:.2f} ms', flush=True)
    del rowptr, colidx, rec, cards, ws, perm
    torch.cuda.empty_cache()
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/exp_relabel.json', 'w'), indent=1)
