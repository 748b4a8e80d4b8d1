"""Times the common-neighbour heuristics kernel (RA) on an OGB-citation2-shaped synthetic graph, next to the
scipy formulation of the reference on a bounded sample (GPU only; measurement aid for SURVEY 8f rank 3)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import heuristics_oracle as ho  # noqa: E402
from subgraph_sketching_b200 import heuristics as bh  # noqa: E402
from subgraph_sketching_b200.graphs import SHAPES, powerlaw_edges, sample_links  # noqa: E402

shape = SHAPES['citation2']
n, L = shape['num_nodes'], 10_000_000
dev = torch.device('cuda', 0)
ei = powerlaw_edges(n, shape['edges'], 0, dev)
links = sample_links(n, ei, L // 2, L - L // 2, 0, dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
adj = bh.SortedAdjacency.from_edge_index(ei, n)
adj.col_sums()
torch.cuda.synchronize()
t_prep = time.perf_counter() - t0
times = []
for _ in range(4):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    scores, _ = bh.RA(adj, links)
    e.record()
    torch.cuda.synchronize()
    times.append(s.elapsed_time(e))
ms = min(times[1:])
deg = (adj.rowptr[1:] - adj.rowptr[:-1]).float()
touched = float((deg[links[:, 0]] + deg[links[:, 1]]).sum()) * 4 + L * 20
print(f'graph: N={n} nnz={adj.colidx.numel()}  links={L}')
print(f'sorted adjacency + column sums (one-off): {t_prep * 1e3:.1f} ms')
print(f'RA kernel: {ms:.2f} ms -> {L / ms / 1e3:.1f} M links/s; adjacency bytes touched {touched / 1e9:.2f} GB '
      f'-> {touched / ms / 1e6:.0f} GB/s')
# CPU: the reference formulation (scipy) on a sample of the same links
sample = links[:200_000].cpu()
A = ho.adjacency(ei.cpu().numpy(), n)
t0 = time.perf_counter()
ref = ho.scores(A, sample.numpy(), 'ra')
t_cpu = time.perf_counter() - t0
err = (scores[:200_000].cpu() - ref).abs().max().item()
print(f'scipy (reference formulation), {len(sample)} links: {t_cpu:.2f} s -> {len(sample) / t_cpu / 1e3:.1f} k links/s; '
      f'max abs diff vs GPU {err:.2e}')
