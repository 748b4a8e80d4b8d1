"""Synthetic graph / link generators for measurement and tests (torch ops on the requested device; data
generation is plumbing, not part of the hot path).  There is no network in the build or GPU containers, so
the OGB-shaped workloads are synthetic graphs with the public node / edge counts of those datasets."""
from __future__ import annotations

import torch


def rmat_edges(scale, edge_factor=16, seed=0, device='cpu', a=0.57, b=0.19, c=0.19, chunk=1 << 26):
    """Graph500-style R-MAT edge list (a, b, c, d = 0.57, 0.19, 0.19, 0.05), symmetrised and de-duplicated
    like PyG's to_undirected -> int64 [2, E] sorted by (src, dst)."""
    n = 1 << scale
    e = n * edge_factor
    g = torch.Generator(device=device).manual_seed(seed)
    keys = []
    for lo in range(0, e, chunk):
        cnt = min(chunk, e - lo)
        src = torch.zeros(cnt, dtype=torch.int64, device=device)
        dst = torch.zeros(cnt, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(cnt, generator=g, device=device)
            sb = (r >= a + b).long()
            db = (((r >= a) & (r < a + b)) | (r >= a + b + c)).long()
            src = src * 2 + sb
            dst = dst * 2 + db
        keys.append(src * n + dst)
        keys.append(dst * n + src)
        del src, dst, r, sb, db
    key = torch.unique(torch.cat(keys))
    del keys
    return torch.stack([key // n, key % n])


def powerlaw_edges(num_nodes, num_directed_edges, seed=0, device='cpu', exponent=1.6):
    """undirected heavy-tailed graph with ~num_directed_edges directed entries: endpoints drawn from a Zipf-like
    node distribution and paired with uniform partners (OGB-shaped workloads)"""
    g = torch.Generator(device=device).manual_seed(seed)
    half = num_directed_edges // 2
    u = torch.rand(half, generator=g, device=device)
    heavy = (num_nodes * u.pow(exponent)).long().clamp_(0, num_nodes - 1)
    perm_mult = 2654435761 % num_nodes | 1  # scatter the heavy ids over the id range
    heavy = (heavy * perm_mult) % num_nodes
    other = torch.randint(0, num_nodes, (half,), generator=g, device=device)
    key = torch.unique(torch.cat([heavy * num_nodes + other, other * num_nodes + heavy]))
    return torch.stack([key // num_nodes, key % num_nodes])


def sample_links(num_nodes, edge_index, n_random, n_edges, seed=0, device='cpu'):
    """candidate links: uniform random pairs followed by a sample of true edges -> int64 [L, 2]"""
    g = torch.Generator(device=device).manual_seed(seed + 12345)
    rnd = torch.randint(0, num_nodes, (n_random, 2), generator=g, device=device)
    if n_edges > 0 and edge_index.shape[1] > 0:
        pick = torch.randint(0, edge_index.shape[1], (n_edges,), generator=g, device=device)
        pos = edge_index[:, pick].t()
        return torch.cat([rnd, pos]).contiguous()
    return rnd


# public dataset shapes (node count, directed edge count after to_undirected) used as synthetic stand-ins
SHAPES = {
    'collab': dict(num_nodes=235_868, edges=2_358_104, hops=2, links=2_664_517),
    'ppa': dict(num_nodes=576_289, edges=42_463_862, hops=2, links=3_000_000),
    'citation2': dict(num_nodes=2_927_963, edges=60_703_760, hops=2, links=10_000_000),
}
