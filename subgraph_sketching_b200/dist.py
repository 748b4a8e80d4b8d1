"""Node-sharded multi-GPU build of the sketch tables (one process per GPU, torch.distributed).

The path shards naturally by destination node: rank r owns a contiguous block of rows of every
hop table and the CSR rows of those destinations.  One hop = local k-hop merge of the owned rows (reads the
full previous-hop table, writes the owned slice of the next one) followed by ONE all-gather of the owned
slices (NCCL over NVLink / NVSwitch), after which every rank holds the full hop table again.  After the last
hop the tables are replicated, so candidate links shard trivially: each rank computes the features of its
slice of the link list with no further communication.

The reference has no distributed code at all (src/hashing.py is single process); results are bit-identical
to the single-GPU engine because min/max merges do not depend on the partition.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .hashing import ElphHashes, HopSketch, SketchTables, build_csr, _to_host


def shard_bounds(num_nodes, world_size, rank):
    """contiguous row block of `rank`: equal blocks of ceil(N / G) rows, the last ones may be short/empty"""
    per = (num_nodes + world_size - 1) // world_size if world_size > 0 else num_nodes
    lo = min(rank * per, num_nodes)
    hi = min(lo + per, num_nodes)
    return per, lo, hi


def link_slice(n_links, world_size, rank):
    per = (n_links + world_size - 1) // world_size
    lo = min(rank * per, n_links)
    return lo, min(lo + per, n_links)


def allgather_rows(full, per, rank, world_size, group=None):
    """in-place all-gather: rank r's rows [r*per, (r+1)*per) of `full` ([G*per, ...]) are sent to everyone"""
    assert full.shape[0] == per * world_size
    mine = full[rank * per:(rank + 1) * per]
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(full, mine, group=group)
    else:  # gloo (CPU tests): list form, out-of-place input
        parts = [full[r * per:(r + 1) * per] for r in range(world_size)]
        dist.all_gather(parts, mine.clone(), group=group)
    return full


class ShardedElphHashes(object):
    """ElphHashes over `world_size` GPUs: same build_hash_tables / get_subgraph_features surface; every rank
    passes the same (replicated) edge_index and link list and gets the full tables plus ITS slice of features
    (`link_slice`)."""

    def __init__(self, args, group=None, **kw):
        self.eh = ElphHashes(args, **kw)
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def build_hash_tables(self, num_nodes, edge_index):
        eh, G, r = self.eh, self.world_size, self.rank
        device = edge_index.device
        assert device.type == 'cuda', 'the sharded build takes device-resident edges'
        per, lo, hi = shard_bounds(num_nodes, G, r)
        K = eh.max_hops
        with torch.cuda.device(device):
            ev = eh._event_begin(device)
            rowptr, colidx, nnz, max_id = build_csr(edge_index, device, num_rows=hi - lo, add_loops=True, row_begin=lo)
            eh._event_end('csr_build', ev, device)
            if max_id >= num_nodes:
                raise IndexError(f'edge_index refers to node {max_id} but num_nodes is {num_nodes}')
            rb = eh._record_bytes()
            n_pad = per * G
            recs = [torch.empty((n_pad, rb), dtype=torch.uint8, device=device) for _ in range(K + 1)]
            cards = torch.zeros((n_pad, K), dtype=torch.float32, device=device)
            ev = eh._event_begin(device)
            eh._init_records(num_nodes, device, out=recs[0][:num_nodes])  # hop 0 is cheap: no exchange needed
            eh._event_end('init_records', ev, device)
            ws = None
            for k in range(1, K + 1):
                if hi > lo:
                    ws = eh._merge(rowptr, colidx, nnz, recs[k - 1], recs[k][lo:hi], cards[lo:hi, k - 1], device, ws)
                ev = eh._event_begin(device)
                allgather_rows(recs[k], per, r, G, self.group)
                eh._event_end('allgather', ev, device)
            allgather_rows(cards, per, r, G, self.group)
            tables = SketchTables({k: HopSketch(recs[k][:num_nodes], eh.num_perm, eh.p, device) for k in range(K + 1)},
                                  eh.num_perm, eh.p)
            return tables, cards[:num_nodes]

    def get_subgraph_features(self, links, hash_table, cards, batch_size=11000000):
        """features of this rank's slice of `links` (rows link_slice(len(links), G, rank))"""
        lo, hi = link_slice(links.shape[0], self.world_size, self.rank)
        return self.eh.get_subgraph_features(links[lo:hi], hash_table, cards, batch_size)
