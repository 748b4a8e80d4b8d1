"""Node-sharded multi-GPU build of the sketch tables (one process per GPU, torch.distributed).

The path shards naturally by destination node: rank r owns a contiguous block of rows of every
hop table and the CSR rows of those destinations.  One hop = local k-hop merge of the owned rows (reads the
full previous-hop table, writes the owned block of the next one) followed by ONE exchange step in which every
rank's block is replicated to all others (NCCL over NVLink / NVSwitch), after which every rank holds the full
hop table again.  After the last hop the tables are replicated, so candidate links shard trivially: each
rank computes the features of its slice of the link list with no further communication.

Row blocks are balanced by NEIGHBOUR COUNT, not by row count: on power-law graphs the low ids are the hubs
(equal row blocks gave rank 0 74 % of the edges of an R-MAT-24 graph at 2 ranks), so the block boundaries are
the quantiles of the global rowptr.  Blocks therefore differ in size and the exchange is one broadcast per
owner block (same bytes on the wire as an all-gather).

Exchange modes
  'p2p'  (default when symmetric memory works): the hop tables live in symmetric memory
         (torch.distributed._symmetric_memory: every rank's buffer is mapped into every process over NVLink)
         and the merge kernel itself stores each finished row into all peer tables (ss_khop_merge_peers), so
         the exchange is fused into the kernel and overlaps the merge row by row; a stream-ordered
         symmetric-memory barrier separates the hops.  No NCCL call on the data path.
  'mc'   like 'p2p', but every store is ONE multimem.st to the NVSwitch multicast address of the buffer: the
         switch replicates the row into all GPUs' tables, so a GPU sends each row once instead of G-1 times.
  'nccl' one torch.distributed broadcast per owner block after the merge kernel (works everywhere).

The reference has no distributed code at all (src/hashing.py is single process); results are bit-identical
to the single-GPU engine because min/max merges do not depend on the partition.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from .hashing import (INGEST_MIN_EDGES, ElphHashes, HopSketch, SketchTables, _edge_source, _env_int, _ptr,
                      _stream_ptr, _streamed_degree_pass, check, lib)


def shard_bounds(num_nodes, world_size, rank):
    """equal contiguous blocks of ceil(N / G) rows (the last ones may be short / empty)"""
    per = (num_nodes + world_size - 1) // world_size if world_size > 0 else num_nodes
    lo = min(rank * per, num_nodes)
    hi = min(lo + per, num_nodes)
    return per, lo, hi


def link_slice(n_links, world_size, rank):
    per = (n_links + world_size - 1) // world_size
    lo = min(rank * per, n_links)
    return lo, min(lo + per, n_links)


def balanced_bounds(rowptr, world_size, row_weight=0.0, shares=None):
    """row boundaries [b_0 = 0, ..., b_G = N] such that block r carries the fraction shares[r] (default 1/G)
    of the total COST, cost(block) = neighbours(block) + row_weight * rows(block).
    One neighbour = one 768-byte gather from HBM; a finished row costs one local store plus, in the fused p2p
    exchange, one 768-byte store per peer over NVLink, so the row term dominates at 8 GPUs and vanishes at 1.
    `rowptr` is the global int64 [N + 1] prefix sum (any device); returns a python list of G + 1 ints."""
    n = rowptr.numel() - 1
    if n <= 0:
        return [0] * (world_size + 1)
    if world_size == 1:
        return [0, n]
    if shares is None:
        shares = [1.0 / world_size] * world_size
    tot_share = float(sum(shares))
    cum, acc = [], 0.0
    for r in range(world_size - 1):
        acc += float(shares[r]) / tot_share
        cum.append(acc)
    cost = rowptr.double() + float(row_weight) * torch.arange(n + 1, device=rowptr.device, dtype=torch.float64)
    total = float(cost[-1])
    targets = torch.tensor([total * c for c in cum], dtype=torch.float64, device=rowptr.device)
    cuts = torch.searchsorted(cost, targets, right=False).clamp_(0, n).tolist()
    bounds = [0] + [int(c) for c in cuts] + [n]
    for i in range(1, len(bounds)):  # monotone, in range
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def default_row_weight(world_size, exchange):
    """cost of one output row in units of one neighbour gather (measured on B200 / NVLink 5)"""
    if world_size <= 1:
        return 0.0
    return 1.5 + (10.0 * (world_size - 1) if exchange != 'nccl' else 4.0)


def exchange_blocks(full, bounds, group=None):
    """replicate every owner block full[bounds[r]:bounds[r+1]] from rank r to all ranks (in place)"""
    works = []
    for r in range(len(bounds) - 1):
        lo, hi = bounds[r], bounds[r + 1]
        if hi > lo:
            src = dist.get_global_rank(group, r) if group is not None else r
            works.append(dist.broadcast(full[lo:hi], src=src, group=group, async_op=True))
    for w in works:
        w.wait()
    return full


class ShardedElphHashes(object):
    """ElphHashes over `world_size` GPUs: same build_hash_tables / get_subgraph_features surface; every rank
    passes the same (replicated) edge_index and link list and gets the full tables plus ITS slice of features
    (`link_slice`)."""

    def __init__(self, args, group=None, exchange='auto', **kw):
        assert exchange in ('auto', 'p2p', 'mc', 'nccl')
        self.eh = ElphHashes(args, **kw)
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bounds = None
        self.local_nnz = None
        self.exchange = exchange
        self._symm = None       # cached symmetric buffers: (key, recs[1..K], cards, handles)
        # adaptive balance: the share of the total cost each rank gets follows its measured merge throughput
        # in the previous build (ranks differ in L2 hit rate and NVLink egress, which no static model captures)
        self.adaptive = True
        self.shares = None
        self._merge_events = []
        self.exchange_error = None
        if self.world_size == 1 or self.world_size - 1 > 7:
            self.exchange = 'nccl'

    # ------------------------------------------------------------------ symmetric-memory tables
    def _symmetric_buffers(self, num_nodes, K, rb, device):
        """hop tables 1..K and cards in symmetric memory (allocated once per shape, reused by later builds);
        returns None (on every rank consistently) when symmetric memory is unavailable"""
        key = (num_nodes, K, rb, str(device))
        if self._symm is not None and self._symm[0] == key:
            return self._symm
        ok = 1
        recs, hdls, cards, chdl = [], [], None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.group if self.group is not None else dist.group.WORLD
            for _ in range(K):
                t = symm_mem.empty(max(num_nodes, 1) * rb, dtype=torch.uint8, device=device)
                hdls.append(symm_mem.rendezvous(t, group=grp))
                recs.append(t.view(max(num_nodes, 1), rb)[:num_nodes])
            c = symm_mem.empty(max(num_nodes, 1) * K, dtype=torch.float32, device=device)
            chdl = symm_mem.rendezvous(c, group=grp)
            cards = c.view(max(num_nodes, 1), K)[:num_nodes]
        except Exception as e:  # noqa: BLE001 -- any failure means: fall back to NCCL, on every rank
            ok = 0
            self.exchange_error = f'{type(e).__name__}: {e}'
        flag = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            return None
        self._symm = (key, recs, cards, hdls, chdl)
        return self._symm

    def _local_csr(self, edge_index, num_nodes, device):
        """global rowptr (every rank computes the same one) -> balanced bounds -> this rank's CSR rows"""
        ei, zero_copy = _edge_source(edge_index, device)
        n_edges = ei.shape[1]
        src, dst = ei[0], ei[1]
        ws_bytes = check(lib.ss_csr_workspace_bytes(num_nodes), 'ss_csr_workspace_bytes')
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=device)
        rowptr_g = torch.empty(num_nodes + 1, dtype=torch.int64, device=device)
        stats = torch.empty(4, dtype=torch.int64, device=device)
        src32 = dst32 = None
        st = _stream_ptr(device)
        G, r = self.world_size, self.rank
        if G > 1 and n_edges >= (1 << 10):
            # sharded first pass: rank r scans only its 1/G slice of the edge list (for a pinned HOST list that is
            # also all it pulls over its own PCIe link).  Prefix sums are linear, so the global rowptr is the SUM
            # over ranks of the partial ones (one all-reduce) plus the self loops; a host-resident list is
            # re-assembled on every GPU from the 32-bit slice copies with one all-gather over NVLink.
            per = (n_edges + G - 1) // G
            e_lo, e_hi = min(r * per, n_edges), min((r + 1) * per, n_edges)
            s32 = d32 = None
            if zero_copy:
                src32 = torch.empty(G * per, dtype=torch.int32, device=device)
                dst32 = torch.empty(G * per, dtype=torch.int32, device=device)
                s32, d32 = src32[r * per:], dst32[r * per:]
            if zero_copy and _env_int('SS_B200_DIST_STREAM', 0) and e_hi - e_lo >= INGEST_MIN_EDGES:
                # EXPERIMENTAL, opt-in, unmeasured: this rank's slice through the DMA staging ring of the
                # single-GPU build (55.6 GB/s) instead of in-place reads by the SMs (42-48 GB/s)
                ring = _streamed_degree_pass(ei, e_hi - e_lo, 0, num_nodes, s32, d32, stats, ws, device, e_lo=e_lo)
                check(lib.ss_csr_rowptr_finish(0, 0, num_nodes, _ptr(rowptr_g), _ptr(stats), _ptr(ws), ws.numel(), st),
                      'ss_csr_rowptr_finish')
                del ring  # freed in stream order: every chunk was consumed by a kernel enqueued above
            else:
                check(lib.ss_csr_rowptr(_ptr(src[e_lo:e_hi]), _ptr(dst[e_lo:e_hi]), e_hi - e_lo, 0, 0, num_nodes,
                                        _ptr(rowptr_g), _ptr(s32), _ptr(d32), _ptr(stats), _ptr(ws), ws.numel(), st),
                      'ss_csr_rowptr')
            dist.all_reduce(rowptr_g, op=dist.ReduceOp.SUM, group=self.group)
            ext = torch.stack([stats[0], -stats[3]])
            dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=self.group)
            max_id, min_id = int(ext[0]), -int(ext[1])
            n_loops = max_id + 1
            if 0 <= max_id < num_nodes:  # self loop of node i < n_loops adds 1 to every prefix entry above i
                rowptr_g += torch.arange(num_nodes + 1, device=device, dtype=torch.int64).clamp_(max=n_loops)
            if zero_copy:
                dist.all_gather_into_tensor(src32, src32[r * per:(r + 1) * per].clone(), group=self.group)
                dist.all_gather_into_tensor(dst32, dst32[r * per:(r + 1) * per].clone(), group=self.group)
        else:
            if zero_copy and n_edges:  # host-resident edge list: keep 32-bit device copies for the fill pass
                src32 = torch.empty(n_edges, dtype=torch.int32, device=device)
                dst32 = torch.empty(n_edges, dtype=torch.int32, device=device)
            check(lib.ss_csr_rowptr(_ptr(src), _ptr(dst), n_edges, -1, 0, num_nodes, _ptr(rowptr_g), _ptr(src32),
                                    _ptr(dst32), _ptr(stats), _ptr(ws), ws.numel(), st), 'ss_csr_rowptr')
            max_id, _, n_loops, min_id = (int(v) for v in stats.tolist())
        if max_id >= num_nodes or (n_edges and min_id < 0):
            raise IndexError(f'edge_index refers to node {max_id if max_id >= num_nodes else min_id} but num_nodes '
                             f'is {num_nodes}')
        bounds = balanced_bounds(rowptr_g, self.world_size, default_row_weight(self.world_size, self.exchange),
                                 self.shares)
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        rowptr = (rowptr_g[lo:hi + 1] - rowptr_g[lo]).contiguous()
        nnz = int(rowptr[-1]) if hi > lo else 0
        colidx = torch.empty(max(nnz, 4), dtype=torch.int32, device=device)
        if hi > lo:
            check(lib.ss_csr_fill(_ptr(src), _ptr(dst), _ptr(src32), _ptr(dst32), n_edges, n_loops, None, lo, hi - lo,
                                  _ptr(rowptr), _ptr(colidx), _ptr(ws), ws.numel(), st), 'ss_csr_fill')
        return rowptr, colidx, nnz, bounds

    def _update_shares(self, device):
        """turn the merge times of the previous build into new cost shares (one tiny all-gather)"""
        if not self.adaptive or self.world_size == 1 or not self._merge_events:
            return
        ms = sum(s.elapsed_time(e) for s, e in self._merge_events)  # events of a finished build: no stall
        self._merge_events = []
        mine = torch.tensor([ms], device=device, dtype=torch.float64)
        allms = torch.empty(self.world_size, device=device, dtype=torch.float64)
        dist.all_gather_into_tensor(allms, mine, group=self.group)
        t = allms.clamp_(min=1e-3).tolist()
        old = self.shares or [1.0 / self.world_size] * self.world_size
        # throughput of rank r = share_r / t_r; next shares proportional to it (damped)
        speed = [o / x for o, x in zip(old, t)]
        tot = sum(speed)
        new = [0.3 * o + 0.7 * (v / tot) for o, v in zip(old, speed)]
        tot = sum(new)
        self.shares = [v / tot for v in new]

    def build_hash_tables(self, num_nodes, edge_index):
        eh, r = self.eh, self.rank
        _lib.require_cuda()
        device = edge_index.device if edge_index.is_cuda else torch.device('cuda', torch.cuda.current_device())
        K = eh.max_hops
        with torch.cuda.device(device):
            self._update_shares(device)
            rb = eh._record_bytes()
            symm = None
            if self.exchange in ('auto', 'p2p', 'mc'):
                symm = self._symmetric_buffers(num_nodes, K, rb, device)
                if symm is None and self.exchange in ('p2p', 'mc'):
                    raise RuntimeError(f'symmetric memory is unavailable: {self.exchange_error}')
                if symm is None:
                    self.exchange = 'nccl'
                elif self.exchange == 'auto':
                    # measured on 8 x B200 (R-MAT 24): multicast 19.8 ms / hop vs 21.3 ms unicast; equal at 2 GPUs
                    has_mc = all(int(h.multicast_ptr) for h in list(symm[3]) + [symm[4]])
                    self.exchange = 'mc' if (has_mc and self.world_size > 2) else 'p2p'
                if self.exchange == 'mc' and not all(int(h.multicast_ptr) for h in list(symm[3]) + [symm[4]]):
                    raise RuntimeError('this system has no NVSwitch multicast support for symmetric memory')
            ev = eh._event_begin(device)
            rowptr, colidx, nnz, bounds = self._local_csr(edge_index, num_nodes, device)
            eh._event_end('csr_build', ev, device)
            self.bounds, self.local_nnz = bounds, nnz
            lo, hi = bounds[r], bounds[r + 1]
            rec0 = torch.empty((num_nodes, rb), dtype=torch.uint8, device=device)
            ev = eh._event_begin(device)
            eh._init_records(num_nodes, device, out=rec0)  # hop 0 is cheap: computed redundantly, no exchange
            eh._event_end('init_records', ev, device)
            ws = None
            if symm is not None:
                _, srecs, cards, hdls, chdl = symm
                recs = [rec0] + list(srecs)
                others = [q for q in range(self.world_size) if q != r]
                hdls[0].barrier()  # no peer still reads these buffers from an earlier build
                for k in range(1, K + 1):
                    if hi > lo:
                        mc_rec = mc_cards = 0
                        peer_recs = peer_cards = None
                        if self.exchange == 'mc':
                            mc_rec = int(hdls[k - 1].multicast_ptr) + lo * rb
                            mc_cards = int(chdl.multicast_ptr) + (lo * K + (k - 1)) * 4
                        else:
                            peer_recs = [int(hdls[k - 1].buffer_ptrs[q]) + lo * rb for q in others]
                            peer_cards = [int(chdl.buffer_ptrs[q]) + (lo * K + (k - 1)) * 4 for q in others]
                        t0 = torch.cuda.Event(enable_timing=True)
                        t0.record()
                        ws = eh._merge(rowptr, colidx, nnz, recs[k - 1], recs[k][lo:hi], cards[lo:hi, k - 1], device,
                                       ws, peer_recs, peer_cards, mc_rec, mc_cards)
                        t1 = torch.cuda.Event(enable_timing=True)
                        t1.record()
                        self._merge_events.append((t0, t1))
                    ev = eh._event_begin(device)
                    hdls[k - 1].barrier()  # every rank's launch (and its peer stores) has completed
                    eh._event_end('exchange', ev, device)
            else:
                recs = [rec0] + [torch.empty((num_nodes, rb), dtype=torch.uint8, device=device) for _ in range(K)]
                cards = torch.zeros((num_nodes, K), dtype=torch.float32, device=device)
                for k in range(1, K + 1):
                    if hi > lo:
                        t0 = torch.cuda.Event(enable_timing=True)
                        t0.record()
                        ws = eh._merge(rowptr, colidx, nnz, recs[k - 1], recs[k][lo:hi], cards[lo:hi, k - 1], device,
                                       ws)
                        t1 = torch.cuda.Event(enable_timing=True)
                        t1.record()
                        self._merge_events.append((t0, t1))
                    ev = eh._event_begin(device)
                    exchange_blocks(recs[k], bounds, self.group)
                    eh._event_end('exchange', ev, device)
                exchange_blocks(cards, bounds, self.group)
            tables = SketchTables({k: HopSketch(recs[k], eh.num_perm, eh.p, device) for k in range(K + 1)},
                                  eh.num_perm, eh.p)
            return tables, cards

    def get_subgraph_features(self, links, hash_table, cards, batch_size=11000000):
        """features of this rank's slice of `links` (rows link_slice(len(links), G, rank))"""
        lo, hi = link_slice(links.shape[0], self.world_size, self.rank)
        return self.eh.get_subgraph_features(links[lo:hi], hash_table, cards, batch_size)
