"""Node-sharded multi-GPU build of the sketch tables (one process per GPU, torch.distributed).

The path shards naturally by destination node: rank r owns a contiguous block of rows of every
hop table and the CSR rows of those destinations.  One hop = local k-hop merge of the owned rows (it gathers
previous-hop rows of arbitrary nodes, writes the owned block of the next table) plus ONE exchange step that puts
the finished rows where the next hop (and the pairwise read) will look for them.  Results are bit-identical to
the single-GPU engine because min / max merges do not depend on the partition.  The reference has no
distributed code at all (src/hashing.py is single process).

Row blocks are balanced by COST, not by row count: on power-law graphs the low ids are the hubs (equal row
blocks gave rank 0 74 % of the edges of an R-MAT-24 graph at 2 ranks), so the block boundaries are quantiles of
cost(block) = neighbours + row_weight * rows, then adapted from the merge times each rank measured.

Exchange modes (`exchange=`)
  'halo' (what 'auto' picks beyond 4 GPUs): the hop tables live in symmetric memory
         (torch.distributed._symmetric_memory: every rank's buffer is mapped into every process over NVLink) and the
         merge kernel itself stores each finished row into the tables of exactly those peers whose neighbour lists
         read it (ss_khop_merge_ex `peer_mask`; the mask comes from one all-gather of per-rank "rows I read" byte
         maps at CSR time).  Rows nobody else reads stay local, and the LAST hop -- which no later hop gathers
         from -- is not exchanged at all: get_subgraph_features reads a record it does not hold from its owner's
         table through the same mapping (ss_link_features_sharded).  On R-MAT-24 over 8 GPUs that is ~30 % of the
         bytes of a full replication for hops 1..K-1 and none for hop K.  The tables it returns are complete only
         for this engine's own get_subgraph_features (each rank: its block + its halo).
  'p2p'  full replication fused into the merge kernel: one store per peer for every finished row.
  'mc'   like 'p2p', but every store is ONE multimem.st to the NVSwitch multicast address of the buffer.
  'nccl' one torch.distributed broadcast per owner block after the merge kernel (works everywhere).
'p2p' / 'mc' / 'nccl' return fully replicated tables (every rank can index any row, as the reference's API
promises); 'auto' replicates up to 4 GPUs (p2p at 2, multicast at 3-4: measured faster there, see build_hash_tables) and
switches to 'halo' beyond.  No NCCL call sits on the per-hop data path of the three
symmetric-memory modes: hops are separated by a stream-ordered symmetric-memory barrier.

Edge lists that are ordered by source and symmetric (PyG coalesce / to_undirected output, what the reference
hands to build_hash_tables) take the STREAMING CSR: the edges of a row block are one contiguous slice of the list,
found by binary search, so every rank reads only its own slice (over its own PCIe link when the list is in pinned
host memory) -- no histogram, no all-reduce of the row pointer, no pass over the whole list.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import MergeDesc, ShardView
from .hashing import (_FP_KEYS, CSR_FAST_MIN_EDGES, INGEST_MIN_EDGES, ElphHashes, HopSketch, SketchTables, _edge_source,
                      _env_int, _ptr, _stream_chunks, _stream_ptr, _streamed_degree_pass, check, lib)


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


def cumulative_shares(world_size, shares=None):
    """[c_1, ..., c_{G-1}]: fraction of the total cost in front of each cut"""
    if shares is None:
        shares = [1.0 / world_size] * world_size
    tot = float(sum(shares))
    cum, acc = [], 0.0
    for r in range(world_size - 1):
        acc += float(shares[r]) / tot
        cum.append(acc)
    return cum


def balanced_bounds(rowptr, world_size, row_weight=0.0, shares=None):
    """row boundaries [b_0 = 0, ..., b_G = N] such that block r carries the fraction shares[r] (default 1/G)
    of the total COST, cost(block) = neighbours(block) + row_weight * rows(block).
    One neighbour = one 768-byte gather from HBM; a finished row costs one local store plus, in the fused
    exchange, stores over NVLink, so the row term grows with the number of peers and vanishes at 1.
    `rowptr` is the global int64 [N + 1] prefix sum (any device); returns a python list of G + 1 ints."""
    n = rowptr.numel() - 1
    if n <= 0:
        return [0] * (world_size + 1)
    if world_size == 1:
        return [0, n]
    cum = cumulative_shares(world_size, shares)
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
    if exchange == 'halo':  # a row is pushed to ~30 % of the peers on average, and not at all in the last hop
        return 1.5 + 2.0 * (world_size - 1)
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


def halo_masks(marks, bounds, rank):
    """from the all-gathered "rows rank q reads" byte maps [G, N] (0 / 1):
    peer_mask uint8 [hi - lo]: bit i set = the i-th OTHER rank (ascending rank order, own rank skipped) reads my row
    local_rows uint8 [N]: rows whose hops 1..K-1 are valid in my copy (my block + every row I read)"""
    world = marks.shape[0]
    lo, hi = bounds[rank], bounds[rank + 1]
    others = [q for q in range(world) if q != rank]
    mask = torch.zeros(hi - lo, dtype=torch.uint8, device=marks.device)
    for i, q in enumerate(others):
        mask |= (marks[q, lo:hi] != 0).to(torch.uint8) << i
    local = (marks[rank] != 0).to(torch.uint8)
    local[lo:hi] = 1
    return mask, local


class ShardedElphHashes(object):
    """ElphHashes over `world_size` GPUs: same build_hash_tables / get_subgraph_features surface; every rank
    passes the same (replicated) edge_index and link list and gets the tables plus ITS slice of features
    (`link_slice`).  reuse_buffers (default): the symmetric-memory tables are allocated once per shape and a
    later build REUSES them -- tables returned by the earlier build are then invalid and raise when touched;
    reuse_buffers=False allocates fresh symmetric buffers for every build (a collective rendezvous each time)."""

    def __init__(self, args, group=None, exchange='auto', reuse_buffers=True, **kw):
        assert exchange in ('auto', 'halo', 'p2p', 'mc', 'nccl')
        self.eh = ElphHashes(args, **kw)
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bounds = None
        self.local_nnz = None
        self.exchange = exchange
        self.reuse_buffers = reuse_buffers
        self._symm = None       # cached symmetric buffers: (key, recs[1..K], cards, handles, cards handle)
        self._lease = None      # [True] while the tables of the last build own the cached buffers
        self._shard = None      # ShardView of the last 'halo' build (+ tensors it points at)
        # adaptive balance: the share of the total cost each rank gets follows its measured merge throughput
        # in the previous build (ranks differ in L2 hit rate and NVLink egress, which no static model captures)
        self.adaptive = True
        self.shares = None
        self._merge_events = []
        self.exchange_error = None
        self.csr_path = None    # 'streaming' | 'histogram' (which CSR build the last call took)
        self._halo_mask = None
        self._fp_keys = None
        self._start_init = None
        self._bounds_dev = None
        if self.world_size == 1 or self.world_size - 1 > 7:
            self.exchange = 'nccl'

    @property
    def halo_fraction(self):
        """pushed (row, peer) pairs / all pairs of the last 'halo' build on this rank (synchronises)"""
        m = self._halo_mask
        if m is None or m.numel() == 0 or self.world_size < 2:
            return None
        pop = torch.tensor([bin(i).count('1') for i in range(256)], device=m.device)
        return float(pop[m.long()].sum()) / (m.numel() * (self.world_size - 1))

    # ------------------------------------------------------------------ symmetric-memory tables
    def _symmetric_buffers(self, num_nodes, K, rb, device):
        """hop tables 1..K and cards in symmetric memory (allocated once per shape, reused by later builds);
        returns None (on every rank consistently) when symmetric memory is unavailable"""
        key = (num_nodes, K, rb, str(device))
        if self.reuse_buffers and self._symm is not None and self._symm[0] == key:
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

    # ------------------------------------------------------------------ CSR of the owned row block
    def _streaming_csr(self, ei, zero_copy, num_nodes, device):
        """row blocks + this rank's CSR from a list ordered by source and symmetric, or None (consistently on every
        rank) when the list is not: see the module docstring"""
        G, r = self.world_size, self.rank
        n_edges = ei.shape[1]
        st = _stream_ptr(device)
        if self._fp_keys is None:
            # the fingerprints of all slices are summed: every rank must hash with the SAME (rank 0's random) keys
            k = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in _FP_KEYS], dtype=torch.int64, device=device)
            dist.broadcast(k, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            self._fp_keys = tuple(int(v) & ((1 << 64) - 1) for v in k.tolist())
        eh = self.eh
        ev = eh._event_begin(device)
        cum = torch.tensor(cumulative_shares(G, self.shares) or [0.5], dtype=torch.float64, device=device)
        cuts = torch.empty(2 * (G + 1), dtype=torch.int64, device=device)
        row_cost = 1.0 + default_row_weight(G, self.exchange)  # + 1: every row carries its self loop
        check(lib.ss_csr_sorted_bounds(_ptr(ei[0]), n_edges, num_nodes, row_cost, _ptr(cum), G - 1, _ptr(cuts),
                                       _ptr(cuts[G + 1:]), st), 'ss_csr_sorted_bounds')
        cl = [int(v) for v in cuts.tolist()]  # host read #1: sizes of the blocks
        eh._event_end('csr.bounds', ev, device)
        if self._start_init is not None:
            self._start_init()
        ev = eh._event_begin(device)
        bounds, eoff = cl[:G + 1], cl[G + 1:]
        self._bounds_dev = cuts[:G + 1]  # cuts of an increasing cost function: already monotone
        lo, hi = bounds[r], bounds[r + 1]
        e_lo, e_hi = eoff[r], eoff[r + 1]
        n_loc, rows = e_hi - e_lo, hi - lo
        cap = n_loc + rows
        colidx = torch.empty(max(cap, 4), dtype=torch.int32, device=device)
        rowptr = torch.empty(rows + 1, dtype=torch.int64, device=device)
        st12 = torch.empty(12, dtype=torch.int64, device=device)
        carry = torch.zeros(2, dtype=torch.int64, device=device)

        def chunk(k_ptr, v_ptr, count, e_base):
            check(lib.ss_csr_sorted_chunk_rows(k_ptr, v_ptr, count, e_base, e_lo, lo, rows, 1, cap, self._fp_keys[0],
                                               self._fp_keys[1], _ptr(rowptr), _ptr(colidx), _ptr(st12), _ptr(carry),
                                               _stream_ptr(device)), 'ss_csr_sorted_chunk_rows')

        ring = None
        if zero_copy and n_loc >= INGEST_MIN_EDGES:
            ring = _stream_chunks(ei, n_loc, device, lambda b0, b1, a, b, c: chunk(_ptr(b0), _ptr(b1), b - a, e_lo + a),
                                  e_lo=e_lo)
        else:
            chunk(_ptr(ei[0, e_lo:e_hi]), _ptr(ei[1, e_lo:e_hi]), n_loc, e_lo)
        eh._event_end('csr.stream', ev, device)
        ev = eh._event_begin(device)
        # the verdict is global: order violations / range errors / fingerprints summed, id range max / min over ranks
        sums = st12[4:11].clone()
        ext = torch.stack([st12[0], -st12[3]])
        # the order check inside a slice misses the pair that straddles two slices: add it here
        if r > 0 and n_loc > 0 and e_lo > 0:
            prev_key = ei[0, e_lo - 1:e_lo].to(device)
            sums[4] += (ei[0, e_lo:e_lo + 1].to(device) < prev_key).long().sum()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=self.group)
        st12[4:11] = sums
        st12[0], st12[3] = ext[0], -ext[1]
        check(lib.ss_csr_sorted_finish_rows(n_loc, lo, rows, 1, cap, _ptr(rowptr), _ptr(colidx), _ptr(st12), _ptr(carry),
                                            st), 'ss_csr_sorted_finish_rows')
        s = [int(v) for v in st12.tolist()]  # host read #2: the verdict (identical on every rank)
        eh._event_end('csr.verdict', ev, device)
        del ring
        max_id, min_id = s[0], s[3]
        ok = s[8] == 0 and s[10] == 0 and (s[4], s[5]) == (s[6], s[7]) and max_id < num_nodes and (n_edges == 0 or min_id >= 0)
        if _env_int('SS_B200_DEBUG', 0):
            print(f'[rank {r}] streaming csr: bounds={bounds} eoff={eoff} stats={s} ok={ok}', flush=True)
        if not ok:
            return None
        return rowptr, colidx, s[1], bounds

    def _local_csr(self, edge_index, num_nodes, device):
        """(rowptr, colidx, nnz, bounds): this rank's CSR rows, the row blocks of all ranks"""
        ei, zero_copy = _edge_source(edge_index, device)
        n_edges = ei.shape[1]
        G, r = self.world_size, self.rank
        if (G > 1 and _env_int('SS_B200_CSR_FAST', 1) and n_edges >= CSR_FAST_MIN_EDGES and num_nodes > 0):
            got = self._streaming_csr(ei, zero_copy, num_nodes, device)
            if got is not None:
                self.csr_path = 'streaming'
                return got
        self.csr_path = 'histogram'
        src, dst = ei[0], ei[1]
        ws_bytes = check(lib.ss_csr_workspace_bytes(num_nodes), 'ss_csr_workspace_bytes')
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=device)
        rowptr_g = torch.empty(num_nodes + 1, dtype=torch.int64, device=device)
        stats = torch.empty(4, dtype=torch.int64, device=device)
        src32 = dst32 = None
        st = _stream_ptr(device)
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
            if zero_copy and e_hi - e_lo >= INGEST_MIN_EDGES:
                # this rank's slice through the DMA staging ring of the single-GPU build
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
        """turn the merge times of the previous build into new cost shares (one tiny all-gather).  Every rank must
        take the same branch: whether the events of the previous build have completed is itself agreed upon."""
        if not self.adaptive or self.world_size == 1:
            return
        ready = bool(self._merge_events) and all(e.query() for _, e in self._merge_events)
        ms = sum(s.elapsed_time(e) for s, e in self._merge_events) if ready else -1.0
        self._merge_events = []
        mine = torch.tensor([ms], device=device, dtype=torch.float64)
        allms = torch.empty(self.world_size, device=device, dtype=torch.float64)
        dist.all_gather_into_tensor(allms, mine, group=self.group)
        t = allms.tolist()
        if min(t) <= 0.0:  # some rank has nothing to report (first build, events not finished): keep the shares
            return
        old = self.shares or [1.0 / self.world_size] * self.world_size
        # throughput of rank r = share_r / t_r; next shares proportional to it (damped)
        speed = [o / max(x, 1e-3) for o, x in zip(old, t)]
        tot = sum(speed)
        new = [0.3 * o + 0.7 * (v / tot) for o, v in zip(old, speed)]
        tot = sum(new)
        self.shares = [v / tot for v in new]

    # ------------------------------------------------------------------ build
    def build_hash_tables(self, num_nodes, edge_index):
        eh, r = self.eh, self.rank
        _lib.require_cuda()
        device = edge_index.device if edge_index.is_cuda else torch.device('cuda', torch.cuda.current_device())
        K = eh.max_hops
        G = self.world_size
        with torch.cuda.device(device):
            self._update_shares(device)
            rb = eh._record_bytes()
            symm = None
            if self.exchange in ('auto', 'halo', 'p2p', 'mc'):
                symm = self._symmetric_buffers(num_nodes, K, rb, device)
                if symm is None and self.exchange in ('halo', 'p2p', 'mc'):
                    raise RuntimeError(f'symmetric memory is unavailable: {self.exchange_error}')
                if symm is None:
                    self.exchange = 'nccl'
                elif self.exchange == 'auto':
                    # measured on R-MAT 24, K=3, 20 M links (profiles/r02_bench_*gpu.json), ms per step:
                    #   GPUs   replicate (p2p / mc)   halo
                    #    2        108.2 (p2p)         138.5    one peer: nearly every row is in its halo anyway, and the
                    #    4         80.0 (mc)           83.9    pairwise kernel's remote reads cost more than replication
                    #    8         75.7 (mc, round 1)  54.0    from 8 GPUs on every GPU ingesting the whole table loses
                    has_mc = all(int(h.multicast_ptr) for h in list(symm[3]) + [symm[4]])
                    if eh.num_perm == 128 and eh.p == 8 and self.world_size > 4:
                        self.exchange = 'halo'
                    else:
                        self.exchange = 'mc' if (has_mc and self.world_size > 2) else 'p2p'
                if self.exchange == 'mc' and not all(int(h.multicast_ptr) for h in list(symm[3]) + [symm[4]]):
                    raise RuntimeError('this system has no NVSwitch multicast support for symmetric memory')
            if self._lease is not None:
                self._lease[0] = False  # the tables of the previous build are about to be overwritten
            lease = [True] if (symm is not None and self.reuse_buffers) else None
            self._lease = lease
            main = torch.cuda.current_stream(device)
            # hop 0 does not depend on the graph: on a side stream, under the CSR build.  It is launched from INSIDE the
            # CSR build, after the row-block cuts were read back: its persistent grid fills every SM for ~4 ms, and a
            # one-block kernel the host is waiting for (the cuts) would otherwise queue behind it
            rec0 = torch.empty((num_nodes, rb), dtype=torch.uint8, device=device)
            side = eh._side_stream(device)
            init_state = {}

            def start_init():
                if init_state:
                    return
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ev_i = eh._event_begin(device)
                    eh._init_records(num_nodes, device, out=rec0)  # cheap: computed redundantly, never exchanged
                    eh._event_end('init_records', ev_i, device)
                    init_state['done'] = torch.cuda.Event()
                    init_state['done'].record(side)

            self._start_init = start_init
            ev = eh._event_begin(device)
            try:
                rowptr, colidx, nnz, bounds = self._local_csr(edge_index, num_nodes, device)
            except BaseException:
                self._start_init = None
                start_init()
                main.wait_event(init_state['done'])  # rec0 must not be recycled under the side stream
                raise
            self._start_init = None
            start_init()  # (the histogram path never called it)
            self.bounds, self.local_nnz = bounds, nnz
            lo, hi = bounds[r], bounds[r + 1]
            others = [q for q in range(G) if q != r]
            peer_mask = zero_mask = local_rows = None
            ev_h = eh._event_begin(device)
            if self.exchange == 'halo' and symm is not None and self.csr_path == 'streaming':
                # the list was verified symmetric: who reads my rows follows from my own CSR rows, no exchange
                peer_mask = torch.zeros((hi - lo + 7) // 4 * 4, dtype=torch.uint8, device=device)[:hi - lo]
                local_rows = torch.zeros(num_nodes, dtype=torch.uint8, device=device)
                check(lib.ss_halo_from_csr(_ptr(rowptr), _ptr(colidx), hi - lo, nnz, _ptr(self._bounds_dev), G, r,
                                           _ptr(peer_mask), _ptr(local_rows), _stream_ptr(device)), 'ss_halo_from_csr')
                local_rows[lo:hi] = 1
                zero_mask = torch.zeros_like(peer_mask)
            elif self.exchange == 'halo' and symm is not None:
                marks = torch.zeros((G, num_nodes), dtype=torch.uint8, device=device)
                mine = torch.zeros(num_nodes, dtype=torch.uint8, device=device)
                check(lib.ss_mark_rows(_ptr(colidx), nnz, _ptr(mine), _stream_ptr(device)), 'ss_mark_rows')
                dist.all_gather_into_tensor(marks.view(-1), mine, group=self.group)
                peer_mask, local_rows = halo_masks(marks, bounds, r)
                zero_mask = torch.zeros_like(peer_mask)
                del marks, mine
            eh._event_end('csr.halo', ev_h, device)
            main.wait_event(init_state['done'])
            eh._event_end('csr_build', ev, device)
            ws = None
            if symm is not None:
                _, srecs, cards, hdls, chdl = symm
                recs = [rec0] + list(srecs)
                hdls[0].barrier()  # no peer still reads these buffers from an earlier build
                for k in range(1, K + 1):
                    if hi > lo:
                        t0 = torch.cuda.Event(enable_timing=True)
                        t0.record()
                        ws = self._merge_exchange(k, K, rowptr, colidx, nnz, recs, cards, lo, hi, rb, hdls, chdl, others,
                                                  peer_mask, zero_mask, ws, device)
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
            self._shard = None
            if self.exchange == 'halo' and symm is not None:
                sv = ShardView()
                sv.n_ranks, sv.rank, sv.last_hop_own_only = G, r, 1
                for q in range(G + 1):
                    sv.bounds[q] = bounds[q]
                for k in range(1, K + 1):
                    for q in range(G):
                        sv.peer_records[k][q] = int(hdls[k - 1].buffer_ptrs[q])
                sv.local_rows = local_rows.data_ptr()
                self._shard = (sv, local_rows, recs)
                self._halo_mask = peer_mask
            tables = SketchTables({k: HopSketch(recs[k], eh.num_perm, eh.p, device, lease=lease if k else None)
                                   for k in range(K + 1)}, eh.num_perm, eh.p)
            tables.sharding = dict(exchange=self.exchange, bounds=list(bounds), rank=r,
                                   complete=self.exchange != 'halo')
            return tables, cards

    def _merge_exchange(self, k, K, rowptr, colidx, nnz, recs, cards, lo, hi, rb, hdls, chdl, others, peer_mask, zero_mask,
                        ws, device):
        """hop k over the owned rows, finished rows stored wherever the exchange mode wants them"""
        eh = self.eh
        if self.exchange != 'halo':
            mc_rec = mc_cards = 0
            peer_recs = peer_cards = None
            if self.exchange == 'mc':
                mc_rec = int(hdls[k - 1].multicast_ptr) + lo * rb
                mc_cards = int(chdl.multicast_ptr) + (lo * K + (k - 1)) * 4
            else:
                peer_recs = [int(hdls[k - 1].buffer_ptrs[q]) + lo * rb for q in others]
                peer_cards = [int(chdl.buffer_ptrs[q]) + (lo * K + (k - 1)) * 4 for q in others]
            return eh._merge(rowptr, colidx, nnz, recs[k - 1], recs[k][lo:hi], cards[lo:hi, k - 1], device, ws, peer_recs,
                             peer_cards, mc_rec, mc_cards)
        need = check(lib.ss_merge_workspace_bytes(nnz, eh.num_perm, eh.p), 'ss_merge_workspace_bytes')
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 16), dtype=torch.uint8, device=device)
        n_peers = len(others)
        pr = (ctypes.c_void_p * n_peers)(*[int(hdls[k - 1].buffer_ptrs[q]) + lo * rb for q in others])
        pc = (ctypes.c_void_p * n_peers)(*[int(chdl.buffer_ptrs[q]) + (lo * K + (k - 1)) * 4 for q in others])
        hc = eh._consts(device)['hc']
        out, cards_col = recs[k][lo:hi], cards[lo:hi, k - 1]
        d = MergeDesc()
        d.rowptr, d.colidx, d.n_rows, d.nnz = _ptr(rowptr).value, _ptr(colidx).value, hi - lo, nnz
        d.rec_in, d.in_rows, d.in_stride = recs[k - 1].data_ptr(), recs[k - 1].shape[0], recs[k - 1].stride(0)
        d.rec_out, d.out_stride = out.data_ptr(), out.stride(0)
        d.num_perm, d.hll_p, d.layout, d.variant = eh.num_perm, eh.p, _lib.SS_LAYOUT_FULL, _lib.MERGE_VARIANTS[eh.merge_variant]
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        d.cards_out, d.cards_stride, d.hc = cards_col.data_ptr(), cards_col.stride(0), ctypes.addressof(hc)
        d.n_peers = n_peers
        d.peer_rec_out = ctypes.cast(pr, ctypes.c_void_p).value
        d.peer_cards_out = ctypes.cast(pc, ctypes.c_void_p).value
        # the last hop is read by the pairwise kernel only, which fetches what it lacks from the owner: push nothing
        d.peer_mask = (zero_mask if k == K else peer_mask).data_ptr()
        ev = eh._event_begin(device)
        check(lib.ss_khop_merge_ex(ctypes.byref(d), _stream_ptr(device)), 'ss_khop_merge_ex')
        eh._event_end('khop_merge', ev, device)
        return ws

    def get_subgraph_features(self, links, hash_table, cards, batch_size=11000000):
        """features of this rank's slice of `links` (rows link_slice(len(links), G, rank))"""
        lo, hi = link_slice(links.shape[0], self.world_size, self.rank)
        shard = None
        if self._shard is not None and getattr(hash_table, 'sharding', {}).get('exchange') == 'halo':
            if dict.__getitem__(hash_table, 1)._records is not self._shard[2][1]:
                raise RuntimeError("tables of a 'halo' build can only be read by the engine that built them, before "
                                   'its next build')
            shard = self._shard[0]
        return self.eh.get_subgraph_features(links[lo:hi], hash_table, cards, batch_size, _shard=shard)

    def owned_rows(self, tables, cards):
        """this rank's own block of every hop table and of the cardinalities (the part of a 'halo' build that is
        authoritative): [(hop, uint8 [rows, record_bytes])], float32 [rows, K]"""
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        return [(k, tables.records(k)[lo:hi]) for k in sorted(tables.keys())], cards[lo:hi]
