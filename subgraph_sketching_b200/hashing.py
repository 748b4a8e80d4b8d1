"""B200-native drop-in for /root/reference/src/hashing.py (`ElphHashes`, `MinhashPropagation`,
`HllPropagation`, `LABEL_LOOKUP`).

Same names, arguments, return types and error behaviour as the reference class; every array computation is
a hand-written sm_100a kernel in libss_b200.so reached through the ctypes C ABI (include/ss_b200.h).  torch
is used for device memory, streams and host<->device copies only.  There is no CPU fallback: the module
raises if the library is missing or no CUDA device is visible.

Device semantics (reference: hashing.py is device-agnostic torch)
  * CUDA inputs  -> kernels run on that device / the current stream, outputs stay there (the ELPH path,
    models/elph.py:190-213, train.py:199-204)
  * CPU inputs   -> transparent offload: pinned H2D copy, kernels, D2H copy; the result lives on the input's
    device exactly as in the reference (the BUDDY preprocessing path, datasets/elph.py:85,200,207)

Internal layout: one compact 768-byte record per node per hop ([128 x uint32 MinHash | 256 x uint8 HLL]).
`build_hash_tables` returns a `SketchTables` mapping that holds the records on the GPU and materialises the
reference's int64 / int8 tensors only when `table[k]['minhash']` / `['hll']` is indexed; it pickles
(`torch.save`) as the reference's plain dict-of-dict of CPU tensors, so on-disk caches interoperate.
"""
from __future__ import annotations

import ctypes
import logging
import os
from time import time

import numpy as np
import torch

from . import _lib
from ._lib import HllConsts, HopView, check, lib

logger = logging.getLogger(__name__)
logger.setLevel(logging.INFO)

# feature index -> (hops from u, hops from v), keyed by max hops  (hashing.py:22-25)
LABEL_LOOKUP = {1: {0: (1, 1), 1: (0, 1), 2: (1, 0)},
                2: {0: (1, 1), 1: (2, 1), 2: (1, 2), 3: (2, 2), 4: (0, 1), 5: (1, 0), 6: (0, 2), 7: (2, 0)},
                3: {0: (1, 1), 1: (2, 1), 2: (1, 2), 3: (2, 2), 4: (3, 1), 5: (1, 3), 6: (3, 2), 7: (2, 3), 8: (3, 3),
                    9: (0, 1), 10: (1, 0), 11: (0, 2), 12: (2, 0), 13: (0, 3), 14: (3, 0)}}

_TABLES_PATH = os.environ.get('SS_B200_HLLPP_TABLES') or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), 'data', 'hllpp_tables.npz')


def _env_int(name, default):
    v = os.environ.get(name)
    return int(v) if v not in (None, '') else default


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _cuda_device(t=None):
    """device the kernels run on for input `t` (its own device if CUDA, else the current CUDA device)"""
    _lib.require_cuda()
    if t is not None and t.is_cuda:
        return t.device
    return torch.device('cuda', torch.cuda.current_device())


def _to_device(t, device):
    if t.device == device:
        return t
    if t.device.type == 'cpu' and t.numel() > (1 << 16) and not t.is_pinned():
        try:
            t = t.pin_memory()
        except RuntimeError:  # pinning can fail on tiny / exotic hosts; the pageable copy is still correct
            pass
    return t.to(device, non_blocking=True)


def _to_host(t):
    """device -> pinned host copy (torch caches pinned blocks, so steady-state calls do not re-pin)"""
    if not t.is_cuda:
        return t
    try:
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:
        return t.cpu()
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out


_warned_tables = set()


def hllpp_tables(p, with_source=False):
    """(threshold, raw_estimate[T], bias[T]) for precision p: from `datasketch` when it is importable
    (what the reference reads, hashing.py:77-80), else from the packaged Monte-Carlo tables
    (tools/gen_hllpp_tables.py) or the file named by SS_B200_HLLPP_TABLES.  The packaged tables are NOT the
    published HLL++ constants: cardinalities between the threshold and 5m then differ from a reference run that
    has `datasketch`, so that fallback warns once per precision (set SS_B200_HLLPP_TABLES or pass `hll_tables`
    to choose tables explicitly).  with_source=True also returns where the tables came from."""
    try:
        from datasketch import hyperloglog_const as hc
        out = (hc._thresholds[p - 4], list(hc._raw_estimate[p - 4]), list(hc._bias[p - 4]))
        src = 'datasketch'
    except ImportError:
        blob = np.load(_TABLES_PATH)
        out = (int(blob['thresholds'][p - 4]), blob[f'raw_estimate_p{p}'], blob[f'bias_p{p}'])
        explicit = bool(os.environ.get('SS_B200_HLLPP_TABLES'))
        src = f'file:{_TABLES_PATH}' if explicit else 'packaged-monte-carlo'
        if not explicit and p not in _warned_tables:
            _warned_tables.add(p)
            import warnings
            warnings.warn(
                f'datasketch is not importable: using the packaged Monte-Carlo HLL++ bias tables for p={p} '
                f'({_TABLES_PATH}).  They are substitutes for datasketch.hyperloglog_const; estimates in the '
                'bias-corrected regime (threshold < e <= 5m) and caches derived from them differ from a reference '
                'run that has datasketch.  Install datasketch, set SS_B200_HLLPP_TABLES, or pass hll_tables=.',
                RuntimeWarning, stacklevel=3)
    return out + (src,) if with_source else out


def hll_alpha(p):
    """HyperLogLog alpha_m (datasketch.HyperLogLogPlusPlus.alpha; hashing.py:72)"""
    if p == 4:
        return 0.673
    if p == 5:
        return 0.697
    if p == 6:
        return 0.709
    return 0.7213 / (1.0 + 1.079 / (1 << p))


def log2_window_table():
    """int32[64]: entry k = largest j >= 0 with ceil(log2(float64(2^k + j))) == k, i.e. how far above a power
    of two the reference's float64 `_np_bit_length` (hashing.py:83-89) still rounds DOWN.  Evaluated with
    numpy itself so that the device rank computation reproduces the reference bit for bit."""
    out = np.zeros(64, dtype=np.int32)
    for k in range(1, 63):
        base = 1 << k

        def rounds_down(j):
            return int(np.ceil(np.log2(np.array([base + j], dtype=np.uint64)))[0]) == k

        if not rounds_down(1):
            continue
        lo, hi = 1, base  # rounds_down(lo) holds; 2^k + base = 2^(k+1) never rounds down
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if rounds_down(mid):
                lo = mid
            else:
                hi = mid
        out[k] = min(lo, np.iinfo(np.int32).max)
    return out


def _edge_source(edge_index, device):
    """(tensor whose storage the CSR kernels read, zero_copy flag).  A pinned host edge_index is NOT copied:
    the kernels read it in place over PCIe with coalesced loads (UVA), so the list crosses the bus once and
    never occupies device memory; pageable host tensors and other dtypes are copied to the device."""
    ei = edge_index
    if ei.dim() != 2 or ei.shape[0] != 2:
        raise ValueError('edge_index must be [2, n_edges]')
    if ei.device.type == 'cpu' and ei.dtype == torch.int64 and ei.is_contiguous() and ei.is_pinned():
        return ei, True
    ei = ei.to(device, non_blocking=True) if ei.device != device else ei
    if ei.dtype != torch.int64:
        ei = ei.long()
    return ei.contiguous(), False


# a pinned host edge list of at least this many edges is STREAMED: DMA copies of INGEST_CHUNK edges into a
# two-slot device staging ring, each chunk histogrammed as it lands (copy engine 55.6 GB/s, against 42-48 GB/s
# for in-place reads of the same list by the SMs -- tools/exp_h2d.py)
INGEST_MIN_EDGES = 1 << 23
INGEST_CHUNK = 1 << 25
_ingest_streams = {}


def _ingest_stream(device):
    key = str(device)
    if key not in _ingest_streams:
        _ingest_streams[key] = torch.cuda.Stream(device=device)
    return _ingest_streams[key]


# streaming CSR of key-ordered lists (ss_csr_sorted_chunk): tried first on lists of at least this many edges
CSR_FAST_MIN_EDGES = 1 << 12
_FP_KEYS = tuple(int.from_bytes(os.urandom(8), 'little') for _ in range(2))  # per-process fingerprint keys


def _sorted_stats_ok(s, n_edges, num_rows, symmetric_needed):
    """accept the speculative streaming CSR?  s = stats_io of ss_csr_sorted_chunk / _finish as a python list"""
    if s[8] != 0 or s[10] != 0 or s[0] >= num_rows or (n_edges and s[3] < 0):
        return False
    return (not symmetric_needed) or (s[4], s[5]) == (s[6], s[7])


def _try_sorted_csr(src, dst, n_edges, num_rows, add_loops, device, streamed_from=None, degree_side=None, chunk_hook=None):
    """speculative one-pass CSR for key-ordered edge lists (coalesced / to_undirected output) -> (rowptr, colidx, nnz,
    max_id) or None when the list is neither (ordered by edge_index[0] AND symmetric) nor ordered by edge_index[1].
    streamed_from: pinned host edge_index to stream through the DMA ring instead of reading src / dst;
    degree_side(buf0, buf1, lo, hi, first): called per landed chunk (the histogram pass of the fallback rides along);
    chunk_hook(rowptr, colidx, capacity, stats, carry, final): called after every absorbed chunk of a streamed list and once
    more (final=True) after the build is complete, BEFORE the verdict is known -- the caller may enqueue speculative
    work on the rows that are complete so far (hop 1 under the ingest stream); chunk_hook(None, ...) reports the verdict."""
    cap = n_edges + (num_rows if add_loops else 0)
    colidx = torch.empty(max(cap, 4), dtype=torch.int32, device=device)
    rowptr = torch.empty(num_rows + 1, dtype=torch.int64, device=device)
    st12 = torch.empty(12, dtype=torch.int64, device=device)
    carry = torch.empty(2, dtype=torch.int64, device=device)
    st = _stream_ptr(device)
    loops = 1 if add_loops else 0
    orientations = ('src', 'dst') if streamed_from is None else ('src',)
    for orient in orientations:
        if streamed_from is None:
            key, val = (src, dst) if orient == 'src' else (dst, src)
            check(lib.ss_csr_sorted_chunk(_ptr(key), _ptr(val), n_edges, 0, num_rows, loops, cap, _FP_KEYS[0], _FP_KEYS[1],
                                          _ptr(rowptr), _ptr(colidx), _ptr(st12), _ptr(carry), st), 'ss_csr_sorted_chunk')
        else:
            ring = _stream_chunks(streamed_from, n_edges, device, lambda b0, b1, lo, hi, c: (
                degree_side(b0, b1, lo, hi, c) if degree_side is not None else None,
                check(lib.ss_csr_sorted_chunk(_ptr(b0), _ptr(b1), hi - lo, lo, num_rows, loops, cap, _FP_KEYS[0],
                                              _FP_KEYS[1], _ptr(rowptr), _ptr(colidx), _ptr(st12), _ptr(carry),
                                              _stream_ptr(device)), 'ss_csr_sorted_chunk'),
                chunk_hook(rowptr, colidx, cap, st12, carry, False) if (chunk_hook is not None and hi < n_edges) else None))
        check(lib.ss_csr_sorted_finish(n_edges, num_rows, loops, cap, _ptr(rowptr), _ptr(colidx), _ptr(st12), _ptr(carry),
                                       st), 'ss_csr_sorted_finish')
        if chunk_hook is not None and streamed_from is not None:
            chunk_hook(rowptr, colidx, cap, st12, carry, True)
        s = [int(v) for v in st12.tolist()]  # the one host synchronisation of the CSR build
        if streamed_from is not None:
            del ring
        accepted = _sorted_stats_ok(s, n_edges, num_rows, symmetric_needed=(orient == 'src'))
        if chunk_hook is not None and streamed_from is not None:
            chunk_hook(None, None, cap, None, None, accepted)  # the verdict
        if accepted:
            return rowptr, colidx, s[1], s[0]
        if not (orient == 'src' and s[9] == 0 and s[10] == 0):
            break  # edge_index[1] is not ordered either
    return None


def _stream_chunks(ei, n_edges, device, consume, e_lo=0):
    """chunked DMA of the edges [e_lo, e_lo + n_edges) of a pinned host edge_index [2, E] into a two-slot device ring on
    the ingest stream; consume(buf_row0, buf_row1, lo, hi, chunk_index) enqueues the kernels of a landed chunk on the
    current stream (lo / hi relative to e_lo).  Returns the ring (keep it alive until the stream has consumed it)."""
    main = torch.cuda.current_stream(device)
    copy = _ingest_stream(device)
    chunk = min(INGEST_CHUNK, n_edges)
    ring = [torch.empty((2, chunk), dtype=torch.int64, device=device) for _ in range(2)]
    copy.wait_stream(main)  # the ring's memory may still be in use by earlier work on this stream
    consumed = [None, None]
    for c, lo in enumerate(range(0, n_edges, chunk)):
        hi = min(lo + chunk, n_edges)
        buf = ring[c & 1]
        with torch.cuda.stream(copy):
            if consumed[c & 1] is not None:
                copy.wait_event(consumed[c & 1])
            buf[0, :hi - lo].copy_(ei[0, e_lo + lo:e_lo + hi], non_blocking=True)  # contiguous row slices: plain DMA
            buf[1, :hi - lo].copy_(ei[1, e_lo + lo:e_lo + hi], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy)
        main.wait_event(ready)
        consume(buf[0], buf[1], lo, hi, c)
        consumed[c & 1] = torch.cuda.Event()
        consumed[c & 1].record(main)
    return ring


def _streamed_degree_pass(ei, n_edges, row_begin, num_rows, src32, dst32, stats, ws, device, e_lo=0):
    """pass 1 of the CSR build over the edges [e_lo, e_lo + n_edges) of a pinned host edge_index [2, E]: chunked
    DMA into a staging ring on the ingest stream, ss_csr_degree_chunk per chunk on the current stream; the
    int32 copies of edge e_lo + i land in src32[i] / dst32[i]"""
    def consume(b0, b1, lo, hi, c):
        check(lib.ss_csr_degree_chunk(_ptr(b0), _ptr(b1), hi - lo, row_begin, num_rows, _ptr(src32[lo:hi]),
                                      _ptr(dst32[lo:hi]), _ptr(stats), _ptr(ws), ws.numel(), 1 if c == 0 else 0,
                                      _stream_ptr(device)), 'ss_csr_degree_chunk')
    return _stream_chunks(ei, n_edges, device, consume, e_lo=e_lo)


def build_csr(edge_index, device, num_rows=None, add_loops=True, row_begin=0, bounds_fn=None, chunk_hook=None):
    """COO edge_index [2, E] -> (rowptr int64 [n_rows+1], colidx int32 [>= nnz], nnz, max_id) keyed by destination
    (PyG flow source -> target, hashing.py:30-35).  With add_loops, a self loop is appended for every node id
    < max(edge_index)+1 (computed on the device), which is add_self_loops(edge_index) without num_nodes
    (hashing.py:148).  One device->host read of the statistics per attempt.  A pinned host edge_index
    never gets a full-size device copy: it is streamed through a staging ring (large lists) or read in place.
    Lists ordered by their CSR key (coalesced / to_undirected output) take the one-pass streaming build
    (ss_csr_sorted_chunk); anything else the histogram + scan + cursor-fill build."""
    ei, zero_copy = _edge_source(edge_index, device)
    n_edges = ei.shape[1]
    if num_rows is None:  # rows = max id + 1: needs the id statistics first
        num_rows = (int(ei.max()) + 1) if n_edges else 0
    src, dst = ei[0], ei[1]
    ws_bytes = check(lib.ss_csr_workspace_bytes(num_rows), 'ss_csr_workspace_bytes')
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=device)
    stats = torch.empty(4, dtype=torch.int64, device=device)
    src32 = dst32 = None
    if zero_copy and n_edges:
        src32 = torch.empty(n_edges, dtype=torch.int32, device=device)
        dst32 = torch.empty(n_edges, dtype=torch.int32, device=device)
    st = _stream_ptr(device)
    loops = -1 if add_loops else 0
    streamed = zero_copy and n_edges >= INGEST_MIN_EDGES
    fast = (_env_int('SS_B200_CSR_FAST', 1) and n_edges >= CSR_FAST_MIN_EDGES and row_begin == 0
            and n_edges + num_rows < (1 << 40))
    histogram_done = False
    if fast:
        side = None
        if streamed:  # the histogram pass of the fallback rides along under the DMA (the SMs are idle anyway)
            def side(b0, b1, lo, hi, c):
                check(lib.ss_csr_degree_chunk(_ptr(b0), _ptr(b1), hi - lo, row_begin, num_rows, _ptr(src32[lo:hi]),
                                              _ptr(dst32[lo:hi]), _ptr(stats), _ptr(ws), ws.numel(), 1 if c == 0 else 0,
                                              _stream_ptr(device)), 'ss_csr_degree_chunk')
        got = _try_sorted_csr(src, dst, n_edges, num_rows, add_loops, device, streamed_from=ei if streamed else None,
                              degree_side=side, chunk_hook=chunk_hook if streamed else None)
        if got is not None:
            return got
        histogram_done = streamed
    rowptr = torch.empty(num_rows + 1, dtype=torch.int64, device=device)
    ring = None
    if streamed:
        if not histogram_done:
            ring = _streamed_degree_pass(ei, n_edges, row_begin, num_rows, src32, dst32, stats, ws, device)
        check(lib.ss_csr_rowptr_finish(loops, row_begin, num_rows, _ptr(rowptr), _ptr(stats), _ptr(ws), ws.numel(), st),
              'ss_csr_rowptr_finish')
    else:
        check(lib.ss_csr_rowptr(_ptr(src), _ptr(dst), n_edges, loops, row_begin, num_rows, _ptr(rowptr), _ptr(src32),
                                _ptr(dst32), _ptr(stats), _ptr(ws), ws.numel(), st), 'ss_csr_rowptr')
    max_id, nnz, n_loops_used, min_id = (int(v) for v in stats.tolist())
    del ring
    if n_edges and min_id < 0:
        raise IndexError(f'edge_index holds a negative node id ({min_id})')
    if max_id >= (1 << 31):
        raise IndexError('node ids must be < 2^31')
    colidx = torch.empty(max(nnz, 4), dtype=torch.int32, device=device)
    if _env_int('SS_B200_CSR_BIN', 0) and n_edges >= _env_int('SS_B200_CSR_BIN_MIN_EDGES', 1 << 22):
        # EXPERIMENTAL, opt-in, measured SLOWER (profiles/r02_experiments_call_a.txt): group the edges by destination
        # block first so that the fill's scattered stores stay inside an L2-resident window of colidx
        n_binned = nnz - min(max(n_loops_used - row_begin, 0), num_rows)
        b_src = torch.empty(max(n_binned, 1), dtype=torch.int32, device=device)
        b_dst = torch.empty(max(n_binned, 1), dtype=torch.int32, device=device)
        bws = torch.zeros(check(lib.ss_csr_bin_workspace_bytes(), 'ss_csr_bin_workspace_bytes'), dtype=torch.uint8,
                          device=device)
        check(lib.ss_csr_bin_edges(_ptr(src), _ptr(dst), _ptr(src32), _ptr(dst32), n_edges, loops, _ptr(stats), row_begin,
                                   num_rows, _ptr(rowptr), _ptr(b_src), _ptr(b_dst), _ptr(bws), bws.numel(), st),
              'ss_csr_bin_edges')
        check(lib.ss_csr_fill(None, None, _ptr(b_src), _ptr(b_dst), n_binned, loops, _ptr(stats), row_begin, num_rows,
                              _ptr(rowptr), _ptr(colidx), _ptr(ws), ws.numel(), st), 'ss_csr_fill')
        return rowptr, colidx, nnz, max_id
    check(lib.ss_csr_fill(_ptr(src), _ptr(dst), _ptr(src32), _ptr(dst32), n_edges, loops, _ptr(stats), row_begin,
                          num_rows, _ptr(rowptr), _ptr(colidx), _ptr(ws), ws.numel(), st), 'ss_csr_fill')
    return rowptr, colidx, nnz, max_id


class _BareOwner(object):
    """what a propagation operator constructed on its own (as the reference's module-level classes allow) needs of an
    ElphHashes: no sketch shape, so only the plain kernels apply"""
    num_perm, p, max_hops, event_log = 0, 0, 0, None

    def __init__(self):
        self._dev = {}

    def _event_begin(self, device):
        return None

    def _event_end(self, name, start, device):
        return None


def _session_of(op):
    from .session import PropagationSession
    if op.owner is not None:
        return op.owner._prop
    if getattr(op, '_own_session', None) is None:
        op._own_session = PropagationSession(_BareOwner())
    return op._own_session


class MinhashPropagation(object):
    """element-wise min over in-neighbours (hashing.py:28-35).  `edge_index` already holds the self loops."""

    def __init__(self, owner=None):
        self.owner = owner

    @torch.no_grad()
    def __call__(self, x, edge_index):
        return self.forward(x, edge_index)

    @torch.no_grad()
    def forward(self, x, edge_index):
        return _propagate(x, edge_index, _session_of(self), True)


class HllPropagation(object):
    """register-wise max over in-neighbours (hashing.py:38-45)"""

    def __init__(self, owner=None):
        self.owner = owner

    @torch.no_grad()
    def __call__(self, x, edge_index):
        return self.forward(x, edge_index)

    @torch.no_grad()
    def forward(self, x, edge_index):
        return _propagate(x, edge_index, _session_of(self), False)


def _propagate(x, edge_index, session, is_min):
    """operator form on the reference's tensors (ELPH calls it 2K times per forward, models/elph.py:209-212): see
    session.PropagationSession -- cached CSR revalidated on the device, pair fusion, sketch reuse, half-record merges"""
    if x.dim() != 2:
        raise ValueError('x must be [n_nodes, width]')
    return session.propagate(x, edge_index, is_min)


class HopSketch(object):
    """one hop of a SketchTables: behaves like {'hll': int8 [N, m], 'minhash': int64 [N, P]}"""

    def __init__(self, records, num_perm, p, out_device, lease=None):
        self._records = records  # uint8 [N, record_bytes] on the GPU
        self._lease = lease      # [True] while the buffer still belongs to this table (see dist.ShardedElphHashes)
        self.num_perm = num_perm
        self.p = p
        self.out_device = out_device
        self._cache = {}

    @property
    def records(self):
        if self._lease is not None and not self._lease[0]:
            raise RuntimeError('these sketch tables live in buffers that a later build_hash_tables of the same sharded '
                               'engine has reused; build with reuse_buffers=False to keep several alive')
        return self._records

    def keys(self):
        return ['hll', 'minhash']

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return 2

    def __contains__(self, key):
        return key in ('hll', 'minhash')

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, key):
        if key not in ('hll', 'minhash'):
            raise KeyError(key)
        if key not in self._cache:
            rec = self.records
            n = rec.shape[0]
            dev = rec.device
            with torch.cuda.device(dev):
                if key == 'minhash':
                    out = torch.empty((n, self.num_perm), dtype=torch.int64, device=dev)
                    check(lib.ss_unpack_records(_ptr(rec), rec.stride(0), n, self.num_perm, self.p, _ptr(out), None,
                                                _stream_ptr(dev)), 'ss_unpack_records')
                else:
                    out = torch.empty((n, 1 << self.p), dtype=torch.int8, device=dev)
                    check(lib.ss_unpack_records(_ptr(rec), rec.stride(0), n, self.num_perm, self.p, None, _ptr(out),
                                                _stream_ptr(dev)), 'ss_unpack_records')
            self._cache[key] = out.to(self.out_device)
        return self._cache[key]

    def as_dict(self, device='cpu'):
        return {'hll': self['hll'].to(device), 'minhash': self['minhash'].to(device)}


class SketchTables(dict):
    """{hop: HopSketch} for hop = 0..K.  Pickles as the reference's plain dict of CPU tensors."""

    def __init__(self, hops, num_perm, p):
        super().__init__(hops)
        self.num_perm = num_perm
        self.p = p

    def records(self, k):
        return dict.__getitem__(self, k).records

    def __reduce__(self):
        # pickle exactly like a plain (Ordered)dict of {'hll', 'minhash'} CPU tensors: loadable by torch.load
        # with weights_only=True and by the reference, without this package
        import collections
        plain = [(k, dict.__getitem__(self, k).as_dict('cpu')) for k in sorted(self.keys())]
        return (collections.OrderedDict, (), None, None, iter(plain))


class ElphHashes(object):
    """
    class to store hashes and retrieve subgraph features -- B200 engine behind the reference's API
    (hashing.py:48-323)
    """

    def __init__(self, args, hll_tables=None, merge_variant='auto'):
        assert args.max_hash_hops in {1, 2, 3}, f'hashing is not implemented for {args.max_hash_hops} hops'
        self.max_hops = args.max_hash_hops
        self.floor_sf = args.floor_sf  # if true set minimum sf to 0 (they're counts, so it should be)
        # minhash params (hashing.py:57-62)
        self._mersenne_prime = np.uint64((1 << 61) - 1)
        self._max_minhash = np.uint64((1 << 32) - 1)
        self._minhash_range = (1 << 32)
        self.minhash_seed = 1
        self.num_perm = args.minhash_num_perm
        from .session import PropagationSession
        self._prop = PropagationSession(self)  # graph cache + memoised sketches shared by both operators
        self.minhash_prop = MinhashPropagation(self)
        # hll params (hashing.py:64-81)
        self.p = args.hll_p
        self.m = 1 << self.p
        self.use_zero_one = args.use_zero_one
        self.label_lookup = LABEL_LOOKUP[self.max_hops]
        if not 4 <= self.p <= 18:
            raise ValueError('p must be in [4, 18]')
        self.hll_hashfunc = None  # the reference stores datasketch's sha1 hashfunc but never calls it
        self.alpha = hll_alpha(self.p)
        self.max_rank = 64 - self.p
        self.hll_size = self.m
        # provenance of the HLL++ tables ('datasketch' | 'packaged-monte-carlo' | 'file:...' | 'caller')
        self.hll_tables_source = 'caller'
        if hll_tables is None:
            *hll_tables, self.hll_tables_source = hllpp_tables(self.p, with_source=True)
        threshold, raw_estimate, bias = hll_tables
        self.hll_threshold = threshold
        self.bias_vector = torch.tensor(np.asarray(bias, dtype=np.float64), dtype=torch.float)
        self.estimate_vector = torch.tensor(np.asarray(raw_estimate, dtype=np.float64), dtype=torch.float)
        self.hll_prop = HllPropagation(self)
        self.merge_variant = merge_variant
        self.validate_links = True  # bounds-check link endpoints (the reference raises IndexError)
        self.event_log = None  # set to a list to record (name, start_event, end_event) around kernels
        # layout / scheduling knobs of build_hash_tables (defaults = measured best, profiles/r01_merge_tuning.txt):
        #   record_stride: bytes between consecutive records of a hop table (None = compact).  1024 keeps every
        #     768-byte record inside one 1 KB-aligned block (fewer DRAM pages per gathered row) for a third more
        #     table memory; used only for tables of at least `padded_tables_min_nodes` rows (below that the tables
        #     are L2-sized anyway) while all K+1 of them stay below `padded_tables_max_frac` of the device memory
        #   overlap_init: hop-0 initialisation (write-bound) on a side stream under the CSR build.  None = auto:
        #     only when the edge list is streamed from pinned host memory (the SMs idle behind the DMA engine, the
        #     3.8 ms are free); against a device-resident list it is neutral (both kernels stretch), so off there
        self.record_stride = _env_int('SS_B200_RECORD_STRIDE', 1024)
        self.padded_tables_max_frac = 0.45
        self.padded_tables_min_nodes = 1 << 20
        self.overlap_init = {None: None, 0: False}.get(_env_int('SS_B200_OVERLAP_INIT', None), True)
        # linear-counting table, evaluated with the reference's own float32 torch expression (hashing.py:195)
        nz = torch.arange(1, self.m + 1, dtype=torch.int64)
        self._lc_host = torch.cat([torch.zeros(1), self.m * torch.log(self.m / nz)]).float()
        self._dev = {}  # per-device constants
        self._twins = {}

    # ------------------------------------------------------------------ device constants
    def _consts(self, device):
        key = str(device)
        d = self._dev.get(key)
        if d is None or d['est_src'] is not self.estimate_vector or d['bias_src'] is not self.bias_vector:
            est = self.estimate_vector.float().contiguous()
            bias = self.bias_vector.float().contiguous()
            if est.numel() != bias.numel() or est.numel() < 6:
                raise ValueError('estimate_vector / bias_vector must have the same length >= 6')
            ab = self._init_permutations(self.num_perm)
            d = {
                'est_src': self.estimate_vector, 'bias_src': self.bias_vector,
                'lc': self._lc_host.to(device), 'est': est.to(device), 'bias': bias.to(device),
                'perm_a': torch.from_numpy(ab[0].astype(np.int64)).to(device),
                'perm_b': torch.from_numpy(ab[1].astype(np.int64)).to(device),
                'window': torch.from_numpy(log2_window_table()).to(device),
            }
            hc = HllConsts()
            hc.p = self.p
            hc.table_len = est.numel()
            hc.monotone = int(bool(torch.all(est[1:] >= est[:-1])))
            hc.threshold = float(np.float32(self.hll_threshold))
            hc.alpha_m2 = float(np.float32(self.alpha * self.m ** 2))
            hc.five_m = float(np.float32(5 * self.m))
            hc.lc_table = d['lc'].data_ptr()
            hc.raw_estimate = d['est'].data_ptr()
            hc.bias = d['bias'].data_ptr()
            d['hc'] = hc
            self._dev[key] = d
        return d

    def _record_bytes(self):
        return check(lib.ss_record_bytes(self.num_perm, self.p), 'ss_record_bytes')

    # ------------------------------------------------------------------ host helpers (numpy, as the reference)
    def _np_bit_length(self, bits):
        """bits needed to represent each int, evaluated like the reference in float64 (hashing.py:83-89)"""
        return np.ceil(np.log2(bits + 1)).astype(int)

    def _get_hll_rank(self, bits):
        """leading-zero rank of each value in a max_rank-bit word (hashing.py:91-104)"""
        bit_length = self._np_bit_length(bits)
        rank = self.max_rank - bit_length + 1
        if min(rank) <= 0:
            raise ValueError("Hash value overflow, maximum size is %d\
                        bits" % self.max_rank)
        return rank

    def _init_permutations(self, num_perm):
        """(a, b) of the affine permutations; legacy RandomState stream, a then b per permutation
        (hashing.py:106-116) -> uint64 [2, num_perm]"""
        gen = np.random.RandomState(self.minhash_seed)
        ab = np.empty((2, num_perm), dtype=np.uint64)
        for j in range(num_perm):
            ab[0, j] = gen.randint(1, self._mersenne_prime, dtype=np.uint64)
            ab[1, j] = gen.randint(0, self._mersenne_prime, dtype=np.uint64)
        return ab

    # ------------------------------------------------------------------ K1
    def _init_records(self, n_nodes, device, first_id=1, out=None):
        d = self._consts(device)
        rb = self._record_bytes()
        rec = out if out is not None else torch.empty((n_nodes, rb), dtype=torch.uint8, device=device)
        check(lib.ss_init_records(n_nodes, first_id, self.num_perm, self.p, _ptr(d['perm_a']), _ptr(d['perm_b']),
                                  _ptr(d['window']), _ptr(rec), rec.stride(0) if n_nodes else rb,
                                  _stream_ptr(device)), 'ss_init_records')
        return rec

    def initialise_minhash(self, n_nodes):
        """hop-0 MinHash signatures, int64 [n, P] on the CPU like the reference (hashing.py:118-124)"""
        device = _cuda_device()
        with torch.cuda.device(device):
            rec = self._init_records(n_nodes, device)
            return HopSketch(rec, self.num_perm, self.p, torch.device('cpu'))['minhash']

    def initialise_hll(self, n_nodes):
        """hop-0 HLL registers, int8 [n, m] on the CPU like the reference (hashing.py:126-137)"""
        device = _cuda_device()
        with torch.cuda.device(device):
            rec = self._init_records(n_nodes, device)
            return HopSketch(rec, self.num_perm, self.p, torch.device('cpu'))['hll']

    # ------------------------------------------------------------------ K2
    def _merge(self, rowptr, colidx, nnz, rec_in, rec_out, cards_col, device, ws=None, peer_recs=None,
               peer_cards=None, mc_rec=0, mc_cards=0):
        """one hop over the rows of `rec_out`; peer_recs / peer_cards: device addresses (ints) of the peers'
        copies of rec_out / cards_col for the fused multi-GPU exchange (ss_khop_merge_peers)"""
        d = self._consts(device)
        n_rows = rec_out.shape[0]
        need = check(lib.ss_merge_workspace_bytes(nnz, self.num_perm, self.p), 'ss_merge_workspace_bytes')
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 16), dtype=torch.uint8, device=device)
        ev = self._event_begin(device)
        n_peers = len(peer_recs) if peer_recs else 0
        if n_peers or mc_rec:
            pr = (ctypes.c_void_p * max(n_peers, 1))(*(peer_recs or [0]))
            pc = (ctypes.c_void_p * max(n_peers, 1))(*(peer_cards or [0] * max(n_peers, 1)))
            check(lib.ss_khop_merge_peers(_ptr(rowptr), _ptr(colidx), n_rows, nnz, _ptr(rec_in), rec_in.shape[0],
                                          rec_in.stride(0), _ptr(rec_out), rec_out.stride(0), self.num_perm, self.p,
                                          _ptr(ws), ws.numel(), _ptr(cards_col),
                                          cards_col.stride(0) if cards_col is not None else 0, ctypes.byref(d['hc']),
                                          _lib.MERGE_VARIANTS[self.merge_variant], n_peers, pr, pc,
                                          ctypes.c_void_p(mc_rec), ctypes.c_void_p(mc_cards),
                                          _stream_ptr(device)), 'ss_khop_merge_peers')
        else:
            check(lib.ss_khop_merge(_ptr(rowptr), _ptr(colidx), n_rows, nnz, _ptr(rec_in), rec_in.shape[0],
                                    rec_in.stride(0), _ptr(rec_out), rec_out.stride(0), self.num_perm, self.p,
                                    _ptr(ws), ws.numel(), _ptr(cards_col),
                                    cards_col.stride(0) if cards_col is not None else 0, ctypes.byref(d['hc']),
                                    _lib.MERGE_VARIANTS[self.merge_variant], _stream_ptr(device)), 'ss_khop_merge')
        self._event_end('khop_merge', ev, device)
        return ws

    def _merge_block(self, rowptr, colidx, n_rows, capacity, rec_in, rec_out, cards_col, ws, block, device):
        """hop merge of the row block described ON THE DEVICE by `block` (ss_khop_merge_ex, blocked launch)"""
        d = _lib.MergeDesc()
        d.rowptr, d.colidx, d.n_rows, d.nnz = rowptr.data_ptr(), colidx.data_ptr(), n_rows, capacity
        d.rec_in, d.in_rows, d.in_stride = rec_in.data_ptr(), rec_in.shape[0], rec_in.stride(0)
        d.rec_out, d.out_stride = rec_out.data_ptr(), rec_out.stride(0)
        d.num_perm, d.hll_p, d.layout, d.variant = self.num_perm, self.p, _lib.SS_LAYOUT_FULL, _lib.SS_MERGE_TMA
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        hc = self._consts(device)['hc']
        d.cards_out, d.cards_stride, d.hc = cards_col.data_ptr(), cards_col.stride(0), ctypes.addressof(hc)
        d.block = block.data_ptr()
        check(lib.ss_khop_merge_ex(ctypes.byref(d), _stream_ptr(device)), 'ss_khop_merge_ex')

    def _event_begin(self, device):
        if self.event_log is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(device))
        return ev

    def _event_end(self, name, start, device):
        if start is None:
            return
        end = torch.cuda.Event(enable_timing=True)
        end.record(torch.cuda.current_stream(device))
        self.event_log.append((name, start, end))

    def _alloc_hop_tables(self, num_nodes, rb, device):
        """K+1 record tables [num_nodes, rb] (uint8), row pitch = record_stride when that is set and affordable"""
        stride = rb
        want = self.record_stride
        if want is not None and want > rb and num_nodes >= self.padded_tables_min_nodes:
            if want % 16:
                raise ValueError('record_stride must be a multiple of 16')
            total = torch.cuda.get_device_properties(device).total_memory
            if (self.max_hops + 1) * num_nodes * want <= self.padded_tables_max_frac * total:
                stride = want
        if stride == rb:
            return [torch.empty((num_nodes, rb), dtype=torch.uint8, device=device) for _ in range(self.max_hops + 1)]
        return [torch.empty((num_nodes, stride), dtype=torch.uint8, device=device)[:, :rb]
                for _ in range(self.max_hops + 1)]

    def build_hash_tables(self, num_nodes, edge_index):
        """
        Generate a hashing table that allows the size of the intersection of two nodes k-hop neighbours to be
        estimated in constant time (hashing.py:139-165)
        @param num_nodes: The number of nodes in the graph
        @param edge_index: Int Tensor [2, edges] edges in the graph
        @return: hashes, cards. Hashes is a mapping {hop: {'hll', 'minhash'}}, cards is a tensor
        [n_nodes, max_hops] on edge_index.device
        """
        device = _cuda_device(edge_index)
        out_device = edge_index.device
        with torch.cuda.device(device):
            start = time()
            rb = self._record_bytes()
            recs = self._alloc_hop_tables(num_nodes, rb, device)
            main = torch.cuda.current_stream(device)
            init_done = None
            overlap = self.overlap_init
            if overlap is None:
                overlap = (edge_index.device.type == 'cpu' and edge_index.dim() == 2 and edge_index.is_pinned()
                           and edge_index.shape[1] >= INGEST_MIN_EDGES)
            if overlap and num_nodes > 0:
                # hop 0 does not depend on the graph: enqueue it first, on the side stream, so that it runs under
                # the CSR build (which also holds the only host synchronisation of this function)
                side = self._side_stream(device)
                side.wait_stream(main)  # the table memory may still be in use by earlier work on this stream
                with torch.cuda.stream(side):
                    ev = self._event_begin(device)
                    self._init_records(num_nodes, device, out=recs[0])
                    self._event_end('init_records', ev, device)
                    init_done = torch.cuda.Event()
                    init_done.record(side)
            cards = torch.zeros((num_nodes, self.max_hops), dtype=torch.float32, device=device)
            # Hop 1 under the ingest stream: in a source-ordered list every row below a landed chunk's last key is
            # complete, so its hop-1 merge (hop 0 does not depend on the graph) runs while the rest of the list is still
            # crossing PCIe.  Speculative like the streaming CSR itself: discarded if the list turns out ineligible.
            hop1 = {'prev': None, 'keep': [], 'ws': None, 'accepted': False, 'ran': False}
            want_hop1 = (init_done is not None and num_nodes > 0 and self.num_perm == 128 and self.p == 8
                         and self.merge_variant in ('auto', 'tma') and _env_int('SS_B200_INGEST_OVERLAP', 1))

            def chunk_hook(rp, ci, cap, st12, carry, final):
                if rp is None:
                    hop1['accepted'] = bool(final) and hop1['ran']
                    return
                if not hop1['ran']:
                    main.wait_event(init_done)  # hop 0 (side stream, ~4 ms) is long done when the first chunk has landed
                    need = check(lib.ss_merge_workspace_bytes(cap, self.num_perm, self.p), 'ss_merge_workspace_bytes') + 4 * rb
                    hop1['ws'] = torch.empty(need, dtype=torch.uint8, device=device)
                    hop1['ran'] = True
                blk = torch.empty(4, dtype=torch.int64, device=device)
                check(lib.ss_csr_sorted_block(_ptr(carry), _ptr(rp), num_nodes, cap, _ptr(st12), 1 if final else 0,
                                              _ptr(hop1['prev']), _ptr(blk), _stream_ptr(device)), 'ss_csr_sorted_block')
                self._merge_block(rp, ci, num_nodes, cap, recs[0], recs[1], cards[:, 0], hop1['ws'], blk, device)
                hop1['prev'] = blk
                hop1['keep'].append(blk)

            ev = self._event_begin(device)
            try:
                rowptr, colidx, nnz, max_id = build_csr(edge_index, device, num_rows=num_nodes, add_loops=True,
                                                        chunk_hook=chunk_hook if want_hop1 else None)
            finally:  # also on the error paths: the table memory must not be recycled under the side stream
                if init_done is not None:
                    main.wait_event(init_done)
            self._event_end('csr_build', ev, device)
            if max_id >= num_nodes:
                raise IndexError(f'edge_index refers to node {max_id} but num_nodes is {num_nodes}')
            ws = None
            for k in range(self.max_hops + 1):
                logger.info(f"Calculating hop {k} hashes")
                if k == 0:
                    if init_done is None:
                        ev = self._event_begin(device)
                        self._init_records(num_nodes, device, out=recs[0])
                        self._event_end('init_records', ev, device)
                elif k == 1 and hop1['accepted']:
                    continue  # merged block by block while the edge list was still arriving
                elif num_nodes > 0:
                    ws = self._merge(rowptr, colidx, nnz, recs[k - 1], recs[k], cards[:, k - 1], device, ws)
            logger.info(f'hash generation enqueued in {time() - start} s')
            tables = SketchTables({k: HopSketch(recs[k], self.num_perm, self.p, out_device)
                                   for k in range(self.max_hops + 1)}, self.num_perm, self.p)
            if out_device == device:
                return tables, cards
            cards_host = _to_host(cards)
            self._remember_twin(cards_host, cards)  # spares get_subgraph_features the re-upload
            return tables, cards_host

    # device twins of host tensors this engine returned.  Kept HERE (id -> (weakref, twin, version)), never on the
    # tensor: torch.save pickles a tensor's __dict__, so an attribute would travel into the reference's cardcache.pt
    def _remember_twin(self, host_tensor, device_tensor):
        import weakref
        key = id(host_tensor)
        twins = self._twins

        def _drop(_ref, key=key):
            twins.pop(key, None)

        twins[key] = (weakref.ref(host_tensor, _drop), device_tensor, host_tensor._version)

    def _twin_of(self, host_tensor, device):
        ent = self._twins.get(id(host_tensor))
        if ent is None:
            return None
        ref, twin, version = ent
        if ref() is not host_tensor or version != host_tensor._version or twin.device != device:
            return None
        return twin

    # ------------------------------------------------------------------ K4
    def _hop_views(self, hash_table, device):
        """HopView[K+1] over compact records; packs reference-layout tensors on the fly"""
        views = (HopView * (self.max_hops + 1))()
        keep = []
        rb = self._record_bytes()
        for k in range(1, self.max_hops + 1):
            entry = dict.__getitem__(hash_table, k) if isinstance(hash_table, SketchTables) else hash_table[k]
            memo = None if isinstance(entry, HopSketch) else self._prop.records_of(entry)
            if isinstance(entry, HopSketch) and entry.records.device == device:
                rec = entry.records
            elif memo is not None and memo.device == device:
                rec = memo  # the dict ELPH.forward assembled from this engine's own operator outputs
            else:
                mh = _to_device(entry['minhash'], device)
                hl = _to_device(entry['hll'], device)
                mh = (mh if mh.dtype == torch.int64 else mh.long()).contiguous()
                hl = (hl if hl.dtype == torch.int8 else hl.to(torch.int8)).contiguous()
                if mh.shape[1] != self.num_perm or hl.shape[1] != self.m or mh.shape[0] != hl.shape[0]:
                    raise ValueError('hash table shapes do not match num_perm / hll_p of this ElphHashes')
                rec = torch.empty((mh.shape[0], rb), dtype=torch.uint8, device=device)
                check(lib.ss_pack_records(_ptr(mh), _ptr(hl), mh.shape[0], self.num_perm, self.p, _ptr(rec),
                                          rec.stride(0), _stream_ptr(device)), 'ss_pack_records')
            keep.append(rec)
            views[k].records = rec.data_ptr()
            views[k].row_stride = rec.stride(0)
            views[k].num_rows = rec.shape[0]
        return views, keep

    def _flags(self):
        return (_lib.SS_FLAG_USE_ZERO_ONE if self.use_zero_one else 0) | (_lib.SS_FLAG_FLOOR if self.floor_sf else 0)

    def _link_source(self, links, device):
        """links as the kernel reads them: device int64 [n, 2], or the caller's PINNED host tensor in place
        (16 bytes per link read over PCIe, 0.3 % of the kernel's traffic -- no staging copy)"""
        if self.max_hops not in (1, 2, 3):
            raise NotImplementedError("Only 1, 2 and 3 hop hashes are implemented")
        if links.dim() != 2 or links.shape[1] != 2:
            raise ValueError('links must be [n_edges, 2]')
        if links.device.type == 'cpu' and links.dtype == torch.int64 and links.is_contiguous() and links.is_pinned():
            return links
        ld = links.to(device, non_blocking=True) if links.device != device else links
        return (ld if ld.dtype == torch.int64 else ld.long()).contiguous()

    def _raise_if_flagged(self, err, views):
        """the kernels never read out of bounds; with validate_links the flag they set becomes the reference's
        IndexError (costs one 4-byte device->host read, i.e. a stream synchronisation)"""
        if self.validate_links and int(err.item()):
            raise IndexError(f'link endpoint out of range [0, {int(views[1].num_rows)})')

    def _get_intersections(self, edge_list, hash_table):
        """
        extract set intersections as jaccard * union (hashing.py:167-189)
        @param edge_list: [n_edges, 2] tensor to get intersections for
        @return: {(k1, k2): float32 [n_edges]} for k1, k2 in 1..max_hops
        """
        device = _cuda_device(edge_list)
        K = self.max_hops
        with torch.cuda.device(device):
            d = self._consts(device)
            ld = self._link_source(edge_list, device)
            views, keep = self._hop_views(hash_table, device)
            n = ld.shape[0]
            inter = torch.empty((n, K * K), dtype=torch.float32, device=device)
            err = torch.zeros(1, dtype=torch.int32, device=device)
            check(lib.ss_link_features(_ptr(ld), n, views, K, self.num_perm, self.p, None, 0, ctypes.byref(d['hc']),
                                       self._flags(), None, _ptr(inter), _ptr(err), _stream_ptr(device)),
                  'ss_link_features')
            self._raise_if_flagged(err, views)
            inter = inter if edge_list.device == device else _to_host(inter)
            return {(k1, k2): inter[:, (k1 - 1) * K + (k2 - 1)] for k1 in range(1, K + 1) for k2 in range(1, K + 1)}

    def get_subgraph_features(self, links, hash_table, cards, batch_size=11000000, _shard=None):
        """
        structural features of each link: hop-wise intersection / difference cardinalities
        (hashing.py:258-323)
        @param links: tensor [n_edges, 2]
        @param hash_table: mapping {hop: {'hll', 'minhash'}} (a SketchTables or the reference's dict of tensors)
        @param cards: Tensor[n_nodes, max_hops] of hll neighbourhood cardinality estimates
        @param batch_size: links per kernel launch; the result does not depend on it
        @return: Tensor[n_edges, max_hops(max_hops+2)] on links.device
        """
        if links.dim() == 1:
            links = links.unsqueeze(0)
        device = _cuda_device(links)
        K = self.max_hops
        F = K * (K + 2)
        with torch.cuda.device(device):
            views, keep = self._hop_views(hash_table, device)
            cd = self._twin_of(cards, device)  # device twin of a cards tensor we returned to the host
            if cd is None:
                cd = cards.to(device, non_blocking=True) if cards.device != device else cards
                cd = (cd if cd.dtype == torch.float32 else cd.float()).contiguous()
            if cd.dim() != 2 or cd.shape[1] < K:
                raise ValueError('cards must be [n_nodes, max_hops]')
            n_table_rows = min(int(views[k].num_rows) for k in range(1, K + 1))
            if cd.shape[0] < n_table_rows:
                # the kernel bounds-checks link endpoints against the hop tables only; the reference would raise
                # IndexError from cards[links[:, 0]] (hashing.py:274)
                raise IndexError(f'cards has {cd.shape[0]} rows but the hash tables have {n_table_rows}')
            ld = self._link_source(links, device)
            n = ld.shape[0]
            d = self._consts(device)
            flags = self._flags()
            err = torch.zeros(1, dtype=torch.int32, device=device)
            batch_size = max(int(batch_size), 1)
            main = torch.cuda.current_stream(device)

            def launch(link_rows, out_rows):
                ev = self._event_begin(device)
                check(lib.ss_link_features_sharded(_ptr(link_rows), link_rows.shape[0], views, K, self.num_perm, self.p,
                                                   _ptr(cd), cd.stride(0), ctypes.byref(d['hc']), flags, _ptr(out_rows),
                                                   None, _ptr(err), ctypes.byref(_shard) if _shard is not None else None,
                                                   _stream_ptr(device)), 'ss_link_features')
                self._event_end('link_features', ev, device)

            if links.device == device:
                out = torch.empty((n, F), dtype=torch.float32, device=device)
                for lo in range(0, n, batch_size):
                    hi = min(lo + batch_size, n)
                    launch(ld[lo:hi], out[lo:hi])
                self._raise_if_flagged(err, views)
                return out
            # host result: a three-stage pipeline over batches of `step` links, double buffered --
            #   ingest stream : DMA of the next batch of a pinned link list into a device slot (a pageable list was
            #                   copied to the device as a whole above)
            #   current stream: the kernel
            #   side stream   : device -> pinned-host copy of the finished batch
            # so the H2D copy of batch b + 1 and the D2H copy of batch b - 1 overlap the kernel of batch b
            try:
                out = torch.empty((n, F), dtype=torch.float32, pin_memory=True)
            except RuntimeError:
                out = torch.empty((n, F), dtype=torch.float32)
            step = max(min(batch_size, 1 << 21), 1)
            side = self._side_stream(device)
            bufs = [torch.empty((min(step, max(n, 1)), F), dtype=torch.float32, device=device) for _ in range(2)]
            freed = [None, None]
            staged = not ld.is_cuda  # pinned host list
            if staged:
                ingest = _ingest_stream(device)
                lbufs = [torch.empty((min(step, max(n, 1)), 2), dtype=torch.int64, device=device) for _ in range(2)]
                ingest.wait_stream(main)
                kernel_done = [None, None]

                def stage_in(b):
                    lo_ = b * step
                    hi_ = min(lo_ + step, n)
                    with torch.cuda.stream(ingest):
                        if kernel_done[b & 1] is not None:
                            ingest.wait_event(kernel_done[b & 1])
                        lbufs[b & 1][:hi_ - lo_].copy_(ld[lo_:hi_], non_blocking=True)
                        ev_in = torch.cuda.Event()
                        ev_in.record(ingest)
                    return ev_in

                n_batches = (n + step - 1) // step
                arrived = stage_in(0) if n_batches else None
            for b, lo in enumerate(range(0, n, step)):
                hi = min(lo + step, n)
                buf = bufs[b & 1][:hi - lo]
                if freed[b & 1] is not None:
                    main.wait_event(freed[b & 1])
                if staged:
                    main.wait_event(arrived)
                    launch(lbufs[b & 1][:hi - lo], buf)
                else:
                    launch(ld[lo:hi], buf)
                done = torch.cuda.Event()
                done.record(main)
                if staged:
                    kernel_done[b & 1] = done
                    if b + 1 < n_batches:
                        arrived = stage_in(b + 1)
                with torch.cuda.stream(side):
                    side.wait_event(done)
                    out[lo:hi].copy_(buf, non_blocking=True)
                    freed[b & 1] = torch.cuda.Event()
                    freed[b & 1].record(side)
            side.synchronize()
            main.synchronize()
            self._raise_if_flagged(err, views)
            return out

    def _side_stream(self, device):
        key = 'side:' + str(device)
        if key not in self._dev:
            self._dev[key] = torch.cuda.Stream(device=device)
        return self._dev[key]

    # ------------------------------------------------------------------ K3 / K5 helpers
    def get_hashval(self, x):
        return x.hashvals

    def _linearcounting(self, num_zero):
        """m * log(m / num_zero) (hashing.py:194-195) via the table built with that expression"""
        return self._lc_host.to(num_zero.device)[num_zero.long()]

    def _estimate_bias(self, e):
        """mean bias of the 6 nearest raw estimates (hashing.py:197-204)"""
        device = _cuda_device(e)
        with torch.cuda.device(device):
            d = self._consts(device)
            ed = _to_device(e, device).float().contiguous()
            out = torch.empty_like(ed)
            check(lib.ss_estimate_bias(_ptr(ed), ed.numel(), ctypes.byref(d['hc']), _ptr(out), _stream_ptr(device)),
                  'ss_estimate_bias')
            return out.to(e.device)

    def _refine_hll_count_estimate(self, estimate):
        """subtract the bias where estimate <= 5m, in place like the reference (hashing.py:206-210)"""
        idx = estimate <= 5 * self.m
        estimate_bias = self._estimate_bias(estimate)
        estimate[idx] = estimate[idx] - estimate_bias[idx]
        return estimate

    def hll_count(self, regs):
        """
        Estimate the size of set unions associated with regs (hashing.py:212-232)
        @param regs: A tensor of registers [n_nodes, register_size] (or one row)
        @return: float32 [n_nodes] on regs.device
        """
        memo = self._prop.cards_of(regs) if regs.dim() == 2 else None
        if memo is not None:  # registers this engine merged itself: the merge epilogue already counted them
            return memo
        if regs.dim() == 1:
            regs = regs.unsqueeze(dim=0)
        if regs.shape[1] != self.m:
            raise ValueError(f'register rows must have {self.m} entries')
        device = _cuda_device(regs)
        with torch.cuda.device(device):
            d = self._consts(device)
            rd = _to_device(regs, device)
            rd = (rd if rd.dtype in (torch.int8, torch.uint8) else rd.to(torch.uint8)).contiguous()
            out = torch.empty(rd.shape[0], dtype=torch.float32, device=device)
            check(lib.ss_hll_count(_ptr(rd), rd.stride(0) if rd.shape[0] else self.m, rd.shape[0],
                                   ctypes.byref(d['hc']), _ptr(out), 1, _stream_ptr(device)), 'ss_hll_count')
            return out.to(regs.device)

    def _hll_merge(self, src, dst):
        if src.shape != dst.shape:
            raise ValueError('source and destination register shapes must be the same')
        device = _cuda_device(src)
        with torch.cuda.device(device):
            a = _to_device(src, device).to(torch.int8).contiguous()
            b = _to_device(dst, device).to(torch.int8).contiguous()
            out = torch.empty_like(a)
            check(lib.ss_max_i8(_ptr(a), _ptr(b), a.numel(), _ptr(out), _stream_ptr(device)), 'ss_max_i8')
            return out.to(src.dtype).to(src.device)

    def _neighbour_merge(self, root, neighbours, is_min):
        device = _cuda_device(root)
        with torch.cuda.device(device):
            want = torch.int64 if is_min else torch.int8
            rows = torch.cat([_to_device(root, device).unsqueeze(dim=0), _to_device(neighbours, device)], dim=0)
            rows = rows.to(want).contiguous()
            rowptr = torch.tensor([0, rows.shape[0]], dtype=torch.int64, device=device)
            out = torch.empty((1, rows.shape[1]), dtype=want, device=device)
            fn = lib.ss_prop_min_i64 if is_min else lib.ss_prop_max_i8
            check(fn(_ptr(rowptr), None, 1, _ptr(rows), _ptr(out), rows.shape[1], _stream_ptr(device)), 'ss_prop')
            return out[0].to(root.dtype).to(root.device)

    def hll_neighbour_merge(self, root, neighbours):
        return self._neighbour_merge(root, neighbours, is_min=False)

    def minhash_neighbour_merge(self, root, neighbours):
        return self._neighbour_merge(root, neighbours, is_min=True)

    def jaccard(self, src, dst):
        """
        get the minhash Jaccard estimate (hashing.py:247-256)
        @param src: tensor [n_edges, num_perms] of hashvalues
        @param dst: tensor [n_edges, num_perms] of hashvalues
        @return: tensor [n_edges] jaccard estimates
        """
        if src.shape != dst.shape:
            raise ValueError('source and destination hash value shapes must be the same')
        device = _cuda_device(src)
        with torch.cuda.device(device):
            width = src.shape[-1]
            a = _to_device(src, device).long().reshape(-1, width).contiguous()
            b = _to_device(dst, device).long().reshape(-1, width).contiguous()
            out = torch.empty(a.shape[0], dtype=torch.float32, device=device)
            # the reference divides by num_perm whatever the row width (hashing.py:256)
            check(lib.ss_jaccard_i64(_ptr(a), _ptr(b), a.shape[0], width, self.num_perm, _ptr(out),
                                     _stream_ptr(device)), 'ss_jaccard_i64')
            return out.reshape(src.shape[:-1]).to(src.device)
