"""B200 versions of the common-neighbour link heuristics CN / AA / RA of /root/reference/src/heuristics.py:11-71
(`HashDataset` computes RA with them when --use_RA, datasets/elph.py:76-77).

Same call signatures: `RA(A, edge_index, batch_size=100000) -> (FloatTensor [n_links], edge_index)` where `A` is
a scipy sparse adjacency matrix (or a `SortedAdjacency` built once with `SortedAdjacency.from_edge_index`) and
`edge_index` is the [n_links, 2] tensor of links the reference passes under that name.  The adjacency rows are
sorted on the GPU once; scores come from one warp-per-link intersection kernel in float64 (as scipy), cast to
float32.  No CPU fallback.
"""
from __future__ import annotations


import numpy as np
import torch

from ._lib import check, lib
from .hashing import _cuda_device, _ptr, _stream_ptr, _to_host


class SortedAdjacency(object):
    """CSR adjacency with sorted, de-duplicated rows and float64 weights on the GPU (the form scipy's
    csr_matrix((w, (row, col))) takes after summing duplicates)"""

    def __init__(self, rowptr, colidx, weights, num_nodes):
        self.rowptr, self.colidx, self.weights, self.num_nodes = rowptr, colidx, weights, num_nodes
        self._colsum = None

    @classmethod
    def from_edge_index(cls, edge_index, num_nodes, edge_weight=None, device=None):
        device = _cuda_device(edge_index) if device is None else torch.device(device)
        ei = edge_index.to(device).long()
        w = torch.ones(ei.shape[1], dtype=torch.float64, device=device) if edge_weight is None \
            else edge_weight.to(device).double().view(-1)
        if ei.numel() and (int(ei.min()) < 0 or int(ei.max()) >= num_nodes):
            raise IndexError('edge_index out of range')
        key, inverse = torch.unique(ei[0] * num_nodes + ei[1], return_inverse=True)  # sorted by (row, col)
        weights = torch.zeros(key.numel(), dtype=torch.float64, device=device).index_add_(0, inverse, w)
        rows = key // num_nodes
        rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=device)
        rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=num_nodes), 0)
        colidx = (key % num_nodes).to(torch.int32)
        unit = bool((weights == 1.0).all()) if weights.numel() else True
        return cls(rowptr, colidx.contiguous(), None if unit else weights.contiguous(), num_nodes)

    @classmethod
    def from_scipy(cls, A, device=None):
        device = _cuda_device() if device is None else torch.device(device)
        A = A.tocsr()
        A.sum_duplicates()
        A.sort_indices()
        if A.shape[0] != A.shape[1]:
            raise ValueError('adjacency matrix must be square')
        w = torch.from_numpy(np.asarray(A.data, dtype=np.float64)).to(device)
        unit = bool((w == 1.0).all()) if w.numel() else True
        return cls(torch.from_numpy(A.indptr.astype(np.int64)).to(device),
                   torch.from_numpy(A.indices.astype(np.int32)).to(device).contiguous(),
                   None if unit else w.contiguous(), A.shape[0])

    def col_sums(self):
        """A.sum(axis=0) as float64 [N] (also the `degrees` of datasets/elph.py:74)"""
        if self._colsum is None:
            dev = self.rowptr.device
            with torch.cuda.device(dev):
                out = torch.empty(self.num_nodes, dtype=torch.float64, device=dev)
                check(lib.ss_col_sums(_ptr(self.colidx), _ptr(self.weights), self.colidx.numel(), self.num_nodes,
                                      _ptr(out), _stream_ptr(dev)), 'ss_col_sums')
            self._colsum = out
        return self._colsum


def _as_adjacency(A):
    if isinstance(A, SortedAdjacency):
        return A
    return SortedAdjacency.from_scipy(A)


def _scores(A, links, mult):
    adj = _as_adjacency(A)
    dev = adj.rowptr.device
    if links.dim() != 2 or links.shape[1] != 2:
        raise ValueError('links must be [n_links, 2]')
    with torch.cuda.device(dev):
        ld = links.to(dev).long().contiguous()
        out = torch.empty(ld.shape[0], dtype=torch.float32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib.ss_common_neighbour_scores(_ptr(adj.rowptr), _ptr(adj.colidx), _ptr(adj.weights), _ptr(mult),
                                             adj.num_nodes, _ptr(ld), ld.shape[0], _ptr(out), _ptr(err),
                                             _stream_ptr(dev)), 'ss_common_neighbour_scores')
        if int(err.item()):
            raise IndexError(f'link endpoint out of range [0, {adj.num_nodes})')
        return out if links.device == dev else _to_host(out)


def CN(A, edge_index, batch_size=100000):
    """Common neighbours (heuristics.py:11-26)"""
    adj = _as_adjacency(A)
    mult = torch.ones(adj.num_nodes, dtype=torch.float64, device=adj.rowptr.device)
    return _scores(adj, edge_index, mult), edge_index


def AA(A, edge_index, batch_size=100000):
    """Adamic Adar (heuristics.py:29-49): multiplier 1 / log(A.sum(axis=0)), infinities -> 0"""
    adj = _as_adjacency(A)
    mult = 1.0 / torch.log(adj.col_sums())
    mult[torch.isinf(mult)] = 0
    return _scores(adj, edge_index, mult), edge_index


def RA(A, edge_index, batch_size=100000):
    """Resource Allocation (heuristics.py:52-71): multiplier 1 / A.sum(axis=0), infinities -> 0"""
    adj = _as_adjacency(A)
    mult = 1.0 / adj.col_sums()
    mult[torch.isinf(mult)] = 0
    return _scores(adj, edge_index, mult), edge_index
