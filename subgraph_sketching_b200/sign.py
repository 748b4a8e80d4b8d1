"""SIGN node-feature pre-propagation on the B200: the drop-in for
`HashDataset._generate_sign_features` / `_preprocess_node_features` of
/root/reference/src/datasets/elph.py:87-134.

    edge_index, w = gcn_norm(edge_index, edge_weight.float(), num_nodes)        # PyG
    x' = torch_sparse.spmm(edge_index, w, N, N, data.x)                         # out[row] += w * x[col]
    sign_k == 0 -> x'            sign_k > 0 -> cat([x, x', x', ...], -1)        # the reference re-propagates
                                                                               # data.x each time (:104-107)

Three kernels of libss_b200.so (csrc/sign.cu): the symmetric-normalisation coefficients per NODE
(`ss_gcn_norm`: deg^-1/2 and the self-loop weight add_remaining_self_loops would give), a CSR of edge positions
keyed by the spmm row (`ss_csr_rowptr` + `ss_sign_fill`), and the row-per-warp SpMM that forms every coefficient
on the fly in the reference's float32 operation order and writes the concatenated blocks directly
(`ss_sign_spmm`).  The normalised edge list and the [E, F] message tensor of the reference never exist.
CPU inputs are offloaded and the result is returned on the input's device; there is no CPU fallback.
"""
from __future__ import annotations

import os
from time import time

import torch

from ._lib import check, lib
from .hashing import _cuda_device, _ptr, _stream_ptr, _to_device, _to_host


def _prepare(edge_index, edge_weight, device):
    ei = _to_device(edge_index, device)
    ei = (ei if ei.dtype == torch.int64 else ei.long()).contiguous()
    if ei.dim() != 2 or ei.shape[0] != 2:
        raise ValueError('edge_index must be [2, n_edges]')
    n_edges = ei.shape[1]
    ew = None
    if edge_weight is not None:
        ew = _to_device(edge_weight, device).reshape(-1)
        ew = (ew if ew.dtype == torch.float32 else ew.float()).contiguous()  # the reference: edge_weight.float()
        if ew.numel() != n_edges:
            raise ValueError('edge_weight must hold one value per edge')
    return ei, ew, n_edges


def gcn_norm_coefficients(edge_index, edge_weight, num_nodes, device=None):
    """per-node form of PyG's gcn_norm: (deg^-1/2 [N], self-loop weight [N]) as float32 device tensors.  The
    normalised weight of edge (r, c, w) is dinv[r] * w * dinv[c]; node i also has the loop (i, i, loop_w[i])."""
    device = _cuda_device(edge_index) if device is None else torch.device(device)
    with torch.cuda.device(device):
        ei, ew, n_edges = _prepare(edge_index, edge_weight, device)
        return _coefficients(ei, ew, n_edges, num_nodes, device)[:2]


def _coefficients(ei, ew, n_edges, num_nodes, device):
    """(dinv, loop_w, workspace, rows_sorted); raises IndexError for ids outside [0, num_nodes) -- the kernel
    validates them in the pass that finds the self loops (one 4-byte read back)"""
    ws_bytes = check(lib.ss_sign_workspace_bytes(num_nodes), 'ss_sign_workspace_bytes')
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=device)
    dinv = torch.empty(num_nodes, dtype=torch.float32, device=device)
    loop_w = torch.empty(num_nodes, dtype=torch.float32, device=device)
    flags = torch.zeros(1, dtype=torch.int32, device=device)
    check(lib.ss_gcn_norm(_ptr(ei[0]) if n_edges else None, _ptr(ei[1]) if n_edges else None, _ptr(ew), n_edges,
                          num_nodes, _ptr(dinv), _ptr(loop_w), _ptr(flags), _ptr(ws), ws.numel(), _stream_ptr(device)),
          'ss_gcn_norm')
    f = int(flags.item())
    if f & 2:
        raise IndexError(f'edge_index refers to a node outside [0, {num_nodes})')
    return dinv, loop_w, ws, not (f & 1)


def sign_features(x, edge_index, edge_weight, sign_k):
    """`_generate_sign_features` on plain tensors: float32 [N, F] for sign_k == 0, else [N, (sign_k + 1) F]
    (elph.py:87-110).  num_nodes = x.size(0) as in the reference.  An edge_index sorted by row (coalesce /
    to_undirected output) gives results bit-identical to the reference's CPU path; otherwise each row's float32
    sum is taken in an unspecified order."""
    if x.dim() != 2:
        raise ValueError('x must be [num_nodes, num_features]')
    if sign_k < 0:
        raise ValueError('sign_k must be >= 0')
    device = _cuda_device(x if x.is_cuda else edge_index)
    N, F = x.shape
    with torch.cuda.device(device):
        xd = _to_device(x, device)
        xd = (xd if xd.dtype == torch.float32 else xd.float()).contiguous()
        ei, ew, n_edges = _prepare(edge_index, edge_weight, device)
        blocks = 1 if sign_k == 0 else sign_k + 1
        out = torch.empty((N, blocks * F), dtype=torch.float32, device=device)
        st = _stream_ptr(device)
        dinv, loop_w, ws, rows_sorted = _coefficients(ei, ew, n_edges, N, device)
        if N and F:
            rowptr = torch.empty(N + 1, dtype=torch.int64, device=device)
            csr_ws = torch.empty(max(check(lib.ss_csr_workspace_bytes(N), 'ss_csr_workspace_bytes'), 256),
                                 dtype=torch.uint8, device=device)
            # histogram + scan of edge_index[0] (the spmm row); no implicit self loops, no id statistics
            check(lib.ss_csr_rowptr(None, _ptr(ei[0]) if n_edges else None, n_edges, 0, 0, N, _ptr(rowptr), None, None,
                                    None, _ptr(csr_ws), csr_ws.numel(), st), 'ss_csr_rowptr')
            perm = None
            if not rows_sorted:
                perm = torch.empty(max(n_edges, 1), dtype=torch.int32, device=device)
                check(lib.ss_sign_fill(_ptr(ei[0]), n_edges, N, _ptr(rowptr), _ptr(perm), _ptr(ws), ws.numel(), st),
                      'ss_sign_fill')
            if sign_k == 0:
                dst, copies = out, 1
            else:
                out[:, :F].copy_(xd)
                dst, copies = out[:, F:], sign_k
            check(lib.ss_sign_spmm(_ptr(rowptr), _ptr(perm), _ptr(ei[1]) if n_edges else None, _ptr(ew), _ptr(dinv),
                                   _ptr(loop_w), _ptr(xd), xd.stride(0), N, F, _ptr(dst), out.stride(0), copies, st),
                  'ss_sign_spmm')
        return out if x.device == device else _to_host(out)


def generate_sign_features(data, edge_index, edge_weight, sign_k):
    """same arguments as HashDataset._generate_sign_features (elph.py:87): `data` only provides `.x`"""
    return sign_features(data.x, edge_index, edge_weight, sign_k)


def feature_cache_name(root, split, sign_k):
    """elph.py:120-123"""
    return f'{root}_{split}_featurecache.pt' if sign_k == 0 else f'{root}_{split}_k{sign_k}_featurecache.pt'


def preprocess_node_features(data, edge_index, edge_weight, sign_k=0, root='.', split='train', load_features=False):
    """HashDataset._preprocess_node_features (elph.py:112-134) as a free function with the reference's cache
    file names: load `{root}_{split}[_k{sign_k}]_featurecache.pt` when asked and present, else compute and,
    when load_features is set, save the CPU tensor there."""
    feature_name = feature_cache_name(root, split, sign_k)
    if load_features and os.path.exists(feature_name):
        print('loading node features from disk')
        return torch.load(feature_name).to(edge_index.device)
    print('constructing node features')
    start_time = time()
    x = generate_sign_features(data, edge_index, edge_weight, sign_k)
    print("Preprocessed features in: {:.2f} seconds".format(time() - start_time))
    if load_features:
        torch.save(x.cpu(), feature_name)
    return x
