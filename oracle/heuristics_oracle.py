"""TEST INFRASTRUCTURE ONLY -- CPU restatement (scipy, float64) of the common-neighbour heuristics of
/root/reference/src/heuristics.py:11-71, used to check subgraph_sketching_b200.heuristics.  Pinned against the
unmodified reference functions in tests/test_oracle.py (build container) and via tests/golden/heuristics.npz."""
import numpy as np
import scipy.sparse as ssp
import torch


def adjacency(edge_index, num_nodes, edge_weight=None):
    """csr_matrix((w, (row, col))) as datasets/elph.py:69-72 builds it (duplicates are summed)"""
    w = np.ones(edge_index.shape[1]) if edge_weight is None else np.asarray(edge_weight, dtype=float)
    return ssp.csr_matrix((w, (np.asarray(edge_index[0]), np.asarray(edge_index[1]))), shape=(num_nodes, num_nodes))


def scores(A, links, kind):
    """sum_w A[u,w] * A_[v,w] with A_ = A scaled column-wise by 1 (cn), 1/log(colsum) (aa), 1/colsum (ra)"""
    links = np.asarray(links)
    if kind == 'cn':
        A_ = A
    else:
        with np.errstate(divide='ignore'):
            col = A.sum(axis=0)
            mult = 1 / (np.log(col) if kind == 'aa' else col)
        mult[np.isinf(mult)] = 0
        A_ = A.multiply(mult).tocsr()
    out = np.array(np.sum(A[links[:, 0]].multiply(A_[links[:, 1]]), 1)).flatten()
    return torch.FloatTensor(out)
