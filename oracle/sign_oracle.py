"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch-CPU, float32) of the SIGN pre-propagation of
/root/reference/src/datasets/elph.py:87-110 and of the two third-party functions it calls.

`torch_geometric` (gcn_norm, add_remaining_self_loops) and `torch_sparse` (spmm) are ABSENT from this image and
unpinned in the reference (README.md:45-56), so their arithmetic is restated here from their published
behaviour (PyG 2.x `torch_geometric/nn/conv/gcn_conv.py::gcn_norm`, `utils/loop.py::add_remaining_self_loops`;
`torch_sparse/spmm.py::spmm`) -- PARITY UNPINNED for that part.  What IS pinned: tests/golden/sign.npz is
produced by the UNMODIFIED reference method `HashDataset._generate_sign_features` (oracle/make_golden.py sign)
with these restatements standing in for the two missing imports, so the reference's own call site (argument
order, `.float()`, concatenation, the sign_k > 0 re-propagation of data.x) is what the goldens record.
"""
import torch


def add_remaining_self_loops(edge_index, edge_attr, fill_value, num_nodes):
    """PyG utils/loop.py: drop every self-loop edge, append one loop per node whose weight is the dropped
    loop's (the last one in edge order when a node has several) or `fill_value`"""
    mask = edge_index[0] != edge_index[1]
    loop_index = torch.arange(num_nodes, dtype=edge_index.dtype)
    loop_attr = torch.full((num_nodes,), float(fill_value), dtype=edge_attr.dtype)
    inv = ~mask
    idx, val = edge_index[0][inv], edge_attr[inv]
    for i, v in zip(idx.tolist(), val.tolist()):  # sequential index_put: the last duplicate wins
        loop_attr[i] = v
    edge_index = torch.cat([edge_index[:, mask], loop_index.unsqueeze(0).repeat(2, 1)], dim=1)
    edge_attr = torch.cat([edge_attr[mask], loop_attr])
    return edge_index, edge_attr


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True,
             flow='source_to_target', dtype=None):
    """PyG gcn_conv.py::gcn_norm for a dense edge_index"""
    fill_value = 2. if improved else 1.
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.shape[1], dtype=dtype or torch.float32)
    if add_self_loops:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill_value, num_nodes)
    row, col = edge_index[0], edge_index[1]
    idx = col if flow == 'source_to_target' else row
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype).scatter_add_(0, idx, edge_weight)
    deg_inv_sqrt = deg.pow_(-0.5)
    deg_inv_sqrt.masked_fill_(deg_inv_sqrt == float('inf'), 0)
    edge_weight = deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]
    return edge_index, edge_weight


def spmm(index, value, m, n, matrix):
    """torch_sparse/spmm.py: out[row] += value * matrix[col] (scatter-add over the edges in order)"""
    assert n == matrix.size(-2)
    row, col = index[0], index[1]
    matrix = matrix if matrix.dim() > 1 else matrix.unsqueeze(-1)
    out = matrix.index_select(-2, col)
    out = out * value.unsqueeze(-1)
    res = torch.zeros((m, out.shape[-1]), dtype=out.dtype)
    return res.index_add_(0, row, out)


def sign_features(x, edge_index, edge_weight, sign_k):
    """elph.py:87-110 on plain tensors"""
    num_nodes = x.size(0)
    ei, ew = gcn_norm(edge_index, edge_weight.float(), num_nodes)
    if sign_k == 0:
        return spmm(ei, ew, x.shape[0], x.shape[0], x)
    xs = [x]
    for _ in range(sign_k):
        xs.append(spmm(ei, ew, x.shape[0], x.shape[0], x))  # data.x again, as the reference
    return torch.cat(xs, dim=-1)
