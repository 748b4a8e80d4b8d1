import torch


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    """PyG semantics: N = num_nodes or max(edge_index)+1; loops appended after the existing edges."""
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1), None


def to_undirected(edge_index):
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    n = int(both.max()) + 1
    key = torch.unique(both[0] * n + both[1])
    return torch.stack([key // n, key % n])
