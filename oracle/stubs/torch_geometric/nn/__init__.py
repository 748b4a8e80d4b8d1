import torch


class MessagePassing(torch.nn.Module):
    """aggr='max' message passing with PyG's default flow (source_to_target): message j -> i for every
    column (j, i) of edge_index; x_j = x[edge_index[0]], reduced at edge_index[1]; rows that receive no
    message are 0 (scatter-max fill value)."""

    def __init__(self, aggr='max'):
        super().__init__()
        if aggr != 'max':
            raise NotImplementedError('stub implements aggr="max" only')
        self.aggr = aggr

    def propagate(self, edge_index, x=None, size=None):
        src, dst = edge_index[0], edge_index[1]
        msgs = x.index_select(0, src)
        out = torch.zeros_like(x)
        idx = dst.view(-1, 1).expand(-1, x.size(1))
        out.scatter_reduce_(0, idx, msgs, reduce='amax', include_self=False)
        return out
