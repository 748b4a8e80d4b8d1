"""GPU parity of the common-neighbour heuristics against the unmodified reference's
outputs (tests/golden/heuristics.npz) and the scipy oracle on a larger random multigraph."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rmat_edges
from oracle import heuristics_oracle as ho
from subgraph_sketching_b200 import heuristics as bh

pytestmark = pytest.mark.gpu


def _close(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.all(np.abs(got - want) <= 1e-6 * np.maximum(1.0, np.abs(want)))


@pytest.mark.parametrize('name', ['ba300', 'multi'])
def test_heuristics_vs_reference_golden(name):
    blob = load_golden('heuristics')
    n = int(blob[f'{name}_n'])
    ei = torch.from_numpy(blob[f'{name}_edge_index'])
    w = torch.from_numpy(blob[f'{name}_weight'])
    links = torch.from_numpy(blob[f'{name}_links'])
    adj = bh.SortedAdjacency.from_edge_index(ei, n, w, device='cuda')
    assert np.array_equal(adj.col_sums().cpu().numpy(), blob[f'{name}_degrees'])
    for kind, fn in (('cn', bh.CN), ('aa', bh.AA), ('ra', bh.RA)):
        scores, back = fn(adj, links)
        assert back is links and scores.dtype == torch.float32 and not scores.is_cuda
        assert _close(scores.numpy(), blob[f'{name}_{kind}']), kind
        assert fn(adj, links.cuda())[0].is_cuda
    # a scipy matrix is accepted like in the reference call RA(self.A, self.links)
    A = ho.adjacency(blob[f'{name}_edge_index'], n, blob[f'{name}_weight'])
    assert _close(bh.RA(A, links)[0].numpy(), blob[f'{name}_ra'])
    with pytest.raises(IndexError):
        bh.RA(adj, torch.tensor([[0, n]]))


def test_heuristics_powerlaw_vs_oracle():
    n = 1 << 12
    ei = rmat_edges(12, 16, 9)
    g = torch.Generator().manual_seed(3)
    links = torch.cat([torch.randint(0, n, (4000, 2), generator=g), ei[:, :4000].t()])
    A = ho.adjacency(ei.numpy(), n)
    adj = bh.SortedAdjacency.from_edge_index(ei.cuda(), n)
    for kind, fn in (('cn', bh.CN), ('aa', bh.AA), ('ra', bh.RA)):
        assert _close(fn(adj, links.cuda())[0].cpu().numpy(), ho.scores(A, links.numpy(), kind).numpy()), kind
