"""shared helpers for the parity tests (TEST CODE: the only place besides bench/smoke that touches oracle/)"""
import glob
import os
from argparse import Namespace

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GRAPH_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz'))
                     if not os.path.basename(p).startswith(('hll_count', 'heuristics', 'sign')))


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))


def make_args(K=2, P=128, p=8, use_zero_one=False, floor_sf=False):
    return Namespace(max_hash_hops=K, floor_sf=floor_sf, minhash_num_perm=P, hll_p=p, use_zero_one=use_zero_one)


def golden_tables(blob):
    """(threshold, raw_estimate, bias) the reference saw when the fixture was generated"""
    return int(blob['hll_threshold']), blob['estimate_vector'], blob['bias_vector']


def float_close(got, want, scale, tol=1e-6):
    """|got - want| <= tol * max(1, scale): features are differences of float32 quantities of magnitude
    ~cards, so the north-star's 1e-6 is relative to that magnitude"""
    got = torch.as_tensor(got, dtype=torch.float64)
    want = torch.as_tensor(want, dtype=torch.float64)
    scale = torch.clamp(torch.as_tensor(scale, dtype=torch.float64), min=1.0)
    while scale.dim() < got.dim():
        scale = scale.unsqueeze(-1)
    err = (got - want).abs() / scale
    return bool((err <= tol).all()), float(err.max()) if err.numel() else 0.0


def link_scale(links, cards):
    links = torch.as_tensor(links).long()
    cards = torch.as_tensor(cards).float()
    return torch.maximum(cards[links[:, 0]].max(dim=1).values, cards[links[:, 1]].max(dim=1).values)


def rmat_edges(scale, edge_factor, seed, device='cpu'):
    from subgraph_sketching_b200.graphs import rmat_edges as gen
    return gen(scale, edge_factor, seed, device)
