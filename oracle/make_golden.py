"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/src/hashing.py behind oracle/stubs) on small seeded inputs.

Run in the build container (the GPU box has no /root/reference):   python -m oracle.make_golden
Every case stores its inputs and the reference's outputs, so the parity tests need nothing but the file.
The HLL++ bias tables the reference saw (packaged Monte-Carlo tables served by the datasketch stub, or
real datasketch if installed -- recorded in `tables_source`) are stored too and injected into the engine
and the oracle restatement by the tests.
"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import ref_loader  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def ring(n):
    i = torch.arange(n)
    return torch.stack([torch.cat([i, (i + 1) % n]), torch.cat([(i + 1) % n, i])])


def barabasi_albert(n, m, seed):
    """undirected BA graph, both directions listed (the reference fixture is a 30-node BA graph,
    test/test_hashing.py:21-33)"""
    rng = np.random.RandomState(seed)
    targets = list(range(m))
    repeated = []
    src, dst = [], []
    for v in range(m, n):
        for t in set(targets):
            src += [v, t]
            dst += [t, v]
        repeated += list(set(targets)) + [v] * m
        targets = [repeated[i] for i in rng.randint(0, len(repeated), size=m)]
    return torch.tensor([src, dst], dtype=torch.int64)


def random_directed(n, e, seed, max_id=None):
    g = torch.Generator().manual_seed(seed)
    hi = n if max_id is None else max_id
    ei = torch.randint(0, hi, (2, e), generator=g)
    return ei


CASES = [
    # name, num_nodes, edge_index, K, P, p, n_links
    ('ring12_k3', 12, ring(12), 3, 128, 8, 40),
    ('ring12_k1', 12, ring(12), 1, 128, 8, 40),
    ('ba30_k2', 30, barabasi_albert(30, 3, 0), 2, 128, 8, 120),
    ('ba300_k3', 300, barabasi_albert(300, 8, 1), 3, 128, 8, 400),
    ('directed_dups_k2', 200, random_directed(200, 1500, 2), 2, 128, 8, 300),
    ('isolated_tail_k3', 260, random_directed(260, 1200, 3, max_id=200), 3, 128, 8, 300),
    ('dense_hub_k2', 3000, torch.cat([random_directed(3000, 9000, 4),
                                      torch.stack([torch.arange(1, 3000), torch.zeros(2999, dtype=torch.int64)]),
                                      torch.stack([torch.zeros(2999, dtype=torch.int64), torch.arange(1, 3000)])],
                                     dim=1), 2, 128, 8, 600),
    ('generic_p4_P8_k2', 120, random_directed(120, 500, 5), 2, 8, 4, 200),
    ('generic_p6_P33_k3', 150, random_directed(150, 900, 6), 3, 33, 6, 200),
    ('generic_p10_P64_k2', 400, barabasi_albert(400, 10, 7), 2, 64, 10, 300),
]


def main():
    ref = ref_loader.load()
    os.makedirs(OUT_DIR, exist_ok=True)
    for name, n, ei, K, P, p, n_links in CASES:
        blob = {'num_nodes': n, 'edge_index': ei.numpy(), 'K': K, 'P': P, 'p': p,
                'tables_source': 'stub' if ref_loader.USING_STUBS.get('datasketch') else 'datasketch'}
        g = torch.Generator().manual_seed(100 + n)
        links = torch.randint(0, n, (n_links, 2), generator=g)
        links[0] = torch.tensor([0, 1])
        links[1] = torch.tensor([3, 3])  # self link
        if ei.shape[1] >= 8:
            links[2:10] = ei[:, :8].t()  # true edges
        blob['links'] = links.numpy()
        for zo in (False, True):
            for fl in (False, True):
                args = Namespace(max_hash_hops=K, floor_sf=fl, minhash_num_perm=P, hll_p=p, use_zero_one=zo)
                eh = ref.ElphHashes(args)
                tables, cards = eh.build_hash_tables(n, ei)
                feats = eh.get_subgraph_features(links, tables, cards)
                blob[f'features_zo{int(zo)}_fl{int(fl)}'] = feats.numpy()
        inter = eh._get_intersections(links, tables)
        blob['intersections'] = torch.stack([inter[(a, b)] for a in range(1, K + 1) for b in range(1, K + 1)],
                                            dim=1).numpy()
        blob['cards'] = cards.numpy()
        for k in range(K + 1):
            blob[f'minhash_{k}'] = tables[k]['minhash'].numpy().astype(np.uint32)
            blob[f'hll_{k}'] = tables[k]['hll'].numpy()
        blob['hll_threshold'] = eh.hll_threshold
        blob['estimate_vector'] = eh.estimate_vector.numpy()
        blob['bias_vector'] = eh.bias_vector.numpy()
        path = os.path.join(OUT_DIR, name + '.npz')
        np.savez_compressed(path, **blob)
        print(f'{name}: N={n} E={ei.shape[1]} K={K} P={P} p={p} L={n_links} -> {os.path.getsize(path)} bytes',
              flush=True)
    # hll_count known answers across regimes (all-equal registers, partial fills, bias regime sweep)
    args = Namespace(max_hash_hops=2, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False)
    eh = ref.ElphHashes(args)
    rng = np.random.RandomState(11)
    rows = [np.full(256, 3), np.concatenate([np.ones(100), np.zeros(156)])]
    for fill in (1, 5, 30, 90, 150, 200, 230, 250, 255, 256):
        for hi in (2, 4, 8, 16, 40):
            r = np.zeros(256)
            idx = rng.permutation(256)[:fill]
            r[idx] = rng.randint(1, hi + 1, size=fill)
            rows.append(r)
    regs = torch.tensor(np.stack(rows), dtype=torch.int8)
    counts = eh.hll_count(regs)
    e = torch.linspace(150., 1500., 400)
    bias = eh._estimate_bias(e)
    np.savez_compressed(os.path.join(OUT_DIR, 'hll_count_p8.npz'), regs=regs.numpy(), counts=counts.numpy(),
                        e=e.numpy(), bias=bias.numpy(), hll_threshold=eh.hll_threshold,
                        estimate_vector=eh.estimate_vector.numpy(), bias_vector=eh.bias_vector.numpy())
    print('hll_count_p8 written')


if __name__ == '__main__' and 'heuristics' not in sys.argv and 'sign' not in sys.argv:
    main()


class _NumpyLinks(object):
    """[n, 2] link list with the two torch-tensor methods heuristics.py uses (`.size(0)`, `[ind, col]`), answering
    with numpy arrays: the scipy in this image (1.18) rejects torch tensors as sparse-matrix indices, which the
    reference's `A[src]` would otherwise hand it.  The reference functions themselves run unmodified."""

    def __init__(self, arr):
        self.arr = np.asarray(arr)

    def size(self, dim):
        return self.arr.shape[dim]

    def __getitem__(self, key):
        ind, col = key
        return self.arr[np.asarray(ind), col]


def heuristics_golden():
    """CN / AA / RA of the unmodified reference on a weighted multigraph and a BA graph"""
    import importlib
    ref_loader.load()
    heur = importlib.import_module('src.heuristics')
    import scipy.sparse as ssp
    blob = {}
    g = torch.Generator().manual_seed(5)
    cases = {'ba300': (300, barabasi_albert(300, 8, 1), None),
             'multi': (150, torch.randint(0, 150, (2, 1500), generator=g), torch.randint(1, 4, (1500,), generator=g))}
    for name, (n, ei, w) in cases.items():
        wv = np.ones(ei.shape[1]) if w is None else w.numpy().astype(float)
        A = ssp.csr_matrix((wv, (ei[0].numpy(), ei[1].numpy())), shape=(n, n))
        links = torch.randint(0, n, (500, 2), generator=g)
        links[:100] = ei[:, :100].t()
        blob[f'{name}_n'] = n
        blob[f'{name}_edge_index'] = ei.numpy()
        blob[f'{name}_weight'] = wv
        blob[f'{name}_links'] = links.numpy()
        for kind, fn in (('cn', heur.CN), ('aa', heur.AA), ('ra', heur.RA)):
            blob[f'{name}_{kind}'] = fn(A, _NumpyLinks(links.numpy()))[0].numpy()
        blob[f'{name}_degrees'] = np.asarray(A.sum(axis=0, dtype=float)).flatten()
    np.savez_compressed(os.path.join(OUT_DIR, 'heuristics.npz'), **blob)
    print('heuristics written')


if __name__ == '__main__' and 'heuristics' in sys.argv:
    heuristics_golden()


def sign_cases():
    """(name, x, edge_index, edge_weight, sign_k): unit weights / integer weights with duplicate edges, duplicate
    self loops and isolated tail nodes / float weights; feature widths on both kernel paths (multiple of 4 or
    not, one or several 128-column tiles)"""
    g = torch.Generator().manual_seed(21)
    out = []
    ei = barabasi_albert(300, 8, 1)
    out.append(('ba300_unit_f16', torch.rand(300, 16, generator=g), ei, torch.ones(ei.shape[1], dtype=torch.int64), (0, 2)))
    ei = torch.randint(0, 150, (2, 1500), generator=g)
    ei[:, :40] = torch.randint(0, 20, (1, 40), generator=g).repeat(2, 1)  # self loops, several per node
    out.append(('multi_int_f7', torch.randn(180, 7, generator=g), ei, torch.randint(1, 4, (1500,), generator=g), (0, 3)))
    ei = torch.randint(0, 400, (2, 6000), generator=g)
    out.append(('float_w_f130', torch.randn(400, 130, generator=g), ei, torch.rand(6000, generator=g) + 0.1, (1,)))
    ei = barabasi_albert(200, 5, 3)
    out.append(('ba200_unit_f256', torch.randn(200, 256, generator=g), ei, torch.ones(ei.shape[1]), (1,)))
    return out


def sign_golden():
    """SIGN pre-propagation through the UNMODIFIED HashDataset._generate_sign_features
    (/root/reference/src/datasets/elph.py:87-110).  torch_geometric / torch_sparse are absent from this image:
    gcn_norm and spmm are served by oracle/sign_oracle.py's restatements (see its header: that part is unpinned)."""
    import importlib
    import types
    from oracle import sign_oracle
    ref_loader.load()

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    importlib.import_module('torch_geometric')
    mod('torch_geometric.data', Dataset=object)
    mod('torch_geometric.nn.conv')
    mod('torch_geometric.nn.conv.gcn_conv', gcn_norm=sign_oracle.gcn_norm)
    mod('torch_sparse', spmm=sign_oracle.spmm, coalesce=None)
    ds = importlib.import_module('src.datasets.elph')
    blob = {}
    for name, x, ei, w, ks in sign_cases():
        blob[f'{name}_x'] = x.numpy()
        blob[f'{name}_edge_index'] = ei.numpy()
        blob[f'{name}_weight'] = w.numpy()
        for k in ks:
            data = types.SimpleNamespace(x=x)
            got = ds.HashDataset._generate_sign_features(None, data, ei, w, k)
            assert torch.equal(got, sign_oracle.sign_features(x, ei, w, k))
            blob[f'{name}_k{k}'] = got.numpy()
    np.savez_compressed(os.path.join(OUT_DIR, 'sign.npz'), **blob)
    print('sign written', os.path.getsize(os.path.join(OUT_DIR, 'sign.npz')))


if __name__ == '__main__' and 'sign' in sys.argv:
    sign_golden()
