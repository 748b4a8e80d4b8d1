"""CPU-only, build container only (/root/reference present): the drop-in mechanism of INTEGRATION.md.
With `sys.modules['src.hashing']` pointing at the B200 module, the UNMODIFIED reference model file binds the
B200 `ElphHashes` and constructs it from the reference's own argument namespace.  (The forward pass needs a
GPU and the reference tree at the same time, which no box has; its call sequence is replayed in
tests/test_gpu_parity.py::test_elph_forward_call_pattern.)"""
import importlib
import sys
import types
import warnings
from argparse import Namespace

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present (GPU box)')


def _stub_pyg_for_models():
    """the trivial extra third-party surface src/models/{elph,gnn}.py import"""
    ref_loader.load()  # puts oracle/stubs (datasketch, torch_geometric base) and /root/reference on sys.path
    tg = importlib.import_module('torch_geometric')
    tnn = importlib.import_module('torch_geometric.nn')

    class _Conv(torch.nn.Module):
        def __init__(self, in_channels, out_channels, **kw):
            super().__init__()
            self.lin = torch.nn.Linear(in_channels, out_channels)

        def forward(self, x, edge_index, *a, **k):
            return self.lin(x)

        def reset_parameters(self):
            self.lin.reset_parameters()

    tnn.GCNConv = getattr(tnn, 'GCNConv', _Conv)
    tnn.SAGEConv = getattr(tnn, 'SAGEConv', _Conv)

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod('torch_geometric.nn.conv', gcn_conv=None)
    mod('torch_geometric.nn.conv.gcn_conv', gcn_norm=lambda *a, **k: (a[0], None))
    mod('torch_geometric.nn.dense')
    mod('torch_geometric.nn.dense.linear', Linear=torch.nn.Linear)
    mod('torch_geometric.typing', Adj=object, OptTensor=object)
    mod('torch_geometric.nn.inits', zeros=lambda t: t.data.zero_() if t is not None else None)
    mod('torch_sparse', SparseTensor=object, spmm=None, coalesce=None)
    return tg


def test_reference_models_bind_b200_engine():
    import subgraph_sketching_b200.hashing as b200
    _stub_pyg_for_models()
    saved = {k: sys.modules.get(k) for k in ('src.hashing', 'src.models.elph', 'src.models.gnn')}
    try:
        for k in ('src.models.elph', 'src.models.gnn'):
            sys.modules.pop(k, None)
        sys.modules['src.hashing'] = b200                      # INTEGRATION.md, option A
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            elph = importlib.import_module('src.models.elph')
        assert elph.ElphHashes is b200.ElphHashes
        args = Namespace(max_hash_hops=2, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False,
                         use_feature=True, feature_prop='gcn', propagate_embeddings=False, sign_k=0,
                         label_dropout=0.5, feature_dropout=0.5, hidden_channels=32, num_negs=1,
                         use_struct_feature=True, add_normed_features=False, use_RA=False, train_node_embedding=False)
        model = elph.ELPH(args, num_features=16)
        assert isinstance(model.elph_hashes, b200.ElphHashes)
        assert model.elph_hashes.max_hops == 2 and model.dim == 8
        assert callable(model.elph_hashes.hll_prop) and callable(model.elph_hashes.minhash_prop)
        assert b200.LABEL_LOOKUP == importlib.import_module('oracle.ref_loader').load().LABEL_LOOKUP
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_cache_file_names_match_reference():
    """subgraph_sketching_b200.cache reproduces HashDataset._generate_file_names (datasets/elph.py:154-173)"""
    import types
    from subgraph_sketching_b200 import cache
    _stub_pyg_for_models()
    tg_data = types.ModuleType('torch_geometric.data')
    tg_data.Dataset = object
    sys.modules.setdefault('torch_geometric.data', tg_data)
    import torch_sparse
    torch_sparse.coalesce = getattr(torch_sparse, 'coalesce', None)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = importlib.import_module('src.datasets.elph')
    for hops in (1, 2, 3):
        for split in ('train', 'valid', 'test'):
            for negs in (1, 5):
                for ds, year in (('Cora', 0), ('ogbl-collab', 2010), ('ogbl-collab', 0)):
                    fake = types.SimpleNamespace(max_hash_hops=hops, split=split, root='/tmp/x/',
                                                 args=types.SimpleNamespace(dataset_name=ds, year=year))
                    want = ref.HashDataset._generate_file_names(fake, negs)
                    assert cache.generate_file_names('/tmp/x/', split, hops, negs, ds, year) == want
