"""On-disk caches of the sketch tables, cardinalities and link features, in the reference's formats and under
the reference's file names, so caches written by either implementation are read by the other
(/root/reference/src/datasets/elph.py:136-222).

    <root><split>[_negsN][year_Y]_[Khop_]subgraph_featurecache.pt   float32 [L, K(K+2)]   (torch.save of a tensor)
    <root><split>[year_Y]_[Khop_]hashcache.pt                        {hop: {'hll': int8 [N, m], 'minhash': int64 [N, P]}}
    <root><split>[year_Y]_[Khop_]cardcache.pt                        float32 [N, K]

`SketchTables` pickles as that plain mapping of CPU tensors, and `ElphHashes.get_subgraph_features` accepts
the plain mapping back (it packs it into compact records on the fly), so nothing else is needed for
`HashDataset` to run unchanged on top of the engine; `preprocess_subgraph_features` below is the same control
flow as a free function for callers that do not use the reference's dataset class.
"""
from __future__ import annotations

import os
from time import time

import torch


def generate_file_names(root, split, max_hash_hops, num_negs=1, dataset_name='', year=0):
    """(feature cache name, year_str, hop_str) exactly as HashDataset._generate_file_names
    (datasets/elph.py:154-173)"""
    hop_str = f'{max_hash_hops}hop_' if max_hash_hops != 2 else ''
    end_str = f'_{hop_str}subgraph_featurecache.pt'
    year_str = f'year_{year}' if (dataset_name == 'ogbl-collab' and year > 0) else ''
    if num_negs == 1 or split != 'train':
        name = f'{root}{split}{year_str}{end_str}'
    else:
        name = f'{root}{split}_negs{num_negs}{year_str}{end_str}'
    return name, year_str, hop_str


def hash_cache_names(root, split, year_str, hop_str):
    """(hashcache, cardcache) names (datasets/elph.py:187-188)"""
    return f'{root}{split}{year_str}_{hop_str}hashcache.pt', f'{root}{split}{year_str}_{hop_str}cardcache.pt'


def preprocess_subgraph_features(elph_hashes, root, split, links, edge_index, num_nodes, num_negs=1,
                                 dataset_name='', year=0, cache_subgraph_features=False, load_hashes=False,
                                 batch_size=11000000, device=None, verbose=False):
    """features of `links`, read from / written to the reference's caches (datasets/elph.py:175-222):
    cached features win; else cached hashes + cards if `load_hashes`; else built; then the reference's second
    application of floor / knock-out.  Returns a float32 [L, K(K+2)] tensor on `device` (default links.device)."""
    K = elph_hashes.max_hops
    device = links.device if device is None else torch.device(device)
    name, year_str, hop_str = generate_file_names(root, split, K, num_negs, dataset_name, year)
    feats = None
    if cache_subgraph_features and os.path.exists(name):
        feats = torch.load(name).to(device)
        assert feats.shape[0] == len(links), ('subgraph features are inconsistent with the link object. Delete '
                                              'subgraph features file and regenerate')
    if feats is None:
        hash_name, cards_name = hash_cache_names(root, split, year_str, hop_str)
        if load_hashes and os.path.exists(hash_name):
            hashes = torch.load(hash_name)
            if not os.path.exists(cards_name):
                raise FileNotFoundError(f'hashes found at {hash_name}, but cards not found. Delete hashes and run again')
            cards = torch.load(cards_name)
        else:
            start = time()
            hashes, cards = elph_hashes.build_hash_tables(num_nodes, edge_index)
            if verbose:
                print('Preprocessed hashes in: {:.2f} seconds'.format(time() - start))
            if load_hashes:
                torch.save(cards.cpu(), cards_name)
                torch.save(hashes, hash_name)  # SketchTables pickles as the reference's plain mapping
        start = time()
        feats = elph_hashes.get_subgraph_features(links, hashes, cards, batch_size).to(device)
        if verbose:
            print('Preprocessed subgraph features in: {:.2f} seconds'.format(time() - start))
        assert feats.shape[0] == len(links)
        if cache_subgraph_features:
            torch.save(feats.cpu(), name)
    if elph_hashes.floor_sf:
        feats[feats < 0] = 0
    if not elph_hashes.use_zero_one:
        if K > 1:
            feats[:, [4, 5]] = 0
        if K == 3:
            feats[:, [11, 12]] = 0
    return feats
