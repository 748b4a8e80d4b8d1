"""GPU parity tests: the CUDA engine (through the ctypes C ABI) against the golden vectors of the
unmodified reference and against the CPU oracle on seeded inputs.  Integer sketches must be bit-exact;
floats within 1e-6 of the magnitude involved (helpers.float_close)."""
import io
import os

import numpy as np
import pytest
import torch

import subgraph_sketching_b200 as ssb
from helpers import (GRAPH_CASES, float_close, golden_tables, link_scale, load_golden, make_args, rmat_edges)
from oracle import sketch_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def engine_for(blob, use_zero_one=False, floor_sf=False, variant='auto'):
    args = make_args(int(blob['K']), int(blob['P']), int(blob['p']), use_zero_one, floor_sf)
    return ssb.ElphHashes(args, hll_tables=golden_tables(blob), merge_variant=variant)


def variants_for(blob):
    return ['auto', 'tma', 'bulk', 'ldg', 'generic'] if (int(blob['P']), int(blob['p'])) == (128, 8) else ['auto']


@pytest.mark.parametrize('name', GRAPH_CASES)
def test_tables_bit_exact_vs_reference_golden(name):
    blob = load_golden(name)
    ei = torch.from_numpy(blob['edge_index']).to(DEV)
    n, K = int(blob['num_nodes']), int(blob['K'])
    for variant in variants_for(blob):
        eh = engine_for(blob, variant=variant)
        tables, cards = eh.build_hash_tables(n, ei)
        assert len(tables) == K + 1 and cards.shape == (n, K) and cards.is_cuda
        for k in range(K + 1):
            mh = tables[k]['minhash']
            hl = tables[k]['hll']
            assert mh.dtype == torch.int64 and hl.dtype == torch.int8 and mh.is_cuda
            assert np.array_equal(mh.cpu().numpy().astype(np.uint32), blob[f'minhash_{k}']), (variant, k)
            assert np.array_equal(hl.cpu().numpy(), blob[f'hll_{k}']), (variant, k)
        ok, err = float_close(cards.cpu(), blob['cards'], torch.from_numpy(blob['cards']))
        assert ok, (variant, err)


@pytest.mark.parametrize('name', GRAPH_CASES)
def test_features_vs_reference_golden(name):
    blob = load_golden(name)
    ei = torch.from_numpy(blob['edge_index']).to(DEV)
    links = torch.from_numpy(blob['links']).to(DEV)
    n, K = int(blob['num_nodes']), int(blob['K'])
    scale = link_scale(blob['links'], blob['cards'])
    eh = engine_for(blob)
    tables, cards = eh.build_hash_tables(n, ei)
    for zo in (False, True):
        for fl in (False, True):
            eh.use_zero_one, eh.floor_sf = zo, fl
            f = eh.get_subgraph_features(links, tables, cards)
            assert f.shape == (links.shape[0], K * (K + 2)) and f.dtype == torch.float32 and f.is_cuda
            ok, err = float_close(f.cpu(), blob[f'features_zo{int(zo)}_fl{int(fl)}'], scale)
            assert ok, (zo, fl, err)
    inter = eh._get_intersections(links, tables)
    assert len(inter) == K * K
    got = torch.stack([inter[(a, b)] for a in range(1, K + 1) for b in range(1, K + 1)], dim=1)
    ok, err = float_close(got.cpu(), blob['intersections'], scale)
    assert ok, err
    # the reference's own dict-of-tensors layout (e.g. a table loaded from a hashcache.pt) gives the same
    plain = {k: {'minhash': torch.from_numpy(blob[f'minhash_{k}'].astype(np.int64)),
                 'hll': torch.from_numpy(blob[f'hll_{k}'])} for k in range(K + 1)}
    eh.use_zero_one, eh.floor_sf = True, False
    f_plain = eh.get_subgraph_features(links.cpu(), plain, torch.from_numpy(blob['cards']))
    assert not f_plain.is_cuda
    ok, err = float_close(f_plain, blob['features_zo1_fl0'], scale)
    assert ok, err


def test_hll_count_and_bias_golden():
    blob = load_golden('hll_count_p8')
    eh = ssb.ElphHashes(make_args(), hll_tables=golden_tables(blob))
    regs = torch.from_numpy(blob['regs'])
    for r in (regs, regs.to(DEV), regs.long(), regs.to(DEV).int()):
        got = eh.hll_count(r)
        assert got.device == r.device and got.dtype == torch.float32
        ok, err = float_close(got.cpu(), blob['counts'], torch.from_numpy(blob['counts']))
        assert ok, err
    assert eh.hll_count(regs[3]).shape == (1,)
    # known answers from the survey (table independent)
    assert abs(float(eh.hll_count(torch.full((256,), 3, dtype=torch.int8))) - 1471.022216797) < 2e-3
    # LC regime is bit-exact (table built with the reference expression)
    lc_rows = blob['counts'] <= 220
    assert np.array_equal(eh.hll_count(regs).numpy()[lc_rows], blob['counts'][lc_rows])
    bias = eh._estimate_bias(torch.from_numpy(blob['e']))
    flips = np.abs(bias.numpy() - blob['bias']) > 1e-5 * np.maximum(1.0, np.abs(blob['bias']))
    # SURVEY "hard part 2": the reference picks the 6 nearest raw estimates with an unstable argsort, so a tie can flip the
    # neighbour set; the observed count is REPORTED (pytest -s / the log), not hidden behind the 1 % allowance
    print(f'[6-NN] {int(flips.sum())} neighbour-set flips out of {flips.size} bias estimates '
          f'(max |delta| = {float(np.abs(bias.numpy() - blob["bias"]).max()):.3e})')
    with open(os.path.join(os.environ.get('SS_TEST_REPORT_DIR', '/tmp'), 'ss_b200_6nn_flips.txt'), 'w') as fh:
        fh.write(f'{int(flips.sum())} flips / {flips.size} estimates\n')
    assert flips.mean() <= 0.01, f'{flips.sum()} 6-NN neighbour-set flips out of {flips.size}'
    # estimates above 5m are untouched by the refinement (test_hashing.py:229-236)
    e = torch.tensor([5 * 256 + 1.0, 4000.0, 1e6])
    assert torch.equal(eh._refine_hll_count_estimate(e.clone()), e)
    small = torch.tensor([300.0, 800.0])
    assert not torch.equal(eh._refine_hll_count_estimate(small.clone()), small)


@pytest.mark.parametrize('K,seed', [(2, 0), (3, 1)])
def test_rmat_with_hubs_vs_oracle(K, seed):
    """power-law graph: hub rows span many ranges of the nnz-split merge; every variant must be bit-exact"""
    scale = 12
    n = 1 << scale
    ei = rmat_edges(scale, 16, seed)
    g = torch.Generator().manual_seed(seed)
    links = torch.cat([torch.randint(0, n, (3000, 2), generator=g), ei[:, :3000].t()])
    o = so.OracleSketches(K, 128, 8, use_zero_one=False, floor_sf=False)
    ot, oc = o.build_hash_tables(n, ei)
    of = o.subgraph_features(links, ot, oc)
    deg = torch.bincount(ei[1], minlength=n)
    assert int(deg.max()) > 512  # a real hub
    for variant in ('tma', 'bulk', 'ldg', 'generic'):
        eh = ssb.ElphHashes(make_args(K), merge_variant=variant)
        tables, cards = eh.build_hash_tables(n, ei.to(DEV))
        for k in range(K + 1):
            assert torch.equal(tables[k]['minhash'].cpu(), ot[k]['minhash']), (variant, k)
            assert torch.equal(tables[k]['hll'].cpu(), ot[k]['hll']), (variant, k)
        ok, err = float_close(cards.cpu(), oc, oc)
        assert ok, (variant, err)
    f = eh.get_subgraph_features(links.to(DEV), tables, cards)
    ok, err = float_close(f.cpu(), of, link_scale(links, oc))
    assert ok, err


def test_init_invariants_and_known_answers():
    """ports of test_initialise_hll / test_initialise_minhash (test/test_hashing.py:338-353) + survey goldens"""
    eh = ssb.ElphHashes(make_args())
    n = 1000
    hll = eh.initialise_hll(n)
    mh = eh.initialise_minhash(n)
    assert not hll.is_cuda and hll.dtype == torch.int8 and hll.shape == (n, 256)
    assert not mh.is_cuda and mh.dtype == torch.int64 and mh.shape == (n, 128)
    assert torch.equal(torch.count_nonzero(hll, dim=1), torch.ones(n, dtype=torch.long))
    assert int(hll.min()) >= 0 and int(hll.max()) <= eh.max_rank + 1
    assert int(mh.max()) <= 2 ** 32 - 1 and int(mh.min()) >= 0
    assert mh[0, :4].tolist() == [4183491429, 1571535613, 1357840683, 3557095531]
    assert mh[1, :4].tolist() == [1090388868, 3382136556, 2349435535, 3020319954]
    got = [(int(torch.nonzero(r)[0]), int(r.max())) for r in hll[:5]]
    assert got == [(229, 2), (138, 1), (240, 4), (20, 1), (220, 1)]
    assert np.array_equal(mh.numpy(), so.minhash_init(n, 128))
    assert np.array_equal(hll.numpy(), so.hll_init(n, 8))
    counts = eh.hll_count(hll)
    assert torch.all((counts - 1).abs() < 0.1)
    for P, p in ((8, 4), (33, 6), (64, 10), (16, 14)):
        e2 = ssb.ElphHashes(make_args(P=P, p=p))
        assert np.array_equal(e2.initialise_minhash(77).numpy(), so.minhash_init(77, P))
        assert np.array_equal(e2.initialise_hll(77).numpy(), so.hll_init(77, p))


def test_init_large_ids_match_oracle_slice():
    """ids far from 0 (sharded init uses first_id = 1 + row offset)"""
    eh = ssb.ElphHashes(make_args())
    dev = torch.device(DEV)
    first = 16_000_000
    rec = eh._init_records(4096, dev, first_id=first)
    hop = ssb.HopSketch(rec, 128, 8, torch.device('cpu'))
    assert np.array_equal(hop['minhash'].numpy(), so.minhash_init(4096, 128, first_id=first))
    assert np.array_equal(hop['hll'].numpy(), so.hll_init(4096, 8, first_id=first))


def test_propagate_operators_two_node_graph():
    """port of test_propagate_minhash (test/test_hashing.py:355-385)"""
    eh = ssb.ElphHashes(make_args())
    ei = torch.tensor([[0, 1, 0, 1], [1, 0, 0, 1]], device=DEV)
    mh = eh.initialise_minhash(2).to(DEV)
    hl = eh.initialise_hll(2).to(DEV)
    out = eh.minhash_prop(mh, ei)
    assert out.dtype == torch.int64 and out.is_cuda
    want = torch.min(mh, dim=0).values
    assert torch.equal(out[0], want) and torch.equal(out[1], want)
    out = eh.hll_prop(hl, ei)
    want = torch.max(hl, dim=0).values
    assert out.dtype == torch.int8 and torch.equal(out[0], want) and torch.equal(out[1], want)
    # CPU tensors are offloaded transparently and come back on the CPU
    out_cpu = eh.minhash_prop(mh.cpu(), ei.cpu())
    assert not out_cpu.is_cuda and torch.equal(out_cpu, eh.minhash_prop(mh, ei).cpu())


def test_propagate_operators_vs_oracle_and_tables():
    n = 700
    g = torch.Generator().manual_seed(9)
    ei = torch.randint(0, 600, (2, 5000), generator=g)  # nodes 600.. have no edges
    ei_loops = so.with_self_loops(ei)
    eh = ssb.ElphHashes(make_args(K=2))
    mh0 = eh.initialise_minhash(n)
    hl0 = eh.initialise_hll(n)
    mh1 = eh.minhash_prop(mh0.to(DEV), ei_loops.to(DEV))
    hl1 = eh.hll_prop(hl0.to(DEV), ei_loops.to(DEV))
    assert torch.equal(mh1.cpu(), so.minhash_propagate(mh0, ei_loops))
    assert torch.equal(hl1.cpu(), so.hll_propagate(hl0, ei_loops))
    assert int(mh1[650].abs().sum()) == 0 and int(hl1[650].abs().sum()) == 0  # no in-edge -> zeros (8a-Q4)
    tables, cards = eh.build_hash_tables(n, ei.to(DEV))
    assert torch.equal(tables[1]['minhash'], mh1) and torch.equal(tables[1]['hll'], hl1)
    # test_hll_counts (test/test_hashing.py:216-227): cards[:, k] == hll_count(table[k+1].hll)
    for k in range(2):
        assert torch.allclose(cards[:, k], eh.hll_count(tables[k + 1]['hll']), atol=1e-8, rtol=0)
        assert torch.allclose(cards[:, k], eh.hll_count(tables[k + 1]['hll'].long()), atol=1e-8, rtol=0)
    assert float(cards[650].abs().sum()) == 0.0


def test_neighbour_merge_port():
    """port of test_neighbour_merge (test/test_hashing.py:313-329): hop-2 row == merge of hop-1 rows"""
    blob = load_golden('ba30_k2')
    ei = torch.from_numpy(blob['edge_index'])
    eh = engine_for(blob)
    tables, _ = eh.build_hash_tables(30, ei)
    assert not tables[1]['hll'].is_cuda  # CPU edge_index -> CPU tables, like the reference
    node = 5
    nbrs = ei[0][ei[1] == node]
    root_h, root_m = tables[1]['hll'][node], tables[1]['minhash'][node]
    merged_h = eh.hll_neighbour_merge(root_h, tables[1]['hll'][nbrs])
    merged_m = eh.minhash_neighbour_merge(root_m, tables[1]['minhash'][nbrs])
    assert torch.equal(merged_h, tables[2]['hll'][node])
    assert torch.equal(merged_m, tables[2]['minhash'][node])


def test_small_helpers():
    eh = ssb.ElphHashes(make_args())
    g = torch.Generator().manual_seed(3)
    a = torch.randint(0, 5, (50, 128), generator=g)
    b = torch.randint(0, 5, (50, 128), generator=g)
    want = torch.count_nonzero(a == b, dim=-1) / 128
    assert torch.equal(eh.jaccard(a, b), want)
    assert torch.equal(eh.jaccard(a.to(DEV), b.to(DEV)).cpu(), want)
    assert float(eh.jaccard(a[0], a[0])) == 1.0 and eh.jaccard(a[0], b[0]).dim() == 0
    h1 = torch.randint(0, 30, (40, 256), generator=g).to(torch.int8)
    h2 = torch.randint(0, 30, (40, 256), generator=g).to(torch.int8)
    assert torch.equal(eh._hll_merge(h1, h2), torch.maximum(h1, h2))
    with pytest.raises(ValueError):
        eh._hll_merge(h1, h2[:, :10])
    with pytest.raises(ValueError):
        eh.jaccard(a, b[:10])


def test_cards_with_fewer_rows_than_the_tables_raise():
    """mismatched caches (cards shorter than the hash tables): IndexError like the reference's cards[links[:, 0]],
    never an out-of-bounds read"""
    n = 512
    ei = rmat_edges(9, 8, 2).to(DEV)
    eh = ssb.ElphHashes(make_args(2))
    tables, cards = eh.build_hash_tables(n, ei)
    links = torch.tensor([[0, 1], [n - 1, 2]], device=DEV)
    with pytest.raises(IndexError):
        eh.get_subgraph_features(links, tables, cards[:n - 10])
    assert eh.get_subgraph_features(links, tables, cards).shape == (2, 8)


def test_feature_api_contract():
    """ports of test_get_subgraph_features (test/test_hashing.py:179-194) and
    test_get_subgraph_features_batched (test/test_elph_datasets.py:69-91)"""
    blob = load_golden('ba300_k3')
    for K in (1, 2, 3):
        eh = ssb.ElphHashes(make_args(K, use_zero_one=True), hll_tables=golden_tables(blob))
        ei = torch.from_numpy(blob['edge_index'])
        links = torch.from_numpy(blob['links'])
        tables, cards = eh.build_hash_tables(300, ei)
        assert not cards.is_cuda
        full = eh.get_subgraph_features(links, tables, cards)
        assert not full.is_cuda and full.shape == (links.shape[0], K * (K + 2))
        assert torch.equal(eh.get_subgraph_features(links, tables, cards, batch_size=3), full)
        assert torch.equal(eh.get_subgraph_features(links, tables, cards, batch_size=7), full)
        one = eh.get_subgraph_features(links[5], tables, cards)  # 1-D link
        assert one.shape == (1, K * (K + 2)) and torch.equal(one[0], full[5])
        eh.use_zero_one = False
        ko = eh.get_subgraph_features(links, tables, cards)
        want = full.clone()
        for col in {1: [], 2: [4, 5], 3: [4, 5, 11, 12]}[K]:
            want[:, col] = 0
        assert torch.equal(ko, want)
        eh.floor_sf = True
        fl = eh.get_subgraph_features(links, tables, cards)
        assert torch.equal(fl, torch.clamp(want, min=0) * 1.0) or torch.equal(fl, torch.where(want < 0, torch.zeros_like(want), want))
        with pytest.raises(IndexError):
            eh.get_subgraph_features(torch.tensor([[0, 300]]), tables, cards)


def test_intersection_symmetry_and_self_links():
    """size-independent property: I(k1,k2)(u,v) == I(k2,k1)(v,u), bit for bit"""
    n = 1 << 14
    ei = rmat_edges(14, 8, 7).to(DEV)
    eh = ssb.ElphHashes(make_args(3))
    tables, cards = eh.build_hash_tables(n, ei)
    g = torch.Generator().manual_seed(1)
    links = torch.randint(0, n, (50000, 2), generator=g).to(DEV)
    a = eh._get_intersections(links, tables)
    b = eh._get_intersections(links.flip(1), tables)
    for k1 in (1, 2, 3):
        for k2 in (1, 2, 3):
            assert torch.equal(a[(k1, k2)], b[(k2, k1)])
    # a self link (u,u): jaccard of hop k with itself is 1 -> I(k,k) == cards[u, k-1]
    u = torch.arange(0, 2000, device=DEV)
    s = eh._get_intersections(torch.stack([u, u], dim=1), tables)
    for k in (1, 2, 3):
        assert torch.equal(s[(k, k)], cards[u, k - 1])


def test_tables_roundtrip_and_torch_save():
    blob = load_golden('ba300_k3')
    eh = engine_for(blob)
    ei = torch.from_numpy(blob['edge_index']).to(DEV)
    tables, cards = eh.build_hash_tables(300, ei)
    buf = io.BytesIO()
    torch.save(tables, buf)
    buf.seek(0)
    loaded = torch.load(buf)  # weights_only default: a plain mapping of CPU tensors
    assert sorted(loaded.keys()) == [0, 1, 2, 3]
    for k in range(4):
        assert np.array_equal(loaded[k]['minhash'].numpy().astype(np.uint32), blob[f'minhash_{k}'])
        assert np.array_equal(loaded[k]['hll'].numpy(), blob[f'hll_{k}'])
    links = torch.from_numpy(blob['links']).to(DEV)
    f1 = eh.get_subgraph_features(links, tables, cards)
    f2 = eh.get_subgraph_features(links, loaded, cards)
    assert torch.equal(f1, f2)


def test_empty_and_degenerate_inputs():
    eh = ssb.ElphHashes(make_args(2))
    # no edges at all: add_self_loops adds nothing -> hop >= 1 rows are all zero
    tables, cards = eh.build_hash_tables(5, torch.zeros((2, 0), dtype=torch.long, device=DEV))
    assert int(tables[1]['minhash'].abs().sum()) == 0 and int(tables[2]['hll'].abs().sum()) == 0
    assert float(cards.abs().sum()) == 0.0
    assert int(tables[0]['hll'].count_nonzero()) == 5
    f = eh.get_subgraph_features(torch.zeros((0, 2), dtype=torch.long, device=DEV), tables, cards)
    assert f.shape == (0, 8)
    # single self loop
    tables, cards = eh.build_hash_tables(3, torch.tensor([[1], [1]], device=DEV))
    assert torch.equal(tables[1]['minhash'][1], tables[0]['minhash'][1])
    assert int(tables[1]['minhash'][2].abs().sum()) == 0
    with pytest.raises(IndexError):
        eh.build_hash_tables(3, torch.tensor([[0, 5], [1, 0]], device=DEV))


def test_elph_forward_call_pattern():
    """the exact engine call sequence of ELPH.forward + train_elph (models/elph.py:186-213, train.py:198-204):
    hop-0 sketches moved to the device, add_self_loops'd edge_index rebuilt on every forward, per-hop
    hll_prop / minhash_prop / hll_count into a CPU `cards`, then get_subgraph_features on a plain dict of
    int64 / int8 CUDA tensors."""
    n, K = 3000, 2
    g = torch.Generator().manual_seed(21)
    ei = torch.randint(0, n - 50, (2, 20000), generator=g)
    links = torch.randint(0, n, (2048, 2), generator=g)
    o = so.OracleSketches(K, 128, 8, use_zero_one=False, floor_sf=False)
    ot, oc = o.build_hash_tables(n, ei)
    of = o.subgraph_features(links, ot, oc)
    eh = ssb.ElphHashes(make_args(K))
    ei_d = ei.to(DEV)
    init_hashes = eh.initialise_minhash(n).to(DEV)
    init_hll = eh.initialise_hll(n).to(DEV)
    for forward_call in range(3):  # later calls hit the graph cache although the loop tensor is a new object
        hash_edge_index = so.with_self_loops(ei_d)
        cards = torch.zeros((n, K))
        table = {}
        for k in range(K + 1):
            if k == 0:
                table[k] = {'minhash': init_hashes, 'hll': init_hll}
            else:
                table[k] = {'hll': eh.hll_prop(table[k - 1]['hll'], hash_edge_index),
                            'minhash': eh.minhash_prop(table[k - 1]['minhash'], hash_edge_index)}
                cards[:, k - 1] = eh.hll_count(table[k]['hll'])
        for k in range(K + 1):
            assert torch.equal(table[k]['minhash'].cpu(), ot[k]['minhash'])
            assert torch.equal(table[k]['hll'].cpu(), ot[k]['hll'])
        feats = eh.get_subgraph_features(links.to(DEV), table, cards)
        assert feats.is_cuda
        ok, err = float_close(feats.cpu(), of, link_scale(links, oc))
        assert ok, err
    # a different graph of the same shape must not be served from the cache
    ei2 = torch.randint(0, n - 50, (2, 20000), generator=g).to(DEV)
    h2 = eh.hll_prop(init_hll, so.with_self_loops(ei2))
    assert torch.equal(h2.cpu(), so.hll_propagate(init_hll.cpu(), so.with_self_loops(ei2.cpu())))


def test_full_size_properties_rmat24():
    """BASELINE.json's scaling-sweep size (R-MAT scale 24, 16.8 M nodes, ~537 M neighbours, K=3) through
    size-independent properties: engine-vs-engine checksum of whole tables, sampled rows against a direct
    gather-reduce, hop monotonicity (self loops), cards == hll_count(table), intersection symmetry."""
    import os
    scale = int(os.environ.get('SS_TEST_FULL_SCALE', '24'))
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9 * (1 << scale) / (1 << 24):
        pytest.skip('not enough free device memory for the full-size case')
    n, K = 1 << scale, 3
    dev = torch.device(DEV)
    ei = rmat_edges(scale, 16, 0, dev)
    eh = ssb.ElphHashes(make_args(K))
    tables, cards = eh.build_hash_tables(n, ei)
    rowptr, colidx, nnz, _ = ssb.build_csr(ei, dev, num_rows=n, add_loops=True)
    del ei
    # (1) checksum of checksums: the per-row bulk-copy engine must reproduce the gather4 engine bit for bit
    alt = ssb.ElphHashes(make_args(K), merge_variant='bulk')
    out = torch.empty_like(tables.records(1))
    c2 = torch.zeros((n, 1), device=dev)
    for k in (1, 2):
        alt._merge(rowptr, colidx, nnz, tables.records(k - 1), out, c2[:, 0], dev)
        assert int(out.view(torch.int32).sum(dtype=torch.int64)) == int(tables.records(k).view(torch.int32).sum(dtype=torch.int64))
        assert torch.equal(c2[:, 0], cards[:, k - 1])
    assert torch.equal(out, tables.records(2))
    del out
    # (2) sampled rows (hubs included) against a direct gather + min/max in torch
    g = torch.Generator().manual_seed(0)
    deg = rowptr[1:] - rowptr[:-1]
    sample = torch.cat([torch.randint(0, n, (300,), generator=g).to(dev), torch.topk(deg, 3).indices])
    for k in (1, 3):
        prev = tables.records(k - 1)
        for r in sample.tolist():
            nb = colidx[int(rowptr[r]):int(rowptr[r + 1])].long()
            rows = prev[nb]
            want_mh = rows[:, :512].contiguous().view(torch.int32).long().bitwise_and(0xffffffff).min(dim=0).values
            want_hl = rows[:, 512:].max(dim=0).values
            got = tables.records(k)[r]
            assert torch.equal(got[:512].view(torch.int32).long().bitwise_and(0xffffffff), want_mh), (k, r)
            assert torch.equal(got[512:], want_hl), (k, r)
    # (3) hop monotonicity on a row sample: every node has a self loop, so sketches only grow
    # (nodes above max(edge_index) get no self loop -- quirk 8a-Q4 -- and are all-zero from hop 1 on)
    max_id = int(colidx[:nnz].max())
    idx = torch.randint(0, max_id + 1, (200000,), generator=g).to(dev)
    if max_id + 1 < n:
        assert int(tables.records(1)[max_id + 1:].count_nonzero()) == 0
    for k in (1, 2, 3):
        a, b = tables.records(k - 1)[idx], tables.records(k)[idx]
        mh_a = a[:, :512].contiguous().view(torch.int32).long().bitwise_and(0xffffffff)
        mh_b = b[:, :512].contiguous().view(torch.int32).long().bitwise_and(0xffffffff)
        assert bool((mh_b <= mh_a).all()) and bool((b[:, 512:] >= a[:, 512:]).all())
        # (4) cards[:, k-1] == hll_count(table[k].hll) on the sample (test_hashing.py:216-227 at full size)
        assert torch.equal(eh.hll_count(b[:, 512:].contiguous()), cards[idx, k - 1])
    # (5) intersection symmetry on 1 M links
    links = torch.randint(0, n, (1_000_000, 2), generator=g).to(dev)
    ia = eh._get_intersections(links, tables)
    ib = eh._get_intersections(links.flip(1), tables)
    for k1 in (1, 2, 3):
        for k2 in (1, 2, 3):
            assert torch.equal(ia[(k1, k2)], ib[(k2, k1)])
    f = eh.get_subgraph_features(links, tables, cards)
    assert f.shape == (1_000_000, 15) and bool(torch.isfinite(f).all())


def test_link_feature_front_ends_agree(monkeypatch):
    """the batched kernel (default: tiles of consecutive links, tails / algebra per batch, source records reused along
    runs of equal sources), the per-link kernel (SS_B200_LINKS=ldg) and the TMA-pair front end (=tma) give the same
    bits -- odd link counts, every tile size, random and source-grouped link lists"""
    n = 1 << 13
    ei = rmat_edges(13, 8, 5).to(DEV)
    g = torch.Generator().manual_seed(2)
    for K in (1, 2, 3):
        eh = ssb.ElphHashes(make_args(K, use_zero_one=True))
        tables, cards = eh.build_hash_tables(n, ei)
        for L in (1, 2, 7, 100, 4097):
            links = torch.randint(0, n, (L, 2), generator=g).to(DEV)
            grouped = links.clone()
            grouped[:, 0] = links[torch.arange(L, device=DEV) // 37, 0]   # runs of 37 links per source
            for lk in (links, grouped):
                monkeypatch.setenv('SS_B200_LINKS', 'ldg')
                a = eh.get_subgraph_features(lk, tables, cards)
                ia = eh._get_intersections(lk, tables)
                monkeypatch.setenv('SS_B200_LINKS', 'tma')
                b = eh.get_subgraph_features(lk, tables, cards)
                ib = eh._get_intersections(lk, tables)
                assert torch.equal(a, b), (K, L)
                assert all(torch.equal(ia[k], ib[k]) for k in ia)
                monkeypatch.setenv('SS_B200_LINKS', 'batched')
                for tile in (None, 3, 8, 24, 32, 96):
                    if tile is None:
                        monkeypatch.delenv('SS_B200_LINK_TILE', raising=False)
                    else:
                        monkeypatch.setenv('SS_B200_LINK_TILE', str(tile))
                    c = eh.get_subgraph_features(lk, tables, cards)
                    ic = eh._get_intersections(lk, tables)
                    assert torch.equal(a, c), (K, L, tile)
                    assert all(torch.equal(ia[k], ic[k]) for k in ia), (K, L, tile)
                monkeypatch.delenv('SS_B200_LINK_TILE', raising=False)
                monkeypatch.delenv('SS_B200_LINKS')
                d = eh.get_subgraph_features(lk, tables, cards)   # the default choice for this K
                assert torch.equal(a, d), (K, L)
    # out-of-range endpoints are flagged by the batched kernel too
    monkeypatch.setenv('SS_B200_LINKS', 'batched')
    with pytest.raises(IndexError):
        eh.get_subgraph_features(torch.tensor([[0, 1], [2, n]], device=DEV), tables, cards)
    with pytest.raises(IndexError):
        eh.get_subgraph_features(torch.tensor([[-1, 1]], device=DEV), tables, cards)


def test_pinned_host_inputs_are_read_in_place(monkeypatch):
    """BUDDY-style CPU inputs: a pinned edge_index / link list is consumed by the kernels over PCIe without a
    staging copy; results equal the device-resident path bit for bit, out-of-range ids still raise"""
    n, K = 6000, 2
    g = torch.Generator().manual_seed(31)
    ei = torch.randint(0, n - 10, (2, 50000), generator=g)
    links = torch.randint(0, n, (70001, 2), generator=g)
    eh = ssb.ElphHashes(make_args(K))
    t_dev, c_dev = eh.build_hash_tables(n, ei.to(DEV))
    f_dev = eh.get_subgraph_features(links.to(DEV), t_dev, c_dev)
    for pin in (True, False):
        e_h = ei.pin_memory() if pin else ei
        l_h = links.pin_memory() if pin else links
        t_h, c_h = eh.build_hash_tables(n, e_h)
        assert not c_h.is_cuda and torch.equal(c_h, c_dev.cpu())
        for k in range(K + 1):
            assert torch.equal(t_h.records(k), t_dev.records(k))
        f_h = eh.get_subgraph_features(l_h, t_h, c_h, batch_size=20000)
        assert not f_h.is_cuda and torch.equal(f_h, f_dev.cpu())
        c_h[0, 0] += 1.0  # a modified cards tensor must not be served from its stale device twin
        f_mod = eh.get_subgraph_features(l_h[:16], t_h, c_h)
        c_ref = c_dev.clone()
        c_ref[0, 0] += 1.0
        assert torch.equal(f_mod, eh.get_subgraph_features(links[:16].to(DEV), t_dev, c_ref).cpu())
    # large pinned lists are streamed (chunked DMA into a staging ring, one degree pass per chunk): force that
    # path with a tiny chunk so that the ring wraps many times and ends on a ragged chunk
    from subgraph_sketching_b200 import hashing as hmod
    monkeypatch.setattr(hmod, 'INGEST_MIN_EDGES', 1)
    monkeypatch.setattr(hmod, 'INGEST_CHUNK', 7001)
    t_s, c_s = eh.build_hash_tables(n, ei.pin_memory())
    assert torch.equal(c_s, c_dev.cpu())
    for k in range(K + 1):
        assert torch.equal(t_s.records(k), t_dev.records(k))
    with pytest.raises(IndexError):
        eh.build_hash_tables(n, torch.tensor([[0, -1], [1, 2]]).pin_memory())
    monkeypatch.undo()
    bad = links.clone()
    bad[5, 1] = n
    with pytest.raises(IndexError):
        eh.get_subgraph_features(bad.pin_memory(), t_dev, c_dev)
    with pytest.raises(IndexError):
        eh.build_hash_tables(n, torch.tensor([[0, -1], [1, 2]]).pin_memory())


def test_operator_forms_use_record_engine_on_large_graphs():
    """hll_prop / minhash_prop called singly on a hub-heavy graph: the half-record TMA merges (registers in place of
    the int8 tensor, MinHash through a 512-byte packed table) and the plain row-per-warp kernels (what an operator
    constructed on its own uses) agree with the oracle and with each other; values outside the sketch range keep the
    reference's exact signed semantics"""
    scale = 13
    n = 1 << scale
    ei = so.with_self_loops(rmat_edges(scale, 16, 11))
    eh = ssb.ElphHashes(make_args(2))
    mh0, hl0 = eh.initialise_minhash(n), eh.initialise_hll(n)
    want_m, want_h = so.minhash_propagate(mh0, ei), so.hll_propagate(hl0, ei)
    ei_d, mh_d, hl_d = ei.to(DEV), mh0.to(DEV), hl0.to(DEV)
    fast_m = eh.minhash_prop(mh_d, ei_d)
    assert eh._prop.stats['single_merges'] == 1 and eh._prop.stats['plain_calls'] == 0
    eh2 = ssb.ElphHashes(make_args(2))
    fast_h = eh2.hll_prop(hl_d, ei_d)
    assert eh2._prop.stats['single_merges'] == 1
    slow_m, slow_h = ssb.MinhashPropagation()(mh_d, ei_d), ssb.HllPropagation()(hl_d, ei_d)
    assert torch.equal(fast_m.cpu(), want_m) and torch.equal(slow_m, fast_m)
    assert torch.equal(fast_h.cpu(), want_h) and torch.equal(slow_h, fast_h)
    # values outside the sketch range (negative / >= 2^32) keep the exact int64 semantics of the reference: the
    # guarded plain kernel takes over on the device, no host round trip decides it
    weird = mh_d.clone()
    weird[5, 7] = -3
    weird[9, 1] = 1 << 40
    assert torch.equal(ssb.ElphHashes(make_args(2)).minhash_prop(weird, ei_d).cpu(), so.minhash_propagate(weird.cpu(), ei))
    # registers are merged as SIGNED bytes, like the reference's int8 scatter-max
    signed = hl_d.clone()
    signed[torch.rand(signed.shape, device=DEV) < 0.01] = -7
    signed[3, :] = -128
    assert torch.equal(ssb.ElphHashes(make_args(2)).hll_prop(signed, ei_d).cpu(), so.hll_propagate(signed.cpu(), ei))
    # other widths / dtypes take the plain kernels
    narrow = torch.randint(-50, 50, (n, 40), device=DEV)
    assert torch.equal(eh.minhash_prop(narrow, ei_d).cpu(), so.minhash_propagate(narrow.cpu(), ei))


def test_elph_session_fusion_reuse_and_graph_changes():
    """the memoised per-batch path (session.py) under the reference's training-loop pattern: a fresh add_self_loops
    tensor every forward, the previous forward's dict still alive while the next one runs.  Every forward is bit-equal
    to the oracle; from the second forward on the two operators of a hop share ONE fused merge; from the third on the
    generations are re-enqueued guarded (graph unchanged -> kernels return at once); a different graph of the same
    shape is noticed on the device and recomputed, without touching tensors the caller still holds"""
    n, K = 4000, 2
    g = torch.Generator().manual_seed(5)
    graphs = [torch.randint(0, n - 50, (2, 30000), generator=g) for _ in range(2)]
    oracles = []
    o = so.OracleSketches(K, 128, 8, use_zero_one=False, floor_sf=False)
    for ei in graphs:
        oracles.append(o.build_hash_tables(n, ei))
    links = torch.randint(0, n, (3000, 2), generator=g)
    eh = ssb.ElphHashes(make_args(K))
    init_hashes = eh.initialise_minhash(n).to(DEV)
    init_hll = eh.initialise_hll(n).to(DEV)
    def forward(gi):
        hash_edge_index = so.with_self_loops(graphs[gi].to(DEV))       # a new tensor object every forward
        cards = torch.zeros((n, K))
        table = {0: {'minhash': init_hashes, 'hll': init_hll}}
        for k in range(1, K + 1):
            table[k] = {'hll': eh.hll_prop(table[k - 1]['hll'], hash_edge_index),
                        'minhash': eh.minhash_prop(table[k - 1]['minhash'], hash_edge_index)}
            cards[:, k - 1] = eh.hll_count(table[k]['hll'])
        feats = eh.get_subgraph_features(links.to(DEV), table, cards)
        return table, cards, feats

    def check(out, gi):
        table, cards, feats = out
        ot, oc = oracles[gi]
        for k in range(K + 1):
            assert torch.equal(table[k]['minhash'].cpu(), ot[k]['minhash']), (gi, k)
            assert torch.equal(table[k]['hll'].cpu(), ot[k]['hll']), (gi, k)
        ok, err = float_close(cards, oc, oc)
        assert ok, err
        ok, err = float_close(feats.cpu(), o.subgraph_features(links, ot, oc), link_scale(links, oc))
        assert ok, err

    sequence = [0, 0, 0, 0, 0, 1, 1, 1, 0, 0]
    prev = None
    for step, gi in enumerate(sequence):
        before = dict(eh._prop.stats)
        cur = forward(gi)
        check(cur, gi)
        if prev is not None:
            check(prev[0], prev[1])           # what the caller still holds was not overwritten
        st = eh._prop.stats
        if step == 0:
            assert st['single_merges'] == 2 * K and st['fused_merges'] == 0
        else:
            assert st['single_merges'] == before['single_merges'], 'every later forward is fused'
            assert st['fused_merges'] == before['fused_merges'] + K
        # two generations alternate (the caller holds the previous forward's): the one of forward step-2 can be
        # re-enqueued guarded iff the graph did not change between step-2 and step-1
        if step >= 3 and sequence[step - 1] == sequence[step - 2]:
            assert st['guarded'] == before['guarded'] + 1, 'generations are re-enqueued guarded'
        prev = (cur, gi)
    # CPU tensors take the same path (transparent offload), results on the CPU
    eh_c = ssb.ElphHashes(make_args(K))
    mh0, hl0 = eh_c.initialise_minhash(n), eh_c.initialise_hll(n)
    for _ in range(2):
        e = so.with_self_loops(graphs[1])
        h1, m1 = eh_c.hll_prop(hl0, e), eh_c.minhash_prop(mh0, e)
        h2, m2 = eh_c.hll_prop(h1, e), eh_c.minhash_prop(m1, e)
        assert not h2.is_cuda and torch.equal(h2, oracles[1][0][2]['hll']) and torch.equal(m2, oracles[1][0][2]['minhash'])
        assert torch.equal(h1, oracles[1][0][1]['hll']) and torch.equal(m1, oracles[1][0][1]['minhash'])
    # node ids are validated on the device and reported at a later call
    bad = so.with_self_loops(graphs[0]).to(DEV)
    bad[0, 17] = n + 3
    eh.hll_prop(init_hll, bad)
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        eh.hll_prop(init_hll, so.with_self_loops(graphs[0].to(DEV)))


def test_link_features_exact_path_for_large_registers():
    """registers above 28 leave the 32-bit fixed-point fast path: arbitrary (synthetic) tables in the reference
    layout, registers up to 57, against the oracle -- for both link-feature front ends"""
    n = 500
    g = torch.Generator().manual_seed(77)
    for K in (1, 2, 3):
        tables = {}
        for k in range(K + 1):
            hll = torch.randint(0, 58, (n, 256), generator=g).to(torch.int8)
            hll[torch.rand((n, 256), generator=g) < 0.3] = 0
            hll[:50] = torch.clamp(hll[:50], max=20)          # some rows stay on the fast path
            tables[k] = {'hll': hll, 'minhash': torch.randint(0, 6, (n, 128), generator=g)}
        cards = torch.rand((n, K), generator=g) * 1000
        links = torch.randint(0, n, (3001, 2), generator=g)
        links[:500] = torch.randint(0, 50, (500, 2), generator=g)
        o = so.OracleSketches(K, 128, 8, use_zero_one=True, floor_sf=False)
        want = o.intersections(links, tables)
        wf = o.subgraph_features(links, tables, cards)
        eh = ssb.ElphHashes(make_args(K, use_zero_one=True))
        got = eh._get_intersections(links.to(DEV), tables)
        for key in want:
            ok, err = float_close(got[key].cpu(), want[key], want[key].abs())
            assert ok, (K, key, err)
        gf = eh.get_subgraph_features(links.to(DEV), tables, cards)
        scale = torch.maximum(link_scale(links, cards), torch.stack([v.abs() for v in want.values()]).max(dim=0).values)
        ok, err = float_close(gf.cpu(), wf, scale)
        assert ok, (K, err)


def test_cache_round_trip_in_reference_formats(tmp_path):
    """features / hashes / cards are cached under the reference's names and formats; a second
    call is served from the caches (first from the hash cache, then from the feature cache) with the same bits"""
    from subgraph_sketching_b200 import cache
    blob = load_golden('ba300_k3')
    ei = torch.from_numpy(blob['edge_index'])
    links = torch.from_numpy(blob['links'])
    root = str(tmp_path) + '/'
    eh = engine_for(blob, use_zero_one=False, floor_sf=True)
    f1 = cache.preprocess_subgraph_features(eh, root, 'train', links, ei, 300, num_negs=1, load_hashes=True)
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ['train_3hop_cardcache.pt', 'train_3hop_hashcache.pt']
    hashes = torch.load(root + 'train_3hop_hashcache.pt')            # the reference's plain mapping of CPU tensors
    assert sorted(hashes.keys()) == [0, 1, 2, 3] and hashes[2]['minhash'].dtype == torch.int64
    assert np.array_equal(hashes[3]['hll'].numpy(), blob['hll_3'])
    saved_cards = torch.load(root + 'train_3hop_cardcache.pt', map_location='cpu', weights_only=True)
    assert np.array_equal(saved_cards.numpy(), eh.build_hash_tables(300, ei)[1].numpy())
    # the cardinalities handed to a CPU caller carry nothing of the engine: a plain CPU tensor with an empty __dict__
    # (torch.save pickles attributes; an embedded device twin would bloat the cache and pin it to one GPU)
    host_cards = eh.build_hash_tables(300, ei)[1]
    assert not host_cards.is_cuda and host_cards.__dict__ == {} and saved_cards.__dict__ == {}
    import os as _os
    assert _os.path.getsize(root + 'train_3hop_cardcache.pt') < 300 * 3 * 4 + 4096
    # ... and the twin that spares get_subgraph_features the re-upload is dropped with the tensor / on modification
    dev0 = torch.device('cuda', torch.cuda.current_device())
    assert eh._twin_of(host_cards, dev0) is not None
    host_cards[0, 0] += 1.0
    assert eh._twin_of(host_cards, dev0) is None
    f2 = cache.preprocess_subgraph_features(eh, root, 'train', links, ei, 300, load_hashes=True,
                                            cache_subgraph_features=True)     # from the hash cache
    assert torch.equal(f1, f2)
    assert (tmp_path / 'train_3hop_subgraph_featurecache.pt').exists()
    f3 = cache.preprocess_subgraph_features(eh, root, 'train', links, None, 300, cache_subgraph_features=True)
    assert torch.equal(f1, f3)                                                 # from the feature cache
    ok, err = float_close(f1, blob['features_zo0_fl1'], link_scale(blob['links'], blob['cards']))
    assert ok, err
    assert cache.generate_file_names(root, 'train', 2, 5)[0].endswith('train_negs5_subgraph_featurecache.pt')


def test_layout_and_scheduling_knobs_are_bit_equal(monkeypatch):
    """build_hash_tables knobs (padded 1024-byte record stride, hop-0 initialisation on a side stream under the
    CSR build, legacy zero-based CSR cursors) change layout / scheduling only: tables, cards and features stay
    bit-identical, strided tables unpack / pickle like compact ones"""
    n, K = 1 << 14, 3
    dev = torch.device(DEV)
    ei = rmat_edges(14, 16, 5, dev)
    g = torch.Generator().manual_seed(2)
    links = torch.randint(0, n, (50_000, 2), generator=g).to(dev)
    base = ssb.ElphHashes(make_args(K))
    base.record_stride, base.overlap_init = None, False
    t0, c0 = base.build_hash_tables(n, ei)
    f0 = base.get_subgraph_features(links, t0, c0)
    for stride, overlap, fill in ((1024, False, 'abs'), (None, True, 'abs'), (1024, True, 'abs'), (896, True, 'legacy')):
        monkeypatch.setenv('SS_B200_CSR_FILL', fill)
        eh = ssb.ElphHashes(make_args(K))
        eh.record_stride, eh.overlap_init, eh.padded_tables_min_nodes = stride, overlap, 0
        for _ in range(2):  # twice: the second build recycles the first one's memory under the side stream
            t1, c1 = eh.build_hash_tables(n, ei)
        if stride:
            assert t1.records(1).stride(0) == stride and t1.records(1).shape[1] == 768
        for k in range(K + 1):
            assert torch.equal(t1.records(k), t0.records(k)), (stride, overlap, fill, k)
        assert torch.equal(c1, c0)
        assert torch.equal(eh.get_subgraph_features(links, t1, c1), f0)
        assert torch.equal(t1[2]['minhash'], t0[2]['minhash']) and torch.equal(t1[K]['hll'], t0[K]['hll'])
    buf = io.BytesIO()
    torch.save(t1, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=True)
    assert torch.equal(back[1]['minhash'], t0[1]['minhash'].cpu())
    with pytest.raises(ValueError):
        eh.record_stride = 1000
        eh.build_hash_tables(n, ei)


def _adjacency_key(csr, n):
    """sorted (row, neighbour) multiset of a (rowptr, colidx, nnz, max_id) tuple"""
    rowptr, colidx, nnz, _ = csr
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == nnz
    rows = torch.repeat_interleave(torch.arange(rowptr.numel() - 1, device=rowptr.device), rowptr[1:] - rowptr[:-1])
    return torch.sort(rows * n + colidx[:nnz].long()).values


def test_streaming_csr_of_key_ordered_lists(monkeypatch):
    """ss_csr_sorted_chunk (one-pass CSR of lists ordered by their CSR key, the default for coalesced / to_undirected
    input) gives the same adjacency as the histogram + fill build: symmetric lists ordered by source, asymmetric lists
    ordered by destination, long runs of isolated rows, ids above max(edge_index), with and without self loops, device
    and streamed pinned-host input; lists that are not eligible (ordered by source but asymmetric, shuffled) fall back"""
    import subgraph_sketching_b200.hashing as H
    dev = torch.device(DEV)
    n = 1 << 15
    sym = rmat_edges(15, 16, 7, dev)                                   # sorted by (src, dst), symmetric
    g = torch.Generator(device=dev).manual_seed(3)
    asym = sym[:, torch.rand(sym.shape[1], generator=g, device=dev) < 0.5]           # ordered by src only
    by_dst = torch.stack([asym[1], asym[0]])                              # asymmetric, ordered by edge_index[1]
    gaps = torch.cat([sym[:, sym[0] < 100], sym[:, sym[0] > 20000] ], dim=1)           # long run of rows without edges
    gaps = gaps[:, (gaps[1] < 100) | (gaps[1] > 20000)]
    shuffled = sym[:, torch.randperm(sym.shape[1], generator=g, device=dev)]
    cases = {'symmetric': (sym, True), 'ordered by dst': (by_dst, True), 'isolated runs': (gaps, True),
             'asymmetric ordered by src': (asym, False), 'shuffled': (shuffled, False)}
    calls = []
    real = H._try_sorted_csr

    def spy(*a, **kw):
        out = real(*a, **kw)
        calls.append(out is not None)
        return out

    monkeypatch.setattr(H, '_try_sorted_csr', spy)
    for name, (ei, eligible) in cases.items():
        for rows, loops in ((n, True), (n + 777, True), (n, False)):
            monkeypatch.setenv('SS_B200_CSR_FAST', '0')
            want = ssb.build_csr(ei, dev, num_rows=rows, add_loops=loops)
            monkeypatch.setenv('SS_B200_CSR_FAST', '1')
            calls.clear()
            got = ssb.build_csr(ei, dev, num_rows=rows, add_loops=loops)
            assert calls == [eligible], (name, calls)
            assert got[2] == want[2] and got[3] == want[3], name
            assert torch.equal(got[0], want[0]), name
            assert torch.equal(_adjacency_key(got, n + 777), _adjacency_key(want, n + 777)), name
    # streamed pinned-host input: chunks of 4096 edges through the DMA ring, histogram pass riding along
    monkeypatch.setattr(H, 'INGEST_MIN_EDGES', 1 << 10)
    monkeypatch.setattr(H, 'INGEST_CHUNK', 1 << 12)
    for name in ('symmetric', 'isolated runs', 'shuffled', 'asymmetric ordered by src'):
        ei, eligible = cases[name]
        monkeypatch.setenv('SS_B200_CSR_FAST', '0')
        want = ssb.build_csr(ei, dev, num_rows=n, add_loops=True)
        monkeypatch.setenv('SS_B200_CSR_FAST', '1')
        calls.clear()
        got = ssb.build_csr(ei.cpu().pin_memory(), dev, num_rows=n, add_loops=True)
        assert calls == [eligible], (name, calls)
        assert got[2] == want[2] and torch.equal(got[0], want[0]), name
        assert torch.equal(_adjacency_key(got, n), _adjacency_key(want, n)), name
    # hop 1 under the ingest stream: a streamed, eligible list has its hop-1 rows merged block by block while later
    # chunks are still arriving (blocked launches described on the device); hubs span several 1024-edge chunks here.
    # Same tables, cards and features as the device-resident build; an ineligible list discards the speculative hop.
    monkeypatch.setattr(H, 'INGEST_CHUNK', 1 << 10)
    monkeypatch.setenv('SS_B200_CSR_FAST', '1')
    g2 = torch.Generator().manual_seed(11)
    lk = torch.randint(0, n, (5000, 2), generator=g2).to(dev)
    for K in (1, 3):
        eh = ssb.ElphHashes(make_args(K))
        t_dev, c_dev = eh.build_hash_tables(n, sym)
        f_dev = eh.get_subgraph_features(lk, t_dev, c_dev)
        blocks = []
        real_block = eh._merge_block
        eh._merge_block = lambda *a, **kw: (blocks.append(1), real_block(*a, **kw))[1]
        for name, expect_blocks in (('symmetric', True), ('isolated runs', True), ('asymmetric ordered by src', True), ('shuffled', True)):
            ei_h = cases[name][0].cpu().pin_memory()
            blocks.clear()
            t_h, c_h = eh.build_hash_tables(n, ei_h)
            assert bool(blocks) == expect_blocks, (name, len(blocks))
            t_ref, c_ref = (t_dev, c_dev) if name == 'symmetric' else eh.build_hash_tables(n, cases[name][0])
            for k in range(K + 1):
                assert torch.equal(t_h.records(k), t_ref.records(k)), (K, name, k)
            assert torch.equal(c_h.to(dev), c_ref), (K, name)
            if name == 'symmetric':
                assert torch.equal(eh.get_subgraph_features(lk, t_h, c_h.to(dev)), f_dev)
        monkeypatch.setenv('SS_B200_INGEST_OVERLAP', '0')
        blocks.clear()
        t_h, c_h = eh.build_hash_tables(n, cases['symmetric'][0].cpu().pin_memory())
        assert not blocks and torch.equal(t_h.records(K), t_dev.records(K))
        monkeypatch.delenv('SS_B200_INGEST_OVERLAP')
    # out-of-range ids still raise through the fallback
    bad = sym.clone()
    bad[1, -1] = n + 5
    with pytest.raises(IndexError):
        ssb.ElphHashes(make_args(1)).build_hash_tables(n, bad)
    # and the tables do not depend on which build produced the adjacency
    eh = ssb.ElphHashes(make_args(2))
    t1, c1 = eh.build_hash_tables(n, sym)
    monkeypatch.setenv('SS_B200_CSR_FAST', '0')
    t0, c0 = eh.build_hash_tables(n, sym)
    for k in range(3):
        assert torch.equal(t1.records(k), t0.records(k))
    assert torch.equal(c1, c0)


def test_experimental_binned_csr_fill_is_equivalent(monkeypatch):
    """SS_B200_CSR_BIN=1 (edges grouped by destination block before the fill; opt-in, measured slower) must give the same
    adjacency: equal rowptr, equal neighbour multiset per row, bit-equal tables -- device and pinned-host inputs, row ranges"""
    monkeypatch.setenv('SS_B200_CSR_FAST', '0')
    n = 1 << 15
    dev = torch.device(DEV)
    ei = rmat_edges(15, 16, 6, dev)
    ref = ssb.build_csr(ei, dev, num_rows=n, add_loops=True)
    monkeypatch.setenv('SS_B200_CSR_BIN', '1')
    monkeypatch.setenv('SS_B200_CSR_BIN_MIN_EDGES', '1')
    for inp in (ei, ei.cpu().pin_memory()):
        got = ssb.build_csr(inp, dev, num_rows=n, add_loops=True)
        assert torch.equal(got[0], ref[0]) and got[2] == ref[2] and got[3] == ref[3]
        rows = torch.repeat_interleave(torch.arange(n, device=dev), ref[0][1:] - ref[0][:-1])
        key_ref = torch.sort(rows * n + ref[1][:ref[2]].long()).values
        key_got = torch.sort(rows * n + got[1][:got[2]].long()).values
        assert torch.equal(key_ref, key_got)
    lo, hi = 1000, 9000  # a row range, as the node-sharded engine builds it
    part = ssb.build_csr(ei, dev, num_rows=hi - lo, add_loops=True, row_begin=lo)
    monkeypatch.delenv('SS_B200_CSR_BIN')
    want = ssb.build_csr(ei, dev, num_rows=hi - lo, add_loops=True, row_begin=lo)
    assert torch.equal(part[0], want[0]) and part[2] == want[2]
    rows = torch.repeat_interleave(torch.arange(hi - lo, device=dev), want[0][1:] - want[0][:-1])
    assert torch.equal(torch.sort(rows * n + part[1][:part[2]].long()).values,
                       torch.sort(rows * n + want[1][:want[2]].long()).values)
    monkeypatch.setenv('SS_B200_CSR_BIN', '1')
    eh = ssb.ElphHashes(make_args(2))
    t1, c1 = eh.build_hash_tables(n, ei)
    monkeypatch.delenv('SS_B200_CSR_BIN')
    t0, c0 = eh.build_hash_tables(n, ei)
    for k in range(3):
        assert torch.equal(t1.records(k), t0.records(k))
    assert torch.equal(c1, c0)
