"""CPU-only: the oracle restatement (oracle/sketch_oracle.py) against the golden vectors produced by the
unmodified reference (tests/golden, oracle/make_golden.py), the survey's known answers, and -- when
/root/reference is present -- the reference itself on fresh random inputs."""
import numpy as np
import pytest
import torch

from helpers import GRAPH_CASES, golden_tables, load_golden
from oracle import ref_loader, sketch_oracle as so


def oracle_for(blob, use_zero_one=False, floor_sf=False):
    p = int(blob['p'])
    thr, est, bias = golden_tables(blob)
    c = so.HllConstants(p, raw_estimate=est, bias=bias, threshold=thr)
    return so.OracleSketches(int(blob['K']), int(blob['P']), p, use_zero_one, floor_sf, constants=c)


@pytest.mark.parametrize('name', GRAPH_CASES)
def test_oracle_matches_reference_golden(name):
    blob = load_golden(name)
    o = oracle_for(blob)
    ei = torch.from_numpy(blob['edge_index'])
    tables, cards = o.build_hash_tables(int(blob['num_nodes']), ei)
    for k in range(int(blob['K']) + 1):
        assert np.array_equal(tables[k]['minhash'].numpy().astype(np.uint32), blob[f'minhash_{k}']), f'minhash hop {k}'
        assert np.array_equal(tables[k]['hll'].numpy(), blob[f'hll_{k}']), f'hll hop {k}'
    assert np.array_equal(cards.numpy(), blob['cards'])
    links = torch.from_numpy(blob['links'])
    for zo in (False, True):
        for fl in (False, True):
            o.use_zero_one, o.floor_sf = zo, fl
            f = o.subgraph_features(links, tables, cards)
            assert np.array_equal(f.numpy(), blob[f'features_zo{int(zo)}_fl{int(fl)}']), (zo, fl)
    inter = o.intersections(links, tables)
    K = int(blob['K'])
    got = torch.stack([inter[(a, b)] for a in range(1, K + 1) for b in range(1, K + 1)], dim=1)
    assert np.array_equal(got.numpy(), blob['intersections'])


def test_oracle_hll_count_golden():
    blob = load_golden('hll_count_p8')
    c = so.HllConstants(8, raw_estimate=blob['estimate_vector'], bias=blob['bias_vector'],
                        threshold=int(blob['hll_threshold']))
    got = so.hll_count(c, torch.from_numpy(blob['regs']))
    assert np.array_equal(got.numpy(), blob['counts'])
    assert np.array_equal(so.bias_of(c, torch.from_numpy(blob['e'])).numpy(), blob['bias'])


def test_survey_known_answers():
    """table-independent values captured from the verbatim reference during the survey"""
    assert so.node_hash64(1, 5).tolist() == [6238072747940578789, 15839785061582574730, 2185194620014831856,
                                             13232826040865663252, 13168350753275463132]
    assert so.node_hash64(0, 1).tolist() == [0]
    a, b = so.permutation_params(4)
    assert a.tolist() == [775169054918279404, 2109959069025162, 401325382989534145, 1130051441076870728]
    assert b.tolist() == [1758426461858698312, 965365488286768773, 1703346441743126657, 1762784241922636284]
    mh = so.minhash_init(2, 128)
    assert mh[0, :4].tolist() == [4183491429, 1571535613, 1357840683, 3557095531]
    assert mh[1, :4].tolist() == [1090388868, 3382136556, 2349435535, 3020319954]
    hll = so.hll_init(5, 8)
    got = [(int(np.nonzero(r)[0][0]), int(r.max())) for r in hll]
    assert got == [(229, 2), (138, 1), (240, 4), (20, 1), (220, 1)]
    c = so.HllConstants(8)
    assert abs(float(so.hll_count(c, torch.full((256,), 3, dtype=torch.int8))[0]) - 1471.022216797) < 1e-3
    regs = torch.cat([torch.ones(100, dtype=torch.int8), torch.zeros(156, dtype=torch.int8)])
    assert abs(float(so.hll_count(c, regs)[0]) - 126.802291870) < 1e-4
    # the float64 bit_length quirk: 2^50 and 2^50+1 report 50 bits, not 51
    assert so.bit_length_f64(np.array([2 ** 50, 2 ** 50 + 1, 2 ** 50 + 2], dtype=np.uint64)).tolist() == [50, 50, 51]
    assert so.bit_length_f64(np.arange(1000, dtype=np.uint64)).tolist() == [int(i).bit_length() for i in range(1000)]


def test_ring_known_answers():
    blob = load_golden('ring12_k3')
    o = oracle_for(blob, use_zero_one=True)
    ei = torch.from_numpy(blob['edge_index'])
    tables, cards = o.build_hash_tables(12, ei)
    assert np.allclose(cards[0].numpy(), [3.017726898, 5.049480915, 7.097474575], atol=1e-6)
    assert tables[1]['minhash'][0, :6].tolist() == [1090388868, 1571535613, 1357840683, 3020319954, 309768754,
                                                    1081767955]
    nz = torch.nonzero(tables[3]['hll'][0]).flatten().tolist()
    assert nz == [20, 109, 121, 138, 188, 229, 240]
    f = o.subgraph_features(torch.tensor([[0, 1]]), tables, cards)[0]
    want = [2.015797138, 0.745637655, 0.903434038, 0.034916878, -0.044432878, 0.075016022, 1.445956945, 1.381957054,
            -0.081554890, 0.300724983, 0.023479700, -0.352553844, -0.011308670, 0.672575474, 0.608575583]
    assert np.allclose(f.numpy(), want, atol=2e-6)


def test_merge_loop_agrees_with_scatter():
    g = torch.Generator().manual_seed(5)
    ei = so.with_self_loops(torch.randint(0, 40, (2, 200), generator=g))
    x = torch.from_numpy(so.minhash_init(50, 16))
    got = so.minhash_propagate(x, ei).numpy()
    want = so.merge_rows_loop(x.numpy(), ei.numpy(), 'min')
    assert np.array_equal(got, want)
    h = torch.from_numpy(so.hll_init(50, 6))
    assert np.array_equal(so.hll_propagate(h, ei).numpy(), so.merge_rows_loop(h.numpy(), ei.numpy(), 'max'))


@pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present (GPU box)')
@pytest.mark.parametrize('K,P,p,seed', [(2, 128, 8, 0), (3, 128, 8, 1), (1, 64, 6, 2), (3, 16, 12, 3)])
def test_oracle_matches_live_reference(K, P, p, seed):
    from argparse import Namespace
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(seed)
    n = 500
    ei = torch.randint(0, 450, (2, 4000), generator=g)
    links = torch.randint(0, n, (700, 2), generator=g)
    eh = ref.ElphHashes(Namespace(max_hash_hops=K, floor_sf=bool(seed & 1), minhash_num_perm=P, hll_p=p,
                                  use_zero_one=bool(seed & 2)))
    rt, rc = eh.build_hash_tables(n, ei)
    rf = eh.get_subgraph_features(links, rt, rc)
    c = so.HllConstants(p, raw_estimate=eh.estimate_vector.numpy(), bias=eh.bias_vector.numpy(),
                        threshold=eh.hll_threshold)
    o = so.OracleSketches(K, P, p, bool(seed & 2), bool(seed & 1), constants=c)
    ot, oc = o.build_hash_tables(n, ei)
    for k in range(K + 1):
        assert torch.equal(ot[k]['minhash'], rt[k]['minhash'])
        assert torch.equal(ot[k]['hll'], rt[k]['hll'])
    assert torch.equal(oc, rc)
    assert torch.equal(o.subgraph_features(links, ot, oc), rf)


def test_heuristics_oracle_matches_reference_golden():
    """CN / AA / RA restatement (oracle/heuristics_oracle.py) vs the unmodified reference (heuristics.py:11-71)"""
    from oracle import heuristics_oracle as ho
    blob = load_golden('heuristics')
    for name in ('ba300', 'multi'):
        A = ho.adjacency(blob[f'{name}_edge_index'], int(blob[f'{name}_n']), blob[f'{name}_weight'])
        for kind in ('cn', 'aa', 'ra'):
            got = ho.scores(A, blob[f'{name}_links'], kind)
            assert np.array_equal(got.numpy(), blob[f'{name}_{kind}']), (name, kind)


def test_sign_oracle_matches_reference_golden_and_scipy():
    """SIGN pre-propagation: the restatement is bit-equal to what the unmodified
    HashDataset._generate_sign_features produced (tests/golden/sign.npz) and agrees with the textbook
    D^-1/2 (A + I) D^-1/2 x evaluated by scipy in float64"""
    import scipy.sparse as ssp
    from oracle import sign_oracle
    blob = load_golden('sign')
    for name, ks in [('ba300_unit_f16', (0, 2)), ('multi_int_f7', (0, 3)), ('float_w_f130', (1,)),
                     ('ba200_unit_f256', (1,))]:
        x = torch.from_numpy(blob[f'{name}_x'])
        ei = torch.from_numpy(blob[f'{name}_edge_index'])
        w = torch.from_numpy(blob[f'{name}_weight'])
        n = x.shape[0]
        for k in ks:
            assert torch.equal(sign_oracle.sign_features(x, ei, w, k), torch.from_numpy(blob[f'{name}_k{k}']))
        r, c, wv = ei[0].numpy(), ei[1].numpy(), w.numpy().astype(np.float64)
        m = r != c
        loop_w = np.ones(n)
        for i, v in zip(r[~m], wv[~m]):
            loop_w[i] = v  # the last self loop of a node keeps its weight
        A = ssp.csr_matrix((wv[m], (r[m], c[m])), shape=(n, n)) + ssp.diags(loop_w)
        deg = np.asarray(A.sum(axis=0)).flatten()
        dinv = np.where(deg > 0, deg ** -0.5, 0.0)
        want = (ssp.diags(dinv) @ A @ ssp.diags(dinv)) @ x.numpy().astype(np.float64)
        got = sign_oracle.sign_features(x, ei, w, 0).numpy()
        assert np.abs(got - want).max() < 5e-6
