"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the subgraph-sketch hot path.

This file restates, in numpy (integer/bit work) and torch-CPU float32 (float tails), what
/root/reference/src/hashing.py computes.  It is the checker for the CUDA engine and the timed CPU
baseline of bench.py; it is never imported by the product package.  It is written from the behaviour of
the reference (cited file:line below), not copied from it: sketches are built by free functions over
arrays, the k-hop merge is a scatter-amax over the COO edge list (what PyG's `aggr='max'` lowers to), and
the inclusion-exclusion feature algebra is table-driven (`FEATURE_RECIPES`).

Parity status (see oracle/__init__.py): pinned against the unmodified reference run in the build
container (tests/golden/*.npz via oracle/make_golden.py, and live in tests/test_oracle_vs_reference.py
when /root/reference is present) and against SURVEY.md section 8(c) known answers.  The HLL++ bias
tables are runtime inputs (datasketch is absent here) -> the bias regime is pinned only relative to the
packaged tables.

Float rules that matter for bit-parity with the reference on CPU (all float32):
  * `python_scalar / tensor` in torch is `tensor.reciprocal() * scalar`  (two roundings)
  * int64 / python_int is a true float32 division
  * the reference's ops are replayed in the same order with the same torch functions.
"""
from __future__ import annotations

import os

import numpy as np
import torch

MERSENNE_61 = (1 << 61) - 1
MASK_32 = (1 << 32) - 1
MINHASH_SEED = 1  # hashing.py:61

_TABLES_PATH = os.environ.get('SS_B200_HLLPP_TABLES') or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), '..', 'subgraph_sketching_b200', 'data', 'hllpp_tables.npz')

# feature index -> (hops from u, hops from v); hashing.py:22-25
LABELS = {
    1: [(1, 1), (0, 1), (1, 0)],
    2: [(1, 1), (2, 1), (1, 2), (2, 2), (0, 1), (1, 0), (0, 2), (2, 0)],
    3: [(1, 1), (2, 1), (1, 2), (2, 2), (3, 1), (1, 3), (3, 2), (2, 3), (3, 3),
        (0, 1), (1, 0), (0, 2), (2, 0), (0, 3), (3, 0)],
}

# Inclusion-exclusion recipes, one row per output feature: (head, [terms subtracted left to right]).
#   head : ('I', k1, k2) intersection estimate | ('cu', k) k-hop cardinality of u | ('cv', k) of v
#   term : j -> feature column j | ('S', w) -> sum of feature columns [0, w)
# Transcribed from the *behaviour* of hashing.py:276-307, including its asymmetries:
#   K=2 row 7 subtracts column 0 twice (hashing.py:287-288), K=3 row 12 uses S5 (contains column 4, the
#   (3,1) feature) where symmetry would suggest column 5 (hashing.py:302-303).
FEATURE_RECIPES = {
    1: [(('I', 1, 1), []),
        (('cv', 1), [0]),
        (('cu', 1), [0])],
    2: [(('I', 1, 1), []),
        (('I', 2, 1), [0]),
        (('I', 1, 2), [0]),
        (('I', 2, 2), [0, 1, 2]),
        (('cv', 1), [('S', 2)]),
        (('cu', 1), [0, 2]),
        (('cv', 2), [('S', 5)]),
        (('cu', 2), [0, ('S', 4), 5])],
    3: [(('I', 1, 1), []),
        (('I', 2, 1), [0]),
        (('I', 1, 2), [0]),
        (('I', 2, 2), [0, 1, 2]),
        (('I', 3, 1), [0, 1]),
        (('I', 1, 3), [0, 2]),
        (('I', 3, 2), [('S', 4), 4]),
        (('I', 2, 3), [('S', 4), 5]),
        (('I', 3, 3), [('S', 8)]),
        (('cv', 1), [0, 1, 4]),
        (('cu', 1), [0, 2, 5]),
        (('cv', 2), [('S', 5), 6, 9]),
        (('cu', 2), [('S', 5), 7, 10]),
        (('cv', 3), [('S', 9), 9, 11]),
        (('cu', 3), [('S', 9), 10, 12])],
}
# columns zeroed when use_zero_one is false (hashing.py:310-318): none for K=1
KNOCKOUT_COLUMNS = {1: [], 2: [4, 5], 3: [4, 5, 11, 12]}


# ------------------------------------------------------------------------------------------------
# integer sketches
# ------------------------------------------------------------------------------------------------
def node_hash64(first_id: int, count: int) -> np.ndarray:
    """64-bit hash of the integer ids first_id .. first_id+count-1.

    Restates pandas.util.hash_array for int64 input (pandas/core/util/hashing.py::_hash_ndarray): a
    splitmix64-style finaliser in wrapping uint64 arithmetic.  The reference hashes ids 1..n
    (hashing.py:121,128) because 0 hashes to 0."""
    v = np.arange(first_id, first_id + count, dtype=np.uint64)
    v ^= v >> np.uint64(30)
    v *= np.uint64(0xBF58476D1CE4E5B9)
    v ^= v >> np.uint64(27)
    v *= np.uint64(0x94D049BB133111EB)
    v ^= v >> np.uint64(31)
    return v


def permutation_params(num_perm: int, seed: int = MINHASH_SEED):
    """(a, b) of the num_perm affine maps; legacy RandomState stream, a then b per permutation
    (hashing.py:106-116)."""
    gen = np.random.RandomState(seed)
    a = np.empty(num_perm, dtype=np.uint64)
    b = np.empty(num_perm, dtype=np.uint64)
    for j in range(num_perm):
        a[j] = gen.randint(1, MERSENNE_61, dtype=np.uint64)
        b[j] = gen.randint(0, MERSENNE_61, dtype=np.uint64)
    return a, b


def minhash_init(n_nodes: int, num_perm: int, first_id: int = 1) -> np.ndarray:
    """hop-0 MinHash signatures, int64 [n, P] (hashing.py:118-124).

    value = (((a*h + b) mod 2^64) mod (2^61-1)) & (2^32-1); the product wraps in uint64 before the
    Mersenne reduction -- that is the reference's bit pattern."""
    a, b = permutation_params(num_perm)
    h = node_hash64(first_id, n_nodes)
    with np.errstate(over='ignore'):
        lin = h[:, None] * a[None, :] + b[None, :]
    sig = (lin % np.uint64(MERSENNE_61)) & np.uint64(MASK_32)
    return sig.astype(np.int64)


def bit_length_f64(bits: np.ndarray) -> np.ndarray:
    """ceil(log2(bits + 1)) evaluated in float64, as the reference does (hashing.py:83-89)."""
    return np.ceil(np.log2(bits + 1)).astype(int)


def hll_init(n_nodes: int, p: int, first_id: int = 1) -> np.ndarray:
    """hop-0 HLL registers, int8 [n, 2^p], exactly one non-zero per row (hashing.py:126-137)."""
    m = 1 << p
    h = node_hash64(first_id, n_nodes)
    slot = (h & np.uint64(m - 1)).astype(np.int64)
    rest = h >> np.uint64(p)
    rank = (64 - p) - bit_length_f64(rest) + 1  # hashing.py:91-104
    if rank.size and rank.min() <= 0:
        raise ValueError('Hash value overflow, maximum size is %d bits' % (64 - p))
    regs = np.zeros((n_nodes, m), dtype=np.int8)
    regs[np.arange(n_nodes), slot] = rank.astype(np.int8)
    return regs


def with_self_loops(edge_index: torch.Tensor) -> torch.Tensor:
    """append (i, i) for i < max(edge_index)+1 -- PyG add_self_loops without num_nodes (hashing.py:148)."""
    n_loop = int(edge_index.max()) + 1 if edge_index.numel() else 0
    loops = torch.arange(n_loop, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, torch.stack([loops, loops])], dim=1)


def scatter_amax(x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """out[i] = max over edges (j -> i) of x[j]; 0 where i has no in-edge (PyG aggr='max')."""
    out = torch.zeros_like(x)
    gathered = x.index_select(0, edge_index[0])
    where = edge_index[1].view(-1, 1).expand(-1, x.size(1))
    out.scatter_reduce_(0, where, gathered, reduce='amax', include_self=False)
    return out


def hll_propagate(regs: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """register-wise max over in-neighbours (hashing.py:38-45)"""
    return scatter_amax(regs, edge_index)


def minhash_propagate(sig: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """element-wise min over in-neighbours, computed as -max(-x) (hashing.py:28-35)"""
    return -scatter_amax(-sig, edge_index)


def merge_rows_loop(x: np.ndarray, edge_index: np.ndarray, op: str) -> np.ndarray:
    """slow, obviously-correct pure-python merge for tiny graphs (independent check of scatter_amax)"""
    out = np.zeros_like(x)
    seen = np.zeros(x.shape[0], dtype=bool)
    for j, i in zip(edge_index[0].tolist(), edge_index[1].tolist()):
        if not seen[i]:
            out[i] = x[j]
            seen[i] = True
        elif op == 'min':
            out[i] = np.minimum(out[i], x[j])
        else:
            out[i] = np.maximum(out[i], x[j])
    return out


# ------------------------------------------------------------------------------------------------
# HLL++ constants and cardinality
# ------------------------------------------------------------------------------------------------
class HllConstants(object):
    """alpha, threshold and the bias-correction tables for one precision p (hashing.py:69-80)"""

    def __init__(self, p: int, raw_estimate=None, bias=None, threshold=None):
        self.p = p
        self.m = 1 << p
        if p == 4:
            self.alpha = 0.673
        elif p == 5:
            self.alpha = 0.697
        elif p == 6:
            self.alpha = 0.709
        else:
            self.alpha = 0.7213 / (1.0 + 1.079 / self.m)
        self.max_rank = 64 - p
        if raw_estimate is None or bias is None or threshold is None:
            t_thr, t_est, t_bias = load_tables(p)
            threshold = t_thr if threshold is None else threshold
            raw_estimate = t_est if raw_estimate is None else raw_estimate
            bias = t_bias if bias is None else bias
        self.threshold = threshold
        self.estimate_vector = torch.as_tensor(np.asarray(raw_estimate), dtype=torch.float32)
        self.bias_vector = torch.as_tensor(np.asarray(bias), dtype=torch.float32)


def load_tables(p: int):
    """(threshold, raw_estimate[T], bias[T]) from datasketch if importable, else the packaged file"""
    try:
        from datasketch import hyperloglog_const as hc  # noqa
        return hc._thresholds[p - 4], list(hc._raw_estimate[p - 4]), list(hc._bias[p - 4])
    except ImportError:
        blob = np.load(_TABLES_PATH)
        return int(blob['thresholds'][p - 4]), blob[f'raw_estimate_p{p}'], blob[f'bias_p{p}']


def linear_counting(c: HllConstants, num_zero: torch.Tensor) -> torch.Tensor:
    """m * log(m / V)  (hashing.py:194-195); `m / V` is reciprocal-then-multiply inside torch"""
    return c.m * torch.log(c.m / num_zero)


def bias_of(c: HllConstants, e: torch.Tensor) -> torch.Tensor:
    """mean bias of the 6 table entries whose raw estimate is nearest to e (hashing.py:197-204)"""
    d2 = (e.unsqueeze(-1) - c.estimate_vector.to(e.device)) ** 2
    nearest = torch.argsort(d2)[:, :6]
    return torch.mean(c.bias_vector.to(e.device)[nearest], dim=1)


def hll_count(c: HllConstants, regs: torch.Tensor) -> torch.Tensor:
    """HyperLogLog++ cardinality of each register row, float32 [n] (hashing.py:212-232, 206-210)"""
    if regs.dim() == 1:
        regs = regs.unsqueeze(0)
    out = torch.ones(regs.shape[0], device=regs.device) * c.threshold + 1
    num_zero = c.m - torch.count_nonzero(regs, dim=1)
    some_zero = num_zero > 0
    out[some_zero] = linear_counting(c, num_zero[some_zero])
    use_raw = out > c.threshold
    e = (c.alpha * c.m ** 2) / torch.sum(2.0 ** (-regs[use_raw]), dim=1)
    small = e <= 5 * c.m
    correction = bias_of(c, e)
    e[small] = e[small] - correction[small]
    out[use_raw] = e
    return out


def jaccard(sig_u: torch.Tensor, sig_v: torch.Tensor, num_perm: int) -> torch.Tensor:
    """fraction of equal MinHash slots (hashing.py:247-256)"""
    if sig_u.shape != sig_v.shape:
        raise ValueError('source and destination hash value shapes must be the same')
    return torch.count_nonzero(sig_u == sig_v, dim=-1) / num_perm


# ------------------------------------------------------------------------------------------------
# the operator-level object (mirrors the reference surface so parity tests read like its tests)
# ------------------------------------------------------------------------------------------------
class OracleSketches(object):
    def __init__(self, max_hops=2, num_perm=128, p=8, use_zero_one=False, floor_sf=False, constants=None):
        assert max_hops in (1, 2, 3)
        self.max_hops = max_hops
        self.num_perm = num_perm
        self.p = p
        self.m = 1 << p
        self.use_zero_one = use_zero_one
        self.floor_sf = floor_sf
        self.c = constants if constants is not None else HllConstants(p)

    # -- tables ------------------------------------------------------------------------------
    def build_hash_tables(self, num_nodes, edge_index):
        """hashing.py:139-165 -> ({k: {'hll', 'minhash'}}, cards f32 [N, K])"""
        ei = with_self_loops(edge_index)
        cards = torch.zeros((num_nodes, self.max_hops))
        tables = {0: {'minhash': torch.from_numpy(minhash_init(num_nodes, self.num_perm)),
                      'hll': torch.from_numpy(hll_init(num_nodes, self.p))}}
        for k in range(1, self.max_hops + 1):
            tables[k] = {'hll': hll_propagate(tables[k - 1]['hll'], ei),
                         'minhash': minhash_propagate(tables[k - 1]['minhash'], ei)}
            cards[:, k - 1] = hll_count(self.c, tables[k]['hll'])
        return tables, cards

    # -- pairwise ----------------------------------------------------------------------------
    def intersections(self, links, tables):
        """{(k1,k2): jaccard * |union|}  (hashing.py:167-189)"""
        u, v = links[:, 0], links[:, 1]
        out = {}
        for k1 in range(1, self.max_hops + 1):
            for k2 in range(1, self.max_hops + 1):
                j = jaccard(tables[k1]['minhash'][u], tables[k2]['minhash'][v], self.num_perm)
                union = torch.maximum(tables[k1]['hll'][u], tables[k2]['hll'][v])
                out[(k1, k2)] = j * hll_count(self.c, union)
        return out

    def subgraph_features(self, links, tables, cards, batch_size=11000000):
        """hashing.py:258-323 -> float32 [L, K(K+2)]"""
        if links.dim() == 1:
            links = links.unsqueeze(0)
        K = self.max_hops
        chunks = []
        for lo in range(0, links.size(0), batch_size):
            part = links[lo:lo + batch_size]
            inter = self.intersections(part, tables)
            cu = cards.to(part.device)[part[:, 0]]
            cv = cards.to(part.device)[part[:, 1]]
            f = torch.zeros((part.size(0), K * (K + 2)), dtype=torch.float32, device=part.device)
            for col, (head, terms) in enumerate(FEATURE_RECIPES[K]):
                if head[0] == 'I':
                    val = inter[(head[1], head[2])]
                elif head[0] == 'cu':
                    val = cu[:, head[1] - 1]
                else:
                    val = cv[:, head[1] - 1]
                for t in terms:
                    if isinstance(t, tuple):
                        val = val - torch.sum(f[:, 0:t[1]], dim=1)
                    else:
                        val = val - f[:, t]
                f[:, col] = val
            if not self.use_zero_one:
                for col in KNOCKOUT_COLUMNS[K]:
                    f[:, col] = 0
            if self.floor_sf:
                f[f < 0] = 0
            chunks.append(f)
        if not chunks:
            return torch.zeros((0, K * (K + 2)), dtype=torch.float32)
        return torch.cat(chunks, dim=0)
