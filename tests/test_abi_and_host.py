"""CPU-only: the C-ABI library loads and exports every symbol include/ss_b200.h declares; host-side logic of
the ElphHashes mirror (constants, permutations, error behaviour).  No compute call is made without a GPU."""
import os
import re
from argparse import Namespace

import numpy as np
import pytest
import torch

import subgraph_sketching_b200 as ssb
from subgraph_sketching_b200 import _lib
from helpers import make_args
from oracle import sketch_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'ss_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ss_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(_lib.lib, s), f'{s} declared in ss_b200.h but not exported by libss_b200.so'
        assert s in _lib.SIGNATURES, f'{s} has no ctypes signature in _lib.py'
    assert sorted(_lib.SIGNATURES) == syms
    assert _lib.lib.ss_version() == _lib.SS_ABI_VERSION


def test_record_geometry_and_argument_errors():
    assert _lib.lib.ss_record_bytes(128, 8) == 768
    assert _lib.lib.ss_record_bytes(33, 6) == 4 * 36 + 64
    assert _lib.lib.ss_record_bytes(8, 4) == 48
    assert _lib.lib.ss_record_bytes(128, 3) < 0
    assert b'unsupported' in _lib.lib.ss_last_error()
    with pytest.raises(ValueError):
        _lib.check(_lib.lib.ss_record_bytes(0, 8), 'ss_record_bytes')
    assert _lib.lib.ss_csr_workspace_bytes(1000) >= 4000
    assert _lib.lib.ss_merge_workspace_bytes(10, 33, 6) == 16


def test_argument_validation_needs_no_gpu():
    """every entry point validates sizes / pointers before it touches CUDA: bad arguments come back as
    SS_ERR_INVALID (-1) with a message, on a machine without a GPU too"""
    lib = _lib.lib
    assert lib.ss_sign_workspace_bytes(1000) >= 8000 and lib.ss_sign_workspace_bytes(-1) < 0
    assert lib.ss_gcn_norm(None, None, None, -1, 10, None, None, None, None, 0, None) == -1
    assert b'negative' in lib.ss_last_error()
    assert lib.ss_gcn_norm(None, None, None, 1 << 31, 10, None, None, None, None, 0, None) == -1
    assert b'2^31' in lib.ss_last_error()
    assert lib.ss_gcn_norm(None, None, None, 5, 10, None, None, None, None, 0, None) == -1      # null outputs
    assert lib.ss_sign_fill(None, 5, 10, None, None, None, 0, None) == -1
    assert lib.ss_sign_spmm(None, None, None, None, None, None, None, 4, 10, 8, None, 8, 1, None) == -1
    assert lib.ss_sign_spmm(None, None, None, None, None, None, None, 8, 10, 8, None, 8, 0, None) == -1  # copies < 1
    assert lib.ss_sign_spmm(None, None, None, None, None, None, None, 8, 0, 8, None, 8, 1, None) == 0   # empty: no-op
    assert lib.ss_csr_degree_chunk(None, None, -1, 0, 10, None, None, None, None, 0, 1, None) == -1
    assert lib.ss_csr_degree_chunk(None, None, 4, 0, 10, None, None, None, None, 0, 1, None) == -1      # null workspace
    assert lib.ss_csr_rowptr_finish(-1, 0, 10, None, None, None, 0, None) == -1
    assert lib.ss_csr_rowptr(None, None, 0, 0, 0, 10, None, None, None, None, None, 0, None) == -1
    assert lib.ss_khop_merge(None, None, 5, 0, None, 5, 768, None, 768, 128, 3, None, 0, None, 0, None, 0, None) == -1
    assert b'unsupported sketch shape' in lib.ss_last_error()
    assert lib.ss_link_features(None, 5, None, 4, 128, 8, None, 0, None, 0, None, None, None, None) == -1
    assert b'1, 2 and 3 hop' in lib.ss_last_error()
    with pytest.raises(ValueError):
        _lib.check(lib.ss_sign_fill(None, -3, 10, None, None, None, 0, None), 'ss_sign_fill')


def test_constructor_mirrors_reference():
    for bad in (0, 4):
        with pytest.raises(AssertionError):
            ssb.ElphHashes(make_args(K=bad))
    eh = ssb.ElphHashes(make_args(K=3, use_zero_one=True))
    assert eh.max_hops == 3 and eh.num_perm == 128 and eh.p == 8 and eh.m == 256
    assert eh.max_rank == 56 and eh.hll_size == 256 and eh.use_zero_one is True and eh.floor_sf is False
    assert eh.label_lookup == ssb.LABEL_LOOKUP[3]
    assert [len(ssb.LABEL_LOOKUP[k]) for k in (1, 2, 3)] == [3, 8, 15]
    assert abs(eh.alpha - 0.7182725932495458) < 1e-15
    assert eh.hll_threshold == 220
    assert eh.estimate_vector.dtype == torch.float32 and eh.bias_vector.shape == eh.estimate_vector.shape
    assert int(eh._max_minhash) == 2 ** 32 - 1 and int(eh._mersenne_prime) == 2 ** 61 - 1
    assert callable(eh.minhash_prop) and callable(eh.hll_prop)


def test_host_helpers_match_oracle():
    eh = ssb.ElphHashes(make_args())
    ab = eh._init_permutations(128)
    a, b = so.permutation_params(128)
    assert ab.dtype == np.uint64 and np.array_equal(ab[0], a) and np.array_equal(ab[1], b)
    bits = np.arange(1000, dtype=np.uint64)
    assert eh._np_bit_length(bits).tolist() == [int(i).bit_length() for i in range(1000)]
    assert eh._get_hll_rank(np.array([1, 2, 3], dtype=np.uint64)).tolist() == [56, 55, 55]
    with pytest.raises(ValueError):
        eh._get_hll_rank(np.array([2 ** 57], dtype=np.uint64))
    # linear-counting table == the reference expression evaluated by torch (hashing.py:195)
    c = so.HllConstants(8)
    nz = torch.arange(1, 257)
    assert torch.equal(eh._linearcounting(nz), so.linear_counting(c, nz))


def test_log2_window_reproduces_float64_bit_length():
    w = ssb.log2_window_table()
    assert w[:49].sum() == 0 and w[50] == 2
    for k in (49, 50, 53, 56, 59):
        base = 2 ** k
        for j in (0, 1, int(w[k]), int(w[k]) + 1, 2 * int(w[k]) + 3):
            want = int(so.bit_length_f64(np.array([base + j - 1], dtype=np.uint64))[0])
            excess = j
            got = k if (excess == 0 or excess <= w[k]) else k + 1
            assert got == want, (k, j)


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    eh = ssb.ElphHashes(make_args())
    with pytest.raises(_lib.SketchLibError):
        eh.build_hash_tables(4, torch.tensor([[0, 1], [1, 0]]))
    with pytest.raises(_lib.SketchLibError):
        eh.hll_count(torch.zeros(256, dtype=torch.int8))
    with pytest.raises(_lib.SketchLibError):
        eh.initialise_minhash(3)


def test_neighbouring_rows_fail_loudly_without_gpu_and_keep_reference_cache_names():
    """SIGN / heuristics: no CPU fallback either; cache file names are the reference's
    (datasets/elph.py:120-123)"""
    from subgraph_sketching_b200 import sign as bs
    assert bs.feature_cache_name('root/ds', 'train', 0) == 'root/ds_train_featurecache.pt'
    assert bs.feature_cache_name('root/ds', 'test', 2) == 'root/ds_test_k2_featurecache.pt'
    with pytest.raises(ValueError):
        bs.sign_features(torch.zeros(4), torch.zeros((2, 0), dtype=torch.int64), None, 0)   # x must be 2-D
    with pytest.raises(ValueError):
        bs.sign_features(torch.zeros(4, 2), torch.zeros((2, 0), dtype=torch.int64), None, -1)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SketchLibError):
            bs.sign_features(torch.zeros(4, 2), torch.tensor([[0, 1], [1, 0]]), None, 1)
        from subgraph_sketching_b200 import heuristics as bh
        with pytest.raises(_lib.SketchLibError):
            bh.SortedAdjacency.from_edge_index(torch.tensor([[0, 1], [1, 0]]), 2)


def test_shape_errors_precede_device_work():
    eh = ssb.ElphHashes(make_args())
    with pytest.raises(ValueError):
        eh.jaccard(torch.zeros(3, 128), torch.zeros(4, 128))
    with pytest.raises(ValueError):
        eh._hll_merge(torch.zeros(3, 256), torch.zeros(3, 255))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'subgraph_sketching_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert '/root/reference' not in text.replace('/root/reference/src', 'REFDOC'), f


def test_packaged_hllpp_tables_warn_and_are_labelled(monkeypatch):
    """without datasketch the Monte-Carlo substitute tables are used: loudly (RuntimeWarning) and labelled on the
    engine; explicit tables are labelled as the caller's"""
    import warnings
    import subgraph_sketching_b200.hashing as H
    try:
        import datasketch  # noqa: F401
        return  # the real constants are importable here: nothing to warn about
    except ImportError:
        pass
    monkeypatch.setattr(H, '_warned_tables', set())
    monkeypatch.delenv('SS_B200_HLLPP_TABLES', raising=False)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        eh = ssb.ElphHashes(make_args())
    assert any(issubclass(x.category, RuntimeWarning) and 'Monte-Carlo' in str(x.message) for x in w)
    assert eh.hll_tables_source == 'packaged-monte-carlo'
    thr, est, bias = H.hllpp_tables(8)
    eh2 = ssb.ElphHashes(make_args(), hll_tables=(thr, est, bias))
    assert eh2.hll_tables_source == 'caller'


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/ss_b200.h compiles as C99 (no C++ in the boundary) and the structs passed by pointer have the same size
    in C and in the ctypes binding (a field added on one side only would shift everything behind it)"""
    import ctypes
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    src = tmp_path / 'abi.c'
    src.write_text('#include "ss_b200.h"\n#include <stdio.h>\nint main(void) { printf("%zu %zu %zu %zu %d\\n", '
                   'sizeof(ss_merge_desc), sizeof(ss_shard_view), sizeof(ss_hll_consts), sizeof(ss_hop_view), '
                   'SS_ABI_VERSION); return 0; }\n')
    exe = tmp_path / 'abi'
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [ctypes.sizeof(_lib.MergeDesc), ctypes.sizeof(_lib.ShardView),
                                     ctypes.sizeof(_lib.HllConsts), ctypes.sizeof(_lib.HopView), _lib.SS_ABI_VERSION]
