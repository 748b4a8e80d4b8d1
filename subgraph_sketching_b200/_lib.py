"""ctypes binding of libss_b200.so (the C ABI declared in include/ss_b200.h).

The library is the product: there is no Python or CPU fallback.  If the shared object has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C subgraph_sketching_b200/csrc`) importing
this module raises, and every compute entry point raises when no CUDA device is present.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libss_b200.so')

SS_ABI_VERSION = 14
SS_FLAG_USE_ZERO_ONE = 1
SS_FLAG_FLOOR = 2
SS_MERGE_AUTO, SS_MERGE_TMA, SS_MERGE_LDG, SS_MERGE_GENERIC, SS_MERGE_BULK = 0, 1, 2, 3, 4
MERGE_VARIANTS = {'auto': SS_MERGE_AUTO, 'tma': SS_MERGE_TMA, 'ldg': SS_MERGE_LDG, 'generic': SS_MERGE_GENERIC,
                  'bulk': SS_MERGE_BULK}

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_ptr = ctypes.c_void_p


class HllConsts(ctypes.Structure):
    """struct ss_hll_consts"""
    _fields_ = [('p', ctypes.c_int32), ('table_len', ctypes.c_int32), ('monotone', ctypes.c_int32),
                ('threshold', ctypes.c_float), ('alpha_m2', ctypes.c_float), ('five_m', ctypes.c_float),
                ('lc_table', c_ptr), ('raw_estimate', c_ptr), ('bias', c_ptr)]


SS_LAYOUT_FULL, SS_LAYOUT_MINHASH, SS_LAYOUT_HLL, SS_LAYOUT_HALF = 0, 1, 2, 3
LAYOUT_BYTES = {SS_LAYOUT_FULL: 768, SS_LAYOUT_MINHASH: 512, SS_LAYOUT_HLL: 256, SS_LAYOUT_HALF: 384}


class MergeDesc(ctypes.Structure):
    """struct ss_merge_desc"""
    _fields_ = [('rowptr', ctypes.c_void_p), ('colidx', ctypes.c_void_p), ('n_rows', ctypes.c_int64), ('nnz', ctypes.c_int64),
                ('rec_in', ctypes.c_void_p), ('in_rows', ctypes.c_int64), ('in_stride', ctypes.c_int64),
                ('rec_out', ctypes.c_void_p), ('out_stride', ctypes.c_int64),
                ('num_perm', ctypes.c_int32), ('hll_p', ctypes.c_int32), ('layout', ctypes.c_int32),
                ('variant', ctypes.c_int32), ('workspace', ctypes.c_void_p), ('workspace_bytes', ctypes.c_int64),
                ('cards_out', ctypes.c_void_p), ('cards_stride', ctypes.c_int64), ('hc', ctypes.c_void_p),
                ('n_peers', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('peer_rec_out', ctypes.c_void_p), ('peer_cards_out', ctypes.c_void_p), ('peer_mask', ctypes.c_void_p),
                ('mc_rec_out', ctypes.c_void_p), ('mc_cards_out', ctypes.c_void_p), ('guard', ctypes.c_void_p),
                ('block', ctypes.c_void_p)]


class ShardView(ctypes.Structure):
    """struct ss_shard_view"""
    _fields_ = [('n_ranks', ctypes.c_int32), ('rank', ctypes.c_int32), ('last_hop_own_only', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('bounds', ctypes.c_int64 * 9), ('peer_records', (ctypes.c_void_p * 8) * 4),
                ('local_rows', ctypes.c_void_p)]


class HopView(ctypes.Structure):
    """struct ss_hop_view"""
    _fields_ = [('records', c_ptr), ('row_stride', c_i64), ('num_rows', c_i64)]


# name -> (restype, argtypes); must list every symbol declared in include/ss_b200.h
SIGNATURES = {
    'ss_version': (c_int, []),
    'ss_last_error': (ctypes.c_char_p, []),
    'ss_device_info': (c_int, [ctypes.POINTER(c_int)] * 3),
    'ss_record_bytes': (c_i64, [c_int, c_int]),
    'ss_init_records': (c_int, [c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    'ss_pack_records': (c_int, [c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_i64, c_ptr]),
    'ss_unpack_records': (c_int, [c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    'ss_pack_records_ex': (c_int, [c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    'ss_unpack_records_ex': (c_int, [c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    'ss_csr_build_nosync': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    'ss_i64_differs': (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    'ss_csr_workspace_bytes': (c_i64, [c_i64]),
    'ss_csr_rowptr': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    'ss_csr_degree_chunk': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_int, c_ptr]),
    'ss_csr_rowptr_finish': (c_int, [c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    'ss_csr_fill': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64,
                            c_ptr]),
    'ss_csr_sorted_chunk': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_int, c_i64, ctypes.c_uint64, ctypes.c_uint64, c_ptr,
                                    c_ptr, c_ptr, c_ptr, c_ptr]),
    'ss_csr_sorted_finish': (c_int, [c_i64, c_i64, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'ss_csr_sorted_chunk_rows': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_i64, ctypes.c_uint64,
                                         ctypes.c_uint64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'ss_csr_sorted_finish_rows': (c_int, [c_i64, c_i64, c_i64, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'ss_csr_sorted_bounds': (c_int, [c_ptr, c_i64, c_i64, ctypes.c_double, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    'ss_csr_sorted_block': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    'ss_mark_rows': (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    'ss_halo_from_csr': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    'ss_csr_bin_workspace_bytes': (c_i64, []),
    'ss_csr_bin_edges': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64,
                                 c_ptr]),
    'ss_merge_workspace_bytes': (c_i64, [c_i64, c_int, c_int]),
    'ss_khop_merge': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_int, c_int, c_ptr, c_i64,
                              c_ptr, c_i64, ctypes.POINTER(HllConsts), c_int, c_ptr]),
    'ss_khop_merge_peers': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_int, c_int, c_ptr,
                                    c_i64, c_ptr, c_i64, ctypes.POINTER(HllConsts), c_int, c_int,
                                    ctypes.POINTER(c_ptr), ctypes.POINTER(c_ptr), c_ptr, c_ptr, c_ptr]),
    'ss_khop_merge_ex': (c_int, [ctypes.POINTER(MergeDesc), c_ptr]),
    'ss_prop_min_i64': (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_ptr]),
    'ss_prop_min_i64_guarded': (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    'ss_prop_max_i8': (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_ptr]),
    'ss_hll_count': (c_int, [c_ptr, c_i64, c_i64, ctypes.POINTER(HllConsts), c_ptr, c_i64, c_ptr]),
    'ss_estimate_bias': (c_int, [c_ptr, c_i64, ctypes.POINTER(HllConsts), c_ptr, c_ptr]),
    'ss_jaccard_i64': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    'ss_max_i8': (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    'ss_col_sums': (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    'ss_common_neighbour_scores': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    'ss_link_features': (c_int, [c_ptr, c_i64, ctypes.POINTER(HopView), c_int, c_int, c_int, c_ptr, c_i64,
                                 ctypes.POINTER(HllConsts), c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    'ss_link_features_sharded': (c_int, [c_ptr, c_i64, ctypes.POINTER(HopView), c_int, c_int, c_int, c_ptr, c_i64,
                                         ctypes.POINTER(HllConsts), c_int, c_ptr, c_ptr, c_ptr, ctypes.POINTER(ShardView), c_ptr]),
    'ss_sign_workspace_bytes': (c_i64, [c_i64]),
    'ss_gcn_norm': (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    'ss_sign_fill': (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    'ss_sign_spmm': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_int,
                             c_ptr]),
}


class SketchLibError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise SketchLibError(
            f'{LIB_PATH} is missing: build the CUDA library first (make -C {os.path.join(_HERE, "csrc")} or '
            '__graft_entry__.build()).  There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    ver = lib.ss_version()
    if ver != SS_ABI_VERSION:
        raise SketchLibError(f'libss_b200.so ABI version {ver} != expected {SS_ABI_VERSION}; rebuild it')
    return lib


# kernels enqueued by one call of each entry point (for the bench's `gpu_launches` claim); entry points
# that return early on empty input are counted by the caller's own bookkeeping
KERNELS_PER_CALL = {
    'ss_init_records': 1, 'ss_pack_records': 1, 'ss_unpack_records': 1, 'ss_csr_rowptr': 4, 'ss_csr_degree_chunk': 1, 'ss_csr_rowptr_finish': 3, 'ss_csr_fill': 2, 'ss_csr_bin_edges': 1, 'ss_csr_sorted_chunk': 1, 'ss_csr_sorted_finish': 1, 'ss_csr_sorted_chunk_rows': 1, 'ss_csr_sorted_finish_rows': 1,
    'ss_csr_sorted_bounds': 1, 'ss_csr_sorted_block': 1, 'ss_mark_rows': 1, 'ss_halo_from_csr': 1, 'ss_link_features_sharded': 1,
    'ss_khop_merge': 2, 'ss_khop_merge_peers': 2, 'ss_khop_merge_ex': 2, 'ss_pack_records_ex': 1, 'ss_unpack_records_ex': 1,
    'ss_csr_build_nosync': 8, 'ss_i64_differs': 1, 'ss_prop_min_i64': 1, 'ss_prop_min_i64_guarded': 1, 'ss_prop_max_i8': 1, 'ss_hll_count': 1, 'ss_estimate_bias': 1,
    'ss_jaccard_i64': 1, 'ss_max_i8': 1, 'ss_link_features': 1, 'ss_col_sums': 1, 'ss_common_neighbour_scores': 1,
    'ss_gcn_norm': 4, 'ss_sign_fill': 1, 'ss_sign_spmm': 1,
}


class _CountingLib(object):
    """thin proxy over the CDLL that counts kernel launches per entry point"""

    def __init__(self, cdll):
        self._cdll = cdll
        self.launches = 0
        self.calls = {}
        for name in SIGNATURES:
            setattr(self, name, self._wrap(name, getattr(cdll, name)))

    def _wrap(self, name, fn):
        k = KERNELS_PER_CALL.get(name, 0)
        if k == 0:
            return fn

        def counted(*args):
            rc = fn(*args)
            if rc == 0:
                self.launches += k
                self.calls[name] = self.calls.get(name, 0) + 1
            return rc
        return counted

    def reset_counters(self):
        self.launches = 0
        self.calls = {}


lib = _CountingLib(_load())


def check(rc, what=''):
    """raise on a negative return code from the library"""
    if rc is not None and rc < 0:
        msg = lib.ss_last_error()
        msg = msg.decode() if msg else ''
        if rc == -1:
            raise ValueError(f'{what}: {msg}')
        raise SketchLibError(f'{what} failed (code {rc}): {msg}')
    return rc


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SketchLibError('subgraph_sketching_b200 needs a CUDA device (sm_100a); none is visible and there is '
                             'no CPU fallback')
