"""Experiment (GPU only): does pinning the hub rows in L2 speed up the k-hop merge?
Relabels an R-MAT graph by descending degree so that the hottest rows are contiguous at the start of the table,
then times one hop (a) as is, (b) with a persisting-L2 access-policy window over the first X MB of the
previous-hop table.  usage: python tools/exp_l2_persist.py [scale] [window_mb ...]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402
from subgraph_sketching_b200.graphs import rmat_edges  # noqa: E402
from argparse import Namespace  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
windows = [float(x) for x in sys.argv[2:]] or [32, 64, 96]
dev = torch.device('cuda', 0)
n = 1 << scale
cudart = ctypes.CDLL('libcudart.so.12')


class Window(ctypes.Structure):
    _fields_ = [('base_ptr', ctypes.c_void_p), ('num_bytes', ctypes.c_size_t), ('hitRatio', ctypes.c_float),
                ('hitProp', ctypes.c_int), ('missProp', ctypes.c_int), ('pad', ctypes.c_char * 36)]


def set_window(stream, ptr, nbytes, ratio=1.0):
    w = Window()
    w.base_ptr, w.num_bytes, w.hitRatio = ptr, nbytes, ratio
    w.hitProp, w.missProp = (2, 1) if nbytes else (0, 0)  # persisting / streaming ; normal / normal
    rc = cudart.cudaStreamSetAttribute(ctypes.c_void_p(stream), 1, ctypes.byref(w))
    assert rc == 0, f'cudaStreamSetAttribute -> {rc}'


def time_hop(eh, rowptr, colidx, nnz, rec_in, rec_out, cards, stream):
    times = []
    with torch.cuda.stream(stream):
        for _ in range(4):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            eh._merge(rowptr, colidx, nnz, rec_in, rec_out, cards[:, 0], dev)
            e.record(stream)
            stream.synchronize()
            times.append(s.elapsed_time(e))
    return min(times[1:])


eh = ssb.ElphHashes(Namespace(max_hash_hops=1, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
ei = rmat_edges(scale, 16, 0, dev)
stream = torch.cuda.Stream(device=dev)
limit = ctypes.c_size_t(0)
cudart.cudaDeviceGetLimit(ctypes.byref(limit), 0x06)
prop_max = ctypes.c_int(0)
cudart.cudaDeviceGetAttribute(ctypes.byref(prop_max), 108, 0)  # cudaDevAttrMaxPersistingL2CacheSize
print(f'persisting L2 limit now {limit.value >> 20} MB, device max {prop_max.value >> 20} MB')
rc = cudart.cudaDeviceSetLimit(0x06, ctypes.c_size_t(prop_max.value))
print('cudaDeviceSetLimit ->', rc)
for relabel in (False, True):
    e2 = ei
    if relabel:
        deg = torch.bincount(ei[1], minlength=n)
        order = torch.argsort(deg, descending=True)
        perm = torch.empty(n, dtype=torch.int64, device=dev)
        perm[order] = torch.arange(n, device=dev)
        e2 = perm[ei]
        del deg, order, perm
    rowptr, colidx, nnz, _ = ssb.build_csr(e2, dev, num_rows=n, add_loops=True)
    rec0 = eh._init_records(n, dev)
    rec1, rec2 = torch.empty_like(rec0), torch.empty_like(rec0)
    cards = torch.zeros((n, 1), device=dev)
    with torch.cuda.stream(stream):
        eh._merge(rowptr, colidx, nnz, rec0, rec1, cards[:, 0], dev)
    stream.synchronize()
    set_window(stream.cuda_stream, 0, 0)
    base = time_hop(eh, rowptr, colidx, nnz, rec1, rec2, cards, stream)
    print(f'relabel={relabel}: no window {base:.2f} ms', flush=True)
    if relabel:
        d = rowptr[1:] - rowptr[:-1]
        for mb in windows:
            rows = int(mb * 1e6 / 768)
            share = float(d[:rows].sum()) / float(d.sum())
            set_window(stream.cuda_stream, rec1.data_ptr(), rows * 768)
            t = time_hop(eh, rowptr, colidx, nnz, rec1, rec2, cards, stream)
            print(f'  window {mb:.0f} MB ({rows} rows, {share:.3f} of neighbour reads): {t:.2f} ms', flush=True)
        set_window(stream.cuda_stream, 0, 0)
    del rowptr, colidx, rec0, rec1, rec2, cards, e2
