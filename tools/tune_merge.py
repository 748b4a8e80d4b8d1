"""Times the k-hop merge engines / TMA configurations on one R-MAT graph (GPU only; tuning aid).
usage: python tools/tune_merge.py [scale] [edge_factor]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402
from subgraph_sketching_b200.graphs import rmat_edges  # noqa: E402
from argparse import Namespace  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device('cuda', 0)
n = 1 << scale
ei = rmat_edges(scale, ef, 0, dev)
eh = ssb.ElphHashes(Namespace(max_hash_hops=1, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
rowptr, colidx, nnz, _ = ssb.build_csr(ei, dev, num_rows=n, add_loops=True)
del ei
stride = int(os.environ.get('TUNE_STRIDE', '768'))  # experiment: padded record stride (DRAM page alignment)
rec0 = eh._init_records(n, dev, out=torch.empty((n, stride), dtype=torch.uint8, device=dev)[:, :768])
rec1 = torch.empty((n, stride), dtype=torch.uint8, device=dev)[:, :768]
rec2 = torch.empty((n, stride), dtype=torch.uint8, device=dev)[:, :768]
cards = torch.zeros((n, 1), device=dev)
bytes_alg = nnz * 768 + n * 768 + 4 * nnz + 8 * (n + 1) + 4 * n
print(f'scale {scale}: N={n} nnz={nnz} algorithmic bytes/hop = {bytes_alg / 1e9:.1f} GB, record stride {stride}')
ref = None
runs = [('tma', 0), ('tma', 1)]
for variant, cfg in runs:
    if cfg is not None:
        os.environ['SS_B200_TMA_CFG'] = str(cfg)
    eh.merge_variant = variant
    # hop 1 then hop 2 (denser sketches); time hop 2
    eh._merge(rowptr, colidx, nnz, rec0, rec1, cards[:, 0], dev)
    times = []
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eh._merge(rowptr, colidx, nnz, rec1, rec2, cards[:, 0], dev)
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    chk = int(rec2.sum(dtype=torch.int64)), float(cards.sum())
    if ref is None:
        ref = chk
    ms = min(times[1:])
    print(f'{variant} cfg={cfg}: {ms:.2f} ms  {bytes_alg / ms / 1e6:.0f} GB/s algorithmic  '
          f'{"OK" if chk == ref else "MISMATCH " + str(chk) + " vs " + str(ref)}', flush=True)
