"""Experiment (GPU only): where is the ceiling of the k-hop merge when DRAM is taken out of the picture?
Runs the production merge kernel over a synthetic CSR whose neighbour ids all fall into the first H rows of the
previous-hop table, for H from L2-resident (16 MB) to DRAM-resident (12.9 GB).  The gather rate at small H is
the L2 -> SM ceiling of this access pattern (768-byte rows through TMA gather4); if it is close to the rate at
full size, the kernel is bound by the L2 fabric / its own issue rate and a better L2 hit rate cannot help.
    python tools/exp_l2_ceiling.py [scale] [avg_degree]      -> stdout + gpurun_out/exp_l2_ceiling.json"""
import json
import os
import sys
from argparse import Namespace

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device('cuda', 0)
n = 1 << scale
nnz = n * deg
eh = ssb.ElphHashes(Namespace(max_hash_hops=1, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
rec_in = eh._init_records(n, dev)
rec_out = torch.empty_like(rec_in)
rowptr = torch.arange(0, nnz + 1, deg, dtype=torch.int64, device=dev)
cards = torch.zeros((n, 1), device=dev)
g = torch.Generator(device=dev).manual_seed(0)
rows = []
for mb in (16, 32, 48, 64, 96, 128, 256, 1024, n * 768 // (1 << 20)):
    h = max(min(mb * (1 << 20) // 768, n), 1)
    colidx = torch.randint(0, h, (nnz,), generator=g, device=dev, dtype=torch.int32)
    times = []
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eh._merge(rowptr, colidx, nnz, rec_in, rec_out, cards[:, 0], dev)
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    ms = min(times[1:])
    row = dict(hot_mb=h * 768 / (1 << 20), ms=ms, g_rows_per_s=nnz / ms / 1e6, l2_to_sm_tbps=nnz * 768 / ms / 1e9)
    rows.append(row)
    print(f'neighbour ids in the first {row["hot_mb"]:9.1f} MB: {ms:7.2f} ms  {row["g_rows_per_s"]:6.2f} G rows/s  '
          f'{row["l2_to_sm_tbps"]:5.2f} TB/s into the SMs', flush=True)
    del colidx
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/exp_l2_ceiling.json', 'w'), indent=1)
