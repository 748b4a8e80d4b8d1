"""Experiment (one GPU): what would COLUMN sharding of the sketches cost per GPU?  (SURVEY 8e "alternative worth
benchmarking": every GPU owns P/C permutations + m/C registers of ALL nodes, no per-hop exchange inside a column group)

The per-GPU merge work of every sharding scheme is a gather of (rows, row bytes):
    node x8            : nnz/8 rows of 768 B  (+ exchange of finished rows)
    hybrid 2 col x 4   : nnz/4 rows of 384 B  (+ exchange inside each group of 4)
    hybrid 4 col x 2   : nnz/2 rows of 192 B
    column x8          : nnz   rows of  96 B  (no exchange at all)
i.e. the same bytes, but 1x / 2x / 4x / 8x the rows.  This script measures the production TMA merge kernel on ONE GPU
over the whole R-MAT graph for the row widths it supports (768 / 512 / 384 / 256 B: ss_khop_merge_ex layouts) and
over 1/2, 1/4, 1/8 of the destination rows, so that the rows-per-second vs row-width trend (is the gather bound by
bytes or by rows?) is measured rather than assumed; 192 B and 96 B are extrapolated from it and labelled so.
usage: python tools/exp_column_sharding.py [scale]      -> stdout + gpurun_out/exp_column_sharding.json"""
import ctypes
import json
import os
import sys
from argparse import Namespace

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402
from subgraph_sketching_b200 import _lib  # noqa: E402
from subgraph_sketching_b200._lib import MergeDesc, check, lib  # noqa: E402
from subgraph_sketching_b200.graphs import rmat_edges  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device('cuda', 0)
n = 1 << scale
eh = ssb.ElphHashes(Namespace(max_hash_hops=1, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
ei = rmat_edges(scale, 16, 0, dev)
rowptr, colidx, nnz, _ = ssb.build_csr(ei, dev, num_rows=n, add_loops=True)
del ei
ws = torch.empty(check(lib.ss_merge_workspace_bytes(nnz, 128, 8)), dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
rows_out = []


def merge(layout, rb, rec_in, rec_out, rp, ci, n_rows, cnt):
    d = MergeDesc()
    d.rowptr, d.colidx, d.n_rows, d.nnz = rp.data_ptr(), ci.data_ptr(), n_rows, cnt
    d.rec_in, d.in_rows, d.in_stride = rec_in.data_ptr(), n, rb
    d.rec_out, d.out_stride = rec_out.data_ptr(), rb
    d.num_perm, d.hll_p, d.layout, d.variant = 128, 8, layout, _lib.SS_MERGE_TMA
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    check(lib.ss_khop_merge_ex(ctypes.byref(d), st), 'ss_khop_merge_ex')


for layout, rb, name in ((_lib.SS_LAYOUT_FULL, 768, 'full record'), (_lib.SS_LAYOUT_MINHASH, 512, 'MinHash half'),
                         (_lib.SS_LAYOUT_HALF, 384, 'column half (64 perms + 128 registers)'),
                         (_lib.SS_LAYOUT_HLL, 256, 'HLL half')):
    rec_in = torch.randint(0, 60, (n, rb), dtype=torch.uint8, device=dev)
    rec_out = torch.empty((n, rb), dtype=torch.uint8, device=dev)
    for frac in (1, 2, 4, 8):
        rows = n // frac  # the hub-heavy low ids: the worst block of a node sharding by equal rows
        # cost-balanced would be nnz / frac neighbours: take the first rows that carry that many
        target = nnz // frac
        rows = int(torch.searchsorted(rowptr, torch.tensor([target], device=dev)).item())
        rows = max(min(rows, n), 1)
        cnt = int(rowptr[rows].item())
        rp = rowptr[:rows + 1].contiguous()
        times = []
        for _ in range(4):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            merge(layout, rb, rec_in, rec_out, rp, colidx, rows, cnt)
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
        ms = min(times[1:])
        row = dict(row_bytes=rb, layout=name, share_of_neighbours=1.0 / frac, dest_rows=rows, neighbours=cnt, ms=ms,
                   g_rows_per_s=cnt / ms / 1e6, tb_per_s=cnt * rb / ms / 1e9)
        rows_out.append(row)
        print(f'{rb:4d} B rows, 1/{frac} of the neighbours ({rows:9d} dest rows): {ms:7.2f} ms  {row["g_rows_per_s"]:6.2f} G rows/s  '
              f'{row["tb_per_s"]:5.2f} TB/s', flush=True)
    del rec_in, rec_out
    torch.cuda.empty_cache()
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows_out, open('gpurun_out/exp_column_sharding.json', 'w'), indent=1)
