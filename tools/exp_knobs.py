"""Whole-step A/B of the build_hash_tables knobs on one R-MAT graph (GPU only; tuning aid).

    python tools/exp_knobs.py [scale] [links]        -> gpurun_out/exp_knobs.json (+ stdout table)

Configurations: CSR fill with zero-based cursors + rowptr read per edge (legacy) vs absolute cursors; 1024-byte
record stride; the experimental destination-block binning in front of the fill (SS_B200_CSR_BIN).  Every configuration is checked
against the first one (hop tables bit-equal, features bit-equal) before it is timed."""
import json
import os
import sys
from argparse import Namespace

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402
from subgraph_sketching_b200.graphs import rmat_edges, sample_links  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << scale
L = int(sys.argv[2]) if len(sys.argv) > 2 else max(int(20_000_000 * n / (1 << 24)), 1000)
K = 3
dev = torch.device('cuda', 0)
ei = rmat_edges(scale, 16, 0, dev).contiguous()
links = sample_links(n, ei, L // 2, L - L // 2, 0, dev)
torch.cuda.synchronize()

CONFIGS = [
    ('legacy_fill', dict(fill='legacy', overlap=False, stride=None)),
    ('abs_fill', dict(fill='abs', overlap=False, stride=None)),
    ('abs_fill+stride1024', dict(fill='abs', overlap=False, stride=1024)),
    ('abs_fill+stride1024+binned_fill', dict(fill='abs', overlap=False, stride=1024, bin=1)),
]


def checksum(tables, feats):
    out = []
    for k in range(K + 1):
        r = tables.records(k)
        out.append(int(r.view(torch.int32).sum(dtype=torch.int64)))
    out.append(int(feats.view(torch.int32).sum(dtype=torch.int64)))
    return out


results = []
ref = None
for name, cfg in CONFIGS:
    os.environ['SS_B200_CSR_FILL'] = cfg['fill']
    os.environ['SS_B200_CSR_BIN'] = str(cfg.get('bin', 0))
    eh = ssb.ElphHashes(Namespace(max_hash_hops=K, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
    eh.overlap_init = cfg['overlap']
    eh.record_stride = cfg['stride']

    def step():
        tables, cards = eh.build_hash_tables(n, ei)
        return tables, cards, eh.get_subgraph_features(links, tables, cards)

    tables, cards, feats = step()
    chk = checksum(tables, feats)
    if ref is None:
        ref = chk
    ok = chk == ref
    del tables, cards, feats
    step()
    eh.event_log = []
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 4
    s.record()
    for _ in range(steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    stages = {}
    for nm, a, b in eh.event_log:
        stages[nm] = stages.get(nm, 0.0) + a.elapsed_time(b) / steps
    eh.event_log = None
    row = dict(config=name, ms_per_step=ms, stages=stages, bit_equal_to_first=ok, scale=scale, links=L)
    results.append(row)
    print(f'{name:36s} {ms:8.2f} ms/step  ' + '  '.join(f'{k}={v:.2f}' for k, v in stages.items()) +
          ('  OK' if ok else '  MISMATCH'), flush=True)
    del eh
    torch.cuda.empty_cache()

os.makedirs('gpurun_out', exist_ok=True)
json.dump(results, open('gpurun_out/exp_knobs.json', 'w'), indent=1)

# ---- merge only: L2 promotion of the gather4 tensor map x record stride (hop 2 timed, best of 3) --------------
eh = ssb.ElphHashes(Namespace(max_hash_hops=1, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
rowptr, colidx, nnz, _ = ssb.build_csr(ei, dev, num_rows=n, add_loops=True)
del ei, links
torch.cuda.empty_cache()
cards = torch.zeros((n, 1), device=dev)
merge_rows = []
ref = None
for stride in (768, 1024):
    rec = [torch.empty((n, stride), dtype=torch.uint8, device=dev)[:, :768] for _ in range(3)]
    eh._init_records(n, dev, out=rec[0])
    for promo in (0, 128, 256):
        os.environ['SS_B200_TMA_L2PROMO'] = str(promo)
        eh._merge(rowptr, colidx, nnz, rec[0], rec[1], cards[:, 0], dev)
        times = []
        for _ in range(4):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            eh._merge(rowptr, colidx, nnz, rec[1], rec[2], cards[:, 0], dev)
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
        chk = (int(rec[2].view(torch.int32).sum(dtype=torch.int64)), float(cards.sum()))
        if ref is None:
            ref = chk
        row = dict(stride=stride, l2_promotion=promo, ms=min(times[1:]), all_ms=times, ok=chk == ref)
        merge_rows.append(row)
        print(f'merge hop 2: stride {stride} L2 promotion {promo:3d}: {row["ms"]:.2f} ms  {"OK" if row["ok"] else "MISMATCH"}',
              flush=True)
    del rec
    torch.cuda.empty_cache()
os.environ.pop('SS_B200_TMA_L2PROMO', None)
json.dump(dict(steps=results, merge=merge_rows), open('gpurun_out/exp_knobs.json', 'w'), indent=1)
