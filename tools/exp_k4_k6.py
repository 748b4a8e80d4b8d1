"""A/B on one R-MAT graph (GPU only; tuning aid):
  K6  streaming CSR of the key-ordered list (default) vs histogram + cursor fill (SS_B200_CSR_FAST=0)
  K4  batched kernel (default) vs per-link kernel (SS_B200_LINKS=ldg); random links and source-grouped links
      (1001 links per source, the reference's ranking evaluation shape); tile sizes
usage: python tools/exp_k4_k6.py [scale] [links]      -> stdout + gpurun_out/exp_k4_k6.json"""
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
dev = torch.device('cuda', 0)
out = {}


def timeit(fn, reps=4):
    fn()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts)


ei = rmat_edges(scale, 16, 0, dev).contiguous()
torch.cuda.synchronize()
for fast in ('1', '0'):
    os.environ['SS_B200_CSR_FAST'] = fast
    ms = timeit(lambda: ssb.build_csr(ei, dev, num_rows=n, add_loops=True))
    out[f'csr_fast={fast}'] = ms
    print(f'build_csr SS_B200_CSR_FAST={fast}: {ms:.2f} ms', flush=True)
os.environ['SS_B200_CSR_FAST'] = '1'
shuffled = ei[:, torch.randperm(ei.shape[1], device=dev)]
ms = timeit(lambda: ssb.build_csr(shuffled, dev, num_rows=n, add_loops=True))
out['csr_shuffled_fast_then_fallback'] = ms
print(f'build_csr on a shuffled list (speculative pass + fallback): {ms:.2f} ms', flush=True)
del shuffled

for K in (3, 2):
    eh = ssb.ElphHashes(Namespace(max_hash_hops=K, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
    tables, cards = eh.build_hash_tables(n, ei)
    links = sample_links(n, ei, L // 2, L - L // 2, 0, dev)
    grouped = links.clone()
    grouped[:, 0] = links[torch.arange(L, device=dev) // 1001, 0]
    ref = {}
    for name, lk in (('random', links), ('grouped1001', grouped)):
        for mode in ('ldg', 'batched'):
            if mode == 'ldg':
                os.environ['SS_B200_LINKS'] = 'ldg'
            else:
                os.environ['SS_B200_LINKS'] = 'batched'
            f = eh.get_subgraph_features(lk, tables, cards)
            if name not in ref:
                ref[name] = f
            ok = torch.equal(f, ref[name])
            ms = timeit(lambda: eh.get_subgraph_features(lk, tables, cards))
            out[f'K{K}_{name}_{mode}'] = dict(ms=ms, links_per_s=L / ms * 1e3, bit_equal=ok)
            print(f'K={K} {name:12s} {mode:8s}: {ms:7.2f} ms  {L / ms / 1e3:7.1f} M links/s  {"OK" if ok else "MISMATCH"}',
                  flush=True)
        if K == 3:
            for tile in (3, 12, 24, 48, 96):
                os.environ['SS_B200_LINK_TILE'] = str(tile)
                ms = timeit(lambda: eh.get_subgraph_features(lk, tables, cards))
                out[f'K{K}_{name}_tile{tile}'] = ms
                print(f'K={K} {name:12s} tile {tile:3d}: {ms:7.2f} ms', flush=True)
            os.environ.pop('SS_B200_LINK_TILE', None)
    del tables, cards, links, grouped, ref, eh
    torch.cuda.empty_cache()
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/exp_k4_k6.json', 'w'), indent=1)
