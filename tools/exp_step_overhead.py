"""Host-side profile of whole steps (build_hash_tables + batched get_subgraph_features) on the OGB-shaped graphs:
wall clock per step against the sum of the kernel stages, and a cProfile of the host code.
usage: python tools/exp_step_overhead.py [collab|ppa|citation2]"""
import cProfile
import os
import pstats
import sys
import time
from argparse import Namespace

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402
from subgraph_sketching_b200.graphs import SHAPES, powerlaw_edges, sample_links  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'collab'
dev = torch.device('cuda', 0)
s = SHAPES[name]
n, L = s['num_nodes'], s['links']
B = {'collab': L, 'ppa': 65_536, 'citation2': 261_424}[name]
ei = powerlaw_edges(n, s['edges'], 0, dev).contiguous()
links = sample_links(n, ei, L // 2, L - L // 2, 0, dev)
eh = ssb.ElphHashes(Namespace(max_hash_hops=2, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))


def step():
    tables, cards = eh.build_hash_tables(n, ei)
    outs = [eh.get_subgraph_features(links[i:i + B], tables, cards) for i in range(0, L, B)]
    return outs


step()
torch.cuda.synchronize()
for label, log in (('events off', None), ('events on', [])):
    eh.event_log = log
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    msg = f'{name} {label}: {dt * 1e3:.2f} ms per step wall clock'
    if log is not None:
        st = {}
        for nm, a, b in log:
            st[nm] = st.get(nm, 0.0) + a.elapsed_time(b) / 5
        msg += '  kernel stages: ' + ', '.join(f'{k}={v:.2f}' for k, v in st.items()) + f'  sum={sum(st.values()):.2f}'
    print(msg, flush=True)
eh.event_log = None
# build only / features only
t0 = time.perf_counter()
for _ in range(5):
    tables, cards = eh.build_hash_tables(n, ei)
torch.cuda.synchronize()
print(f'build_hash_tables alone: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms', flush=True)
t0 = time.perf_counter()
for _ in range(5):
    outs = [eh.get_subgraph_features(links[i:i + B], tables, cards) for i in range(0, L, B)]
torch.cuda.synchronize()
print(f'feature calls alone: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms', flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
