"""Where does the time of a SMALL get_subgraph_features call go?  (ELPH / BUDDY-style 65,536-link batches on a
ppa-shaped graph: the kernel takes ~75 us, the call much longer.)  cProfile over 200 calls + wall-clock per call with and
without link validation.  usage: python tools/exp_call_overhead.py"""
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

dev = torch.device('cuda', 0)
s = SHAPES['ppa']
n = s['num_nodes']
ei = powerlaw_edges(n, s['edges'], 0, dev).contiguous()
links = sample_links(n, ei, 1_500_000, 1_500_000, 0, dev)
eh = ssb.ElphHashes(Namespace(max_hash_hops=2, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False))
tables, cards = eh.build_hash_tables(n, ei)
B = 65_536
batches = [links[i:i + B] for i in range(0, links.shape[0], B)]


def run(reps):
    for r in range(reps):
        for b in batches:
            eh.get_subgraph_features(b, tables, cards)
    torch.cuda.synchronize()


for validate in (True, False):
    eh.validate_links = validate
    run(1)
    t0 = time.perf_counter()
    run(5)
    dt = time.perf_counter() - t0
    print(f'validate_links={validate}: {dt / (5 * len(batches)) * 1e6:.1f} us per call ({len(batches)} calls of {B} links)', flush=True)
eh.validate_links = True
pr = cProfile.Profile()
pr.enable()
run(4)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
