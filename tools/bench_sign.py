"""Times the SIGN pre-propagation (gcn_norm + SpMM + concat) on an OGB-citation2-shaped
synthetic graph with 128 float32 features and sign_k = 3, next to the reference's formulation (oracle port of
gcn_norm + torch_sparse.spmm, torch-CPU, all host cores) on the same inputs (GPU only; measurement aid).
    python tools/bench_sign.py  -> stdout + gpurun_out/bench_sign.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import sign_oracle  # noqa: E402
from subgraph_sketching_b200 import sign as bs  # noqa: E402
from subgraph_sketching_b200.graphs import SHAPES, powerlaw_edges  # noqa: E402

shape = SHAPES['citation2']
n, F, K = shape['num_nodes'], 128, 3
dev = torch.device('cuda', 0)
ei = powerlaw_edges(n, shape['edges'], 0, dev).contiguous()   # sorted by row, like to_undirected output
E = ei.shape[1]
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, F, generator=g, device=dev)
w = torch.ones(E, device=dev)
torch.cuda.synchronize()
times = []
for _ in range(4):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = bs.sign_features(x, ei, w, K)
    e.record()
    torch.cuda.synchronize()
    times.append(s.elapsed_time(e))
ms = min(times[1:])
alg = E * (4 * F + 12) + n * (4 * F * (1 + K) + 16) + n * 4 * F  # gathers + metadata, x row, K output blocks + x copy
res = {'num_nodes': n, 'edges': E, 'features': F, 'sign_k': K, 'ms': ms, 'all_ms': times,
       'algorithmic_GB': alg / 1e9, 'GBps': alg / ms / 1e6}
print(f'graph: N={n} E={E} F={F} sign_k={K}')
print(f'sign_features (gcn_norm + CSR + SpMM + concat): {ms:.2f} ms; algorithmic {alg / 1e9:.1f} GB -> {alg / ms / 1e6:.0f} GB/s',
      flush=True)
# unsorted edge order exercises the atomics-built edge-position CSR
perm = torch.randperm(E, device=dev, generator=g)
ei_u = ei[:, perm].contiguous()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
bs.sign_features(x, ei_u, w, K)
s.record()
out_u = bs.sign_features(x, ei_u, w, K)
e.record()
torch.cuda.synchronize()
res['ms_unsorted'] = s.elapsed_time(e)
res['max_abs_diff_sorted_vs_unsorted'] = float((out_u - out).abs().max())
print(f'same graph, shuffled edge order: {res["ms_unsorted"]:.2f} ms; max |diff| vs sorted {res["max_abs_diff_sorted_vs_unsorted"]:.2e}',
      flush=True)
# CPU: the reference's formulation on a bounded sample (a 1/10-scale graph of the same shape), checked bit for bit
torch.set_num_threads(os.cpu_count() or 1)
ns = n // 10
eis = powerlaw_edges(ns, shape['edges'] // 10, 1, dev).contiguous()
xs = torch.randn(ns, F, generator=g, device=dev)
wsm = torch.ones(eis.shape[1], device=dev)
outs = bs.sign_features(xs, eis, wsm, K)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
outs = bs.sign_features(xs, eis, wsm, K)
e.record()
torch.cuda.synchronize()
xc, eic, wc = xs.cpu(), eis.cpu(), wsm.cpu()
t0 = time.perf_counter()
ref = sign_oracle.sign_features(xc, eic, wc, K)
t_cpu = time.perf_counter() - t0
res['sample'] = {'num_nodes': ns, 'edges': int(eis.shape[1]), 'gpu_ms': s.elapsed_time(e), 'cpu_s': t_cpu,
                 'cpu_threads': torch.get_num_threads(), 'bit_identical_to_cpu': bool(torch.equal(outs.cpu(), ref)),
                 'max_abs_diff_vs_cpu': float((outs.cpu() - ref).abs().max())}
print(f'1/10-scale sample (N={ns}, E={eis.shape[1]}): GPU {res["sample"]["gpu_ms"]:.2f} ms; reference formulation on '
      f'{torch.get_num_threads()} host threads {t_cpu:.2f} s; GPU result bit-identical: {res["sample"]["bit_identical_to_cpu"]} '
      f'(max |diff| {res["sample"]["max_abs_diff_vs_cpu"]:.2e})')
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/bench_sign.json', 'w'), indent=1)
