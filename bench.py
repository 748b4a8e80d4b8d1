#!/usr/bin/env python
"""bench.py -- link structural features/sec through the B200 sketch engine (BASELINE.json metric).

One STEP = one pass of the hot path over one synthetic batch:
    build_hash_tables(N, edge_index)   COO -> CSR, hop-0 sketches, K x (k-hop merge + HLL++ cardinalities)
    get_subgraph_features(links, ...)  K(K+2) features for each of L candidate links
value = L * steps / time (whole job, inputs resident in HBM); e2e = the same through the public API with HOST
(pinned) edge_index / links and host results, copies inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                           the reference's CPU algorithm (oracle port)

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

_OUT = sys.stdout
METRIC = 'link structural features/sec'
UNIT = 'links/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='rmat', choices=['rmat', 'collab', 'ppa', 'citation2'])
    ap.add_argument('--scale', type=int, default=24, help='R-MAT scale (N = 2^scale)')
    ap.add_argument('--edge-factor', type=int, default=16)
    ap.add_argument('--hops', type=int, default=None)
    ap.add_argument('--links', type=int, default=None, help='candidate links per step')
    ap.add_argument('--merge-variant', default='auto', choices=['auto', 'tma', 'ldg', 'generic'])
    ap.add_argument('--cpu-scale', type=int, default=18, help='R-MAT scale of the bounded CPU-baseline sample')
    ap.add_argument('--ref-scale', type=int, default=17, help='R-MAT scale of each --impl reference step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--exchange', default='auto', choices=['auto', 'p2p', 'mc', 'nccl'], help='multi-GPU exchange mode')
    ap.add_argument('--seed', type=int, default=0)
    return ap.parse_args()


def workload_spec(a):
    from subgraph_sketching_b200.graphs import SHAPES
    if a.workload == 'rmat':
        n = 1 << a.scale
        hops = a.hops or 3
        links = a.links if a.links is not None else max(int(20_000_000 * n / (1 << 24)), 1000)
        name = f'rmat{a.scale}_ef{a.edge_factor}_k{hops}'
        return dict(kind='rmat', name=name, num_nodes=n, hops=hops, links=links)
    s = SHAPES[a.workload]
    hops = a.hops or s['hops']
    links = a.links if a.links is not None else s['links']
    return dict(kind='powerlaw', name=f'ogbl-{a.workload}-shaped_k{hops}', num_nodes=s['num_nodes'], edges=s['edges'],
                hops=hops, links=links)


def make_inputs(spec, a, device):
    from subgraph_sketching_b200.graphs import powerlaw_edges, rmat_edges, sample_links
    if spec['kind'] == 'rmat':
        scale = spec['num_nodes'].bit_length() - 1
        ei = rmat_edges(scale, a.edge_factor, a.seed, device)
    else:
        ei = powerlaw_edges(spec['num_nodes'], spec['edges'], a.seed, device)
    links = sample_links(spec['num_nodes'], ei, spec['links'] // 2, spec['links'] - spec['links'] // 2, a.seed, device)
    return ei.contiguous(), links


def engine_args(hops):
    from argparse import Namespace
    return Namespace(max_hash_hops=hops, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.proc = None
        self.path = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            self.sel = uuid if uuid.startswith('GPU-') else 'GPU-' + uuid
        except Exception:
            self.sel = str(device_index)

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', self.sel, f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(',')]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith('active'):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_pass(scale, edge_factor, hops, seed, links_per_node):
    """one pass of the reference's algorithm on the host cores (oracle port: the same torch-CPU ops the
    reference issues -- scatter-amax propagate, [n, m] float pow/sum, [n, T] argsort -- bit-equal to it)"""
    from oracle import sketch_oracle as so
    from subgraph_sketching_b200.graphs import rmat_edges, sample_links
    n = 1 << scale
    ei = rmat_edges(scale, edge_factor, seed, 'cpu')
    L = max(int(links_per_node * n), 1000)
    links = sample_links(n, ei, L // 2, L - L // 2, seed, 'cpu')
    o = so.OracleSketches(hops, 128, 8, use_zero_one=False, floor_sf=False)
    t0 = time.perf_counter()
    tables, cards = o.build_hash_tables(n, ei)
    t1 = time.perf_counter()
    feats = o.subgraph_features(links, tables, cards)
    t2 = time.perf_counter()
    return dict(links=L, seconds=t2 - t0, build_s=t1 - t0, features_s=t2 - t1, nnz=int(ei.shape[1]) + n,
                checksum=float(feats.sum()))


def cpu_sample_desc(scale, edge_factor, hops, r):
    return (f'R-MAT scale {scale} (N={1 << scale}, nnz={r["nnz"]}), edge factor {edge_factor}, K={hops}, '
            f'{r["links"]} links; build_hash_tables {r["build_s"]:.2f} s + get_subgraph_features {r["features_s"]:.2f} s')


def run_reference(a, spec):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lpn = spec['links'] / spec['num_nodes']
    for _ in range(a.warmup):
        cpu_pass(a.ref_scale, a.edge_factor, spec['hops'], a.seed, lpn)
    links, dt, last = 0, 0.0, None
    for _ in range(a.steps):
        last = cpu_pass(a.ref_scale, a.edge_factor, spec['hops'], a.seed, lpn)  # graph generation is not timed
        links += last['links']
        dt += last['seconds']
    value = links / dt
    sample = cpu_sample_desc(a.ref_scale, a.edge_factor, spec['hops'], last)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3 * dt / max(a.steps, 1), 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'u32/u8 sketches, f32 estimates', 'data': 'synthetic',
        'config': {'workload': spec['name'], 'sample': sample, 'hops': spec['hops'], 'num_perm': 128, 'hll_p': 8},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    # the contract is ONE JSON line on stdout: keep a private handle to the real stdout and send everything
    # else that writes to fd 1 (NCCL's version banner, library chatter) to stderr
    global _OUT
    _OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    a = parse()
    spec = workload_spec(a)
    if a.impl == 'reference':
        run_reference(a, spec)
        return

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if a.gpus > 1 and world != a.gpus:
        raise SystemExit(f'--gpus {a.gpus} must be launched as: python -m torch.distributed.run --nnodes=1 '
                         f'--nproc-per-node {a.gpus} --master-addr 127.0.0.1 --master-port P bench.py --gpus {a.gpus} ...')
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback)'
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    import torch.distributed as dist
    distributed = world > 1
    if distributed:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)

    import subgraph_sketching_b200 as ssb
    from subgraph_sketching_b200 import _lib
    from subgraph_sketching_b200.dist import ShardedElphHashes, link_slice

    N, K, L = spec['num_nodes'], spec['hops'], spec['links']
    F = K * (K + 2)
    ei, links = make_inputs(spec, a, device)
    torch.cuda.synchronize()
    n_edges = int(ei.shape[1])
    nnz = n_edges + min(int(ei.max()) + 1, N)
    if distributed:
        eng = ShardedElphHashes(engine_args(K), merge_variant=a.merge_variant, exchange=a.exchange)
        eh = eng.eh
    else:
        eng = eh = ssb.ElphHashes(engine_args(K), merge_variant=a.merge_variant)
    eh.validate_links = True

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(edge_index, link_list):
        tables, cards = eng.build_hash_tables(N, edge_index)
        feats = eng.get_subgraph_features(link_list, tables, cards)
        return feats

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time by CUDA events, max over ranks"""
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        acc = 0.0
        for _ in range(steps):
            out = fn()
            acc += float(out.shape[0])
            del out
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=device)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm -------------------------------------------------------------------
    for _ in range(max(a.warmup, 0)):
        one_step(ei, links)
    eh.event_log = []
    _lib.lib.reset_counters()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(lambda: one_step(ei, links), a.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.lib.launches
    log, eh.event_log = eh.event_log, None
    stage_ms = {}
    for name, s, e in log:
        stage_ms.setdefault(name, []).append(s.elapsed_time(e))
    ms_per_step = total_ms / a.steps
    value = L / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (k-hop merge) -----------------------------------------------
    R = 768
    if distributed:
        rows_local = eng.bounds[rank + 1] - eng.bounds[rank]
        nnz_local = eng.local_nnz
    else:
        rows_local, nnz_local = N, nnz
    merge_bytes = nnz_local * R + rows_local * R + 4 * nnz_local + 8 * (rows_local + 1) + 4 * rows_local
    merge_ms = statistics.mean(stage_ms['khop_merge']) if stage_ms.get('khop_merge') else None
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'merge_traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(spec['name'])
        except Exception:
            traffic = None
    achieved = merge_bytes / (merge_ms * 1e-3) / 1e9 if merge_ms else None
    roofline = {'bound': 'hbm', 'kernel': 'ss_khop_merge (merge_tma/ldg_kernel + merge_fixup_kernel)',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak if achieved else None,
                'traffic': traffic, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': merge_bytes,
                'avg_launch_ms': merge_ms, 'launches_timed': len(stage_ms.get('khop_merge', []))}
    if traffic and merge_ms and not distributed:
        # the same launch time against the DRAM bytes ncu measured for this kernel and workload: `frac` above can
        # exceed 1 because the algorithmic model counts every neighbour record once per edge while hub records hit in L2
        roofline['dram_achieved'] = traffic / (merge_ms * 1e-3) / 1e9
        roofline['dram_frac'] = roofline['dram_achieved'] / peak
    L_local = (link_slice(L, world, rank)[1] - link_slice(L, world, rank)[0]) if distributed else L
    link_bytes = L_local * (2 * K * R + 16 + 8 * K + 4 * F)
    lf_ms = sum(stage_ms.get('link_features', [])) / a.steps if stage_ms.get('link_features') else None
    link_roofline = {'bound': 'hbm', 'kernel': 'link_features_kernel', 'algorithmic_bytes_per_step': link_bytes,
                     'ms_per_step': lf_ms, 'achieved': link_bytes / (lf_ms * 1e-3) / 1e9 if lf_ms else None,
                     'links_per_s_kernel_only': L_local / (lf_ms * 1e-3) if lf_ms else None}
    if link_roofline['achieved']:
        link_roofline['frac'] = link_roofline['achieved'] / peak

    # ---- end-to-end arm: host buffers through the public API -----------------------------------------
    e2e = None
    if not a.no_e2e:
        ei_h = torch.empty(ei.shape, dtype=ei.dtype, pin_memory=True).copy_(ei)
        links_h = torch.empty(links.shape, dtype=links.dtype, pin_memory=True).copy_(links)
        torch.cuda.synchronize()

        def e2e_step():
            if distributed:
                lo_, hi_ = link_slice(L, world, rank)
                tables, cards = eng.build_hash_tables(N, ei_h)  # pinned host edges, read in place
                f = eng.eh.get_subgraph_features(links_h[lo_:hi_], tables, cards)
            else:
                tables, cards = eh.build_hash_tables(N, ei_h)          # cards come back to the host
                f = eh.get_subgraph_features(links_h, tables, cards)   # features come back to the host
            assert not f.is_cuda
            return f

        e2e_step()
        e2e_steps = max(1, min(a.steps, 3))
        eh.event_log = []
        e2e_ms = timed(e2e_step, e2e_steps) / e2e_steps
        e2e_log, eh.event_log = eh.event_log, None
        e2e_stage = {}
        for name, s_, e_ in e2e_log:
            e2e_stage[name] = e2e_stage.get(name, 0.0) + s_.elapsed_time(e_) / e2e_steps
        if distributed:
            h2d = ei_h.numel() * 8 + L_local * 16
            d2h = L_local * F * 4
        else:
            h2d = ei_h.numel() * 8 + links_h.numel() * 8  # read in place by the kernels; cards keep a device twin
            d2h = N * K * 4 + L * F * 4
        e2e = {'value': L / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'ms_per_step': e2e_ms, 'steps': e2e_steps, 'stage_ms_per_step': e2e_stage}
        del ei_h, links_h

    # ---- bounded CPU baseline (rank 0, N = 1 only) ----------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        r = cpu_pass(a.cpu_scale, a.edge_factor, K, a.seed, L / N)
        cpu_baseline = {'value': r['links'] / r['seconds'], 'unit': UNIT, 'cores': torch.get_num_threads(),
                        'kind': 'port', 'sample': cpu_sample_desc(a.cpu_scale, a.edge_factor, K, r)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'u32/u8 sketches, f32 estimates', 'data': 'synthetic',
            'config': {'workload': spec['name'], 'num_nodes': N, 'directed_edges': n_edges, 'nnz_with_self_loops': nnz,
                       'hops': K, 'num_perm': 128, 'hll_p': 8, 'links_per_step': L, 'features_per_link': F,
                       'merge_variant': a.merge_variant, 'partition': (f'node-sharded x{world} (row blocks balanced by neighbour count), exchange={eng.exchange}'
                                     if distributed else 'single'),
                       'l2': 'inputs larger than L2 (each hop table is N*768 B), no explicit flush'},
            'features_per_s': value * F,
            'stage_ms_per_step': {k: sum(v) / a.steps for k, v in stage_ms.items()},
            'roofline': roofline, 'link_features_roofline': link_roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e,
            'gpu_launches': launches, 'clocks': clocks,
        }
        print(json.dumps(line), file=_OUT, flush=True)
    if distributed:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
