#!/usr/bin/env python
"""bench.py -- link structural features/sec through the B200 sketch engine (BASELINE.json metric).

One STEP = one pass of the hot path over one synthetic batch:
    build_hash_tables(N, edge_index)   COO -> CSR, hop-0 sketches, K x (k-hop merge + HLL++ cardinalities)
    get_subgraph_features(links, ...)  K(K+2) features for each of L candidate links
value = L * steps / time (whole job, inputs resident in HBM); e2e = the same through the public API with HOST
(pinned) edge_index / links and host results, copies inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                           the reference's CPU algorithm (oracle port)

Prints ONE JSON line on rank 0.  Besides the headline workload (R-MAT scale 24, K = 3: BASELINE.json configs[4]) the
line carries
    configs       one record per OGB-shaped BASELINE configuration: ogbl-collab-shaped (bulk), ogbl-ppa-shaped in
                  65,536-link batches plus the ELPH per-batch call pattern, ogbl-citation2-shaped in 261,424-link
                  batches plus a source-grouped ranking set (1 positive + 1000 negatives per source); under --gpus N
                  the citation2-shaped graph node-sharded over the N GPUs
    parity_check  (N > 1) every rank's block of every hop table, the cardinalities and its feature slice compared
                  bit for bit with a single-GPU build on rank 0, outside the timed region
    sampled_check (N = 1) hop-0 rows against the oracle's initialisation, sampled rows of every hop against a direct
                  min / max over the engine's own previous-hop rows, sampled links against the oracle's feature
                  arithmetic on the engine's rows
"""
import argparse
import importlib.util
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
R = 768  # bytes of one compact record


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
    ap.add_argument('--cpu-scale', type=int, default=19, help='R-MAT scale of the bounded CPU-baseline sample')
    ap.add_argument('--ref-scale', type=int, default=None,
                    help='R-MAT scale of each --impl reference step (default: the largest whose steps + warm-up fit '
                         '--ref-budget seconds)')
    ap.add_argument('--ref-budget', type=float, default=240.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the OGB-shaped configuration records')
    ap.add_argument('--no-checks', action='store_true', help='skip parity_check / sampled_check')
    ap.add_argument('--exchange', default='auto', choices=['auto', 'halo', 'p2p', 'mc', 'nccl'], help='multi-GPU exchange mode')
    ap.add_argument('--seed', type=int, default=0)
    return ap.parse_args()


def graphs_module():
    """the synthetic generators, loaded by path: importing the PACKAGE would map libss_b200.so into the process, which
    the reference arm must not do"""
    spec = importlib.util.spec_from_file_location('_ss_b200_graphs', os.path.join(ROOT, 'subgraph_sketching_b200', 'graphs.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_spec(a, g):
    if a.workload == 'rmat':
        n = 1 << a.scale
        hops = a.hops or 3
        links = a.links if a.links is not None else max(int(20_000_000 * n / (1 << 24)), 1000)
        name = f'rmat{a.scale}_ef{a.edge_factor}_k{hops}'
        return dict(kind='rmat', name=name, num_nodes=n, hops=hops, links=links)
    s = g.SHAPES[a.workload]
    hops = a.hops or s['hops']
    links = a.links if a.links is not None else s['links']
    return dict(kind='powerlaw', name=f'ogbl-{a.workload}-shaped_k{hops}', num_nodes=s['num_nodes'], edges=s['edges'],
                hops=hops, links=links)


def make_inputs(spec, a, device, g):
    if spec['kind'] == 'rmat':
        scale = spec['num_nodes'].bit_length() - 1
        ei = g.rmat_edges(scale, a.edge_factor, a.seed, device)
    else:
        ei = g.powerlaw_edges(spec['num_nodes'], spec['edges'], a.seed, device)
    links = g.sample_links(spec['num_nodes'], ei, spec['links'] // 2, spec['links'] - spec['links'] // 2, a.seed, device)
    return ei.contiguous(), links


def engine_args(hops):
    from argparse import Namespace
    return Namespace(max_hash_hops=hops, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=False)


def link_bytes(K, group=None):
    """algorithmic bytes per link of the pairwise kernel; `group` = links per source when the list is
    source-grouped: u's K records are read once per group"""
    F = K * (K + 2)
    rows = 2 * K * R if not group else K * R * (1.0 + 1.0 / group)
    return rows + 16 + 8 * K + 4 * F


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
# The ONLY part of this file that touches oracle/ (test infrastructure): the reference's algorithm timed on the host
# cores, and -- as the checker, never as the thing measured -- the spot checks of the GPU results below.
def cpu_pass(g, scale, edge_factor, hops, seed, links_per_node):
    """one pass of the reference's algorithm on the host cores (oracle port: the same torch-CPU ops the
    reference issues -- scatter-amax propagate, [n, m] float pow/sum, [n, T] argsort -- bit-equal to it)"""
    from oracle import sketch_oracle as so
    n = 1 << scale
    ei = g.rmat_edges(scale, edge_factor, seed, 'cpu')
    L = max(int(links_per_node * n), 1000)
    links = g.sample_links(n, ei, L // 2, L - L // 2, seed, 'cpu')
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


def run_reference(a, spec, g):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lpn = spec['links'] / spec['num_nodes']
    warmup = min(a.warmup, 1)  # a CPU pass needs no more than one warm-up (thread pool, page faults)
    scale = a.ref_scale
    note = None
    if scale is None:
        # largest sample whose steps + warm-up fit the time budget: one pass costs ~6 s at scale 17 on 16 cores and
        # doubles per scale (the reference materialises nnz x 1280 B of messages per hop: 567 GB at scale 24)
        t0 = time.perf_counter()
        probe = cpu_pass(g, 14, a.edge_factor, spec['hops'], a.seed, lpn)
        per14 = max(time.perf_counter() - t0, probe['seconds'])
        scale = 14
        while scale < 20 and per14 * (2.15 ** (scale + 1 - 14)) * (a.steps + warmup) <= a.ref_budget:
            scale += 1
        note = (f'scale chosen so that {a.steps} steps + {warmup} warm-up fit {a.ref_budget:.0f} s on {cores} cores '
                f'(probe: scale 14 = {per14:.2f} s per pass)')
    for _ in range(warmup):
        cpu_pass(g, scale, a.edge_factor, spec['hops'], a.seed, lpn)
    links, dt, last = 0, 0.0, None
    for _ in range(a.steps):
        last = cpu_pass(g, scale, a.edge_factor, spec['hops'], a.seed, lpn)  # graph generation is not timed
        links += last['links']
        dt += last['seconds']
    value = links / dt
    sample = cpu_sample_desc(scale, a.edge_factor, spec['hops'], last)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3 * dt / max(a.steps, 1), 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'u32/u8 sketches, f32 estimates', 'data': 'synthetic',
        'config': {'workload': spec['name'], 'sample': sample, 'sample_scale': scale, 'sample_choice': note,
                   'hops': spec['hops'], 'num_perm': 128, 'hll_p': 8, 'warmup_passes_run': warmup,
                   'implementation': 'oracle port (oracle/sketch_oracle.py: the torch-CPU ops of src/hashing.py, bit-equal to '
                                     'the unmodified reference in tests/test_oracle.py); /root/reference itself is absent '
                                     'on the GPU box and pure Python, so there is no oracle/_ref binary to run instead',
                   'native_libraries': 'none of this repository (the generators are loaded by file path)'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    try:  # self-report: shared objects of this repository mapped into the process (must be none)
        line['native_so_loaded'] = sorted({ln.split()[-1] for ln in open('/proc/self/maps') if ROOT in ln and '.so' in ln})
    except OSError:
        line['native_so_loaded'] = None
    print(json.dumps(line), file=_OUT, flush=True)


def sampled_check(eh, n, ei, tables, cards, links, feats, K, seed=0, n_rows=512, n_links=512):
    """spot check of a GPU result against the oracle (checker only): see the module docstring"""
    from oracle import sketch_oracle as so
    dev = ei.device
    g = torch.Generator().manual_seed(seed + 99)
    rows = torch.randint(0, n, (n_rows,), generator=g)
    out = {'rows': n_rows, 'links': n_links}
    # hop 0: the oracle's initialisation of the sampled ids
    mh0 = torch.stack([torch.from_numpy(so.minhash_init(1, 128, first_id=int(r) + 1)[0].astype('int64')) for r in rows[:64]])
    hl0 = torch.stack([torch.from_numpy(so.hll_init(1, 8, first_id=int(r) + 1)[0]) for r in rows[:64]])
    rec0 = tables.records(0)[rows[:64].to(dev)]
    out['hop0_bit_equal'] = bool(torch.equal(rec0[:, :512].contiguous().view(torch.int32).cpu().long() & 0xffffffff, mh0)
                                 and torch.equal(rec0[:, 512:768].cpu().view(torch.int8), hl0))
    # hops 1..K: each sampled row = min / max over its in-neighbours' previous-hop rows (+ itself when it has a self loop)
    src, dst = ei[0], ei[1]
    max_id = int(ei.max())
    ok = True
    rows_d = rows.to(dev)
    order = None
    for k in range(1, K + 1):
        prev, cur = tables.records(k - 1), tables.records(k)
        for r in rows_d[:128].tolist():
            nb = src[dst == r]
            if r <= max_id:
                nb = torch.cat([nb, torch.tensor([r], device=dev)])
            got = cur[r]
            if nb.numel() == 0:
                ok = ok and bool((got == 0).all())
                continue
            p = prev[nb]
            want_mh = p[:, :512].contiguous().view(torch.int32).long().__and__(0xffffffff).min(dim=0).values
            want_hl = p[:, 512:768].max(dim=0).values
            ok = ok and bool(torch.equal(got[:512].view(torch.int32).long() & 0xffffffff, want_mh)) \
                and bool(torch.equal(got[512:768], want_hl))
    del order
    out['hop_rows_bit_equal'] = ok
    # cardinalities and link features: the oracle's float arithmetic on the ENGINE's rows of the sampled nodes
    pick = torch.randint(0, links.shape[0], (n_links,), generator=g)
    lk = links[pick.to(links.device)].cpu()
    nodes, inv = torch.unique(lk.reshape(-1), return_inverse=True)
    mini = {}
    for k in range(K + 1):
        rec = tables.records(k)[nodes.to(dev)].cpu()
        mini[k] = {'minhash': rec[:, :512].contiguous().view(torch.int32).long() & 0xffffffff,
                   'hll': rec[:, 512:768].contiguous().view(torch.int8)}
    o = so.OracleSketches(K, 128, 8, use_zero_one=False, floor_sf=False,
                          constants=so.HllConstants(8, raw_estimate=eh.estimate_vector.numpy(), bias=eh.bias_vector.numpy(),
                                                    threshold=eh.hll_threshold))
    mini_cards = torch.stack([so.hll_count(o.c, mini[k]['hll']) for k in range(1, K + 1)], dim=1)
    got_cards = cards[nodes.to(cards.device)].cpu()
    scale = torch.clamp(mini_cards, min=1.0)
    out['cards_max_rel_err'] = float(((got_cards - mini_cards).abs() / scale).max())
    want = o.subgraph_features(inv.reshape(-1, 2), mini, mini_cards)
    got = feats[pick.to(feats.device)].cpu()
    fscale = torch.clamp(torch.maximum(mini_cards[inv.reshape(-1, 2)[:, 0]].max(1).values,
                                       mini_cards[inv.reshape(-1, 2)[:, 1]].max(1).values), min=1.0)
    out['features_max_rel_err'] = float(((got - want).abs() / fscale[:, None]).max())
    out['ok'] = bool(out['hop0_bit_equal'] and out['hop_rows_bit_equal'] and out['cards_max_rel_err'] <= 1e-6
                     and out['features_max_rel_err'] <= 1e-6)
    out['tolerance'] = '1e-6 relative to max(1, cards) (SURVEY 8c); integer rows bit-exact'
    return out


# ---------------------------------------------------------------------------------------------- GPU arm helpers
def table_checksums(rec, lo, hi, chunk=1 << 20):
    """two order-sensitive 64-bit checksums of rows [lo, hi) of a record table (wrapping int64 arithmetic)"""
    s1 = torch.zeros((), dtype=torch.int64, device=rec.device)
    s2 = torch.zeros((), dtype=torch.int64, device=rec.device)
    cols = torch.arange(1, rec.shape[1] // 4 + 1, device=rec.device, dtype=torch.int64) * 0x9E3779B1
    for a in range(lo, hi, chunk):
        b = min(a + chunk, hi)
        x = rec[a:b].view(torch.int32).to(torch.int64)
        s1 += x.sum()
        rows = (torch.arange(a, b, device=rec.device, dtype=torch.int64) * 2654435761 + 12345) | 1
        s2 += ((x * cols[None, :]).sum(dim=1) * rows).sum()
    return [s1, s2]


def float_checksums(t, lo, hi, chunk=1 << 22):
    x = t[lo:hi].contiguous().view(torch.int32).to(torch.int64)
    flat = x.reshape(-1)
    w = (torch.arange(flat.numel(), device=t.device, dtype=torch.int64) * 2654435761 + 777) | 1
    return [flat.sum(), (flat * w).sum()]


def measure_h2d_gbs(device, nbytes=1 << 30):
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    dst.copy_(src, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    return nbytes / (s.elapsed_time(e) * 1e-3) / 1e9


def stage_totals(log, steps):
    out = {}
    for name, s, e in log:
        out[name] = out.get(name, 0.0) + s.elapsed_time(e) / steps
    return out


def torch_scatter_forward(init_mh, init_hl, ei_loops, K):
    """the reference's own ops on CUDA tensors (what ELPH.forward costs per batch without this engine): PyG's
    MessagePassing(aggr='max') = index_select + scatter_reduce(amax) over [nnz, width] messages (hashing.py:28-45)"""
    def amax(x):
        out = torch.zeros_like(x)
        out.scatter_reduce_(0, ei_loops[1].view(-1, 1).expand(-1, x.size(1)), x.index_select(0, ei_loops[0]), reduce='amax',
                            include_self=False)
        return out
    mh, hl = init_mh, init_hl
    for _ in range(K):
        hl = amax(hl)
        mh = -amax(-mh)
    return mh, hl


def run_ogb_config(name, a, g, device, peak, do_check, rank, world, dist_engine=None):
    """one BASELINE configuration on an OGB-shaped synthetic graph: the build + per-batch feature calls, timed"""
    import subgraph_sketching_b200 as ssb
    from subgraph_sketching_b200.dist import link_slice
    shape = g.SHAPES[name]
    batch = {'collab': None, 'ppa': 65_536, 'citation2': 261_424}[name]
    n, K, L = shape['num_nodes'], shape['hops'], shape['links']
    F = K * (K + 2)
    ei = g.powerlaw_edges(n, shape['edges'], a.seed, device).contiguous()
    links = g.sample_links(n, ei, L // 2, L - L // 2, a.seed, device)
    n_edges = int(ei.shape[1])
    if dist_engine is not None:
        eng, eh = dist_engine, dist_engine.eh
        lo, hi = link_slice(L, world, rank)
    else:
        eng = eh = ssb.ElphHashes(engine_args(K))
        lo, hi = 0, L
    my_links = links  # (the sharded engine slices the full list itself)
    bsz = batch or L

    def features(tables, cards, lk):
        if dist_engine is not None:
            return eng.get_subgraph_features(lk, tables, cards)
        outs = [eh.get_subgraph_features(lk[s:s + bsz], tables, cards) for s in range(0, lk.shape[0], bsz)]
        return outs[0] if len(outs) == 1 else torch.cat(outs)

    def step():
        tables, cards = eng.build_hash_tables(n, ei)
        return tables, cards, features(tables, cards, my_links)

    import torch.distributed as dist
    def barrier():
        if dist_engine is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        step()
    steps = 3
    eh.event_log = []
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        out = step()
        del out  # nothing of the previous step is held while the next one allocates
    e.record()
    barrier()
    ms = torch.tensor([s.elapsed_time(e) / steps], device=device)
    if dist_engine is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    stages = stage_totals(eh.event_log, steps)
    eh.event_log = None
    tables, cards, feats = step()  # kept for the checks / the grouped set below
    lf_ms = stages.get('link_features')
    rec = {'workload': f'ogbl-{name}-shaped', 'num_nodes': n, 'directed_edges': n_edges, 'hops': K, 'links_per_step': L,
           'link_batch': bsz, 'feature_calls_per_step': (L + bsz - 1) // bsz if dist_engine is None else 1,
           'n_gpus': world, 'ms_per_step': ms, 'links_per_s': L / (ms * 1e-3), 'stage_ms_per_step': stages}
    if lf_ms:
        ach = (hi - lo) * link_bytes(K) / (lf_ms * 1e-3) / 1e9
        rec['link_features_roofline'] = {'bound': 'hbm', 'bytes_per_link': link_bytes(K), 'achieved': ach, 'peak': peak,
                                         'unit': 'GB/s', 'frac': ach / peak, 'links_per_s_kernel_only': (hi - lo) / (lf_ms * 1e-3)}
    if dist_engine is not None:
        rec['partition'] = f'node-sharded x{world}, exchange={eng.exchange}, csr={eng.csr_path}'
        hf = eng.halo_fraction
        if hf is not None:
            rec['halo_fraction_rank0'] = hf
    if name == 'citation2' and dist_engine is None:
        # the ranking evaluation shape (data.py:226-230): every source with its positive and 1000 negatives
        sources, per = 2000, 1001
        gen = torch.Generator(device=device).manual_seed(a.seed + 5)
        srcs = torch.randint(0, n, (sources,), generator=gen, device=device)
        grouped = torch.stack([srcs.repeat_interleave(per),
                               torch.randint(0, n, (sources * per,), generator=gen, device=device)], dim=1).contiguous()
        eb = 522_848
        for _ in range(2):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for s0 in range(0, grouped.shape[0], eb):
                eh.get_subgraph_features(grouped[s0:s0 + eb], tables, cards)
            t1.record()
            torch.cuda.synchronize()
        gms = t0.elapsed_time(t1)
        gl = grouped.shape[0]
        ach = gl * link_bytes(K, per) / (gms * 1e-3) / 1e9
        rec['grouped_eval'] = {'sources': sources, 'links_per_source': per, 'links': gl, 'eval_batch': eb, 'ms': gms,
                               'links_per_s': gl / (gms * 1e-3),
                               'roofline': {'bound': 'hbm', 'bytes_per_link': link_bytes(K, per),
                                            'formula': 'K*R*(1 + 1/g) + 16 + 8K + 4F, g = links per source (SURVEY 8d)',
                                            'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak}}
    if name == 'ppa' and dist_engine is None:
        rec['elph_forward'] = elph_forward_record(eh, n, K, ei, links, device)
    if do_check and dist_engine is None:
        rec['sampled_check'] = sampled_check(eh, n, ei, tables, cards, links, feats, K, seed=a.seed)
    del tables, cards, feats, ei, links
    torch.cuda.empty_cache()
    return rec


def elph_forward_record(eh, n, K, ei, links, device, batches=8, batch=65_536):
    """ELPH's per-batch call pattern (models/elph.py:186-216, train.py:198-204): a fresh add_self_loops tensor every
    forward, 2K operator calls + K hll_count, get_subgraph_features on the assembled dict -- through this engine's
    operator API, and once with the reference's own torch ops on CUDA tensors as the stated baseline"""
    loops = torch.arange(n, device=device)
    init_mh = eh.initialise_minhash(n).to(device)
    init_hl = eh.initialise_hll(n).to(device)

    def forward(b):
        hash_edge_index = torch.cat([ei, torch.stack([loops, loops])], dim=1)  # add_self_loops: a new tensor each time
        table = {0: {'minhash': init_mh, 'hll': init_hl}}
        cards = torch.zeros((n, K), device=device)
        for k in range(1, K + 1):
            table[k] = {'hll': eh.hll_prop(table[k - 1]['hll'], hash_edge_index),
                        'minhash': eh.minhash_prop(table[k - 1]['minhash'], hash_edge_index)}
            cards[:, k - 1] = eh.hll_count(table[k]['hll'])
        lk = links[b * batch:(b + 1) * batch]
        return table, eh.get_subgraph_features(lk, table, cards)

    held = None
    times = []
    for b in range(batches + 3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        cur = forward(b % 4)
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
        held = cur  # the training loop holds the previous forward's tensors while the next one runs
    del held
    steady = times[3:]
    # the reference's ops on CUDA: [nnz, 128] int64 + [nnz, 256] int8 message tensors per hop
    base_ms = None
    try:
        hash_edge_index = torch.cat([ei, torch.stack([loops, loops])], dim=1)
        torch_scatter_forward(init_mh, init_hl, hash_edge_index, 1)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch_scatter_forward(init_mh, init_hl, hash_edge_index, K)
        e.record()
        torch.cuda.synchronize()
        base_ms = s.elapsed_time(e)
    except RuntimeError as ex:  # out of memory on the message tensors
        base_ms = None
        torch.cuda.empty_cache()
        _ = ex
    st = dict(eh._prop.stats)
    return {'links_per_batch': batch, 'first_forward_ms': times[0], 'second_forward_ms': times[1],
            'steady_forward_ms': statistics.mean(steady), 'steady_forward_ms_min': min(steady),
            'what': 'K x (hll_prop + minhash_prop + hll_count) + get_subgraph_features of one batch, per forward',
            'torch_scatter_reduce_cuda_propagate_ms': base_ms,
            'torch_baseline_note': 'propagation only (index_select + scatter_reduce amax over nnz x width messages), '
                                   'the reference ops ELPH.forward issues on CUDA tensors',
            'session': st}


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    # the contract is ONE JSON line on stdout: keep a private handle to the real stdout and send everything
    # else that writes to fd 1 (NCCL's version banner, library chatter) to stderr
    global _OUT
    _OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    a = parse()
    g = graphs_module()
    spec = workload_spec(a, g)
    if a.impl == 'reference':
        run_reference(a, spec, g)
        return

    import warnings
    warnings.filterwarnings('ignore', message='datasketch is not importable')
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
    ei, links = make_inputs(spec, a, device, g)
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
    achieved = merge_bytes / (merge_ms * 1e-3) / 1e9 if merge_ms else None
    per_rank = None
    if distributed:
        # the slowest rank bounds the hop: report ITS figure (not rank 0's) and list all of them
        mine = torch.tensor([float(merge_bytes), float(merge_ms or 0.0)], device=device, dtype=torch.float64)
        allr = torch.empty((world, 2), device=device, dtype=torch.float64)
        dist.all_gather_into_tensor(allr.view(-1), mine)
        per_rank = [{'rank': q, 'algorithmic_bytes_per_launch': int(b), 'avg_launch_ms': m,
                     'achieved': (b / (m * 1e-3) / 1e9) if m > 0 else None} for q, (b, m) in enumerate(allr.tolist())]
        slow = max(per_rank, key=lambda x: x['avg_launch_ms'])
        merge_bytes, merge_ms, achieved = slow['algorithmic_bytes_per_launch'], slow['avg_launch_ms'], slow['achieved']
    roofline = {'bound': 'hbm', 'kernel': 'ss_khop_merge (merge_tma_kernel + merge_fixup_kernel)',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak if achieved else None,
                'traffic': None, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': merge_bytes,
                'avg_launch_ms': merge_ms, 'launches_timed': len(stage_ms.get('khop_merge', [])),
                'algorithmic_model': 'nnz*R + rows*R + 4*nnz + 8*(rows+1) + 4*rows, R = 768: every neighbour record counted '
                                     'once per incident edge (no-reuse gather model, SURVEY 8d)'}
    if distributed:
        roofline['per_rank'] = per_rank
        roofline['note'] = ('slowest rank shown; its launch also pushes finished rows to the peers over NVLink, which the '
                            'algorithmic bytes do not count; traffic (DRAM bytes) is only captured at N=1')
    else:
        tpath = os.path.join(ROOT, 'profiles', 'merge_traffic.json')
        traffic = None
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(spec['name'])
            except Exception:
                traffic = None
        if traffic and merge_ms:
            # `frac` can exceed 1: the algorithmic model counts every neighbour record once per edge while the
            # records of hubs hit in the 126 MB L2.  The HBM fraction proper is DRAM bytes / time:
            roofline['traffic'] = traffic
            roofline['traffic_source'] = ('static: dram__bytes_read.sum + dram__bytes_write.sum of ONE merge launch of this '
                                          'workload from a committed ncu --set full capture (profiles/), not re-measured in '
                                          'this run')
            roofline['dram_achieved'] = traffic / (merge_ms * 1e-3) / 1e9
            roofline['dram_frac'] = roofline['dram_achieved'] / peak
            roofline['frac_note'] = ('frac = algorithmic bytes / time / peak (contract); dram_frac = measured DRAM bytes / '
                                     'time / peak is the HBM utilisation (the difference is L2 reuse of hub records)')
    L_local = (link_slice(L, world, rank)[1] - link_slice(L, world, rank)[0]) if distributed else L
    lf_ms = sum(stage_ms.get('link_features', [])) / a.steps if stage_ms.get('link_features') else None
    lbytes = L_local * link_bytes(K)
    link_roofline = {'bound': 'hbm', 'kernel': ('link_features_batched_kernel (sharded tables: remote records over NVLink)' if distributed
                                                else ('link_features_kernel (per link)' if K == 3 else 'link_features_batched_kernel')),
                     'algorithmic_bytes_per_step': lbytes,
                     'ms_per_step': lf_ms, 'achieved': lbytes / (lf_ms * 1e-3) / 1e9 if lf_ms else None,
                     'links_per_s_kernel_only': L_local / (lf_ms * 1e-3) if lf_ms else None, 'peak': peak}
    if link_roofline['achieved']:
        link_roofline['frac'] = link_roofline['achieved'] / peak

    # ---- parity of the sharded build (N > 1), outside the timed region ----------------------------------------
    parity = None
    if distributed and not a.no_checks:
        tables, cards = eng.build_hash_tables(N, ei)
        feats = eng.get_subgraph_features(links, tables, cards)
        lo, hi = eng.bounds[rank], eng.bounds[rank + 1]
        mine = []
        for k in range(K + 1):
            mine += table_checksums(tables.records(k), lo, hi)
        mine += float_checksums(cards, lo, hi)
        mine += float_checksums(feats, 0, feats.shape[0])
        mine = torch.stack(mine)
        allc = torch.empty((world, mine.numel()), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allc.view(-1), mine)
        bounds = list(eng.bounds)
        exchange_mode, csr_path, halo_frac = eng.exchange, eng.csr_path, eng.halo_fraction
        del tables, cards, feats
        torch.cuda.empty_cache()
        if rank == 0:
            one = ssb.ElphHashes(engine_args(K), merge_variant=a.merge_variant)
            one.record_stride = None
            t1, c1 = one.build_hash_tables(N, ei)
            f1 = one.get_subgraph_features(links, t1, c1)
            ok_t, ok_c, ok_f = True, True, True
            for q in range(world):
                want = []
                for k in range(K + 1):
                    want += table_checksums(t1.records(k), bounds[q], bounds[q + 1])
                want += float_checksums(c1, bounds[q], bounds[q + 1])
                flo, fhi = link_slice(L, world, q)
                want += float_checksums(f1, flo, fhi)
                want = torch.stack(want)
                eq = (want == allc[q]).tolist()
                nt = 2 * (K + 1)
                ok_t = ok_t and all(eq[:nt])
                ok_c = ok_c and all(eq[nt:nt + 2])
                ok_f = ok_f and all(eq[nt + 2:])
            parity = {'tables_bit_equal': ok_t, 'cards_bit_equal': ok_c, 'features_bit_equal': ok_f,
                      'what': 'order-sensitive 64-bit checksums of every rank\'s own row block of hop tables 0..K, of its '
                              'block of cards and of its feature slice, against a single-GPU ElphHashes build on rank 0',
                      'exchange': exchange_mode, 'csr': csr_path, 'row_blocks': bounds, 'halo_fraction_rank0': halo_frac}
            del one, t1, c1, f1
            torch.cuda.empty_cache()
        barrier()

    # ---- end-to-end arm: host buffers through the public API -----------------------------------------
    e2e = None
    if not a.no_e2e:
        ei_h = torch.empty(ei.shape, dtype=ei.dtype, pin_memory=True).copy_(ei)
        links_h = torch.empty(links.shape, dtype=links.dtype, pin_memory=True).copy_(links)
        torch.cuda.synchronize()

        def e2e_step():
            if distributed:
                tables, cards = eng.build_hash_tables(N, ei_h)  # pinned host edges: every rank streams its own slice
                f = eng.get_subgraph_features(links_h, tables, cards)  # this rank's slice, back on the host
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
        e2e_stage = stage_totals(e2e_log, e2e_steps)
        if distributed:
            h2d = ei_h.numel() * 8 // world + L_local * 16   # approximately: each rank pulls its own slice of the list
            d2h = L_local * F * 4
        else:
            h2d = ei_h.numel() * 8 + links_h.numel() * 8
            d2h = N * K * 4 + L * F * 4
        h2d_gbs = measure_h2d_gbs(device) if rank == 0 else None
        e2e = {'value': L / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'ms_per_step': e2e_ms, 'steps': e2e_steps, 'stage_ms_per_step': e2e_stage}
        if h2d_gbs:
            e2e['pcie_h2d_gbs_measured'] = h2d_gbs
            e2e['pcie_floor_ms'] = h2d / h2d_gbs / 1e6
            e2e['note'] = ('pcie_floor_ms = h2d bytes of this rank / the pinned H2D rate measured in this run: the part of '
                           'ms_per_step no kernel can remove')
        del ei_h, links_h

    # ---- spot check + bounded CPU baseline (rank 0, N = 1 only) -------------------------------------------
    cpu_baseline = None
    main_check = None
    if rank == 0 and world == 1:
        if not a.no_checks:
            tables, cards = eh.build_hash_tables(N, ei)
            feats = eh.get_subgraph_features(links, tables, cards)
            main_check = sampled_check(eh, N, ei, tables, cards, links, feats, K, seed=a.seed)
            del tables, cards, feats
            torch.cuda.empty_cache()
        if not a.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            r = cpu_pass(g, a.cpu_scale, a.edge_factor, K, a.seed, L / N)
            cpu_baseline = {'value': r['links'] / r['seconds'], 'unit': UNIT, 'cores': torch.get_num_threads(),
                            'kind': 'port', 'sample': cpu_sample_desc(a.cpu_scale, a.edge_factor, K, r)}

    # ---- the OGB-shaped BASELINE configurations -------------------------------------------------------------
    del ei, links
    torch.cuda.empty_cache()
    configs = None
    if not a.no_configs and a.workload == 'rmat':
        configs = []
        if distributed:
            sh = ShardedElphHashes(engine_args(2), exchange=a.exchange)
            rec = run_ogb_config('citation2', a, g, device, peak, False, rank, world, dist_engine=sh)
            configs.append(rec)
            del sh
        else:
            for name in ('collab', 'ppa', 'citation2'):
                configs.append(run_ogb_config(name, a, g, device, peak, not a.no_checks, rank, world))

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'u32/u8 sketches, f32 estimates', 'data': 'synthetic',
            'config': {'workload': spec['name'], 'num_nodes': N, 'directed_edges': n_edges, 'nnz_with_self_loops': nnz,
                       'hops': K, 'num_perm': 128, 'hll_p': 8, 'links_per_step': L, 'features_per_link': F,
                       'merge_variant': a.merge_variant,
                       'partition': (f'node-sharded x{world} (row blocks balanced by cost), exchange={eng.exchange}, '
                                     f'csr={eng.csr_path}' if distributed else 'single'),
                       'hll_tables': eh.hll_tables_source,
                       'l2': 'inputs larger than L2 (each hop table is N*768 B), no explicit flush'},
            'features_per_s': value * F,
            'stage_ms_per_step': {k: sum(v) / a.steps for k, v in stage_ms.items()},
            'roofline': roofline, 'link_features_roofline': link_roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e,
            'gpu_launches': launches, 'clocks': clocks, 'parity_check': parity, 'sampled_check': main_check,
            'configs': configs,
        }
        print(json.dumps(line), file=_OUT, flush=True)
    if distributed:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
