"""Small end-to-end pass over every kernel family, sized for compute-sanitizer (racecheck / synccheck / memcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Covers the TMA / mbarrier kernels (gather4 merge, bulk merge, batched link kernel with bulk copies + cp.async, TMA-pair
link kernel), the half-record layouts, the streaming and histogram CSR builds, the guarded ELPH session and the halo
mask kernel.  Results are compared with each other (engine vs engine), so a data race that changes bits also fails."""
import os
import sys
from argparse import Namespace

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402
from subgraph_sketching_b200 import _lib  # noqa: E402
from subgraph_sketching_b200._lib import check, lib  # noqa: E402
from subgraph_sketching_b200.graphs import rmat_edges  # noqa: E402

dev = torch.device('cuda', 0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 11
n = 1 << scale
ei = rmat_edges(scale, 8, 1, dev)
g = torch.Generator().manual_seed(0)
links = torch.randint(0, n, (777, 2), generator=g).to(dev)
grouped = links.clone()
grouped[:, 0] = links[torch.arange(777, device=dev) // 50, 0]
ref = None
for K in (1, 2, 3):
    args = Namespace(max_hash_hops=K, floor_sf=False, minhash_num_perm=128, hll_p=8, use_zero_one=True)
    base = None
    for variant in ('tma', 'bulk', 'ldg'):
        for fast in ('1', '0'):
            os.environ['SS_B200_CSR_FAST'] = fast
            eh = ssb.ElphHashes(args, merge_variant=variant)
            t, c = eh.build_hash_tables(n, ei)
            if base is None:
                base = (t, c)
            for k in range(K + 1):
                assert torch.equal(t.records(k), base[0].records(k)), (K, variant, fast, k)
            assert torch.equal(c, base[1])
    os.environ['SS_B200_CSR_FAST'] = '1'
    eh = ssb.ElphHashes(args)
    feats = {}
    for mode in ('ldg', 'batched', 'tma'):
        os.environ['SS_B200_LINKS'] = mode
        feats[mode] = [eh.get_subgraph_features(lk, *base) for lk in (links, grouped)]
    os.environ.pop('SS_B200_LINKS')
    for mode in ('batched', 'tma'):
        assert all(torch.equal(a, b) for a, b in zip(feats[mode], feats['ldg'])), (K, mode)
    # the ELPH session: singles, pair fusion, guarded re-enqueue
    init_m, init_h = eh.initialise_minhash(n).to(dev), eh.initialise_hll(n).to(dev)
    loops = torch.arange(int(ei.max()) + 1, device=dev)  # add_self_loops(edge_index) without num_nodes (hashing.py:148)
    prev = None
    for fwd in range(4):
        he = torch.cat([ei, torch.stack([loops, loops])], dim=1)
        m, h = init_m, init_h
        for k in range(1, K + 1):
            h, m = eh.hll_prop(h, he), eh.minhash_prop(m, he)
            eh.hll_count(h)
        assert torch.equal(m, base[0][K]['minhash']) and torch.equal(h, base[0][K]['hll']), (K, fwd)
        prev = (m, h)
# halo masks of a symmetric graph from one rank's CSR rows (rank 1 of 3)
rowptr, colidx, nnz, _ = ssb.build_csr(ei, dev, num_rows=n, add_loops=True)
bounds = torch.tensor([0, n // 5, n // 2, n], dtype=torch.int64, device=dev)
lo, hi = n // 5, n // 2
rp = (rowptr[lo:hi + 1] - rowptr[lo]).contiguous()
ci = colidx[int(rowptr[lo]):int(rowptr[hi])].contiguous()
mask = torch.zeros((hi - lo + 7) // 4 * 4, dtype=torch.uint8, device=dev)
mark = torch.zeros(n, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
check(lib.ss_halo_from_csr(rp.data_ptr(), ci.data_ptr(), hi - lo, ci.numel(), bounds.data_ptr(), 3, 1, mask.data_ptr(),
                           mark.data_ptr(), st), 'ss_halo_from_csr')
owner = torch.bucketize(ci.long(), bounds[1:-1], right=True)
rows = torch.repeat_interleave(torch.arange(hi - lo, device=dev), rp[1:] - rp[:-1])
want = torch.zeros(hi - lo, dtype=torch.int64, device=dev)
for q, bit in ((0, 1), (2, 2)):
    tmp = torch.zeros(hi - lo, dtype=torch.int64, device=dev)
    tmp[rows[owner == q]] = bit
    want |= tmp
assert torch.equal(mask[:hi - lo].long(), want), 'halo mask differs'
wm = torch.zeros(n, dtype=torch.uint8, device=dev)
wm[ci.long()] = 1
assert torch.equal(mark, wm), 'marks differ'
torch.cuda.synchronize()
print(f'sanitize_smoke OK: scale {scale}, {nnz} neighbours, launches = {_lib.lib.launches}')
