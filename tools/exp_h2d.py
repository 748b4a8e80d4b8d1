"""Host -> device ingestion probe (GPU only; tuning aid): how fast can a pinned int64 [2, E] edge list reach the
GPU?  Compares one DMA copy, chunked DMA copies, and the in-place (zero-copy) kernel read the CSR build uses.
    python tools/exp_h2d.py [n_edges]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import subgraph_sketching_b200 as ssb  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 260_000_000
dev = torch.device('cuda', 0)
n = 1 << 24
g = torch.Generator().manual_seed(0)
host = torch.empty((2, E), dtype=torch.int64, pin_memory=True)
host.random_(0, n, generator=g)
gb = host.numel() * 8 / 1e9
dst = torch.empty_like(host, device=dev)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


ms = timed(lambda: dst.copy_(host, non_blocking=True))
print(f'one DMA copy of {gb:.2f} GB: {ms:.1f} ms = {gb / ms * 1e3:.1f} GB/s', flush=True)
for chunk_mb in (64, 256):
    ce = chunk_mb * (1 << 20) // 8

    def chunked():
        for lo in range(0, E, ce):
            hi = min(lo + ce, E)
            dst[:, lo:hi].copy_(host[:, lo:hi], non_blocking=True)
    ms = timed(chunked)
    print(f'chunked DMA ({chunk_mb} MB per row slice): {ms:.1f} ms = {gb / ms * 1e3:.1f} GB/s', flush=True)
ms = timed(lambda: ssb.build_csr(host, dev, num_rows=n, add_loops=True))
print(f'build_csr reading the pinned list in place: {ms:.1f} ms = {gb / ms * 1e3:.1f} GB/s (whole CSR build)', flush=True)
ms = timed(lambda: ssb.build_csr(dst, dev, num_rows=n, add_loops=True))
print(f'build_csr from device memory: {ms:.1f} ms', flush=True)
