#!/usr/bin/env python
"""Order-free impulse-response binning (rfrt_bin_ir, deterministic = 0): the shared-memory privatised histogram
(k_bin_privatised: <= 4 receivers, window in shared memory, warp-aggregated adds) against one global fp64 atomic per
record (k_bin_atomic) on the same records — as many as C2's map produces (50 M), clustered the way arrivals are
(most records of a receiver share a few delay bins).  A fifth, empty receiver row switches the library to the atomic
kernel without touching the records.  Usage: bin_bench.py [n_records]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
n_rx, n_bins = 4, 10_000
lib = _lib.load()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
rx = torch.randint(0, n_rx, (n,), device=dev, generator=g, dtype=torch.int32)
# 80 % of a receiver's records in 8 "line of sight / first order" bins, the rest spread over the window
hot = torch.randint(0, 8, (n,), device=dev, generator=g) * 997 + rx.long() * 13 + 100
cold = torch.randint(0, n_bins, (n,), device=dev, generator=g)
pick = torch.rand(n, device=dev, generator=g) < 0.8
bins = torch.where(pick, hot, cold).to(torch.int64)
amp = torch.rand(n, device=dev, generator=g, dtype=torch.float64)
s = torch.cuda.current_stream().cuda_stream
out = {}
for name, rows in (("k_bin_privatised", n_rx), ("k_bin_atomic", n_rx + 1)):
    ir = torch.zeros((rows, n_bins), dtype=torch.float64, device=dev)
    best = 1e9
    for it in range(4):
        ir.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.rfrt_bin_ir(rx.data_ptr(), bins.data_ptr(), amp.data_ptr(), n, None, rows, n_bins, 0, ir.data_ptr(), s),
                   "rfrt_bin_ir")
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    out[name] = ir[:n_rx].clone()
    print(f"{name:18s} {best:8.3f} ms  {n / best / 1e6:8.2f} G records/s  {20 * n / best / 1e6:8.1f} GB/s of record reads", flush=True)
a, b = out["k_bin_privatised"], out["k_bin_atomic"]
print("max relative difference of the two results:", float(((a - b).abs() / b.abs().clamp_min(1e-300)).max()))
