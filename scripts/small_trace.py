#!/usr/bin/env python
"""C4's environment trace alone (room.stl, 8 bounces, the bench's 16 receivers): k_trace_small timing / ncu target.
Usage: small_trace.py [log2 rays] [receivers 0|1]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, load_mesh  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
with_rx = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = 1 << lg
tx = [10.0, 0.0, 5.0]
rx = np.array([[-14 + 28 * k / 15, 6.0, 5.0] for k in range(16)])
tr = Tracer(load_mesh(os.path.join(ROOT, "models", "room.stl")), 2.998e8, 1e12, 200e-9, 8, n, chunk_rays=n)
job = tr.make_job(rx, 0.1) if with_rx else None
dirs = torch.empty((n, 4), dtype=torch.float32, device=tr.device)
tr.ray_directions(0, n, out=dirs)
from rf_ray_tracing_warp_b200 import _lib  # noqa: E402
from rf_ray_tracing_warp_b200._lib import check, float3  # noqa: E402
counters = torch.zeros(_lib.CTR_COUNT, dtype=torch.int64, device=tr.device)
best = 1e9
for it in range(5):
    counters.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(tr._lib.rfrt_trace(tr._env, job.rxset if job else 0, float3(tx), 8, 0, n, _lib.FLAG_DIRS_READY, dirs.data_ptr(), n,
                             counters.data_ptr(), job.cands.data_ptr() if job else None, job.cand_capacity if job else 0, None, None,
                             torch.cuda.current_stream().cuda_stream), "rfrt_trace")
    e1.record()
    torch.cuda.synchronize()
    if it:
        best = min(best, e0.elapsed_time(e1))
seg = int(counters[_lib.CTR_SEGMENTS].item())
print(f"k_trace_small: {seg} segments in {best:.3f} ms -> {seg / best * 1e3:.4e} segments/s", flush=True)
