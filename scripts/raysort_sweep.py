#!/usr/bin/env python
"""Direction-coherent ray order of the BVH scenes: radix passes (8 bits each, from the top of the 24-bit direction code)
against trace time on the synthetic terrain.  Usage: raysort_sweep.py [n_grid] [n_rays]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, synthetic_terrain  # noqa: E402

n_grid = int(sys.argv[1]) if len(sys.argv) > 1 else 3162
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
tr = Tracer(synthetic_terrain(n_grid, 20.0, 17), 2.998e8, 100e9, 100e-9, 6, n_rays)
for passes in (3, 2, 1):
    os.environ["RFRT_RAY_SORT_PASSES"] = str(passes)
    best = 1e9
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = tr.trace_segments([10, 0, 4.5])
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    print(f"{passes} passes: {best:.2f} ms  {out['segments'] / best * 1e3:.3e} segments/s", flush=True)
