#!/usr/bin/env python
"""Physical-mode throughput on the bench workload's scene (room.stl, 16 receivers, 8 bounces)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, load_mesh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
rx = np.array([[-14.0 + 28.0 * k / 15.0, 6.0, 5.0] for k in range(16)])
tr = Tracer(load_mesh(os.path.join(ROOT, "models/room.stl")), 2.998e8, 1e12, 200e-9, 8, n)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = tr.trace_physical([10, 0, 5], 1.0, rx, 0.1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"physical: {out['stats']['segments']} segments, {out['stats']['arrivals']} arrivals in {dt * 1e3:.1f} ms -> "
          f"{out['stats']['segments'] / dt:.3e} segments/s; max {out['dbm'].max():.2f} dBm")
