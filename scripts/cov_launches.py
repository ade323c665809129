#!/usr/bin/env python
"""Reference-faithful coverage maps of C2 or C3 (optionally at a fraction of the rays) for launch lists / ncu captures.
Usage: cov_launches.py C2|C3 [scale]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from rf_ray_tracing_warp_b200 import Tracer, load_mesh, synthetic_terrain  # noqa: E402
from rf_ray_tracing_warp_b200.coverage import plane_lattice  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
if cfg == "C2":
    mesh, tx, n, B, grid, z, cap = load_mesh(os.path.join(ROOT, "models/almost_empty.stl")), [1, 0, 1], 1 << 24, 4, 256, 2.0, 1 << 26
else:
    mesh, tx, n, B, grid, z, cap = synthetic_terrain(1024, 20.0, 17), [10, 0, 4.5], 1 << 26, 6, 1024, 4.8, 1 << 28
n = int(n * scale)
tr = Tracer(mesh, 2.998e8, 100e9, 100e-9, B, n, max_candidates=cap, max_records=cap)
rx = plane_lattice(grid, grid, z=z)
for it in range(3):
    t0 = time.perf_counter()
    cov = tr.coverage(tx, 1, rx, 0.1)
    torch.cuda.synchronize()
    print(f"map {1e3 * (time.perf_counter() - t0):.1f} ms", cov["stats"], flush=True)
    del cov
