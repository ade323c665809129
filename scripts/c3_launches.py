#!/usr/bin/env python
"""One reference-faithful coverage map of C3 at a fraction of its rays (for launch lists / ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from rf_ray_tracing_warp_b200 import Tracer, synthetic_terrain  # noqa: E402
from rf_ray_tracing_warp_b200.coverage import plane_lattice  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0625
n = int((1 << 26) * scale)
tr = Tracer(synthetic_terrain(1024, 20.0, 17), 2.998e8, 100e9, 100e-9, 6, n, max_candidates=1 << 28, max_records=1 << 28)
rx = plane_lattice(1024, 1024, z=4.8)
import time  # noqa: E402
for it in range(2):
    t0 = time.perf_counter()
    cov = tr.coverage([10, 0, 4.5], 1, rx, 0.1)
    torch.cuda.synchronize()
    print(f"map {1e3 * (time.perf_counter() - t0):.1f} ms", cov["stats"], flush=True)
    del cov
