#!/usr/bin/env python
"""Environment-only trace on the synthetic terrain (configs C3 / C5 meshes): BVH build time + segments/s."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, synthetic_terrain  # noqa: E402

n_grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
tx = [10, 0, 4.5] if len(sys.argv) <= 3 else [float(v) for v in sys.argv[3].split(",")]
t0 = time.perf_counter()
mesh = synthetic_terrain(n_grid, 20.0, 17)
t1 = time.perf_counter()
tr = Tracer(mesh, 2.998e8, 100e9, 100e-9, 6, n_rays)
torch.cuda.synchronize()
t2 = time.perf_counter()
print("mesh gen %.2fs, upload+build %.3fs" % (t1 - t0, t2 - t1), tr.mesh_info())
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = tr.trace_segments(tx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"trace: {out['segments']} segments, {out['env_hits']} hits in {ms:.1f} ms -> {out['segments'] / ms * 1e3:.3e} segments/s")
out = tr.trace_segments(tx, checksum=True)
print(f"checksum {out['checksum']:016x}; per segment: {out['node_visits'] / out['segments']:.2f} internal nodes fetched, "
      f"{out['tri_tests'] / out['segments']:.2f} triangles tested")
