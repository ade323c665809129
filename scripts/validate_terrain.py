#!/usr/bin/env python
"""One-off validation at a size the test suite does not afford: the BVH walk on a synthetic terrain of n_grid^2 quads
against the CPU restatement (its own median-split BVH == its brute force, tested on the CPU) — hit triangle and hit
distance of every (ray, bounce) bit for bit.  Usage: validate_terrain.py [n_grid] [log2 rays]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu  # noqa: E402
from rf_ray_tracing_warp_b200 import Tracer, synthetic_terrain  # noqa: E402

n_grid = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
B, tx = 6, [10, 0, 4.5]
mesh = synthetic_terrain(n_grid, 20.0, 17)
soup = mesh.triangles.astype(np.float32)
tr = Tracer(mesh, 2.998e8, 100e9, 100e-9, B, n)
out = tr.trace_segments(tx, dump=True)
t0 = time.perf_counter()
seg, tri, t = cpu.trace_env(soup, tx, B, 0, n, bvh=cpu.Bvh(soup))
print(f"oracle: {seg} segments in {time.perf_counter() - t0:.1f} s on {cpu.max_threads()} threads; GPU: {out['segments']} segments, "
      f"tree depth {tr.mesh_info()['max_depth']}")
g_tri, g_t = out["hit_tri"].cpu().numpy(), out["hit_t"].cpu().numpy()
assert out["segments"] == seg
assert np.array_equal(g_tri, tri), f"{(g_tri != tri).sum()} hit triangles differ"
assert np.array_equal(g_t.view(np.uint32), t.view(np.uint32)), "hit distances differ"
print(f"OK: {len(soup)} triangles, {n} rays x {B} bounces: hit triangle and distance identical on all {seg} segments")
