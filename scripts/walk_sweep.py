#!/usr/bin/env python
"""Tuning sweep of the BVH-scene trace kernel (k_trace_walk) on the synthetic terrain: refill / node-loop thresholds
(RFRT_WALK_REFILL, RFRT_WALK_NODE_MIN).  Usage: walk_sweep.py [n_grid] [n_rays]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, synthetic_terrain  # noqa: E402

n_grid = int(sys.argv[1]) if len(sys.argv) > 1 else 3162
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
tx = [10, 0, 4.5]
tr = Tracer(synthetic_terrain(n_grid, 20.0, 17), 2.998e8, 100e9, 100e-9, 6, n_rays)
print(tr.mesh_info())


def run(tag):
    best = 1e9
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = tr.trace_segments(tx)
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    c = tr.trace_segments(tx, checksum=True)
    print(f"{tag:28s} {best:7.2f} ms  {out['segments'] / best * 1e3:.3e} segments/s  checksum {c['checksum']:016x} "
          f"nodes/seg {c['node_visits'] / c['segments']:.2f} tris/seg {c['tri_tests'] / c['segments']:.2f}", flush=True)


run("defaults")
for refill in (8, 16, 24):
    for node_min in (4, 8, 12):
        os.environ["RFRT_WALK_REFILL"], os.environ["RFRT_WALK_NODE_MIN"] = str(refill), str(node_min)
        run(f"walk refill={refill} node_min={node_min}")
