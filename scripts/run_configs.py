#!/usr/bin/env python
"""Runs the BASELINE.md configs C1..C3 (C4 is bench.py) through the public API and prints timings + counters."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
C = 2.998e8


def main():
    import torch
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh, synthetic_terrain
    from rf_ray_tracing_warp_b200.coverage import plane_lattice
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C1", "C2"])
    ap.add_argument("--scale", type=float, default=1.0, help="scale the ray count")
    args = ap.parse_args()
    for cfg in args.configs:
        t0 = time.perf_counter()
        if cfg == "C1":
            tr = Tracer(load_mesh(os.path.join(ROOT, "models/room.stl")), C, 100e9, 200e-9, 3, int(1_000_000 * args.scale))
            torch.cuda.synchronize(); t1 = time.perf_counter()
            paths, ir = tr.compute_cir([10, 0, 5], 1, [-10, 0, 5], 0.1)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            out = dict(cfg=cfg, build_s=t1 - t0, run_s=t2 - t1, paths=len(paths), nonzero_bins=int(np.count_nonzero(ir)), **tr.last_stats)
        elif cfg in ("C2", "C3"):
            if cfg == "C2":
                mesh, tx, n, B, rx = load_mesh(os.path.join(ROOT, "models/almost_empty.stl")), [1, 0, 1], 1 << 24, 4, plane_lattice(256, 256, z=2.0)
            else:
                mesh, tx, n, B, rx = synthetic_terrain(1024, 20.0, 17), [10, 0, 4.5], 1 << 26, 6, plane_lattice(1024, 1024, z=4.8)
            n = int(n * args.scale)
            tr = Tracer(mesh, C, 100e9, 100e-9, B, n, max_candidates=1 << 26, max_records=1 << 26)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            cov = tr.coverage(tx, 1, rx, 0.1)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            cov2 = tr.coverage(tx, 1, rx, 0.1)
            torch.cuda.synchronize(); t3 = time.perf_counter()
            out = dict(cfg=cfg, build_s=t1 - t0, run_s=t2 - t1, run2_s=t3 - t2, covered=int(np.isfinite(cov["dbm"]).sum()),
                       max_dbm=float(np.nanmax(cov["dbm"])), mesh=tr.mesh_info(), **cov["stats"])
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
