#!/usr/bin/env python
"""Coverage-map milliseconds (BASELINE.json's second metric) for C2 / C3, with the CPU restatement beside them.

    python scripts/coverage_bench.py [C2] [C3] [--cpu-rays 262144]

reference mode : Tracer.coverage — identical per receiver to coverage.py:38-55 (one trace, literal replay of every
                 (ray, receiver) pair, 10 000-bin impulse responses, convolution power)
physical mode  : Tracer.trace_physical — one pass, coherent field per receiver
cpu            : the oracle restatement of what the reference does for ONE receiver of the lattice (full re-trace of a
                 bounded ray sample through kernel.py:38-98 + tracer.py:84-117 + coverage.py:45-55), on all host threads;
                 the reference repeats that for every receiver, so its map time is that x receivers x (N / sample).
"""
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
    ap.add_argument("configs", nargs="*", default=["C2"])
    ap.add_argument("--cpu-rays", type=int, default=1 << 18)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reference-at-full-size", action="store_true", help="also time the reference-faithful C3 map at full size (~10 s per run)")
    args = ap.parse_args()
    for cfg in args.configs:
        if cfg == "REF":  # the reference's own coverage.py:15-16,38-40 workload (room.stl, tx as main.py:30)
            from rf_ray_tracing_warp_b200.coverage import reference_lattice
            mesh, tx, n, B, grid = load_mesh(os.path.join(ROOT, "models/room.stl")), [10, 0, 5], 1_000_000, 2, 16
            rx = reference_lattice()
        elif cfg == "C2":
            mesh, tx, n, B, grid, z = load_mesh(os.path.join(ROOT, "models/almost_empty.stl")), [1, 0, 1], 1 << 24, 4, 256, 2.0
        else:
            mesh, tx, n, B, grid, z = synthetic_terrain(1024, 20.0, 17), [10, 0, 4.5], 1 << 26, 6, 1024, 4.8
        n = int(n * args.scale)
        if cfg != "REF":
            rx = plane_lattice(grid, grid, z=z)
        cap = (1 << 28) if cfg == "C3" else (1 << 26)
        tr = Tracer(mesh, C, 100e9, 100e-9, B, n, max_candidates=cap, max_records=cap)
        out = dict(cfg=cfg, rays=n, bounces=B, receivers=int(rx.shape[0]), triangles=tr.mesh_info()["n_triangles"])

        def timed(fn, reps):
            fn(); torch.cuda.synchronize()
            ts, r = [], None
            for _ in range(reps):
                r = None  # (a reference-mode result holds the dense impulse responses: 84 GB for C3)
                t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            r.pop("impulse_response", None)
            return min(ts) * 1e3, r

        ms, cov = timed(lambda: tr.trace_physical(tx, 1.0, rx, 0.1), 3)
        out.update(physical_ms=ms, physical_arrivals=cov["stats"]["arrivals"], physical_segments=cov["stats"]["segments"],
                   physical_max_dbm=float(np.nanmax(np.where(cov["power"] > 0, cov["dbm"], np.nan))))
        if cfg in ("C2", "REF") or args.scale <= 0.25 or args.reference_at_full_size:
            ms, cov = timed(lambda: tr.coverage(tx, 1, rx, 0.1), 1 if cfg == "C3" else 2)
            out.update(reference_ms=ms, reference_records=cov["stats"]["records"], reference_max_dbm=float(np.nanmax(cov["dbm"])))
        if args.cpu_rays > 0:
            from oracle import cpu, geometry, post
            soup = mesh.triangles.astype(np.float32)
            bvh = cpu.Bvh(soup)
            k = rx.shape[0] // 2 + grid // 2
            t0 = time.perf_counter()
            tids, rows = cpu.trace_received(soup, geometry.rx_soup(rx[k], 0.1), tx, B, 0, args.cpu_rays, bvh=bvh)
            ir = post.impulse_response(post.clean_paths(rows, np.ones(len(rows), np.uint32)), 1, n, C, 100e9, 100e-9)
            post.rx_power(ir, 100e-9)
            dt = time.perf_counter() - t0
            out.update(cpu_one_receiver_sample_s=dt, cpu_sample_rays=args.cpu_rays, cpu_threads=cpu.max_threads(),
                       cpu_map_extrapolated_s=dt * rx.shape[0] * n / args.cpu_rays)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
