#!/usr/bin/env python
"""BASELINE.md config C1 (the reference's own CPU-runnable case: room.stl, 1 M rays, 3 bounces, 1 receiver, impulse
response as main.py) — Tracer.compute_cir on the GPU next to the CPU restatement on 1 thread and on all host threads."""
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
    from oracle import cpu, geometry, post
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    stl = os.path.join(ROOT, "models/room.stl")
    n, B, tx, rx, r = 1_000_000, 3, [10, 0, 5], [-10, 0, 5], 0.1
    tr = Tracer(load_mesh(stl), C, 100e9, 200e-9, B, n)
    tr.compute_cir(tx, 1, rx, r); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); paths, ir = tr.compute_cir(tx, 1, rx, r); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    soup = geometry.load_stl_soup(stl)
    res = {}
    for name, nt in (("cpu_1_thread", 1), ("cpu_all_threads", 0)):
        t0 = time.perf_counter()
        tids, rows = cpu.trace_received(soup, geometry.rx_soup(rx, r), tx, B, 0, n, nthreads=nt)
        o_ir = post.impulse_response(post.clean_paths(rows, np.ones(len(rows), np.uint32)), 1, n, C, 100e9, 200e-9)
        res[name + "_s"] = time.perf_counter() - t0
    assert np.array_equal(ir != 0, o_ir != 0) and np.allclose(ir, o_ir, rtol=1e-5, atol=0)
    print(json.dumps(dict(cfg="C1", gpu_compute_cir_ms=1e3 * min(ts), segments=tr.last_stats["segments"], received_paths=len(paths),
                          cpu_threads=cpu.max_threads(), **res)), flush=True)


if __name__ == "__main__":
    main()
