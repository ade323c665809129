#!/usr/bin/env python
"""Stage timing of the dense coverage path (CUDA events) on C2 or C3 (scaled)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, load_mesh, synthetic_terrain  # noqa: E402
from rf_ray_tracing_warp_b200.coverage import plane_lattice  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.125
C = 2.998e8
if cfg == "C2":
    mesh, tx, n, B, rx = load_mesh(os.path.join(ROOT, "models/almost_empty.stl")), [1, 0, 1], 1 << 24, 4, plane_lattice(256, 256, z=2.0)
else:
    mesh, tx, n, B, rx = synthetic_terrain(1024, 20.0, 17), [10, 0, 4.5], 1 << 26, 6, plane_lattice(1024, 1024, z=4.8)
n = int(n * scale)
tr = Tracer(mesh, C, 100e9, 100e-9, B, n, max_candidates=1 << 26, max_records=1 << 26)
L = 10000
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ir = torch.zeros((rx.shape[0], L), dtype=torch.float64, device=tr.device)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    job = tr.make_job(rx, 0.1)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    t_enq = t_bin = 0.0
    chunk = (1 << 19) if cfg == "C3" else (1 << 24)
    pos = 0
    while pos < n:
        hi = min(n, pos + chunk)
        a = time.perf_counter()
        job.enqueue(tx, 1.0, ray_range=(pos, hi))
        c = job.counters()
        b = time.perf_counter()
        assert c["candidates"] <= job.cand_capacity, c
        job.bin_into(ir)
        torch.cuda.synchronize()
        t_bin += time.perf_counter() - b
        t_enq += b - a
        pos = hi
    t3 = time.perf_counter()
    p = tr.rx_power_dense(ir)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    job.close()
    print(dict(zero_ir=round(t1 - t0, 3), rxset=round(t2 - t1, 3), trace_receive=round(t_enq, 3), bin=round(t_bin, 3),
               power=round(t4 - t3, 3), records=c["records"]))
    del ir
