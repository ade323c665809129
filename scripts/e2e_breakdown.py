#!/usr/bin/env python
"""Stage timing of the end-to-end call bench.py reports as `e2e` (Tracer.compute_cir_multi on the C4 workload)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, load_mesh
from rf_ray_tracing_warp_b200 import sharding
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
rx = np.array([[-14.0 + 28.0 * k / 15.0, 6.0, 5.0] for k in range(16)])
tr = Tracer(load_mesh(os.path.join(ROOT, "models/room.stl")), 2.998e8, 1e12, 200e-9, 8, R, chunk_rays=chunk)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T()
    job = tr.make_job(rx, 0.1, want_paths=False); t1 = T()
    job.enqueue([10, 0, 5], 1.0); t2 = T()
    c = job.counters(); t3 = T()
    n = c["records"]
    rec = {k: (v[:n] if v is not None else None) for k, v in job.rec.items()}
    rec = sharding.sort_records(rec); t4 = T()
    ir = tr._dense_ir(rec, 16); t5 = T()
    h = ir.cpu().numpy(); t6 = T()
    job.close(); t7 = T()
    print(dict(make_job=round(1e3*(t1-t0),2), enqueue=round(1e3*(t2-t1),2), counters=round(1e3*(t3-t2),2), sort=round(1e3*(t4-t3),2),
               dense_ir=round(1e3*(t5-t4),2), d2h=round(1e3*(t6-t5),2), close=round(1e3*(t7-t6),2), total=round(1e3*(t7-t0),2)))
t0 = T(); out = tr.compute_cir_multi(np.array([10., 0, 5]), 1.0, rx, 0.1); h = out["impulse_response"].cpu().numpy(); t1 = T()
print("compute_cir_multi + d2h ms", round(1e3*(t1-t0), 2))
