#!/usr/bin/env python
"""Stage timing of Tracer.coverage on config C2 (CUDA events)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, load_mesh, sharding  # noqa: E402
from rf_ray_tracing_warp_b200._lib import check, float3  # noqa: E402
from rf_ray_tracing_warp_b200.coverage import plane_lattice  # noqa: E402

C = 2.998e8
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
tr = Tracer(load_mesh(os.path.join(ROOT, "models/almost_empty.stl")), C, 100e9, 100e-9, 4, n, max_candidates=1 << 26, max_records=1 << 26)
rx = plane_lattice(256, 256, z=2.0)
for it in range(2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    ev[0].record()
    job = tr.make_job(rx, 0.1)
    ev[1].record()
    lib, t = tr._lib, tr
    tx = float3([1, 0, 1])
    job.counters_t.zero_()
    check(lib.rfrt_trace(t._env, job.rxset, tx, 4, 0, n, 0, t._scratch(n).data_ptr(), t.chunk_rays, job.counters_t.data_ptr(),
                         job.cands.data_ptr(), job.cand_capacity, None, None, torch.cuda.current_stream().cuda_stream), "trace")
    ev[2].record()
    r = job.rec
    check(lib.rfrt_trace_receive(t._env, job.rxset, tx, 4, job.cands.data_ptr(), job.cand_capacity, job.counters_t.data_ptr(),
                                 1.0 / n, C, 100e9, r["ray"].data_ptr(), r["rx"].data_ptr(), r["nverts"].data_ptr(),
                                 r["bin"].data_ptr(), r["amp"].data_ptr(), r["dist"].data_ptr(), None, job.rec_capacity,
                                 torch.cuda.current_stream().cuda_stream), "receive")
    ev[3].record()
    c = job.counters()
    rec = {k: (v[:c["records"]] if v is not None else None) for k, v in r.items()}
    ev[4].record()
    rec = sharding.sort_records(rec)
    ev[5].record()
    p = tr.rx_power(rec, rx.shape[0])
    ev[6].record()
    torch.cuda.synchronize()
    names = ["rxset build", "trace_env", "trace_receive", "counters", "sort records", "rx_power (torch csr + kernel)"]
    print(c, {nm: round(ev[i].elapsed_time(ev[i + 1]), 2) for i, nm in enumerate(names)})
    job.close()
