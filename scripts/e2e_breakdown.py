#!/usr/bin/env python
"""Stage timing (CUDA events, max over ranks) of the end-to-end call bench.py reports as `e2e`
(Tracer.compute_cir_multi on the C4 workload), at 1 GPU or under torchrun at N GPUs:

    python scripts/e2e_breakdown.py [rays_per_gpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/e2e_breakdown.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, _lib, load_mesh  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
rx = np.array([[-14.0 + 28.0 * k / 15.0, 6.0, 5.0] for k in range(16)])
tr = Tracer(load_mesh(os.path.join(ROOT, "models/room.stl")), 2.998e8, 1e12, 200e-9, 8, R * world, device=dev,
            ray_range=(rank * R, (rank + 1) * R), shard=world > 1)
L = 200000
ir = torch.empty((16, L), dtype=torch.float64, device=dev)
pinned = torch.empty((16, L), dtype=torch.float64, pin_memory=True)
seg_cap = 1 << 18
names = ["rxset build + buffers", "trace + replay", "pack", "all-gather", "sort", "ordered IR", "D2H IR + summary"]
rows = []
for it in range(6):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    job = tr.make_job(rx, 0.1, want_paths=False)
    ev[1].record()
    job.enqueue([10, 0, 5], 1.0)
    ev[2].record()
    # collect(), stage by stage
    lib, r = tr._lib, job.rec
    job.collect(seg_cap, ir=None) if it == 0 else None   # allocates the exchange buffers once per job shape
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.rfrt_records_pack(job.counters_t.data_ptr(), r["ray"].data_ptr(), r["rx"].data_ptr(), r["nverts"].data_ptr(),
                                     r["bin"].data_ptr(), r["amp"].data_ptr(), r["dist"].data_ptr(), None, job.rec_capacity, 0,
                                     job._seg_local.data_ptr(), seg_cap, s), "pack") if it else None
    ev[3].record()
    if world > 1 and it:
        dist.all_gather_into_tensor(job._segs, job._seg_local)
    ev[4].record()
    if it:
        o, n = job._sorted, world * seg_cap
        ws = tr._workspace(n)
        _lib.check(lib.rfrt_records_sort(job._segs.data_ptr(), world, seg_cap, 0, 16, o["ray"].data_ptr(), o["rx"].data_ptr(),
                                         o["nverts"].data_ptr(), o["bin"].data_ptr(), o["amp"].data_ptr(), o["dist"].data_ptr(), None,
                                         job._summary.data_ptr(), ws.data_ptr(), ws.numel(), s), "sort")
    ev[5].record()
    if it:
        ir.zero_()
        _lib.check(lib.rfrt_arrivals_build(o["rx"].data_ptr(), o["bin"].data_ptr(), o["amp"].data_ptr(), n, job._summary.data_ptr(), 16, L,
                                           None, None, None, ir.data_ptr(), ws.data_ptr(), ws.numel(), s), "arrivals")
    ev[6].record()
    if it:
        pinned.copy_(ir, non_blocking=True)
        c = job._summary.cpu()
    ev[7].record()
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t0)
    job.close()
    if it >= 2:
        rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))] + [wall])
t = torch.tensor(rows, dtype=torch.float64, device=dev).mean(0)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
t0 = time.perf_counter()
for _ in range(3):
    out = tr.compute_cir_multi(np.array([10., 0, 5]), 1.0, rx, 0.1)
    pinned.copy_(out["impulse_response"], non_blocking=True)
    torch.cuda.synchronize()
api = torch.tensor([1e3 * (time.perf_counter() - t0) / 3], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(api, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps(dict(n_gpus=world, rays_per_gpu=R, records=int(c[_lib.SUM_RECORDS]),
                          stage_ms={k: round(float(v), 3) for k, v in zip(names + ["wall"], t.tolist())},
                          compute_cir_multi_plus_d2h_ms=round(float(api.item()), 3))), flush=True)
if world > 1:
    dist.destroy_process_group()
