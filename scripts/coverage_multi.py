#!/usr/bin/env python
"""C2's reference-faithful coverage map with the rays sharded over the GPUs of one box (strong scaling): trace + replay
per rank, one reduce-scatter of the impulse-response rows, power per row block, all-gather of the powers.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/coverage_multi.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, load_mesh  # noqa: E402
from rf_ray_tracing_warp_b200.coverage import plane_lattice  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, B = 1 << 24, 4
tr = Tracer(load_mesh(os.path.join(ROOT, "models/almost_empty.stl")), 2.998e8, 100e9, 100e-9, B, n, shard=world > 1,
            max_candidates=1 << 26, max_records=1 << 26)
rx = plane_lattice(256, 256, z=2.0)
ms = []
for it in range(4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    cov = tr.coverage([1, 0, 1], 1, rx, 0.1)
    torch.cuda.synchronize()
    ms.append(1e3 * (time.perf_counter() - t0))
    power = cov["power"]
    del cov
t = torch.tensor([min(ms[1:])], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps(dict(cfg="C2", n_gpus=world, map_ms=float(t.item()), map_ms_all=ms, cells_with_signal=int(np.isfinite(power).sum()),
                          power_sum=float(np.nansum(power)))), flush=True)
if world > 1:
    dist.destroy_process_group()
