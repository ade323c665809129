#!/usr/bin/env python
"""synthetic terrain of ~20 M triangles, 1024 x 1024 receiver lattice, rays sharded over the
GPUs of one box, per-GPU grids combined with ONE all-reduce (NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_c5.py [--rays R]

Prints one JSON line from rank 0: environment-only trace (segments/s), physical-mode coverage map (ms, incl. the
all-reduce) and an invariance checksum of the grid (sum of |field|^2), which must not depend on N beyond fp64
summation order."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rf_ray_tracing_warp_b200 import Tracer, synthetic_terrain  # noqa: E402
from rf_ray_tracing_warp_b200.coverage import plane_lattice  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 30)
    ap.add_argument("--grid", type=int, default=3162, help="terrain quads per side (3162 -> 19 996 488 triangles)")
    ap.add_argument("--lattice", type=int, default=1024)
    ap.add_argument("--bounces", type=int, default=6)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mesh = synthetic_terrain(args.grid, 20.0, 17)
    tr = Tracer(mesh, 2.998e8, 100e9, 100e-9, args.bounces, args.rays, shard=world > 1)
    rx = plane_lattice(args.lattice, args.lattice, z=4.8)
    tx = [10.0, 0.0, 4.5]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    tr.trace_segments(tx, ray_range=(tr.ray_range[0], tr.ray_range[0] + (1 << 20)))  # warm-up (sort workspace, clocks)
    sync(); t0 = time.perf_counter()
    seg = tr.trace_segments(tx)
    sync(); t_env = time.perf_counter() - t0
    segs = torch.tensor([seg["segments"]], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(segs)
    sync(); t0 = time.perf_counter()
    cov = tr.trace_physical(tx, 1.0, rx, 0.1)
    sync(); t_cov = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps(dict(cfg="C5", n_gpus=world, rays=args.rays, triangles=tr.mesh_info()["n_triangles"],
                              build_ms=tr.mesh_info()["build_ms"], env_segments=int(segs.item()), env_seconds=t_env,
                              env_segments_per_s=int(segs.item()) / t_env, coverage_physical_ms=1e3 * t_cov,
                              arrivals=cov["stats"]["arrivals"], grid_power_sum=float(np.sum(cov["power"])),
                              grid_max_dbm=float(np.max(cov["dbm"])))), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
