#!/usr/bin/env python
"""bench.py — traced ray-segments/s of the hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rays R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): C4 of BASELINE.md — models/room.stl (44 triangles), tx (10,0,5), 16 receivers
(r = 0.1) on the line y=6, z=5, x=-14..14, 8 bounces, 1 TS/s x 200 ns = 200 000 delay bins, 268 435 456 rays
PER GPU (weak scaling: rank g traces global ray ids [g*R, (g+1)*R)), synthetic = seeded ray directions.
One step = the whole hot path over that batch: directions -> environment trace (+ receiver tests) -> literal
replay of the received candidates -> impulse-response binning.

`value`  : segments (device counter, summed over ranks) / device time of K steps (CUDA events, max over ranks),
           inputs resident in HBM.
`e2e`    : same metric through the public API (Tracer.compute_cir_multi) with HOST inputs (tx / receiver
           positions as NumPy) and the impulse responses copied back to the host inside the timed region.
`--impl reference` : the CPU restatement of the reference (oracle/, Warp + trimesh are not installable) on
           all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C = 2.998e8
WORKLOAD = dict(stl="models/room.stl", tx=[10.0, 0.0, 5.0], bounces=8, rate=1e12, window=200e-9, radius=0.1,
                rx=[[-14.0 + 28.0 * k / 15.0, 6.0, 5.0] for k in range(16)], rays_per_gpu=1 << 28)
BYTES_PER_SEGMENT = 32 * 6 + 48  # 32*ceil(log2 F) + 48, F = 44 -> 240 B


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (pynvml; nvidia-smi semantics)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_reference(args):
    """Reference arm: the CPU restatement of kernel.py + tracer.py on all host threads.  Each step re-traces a
    bounded sample of the workload once PER RECEIVER, as the reference does (coverage.py:38-43 calls compute_cir
    per receiver), and post-processes it with the NumPy restatement of tracer.py:84-117."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu, geometry, post
    w = WORKLOAD
    soup = geometry.load_stl_soup(os.path.join(ROOT, w["stl"]))
    bvh = cpu.Bvh(soup)
    rx_soups = [geometry.rx_soup(c, w["radius"]) for c in w["rx"]]
    sample = args.ref_rays
    threads = cpu.max_threads()

    def step(begin):
        seg, _, _ = cpu.trace_env(soup, w["tx"], w["bounces"], begin, sample, instrument=False, bvh=bvh)
        for rxs in rx_soups:  # the reference re-traces every ray for every receiver
            tids, rows = cpu.trace_received(soup, rxs, w["tx"], w["bounces"], begin, sample, bvh=bvh)
            paths = post.clean_paths(rows, np.ones(len(rows), dtype=np.uint32))
            post.impulse_response(paths, 1.0, w["rays_per_gpu"], C, w["rate"], w["window"])
        return seg

    for i in range(args.warmup):
        step(i * sample)
    t0 = time.perf_counter()
    segs = 0
    for i in range(args.steps):
        segs += step((args.warmup + i) * sample)
    dt = time.perf_counter() - t0
    value = segs / dt
    desc = (f"{sample} rays/step x {len(rx_soups)} receivers, literal per-receiver re-trace (kernel.py:38-98) + "
            f"NumPy post (tracer.py:84-117); segments counted once per ray; {threads} OpenMP threads; the per-pass rate is "
            f"{value * (len(rx_soups) + 1):.3e} segment-queries/s")
    line = {"impl": "reference", "metric": "traced ray-segments/s", "value": value, "unit": "segments/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "sample_rays_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "segments/s", "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name():
    w = WORKLOAD
    return (f"C4 room.stl dense impulse response: {w['rays_per_gpu']} rays/GPU x {w['bounces']} bounces, "
            f"{len(w['rx'])} receivers r={w['radius']}, L={int(w['rate'] * w['window'])} bins, tx={w['tx']}")


def cpu_baseline(sample_rays):
    from oracle import cpu, geometry, post
    w = WORKLOAD
    soup = geometry.load_stl_soup(os.path.join(ROOT, w["stl"]))
    bvh = cpu.Bvh(soup)
    threads = cpu.max_threads()
    t0 = time.perf_counter()
    seg, _, _ = cpu.trace_env(soup, w["tx"], w["bounces"], 0, sample_rays, instrument=False, bvh=bvh)
    t_env = time.perf_counter() - t0
    t0 = time.perf_counter()
    for c in w["rx"]:
        tids, rows = cpu.trace_received(soup, geometry.rx_soup(c, w["radius"]), w["tx"], w["bounces"], 0, sample_rays,
                                        bvh=bvh)
        post.impulse_response(post.clean_paths(rows, np.ones(len(rows), dtype=np.uint32)), 1.0, w["rays_per_gpu"], C,
                              w["rate"], w["window"])
    t_ref = time.perf_counter() - t0
    return {"value": seg / t_ref, "unit": "segments/s", "cores": threads, "kind": "port",
            "sample": (f"ray ids [0,{sample_rays}) of the same workload: {len(w['rx'])} literal per-receiver passes of "
                       f"kernel.py:38-98 + tracer.py:84-117 in {t_ref:.2f} s on {threads} threads (segments counted once "
                       f"per ray); environment-only single pass: {seg / t_env:.3e} segments/s")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=WORKLOAD["rays_per_gpu"], help="rays per GPU per step")
    ap.add_argument("--ref-rays", type=int, default=1 << 19, help="rays per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=1 << 21, help="rays of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    from rf_ray_tracing_warp_b200._lib import load as load_lib

    load_lib()  # fail loudly if the CUDA extension is missing
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = dict(WORKLOAD)
    w["rays_per_gpu"] = args.rays
    R = w["rays_per_gpu"]
    n_total = R * world
    ray_range = (rank * R, (rank + 1) * R)
    mesh = load_mesh(os.path.join(ROOT, w["stl"]))
    tracer = Tracer(mesh, C, w["rate"], w["window"], w["bounces"], n_total, device=dev, ray_range=ray_range,
                    shard=world > 1)
    L = int(w["rate"] * w["window"])
    n_rx = len(w["rx"])
    job = tracer.make_job(np.asarray(w["rx"]), w["radius"], want_paths=False)
    dirs = torch.empty((R, 4), dtype=torch.float32, device=dev)
    ir = torch.zeros((n_rx, L), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    trace_ms, gen_ms = [], []

    def device_step(timed):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        tracer.ray_directions(ray_range[0], ray_range[1], out=dirs)
        e1.record()
        # environment trace (the dominant kernel) is timed on its own stream position for the roofline
        job.enqueue(w["tx"], 1.0, ray_range=ray_range, dirs=dirs, ir=ir)
        e2.record()
        if timed:
            gen_ms.append((e0, e1))
        return e0, e2

    # the trace kernel alone (for roofline.achieved): bracket ONLY rfrt_trace with events
    def trace_only_ms():
        from rf_ray_tracing_warp_b200._lib import check, float3
        from rf_ray_tracing_warp_b200 import _lib as L_
        a, b = ev(), ev()
        job.counters_t.zero_()
        a.record()
        check(tracer._lib.rfrt_trace(tracer._env, job.rxset, float3(w["tx"]), w["bounces"], ray_range[0], ray_range[1], 1,
                                     dirs.data_ptr(), R, job.counters_t.data_ptr(), job.cands.data_ptr(),
                                     job.cand_capacity, None, None, torch.cuda.current_stream().cuda_stream), "rfrt_trace")
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b), int(job.counters_t[L_.CTR_SEGMENTS].item())

    for _ in range(args.warmup):
        device_step(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    step_events, seg_total = [], 0
    launches0 = job.kernel_launches
    for _ in range(args.steps):
        flush.fill_(1)  # evict L2 between timed steps (not timed)
        e0, e2 = device_step(True)
        step_events.append((e0, e2))
    torch.cuda.synchronize(dev)
    counters = job.counters()
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in step_events]
    total_ms = float(sum(step_ms))
    launches = (job.kernel_launches - launches0) + args.steps  # + rfrt_ray_directions per step
    seg_step = counters["segments"]
    assert counters["candidates"] <= job.cand_capacity and counters["records"] <= job.rec_capacity

    # dominant kernel alone, same inputs (roofline)
    t_ms = [trace_only_ms() for _ in range(3)]
    k_ms = float(np.mean([t for t, _ in t_ms[1:]]))
    assert all(s == seg_step for _, s in t_ms)
    g_ms = float(np.mean([a.elapsed_time(b) for a, b in gen_ms])) if gen_ms else 0.0

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    segs = torch.tensor([seg_step * args.steps], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(segs, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    total_segs = int(segs.item())
    value = total_segs / (total_ms * 1e-3)

    # ---- e2e: public API, host inputs, host outputs ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        tx_host = np.asarray(w["tx"], dtype=np.float64)
        rx_host = np.asarray(w["rx"], dtype=np.float64)

        ir_pinned = torch.empty((n_rx, L), dtype=torch.float64, pin_memory=True)

        def api_step():
            out = tracer.compute_cir_multi(tx_host, 1.0, rx_host, w["radius"], return_paths=False, dense=True)
            ir_pinned.copy_(out["impulse_response"], non_blocking=True)  # device -> host read of the step's result
            torch.cuda.synchronize(dev)
            return out["stats"]["segments"], ir_pinned.numpy()

        api_step()
        barrier()
        t0 = time.perf_counter()
        e_segs = 0
        e_steps = max(1, min(args.steps, 3))
        for _ in range(e_steps):
            s, ir_host = api_step()
            e_segs += s  # already summed over ranks by the record exchange
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": e_segs / float(dt.item()), "unit": "segments/s",
               "h2d_bytes_per_step": int(tx_host.nbytes // 2 + rx_host.nbytes + 42 * 3 * 8 + 80 * 3 * 4),
               "d2h_bytes_per_step": int(ir_host.nbytes + 8 * 8), "steps": e_steps,
               "api": "Tracer.compute_cir_multi (receiver-set build + trace + replay + ordered binning + D2H of the impulse responses)"}

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = seg_step * BYTES_PER_SEGMENT / (k_ms * 1e-3) / 1e9
        traffic, issue_pct = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
                ncu = json.load(f).get("k_trace_small", {})
                traffic, issue_pct = ncu.get("dram_bytes_per_launch"), ncu.get("issue_active_pct")
            if traffic is not None:
                traffic = int(traffic * R / (1 << 28))  # captured at 2^28 rays; it is the direction buffer, linear in rays
        except Exception:
            pass
        line = {"metric": "traced ray-segments/s", "value": value, "unit": "segments/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name().replace(str(WORKLOAD["rays_per_gpu"]), str(R)),
                           "segments_per_step_per_gpu": seg_step, "received_records_per_step": counters["records"],
                           "l2": "256 MiB flush between timed steps; the 4 GiB direction buffer streamed by every step exceeds the 126 MB L2",
                           "parallelism": f"ray-range sharding x{world}, BVH replicated, sparse record all-gather"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "k_trace_small", "kernel_ms": k_ms, "peak_source": peak_src,
                             "algorithmic_bytes_per_segment": BYTES_PER_SEGMENT,
                             "note": "the 44-triangle scene is resident in shared memory: HBM traffic is only the direction "
                                     "buffer, the binding limit is instruction issue (profiles/README.md), so frac can exceed 1",
                             "traffic_note": "ncu dram__bytes_read+write per launch at 268435456 rays (profiles/ncu_summary.json)",
                             "issue_active_pct": issue_pct},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches * 1,
                "kernel_ms": {"k_gen_dirs": g_ms, "k_trace_small": k_ms, "step": total_ms / args.steps}}
        if args.cpu_sample > 0:
            line["cpu_baseline"] = cpu_baseline(args.cpu_sample)
        print(json.dumps(line), flush=True)
    job.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
