#!/usr/bin/env python
"""bench.py — traced ray-segments/s of the hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rays R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): C4 of BASELINE.md — models/room.stl (44 triangles), tx (10,0,5), 16 receivers
(r = 0.1) on the line y=6, z=5, x=-14..14, 8 bounces, 1 TS/s x 200 ns = 200 000 delay bins, 268 435 456 rays
PER GPU (weak scaling: rank g traces global ray ids [g*R, (g+1)*R)), synthetic = seeded ray directions.
One step = the whole hot path over that batch, ending where SURVEY.md 8(d) ends the metric: directions ->
environment trace (+ receiver tests) -> literal replay of the received candidates -> record exchange (pack + ONE
all-gather of the ranks' record segments, NCCL) -> library sort by (receiver, ray id) -> ordered impulse responses.

`value`  : segments (device counters, summed over ranks by the exchange itself) / device time of K steps (CUDA events
           around the whole step incl. the collective, max over ranks), inputs resident in HBM.
`e2e`    : same metric through the public API (Tracer.compute_cir_multi) with HOST inputs (tx / receiver positions as
           NumPy) and the impulse responses copied back to the host inside the timed region, over the same K steps.
`coverage`: the second half of BASELINE.json's metric — the C2 coverage map (almost_empty.stl, 256 x 256 receivers,
           16.8 M rays, 4 bounces) through Tracer.coverage, host in / host grid out, in ms, beside the CPU port's time
           for ONE receiver of that map (the reference re-traces per receiver: coverage.py:38-43).  Rank 0, N = 1 only.
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
COVERAGE = dict(stl="models/almost_empty.stl", tx=[1.0, 0.0, 1.0], bounces=4, rate=100e9, window=100e-9, radius=0.1,
                grid=256, extent=15.0, z=2.0, rays=1 << 24)
BYTES_PER_SEGMENT = 32 * 6 + 48  # 32*ceil(log2 F) + 48, F = 44 -> 240 B


def host_threads():
    """Threads the CPU arms may use: the affinity mask of this process — NOT omp_get_max_threads(), which torchrun
    pins to 1 through OMP_NUM_THREADS."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_ncu(kernel):
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return json.load(f).get(kernel, {})
    except Exception:
        return {}


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


def workload_name(rays=None):
    w = WORKLOAD
    return (f"C4 room.stl dense impulse response: {rays or w['rays_per_gpu']} rays/GPU x {w['bounces']} bounces, "
            f"{len(w['rx'])} receivers r={w['radius']}, L={int(w['rate'] * w['window'])} bins, tx={w['tx']}")


def reference_pass(cpu, geometry, post, soup, bvh, rx_soups, begin, sample, threads):
    """What the reference does with `sample` rays and R receivers: one full trace per receiver (coverage.py:38-43 calls
    compute_cir per receiver; kernel.py:38-98) + the NumPy post-processing of tracer.py:84-117.  Returns
    (segments counted once per ray, seconds of the environment-only pass, seconds of the R reference passes)."""
    w = WORKLOAD
    t0 = time.perf_counter()
    seg, _, _ = cpu.trace_env(soup, w["tx"], w["bounces"], begin, sample, instrument=False, bvh=bvh, nthreads=threads)
    t_env = time.perf_counter() - t0
    t0 = time.perf_counter()
    for rxs in rx_soups:  # the reference re-traces every ray for every receiver
        tids, rows = cpu.trace_received(soup, rxs, w["tx"], w["bounces"], begin, sample, bvh=bvh, nthreads=threads)
        paths = post.clean_paths(rows, np.ones(len(rows), dtype=np.uint32))
        post.impulse_response(paths, 1.0, w["rays_per_gpu"], C, w["rate"], w["window"])
    return seg, t_env, time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the CPU restatement of kernel.py + tracer.py on all host threads, each step a bounded sample of
    the workload.  Rank 0 alone runs it."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import cpu, geometry, post
    w = WORKLOAD
    soup = geometry.load_stl_soup(os.path.join(ROOT, w["stl"]))
    bvh = cpu.Bvh(soup)
    rx_soups = [geometry.rx_soup(c, w["radius"]) for c in w["rx"]]
    sample, threads = args.ref_rays, host_threads()
    for i in range(args.warmup):
        reference_pass(cpu, geometry, post, soup, bvh, rx_soups, i * sample, sample, threads)
    segs = dt = dt_env = 0.0
    for i in range(args.steps):
        s, t_env, t_ref = reference_pass(cpu, geometry, post, soup, bvh, rx_soups, (args.warmup + i) * sample, sample, threads)
        segs += s
        dt += t_ref
        dt_env += t_env
    value = segs / dt
    desc = (f"{sample} rays/step x {len(rx_soups)} receivers, literal per-receiver re-trace (kernel.py:38-98) + NumPy post "
            f"(tracer.py:84-117); segments counted once per ray; {threads} OpenMP threads; a single environment-only pass "
            f"runs at {segs / dt_env:.3e} segments/s")
    line = {"impl": "reference", "metric": "traced ray-segments/s", "value": value, "unit": "segments/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.rays), "sample_rays_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "segments/s", "cores": threads, "kind": "port", "sample": desc,
                             "single_pass_value": segs / dt_env},
            "e2e": {"value": value, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(sample_rays):
    from oracle import cpu, geometry, post
    w = WORKLOAD
    soup = geometry.load_stl_soup(os.path.join(ROOT, w["stl"]))
    bvh = cpu.Bvh(soup)
    threads = host_threads()
    rx_soups = [geometry.rx_soup(c, w["radius"]) for c in w["rx"]]
    seg, t_env, t_ref = reference_pass(cpu, geometry, post, soup, bvh, rx_soups, 0, sample_rays, threads)
    return {"value": seg / t_ref, "unit": "segments/s", "cores": threads, "kind": "port", "single_pass_value": seg / t_env,
            "sample": (f"ray ids [0,{sample_rays}) of the same workload: {len(w['rx'])} literal per-receiver passes of "
                       f"kernel.py:38-98 + tracer.py:84-117 in {t_ref:.2f} s on {threads} threads (segments counted once "
                       f"per ray); environment-only single pass: {seg / t_env:.3e} segments/s")}


def coverage_record(torch, dev, cpu_rays):
    """BASELINE.json's "coverage-map ms": config C2 through Tracer.coverage (host positions in, host power grid out),
    timed with the wall clock around the call (it synchronises), beside the CPU port's time for ONE receiver."""
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    from rf_ray_tracing_warp_b200.coverage import plane_lattice
    cv = COVERAGE
    lattice = plane_lattice(cv["grid"], cv["grid"], cv["extent"], cv["z"])
    tr = Tracer(load_mesh(os.path.join(ROOT, cv["stl"])), C, cv["rate"], cv["window"], cv["bounces"], cv["rays"], device=dev,
                max_candidates=1 << 25, max_records=1 << 25)
    ms = []
    for _ in range(4):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        cov = tr.coverage(np.asarray(cv["tx"]), 1.0, lattice, cv["radius"])
        torch.cuda.synchronize(dev)
        ms.append(1e3 * (time.perf_counter() - t0))
    rec = {"config": f"C2 almost_empty.stl, {cv['grid']}x{cv['grid']} receivers r={cv['radius']} z={cv['z']}, {cv['rays']} rays x "
                     f"{cv['bounces']} bounces, L={int(cv['rate'] * cv['window'])}",
           "map_ms": float(min(ms[1:])), "map_ms_all": [float(x) for x in ms], "api": "Tracer.coverage (host in, host grid out)",
           "segments": cov["stats"]["segments"], "received_records": cov["stats"]["records"],
           "cells_with_signal": int(np.isfinite(cov["dbm"]).sum())}
    if cpu_rays > 0:
        from oracle import cpu, geometry, post
        soup = geometry.load_stl_soup(os.path.join(ROOT, cv["stl"]))
        threads = host_threads()
        k = (cv["grid"] // 2) * cv["grid"] + cv["grid"] // 3
        t0 = time.perf_counter()
        tids, rows = cpu.trace_received(soup, geometry.rx_soup(lattice[k], cv["radius"]), cv["tx"], cv["bounces"], 0, cpu_rays,
                                        nthreads=threads)
        ir = post.impulse_response(post.clean_paths(rows, np.ones(len(rows), dtype=np.uint32)), 1.0, cv["rays"], C, cv["rate"],
                                   cv["window"])
        post.rx_power(ir, cv["window"])
        dt = time.perf_counter() - t0
        per_rx_ms = 1e3 * dt * cv["rays"] / cpu_rays
        rec["cpu_port"] = {"one_receiver_ms_measured": 1e3 * dt, "rays_measured": cpu_rays, "cores": threads,
                           "one_receiver_ms_at_full_rays": per_rx_ms,
                           "map_ms_extrapolated": per_rx_ms * cv["grid"] * cv["grid"],
                           "how": "coverage.py:38-43 runs the full trace + post once per receiver: measured for one receiver "
                                  "on a ray sample, scaled linearly to the map's rays and receivers (extrapolation, labelled)"}
    return rec


def invariance_check(torch, dist, mesh, dev, rank, world):
    """A 2^22-ray job sharded over the ranks against the same job on rank 0 alone: impulse responses bit-identical."""
    from rf_ray_tracing_warp_b200 import Tracer
    w = WORKLOAD
    n = 1 << 22
    sharded = Tracer(mesh, C, w["rate"], w["window"], w["bounces"], n, device=dev, shard=True)
    ir_s = sharded.compute_cir_multi(w["tx"], 1.0, np.asarray(w["rx"]), w["radius"], dense=True)["impulse_response"]
    ok = torch.ones(1, dtype=torch.int64, device=dev)
    if rank == 0:
        alone = Tracer(mesh, C, w["rate"], w["window"], w["bounces"], n, device=dev, shard=False)
        ir_1 = alone.compute_cir_multi(w["tx"], 1.0, np.asarray(w["rx"]), w["radius"], dense=True)["impulse_response"]
        same = torch.equal(ir_s.view(torch.int64), ir_1.view(torch.int64)) and bool((ir_1 != 0).sum() > 100) and \
            sharded.last_stats == alone.last_stats
        ok[0] = 1 if same else 0
    dist.broadcast(ok, 0)
    if not bool(ok.item()):
        raise SystemExit("bench.py: sharded impulse responses differ from the single-GPU ones")
    return "bit-identical"


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
    ap.add_argument("--no-coverage", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from rf_ray_tracing_warp_b200 import Tracer, _lib, load_mesh

    _lib.load()  # fail loudly if the CUDA extension is missing
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
    invariance = invariance_check(torch, dist, mesh, dev, rank, world) if world > 1 else None
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
    seg_cap = [tracer.exchange_records]
    marks = []

    def device_step(timed):
        """the whole hot path of one batch on the current stream, no host synchronisation inside"""
        e = [ev() for _ in range(4)]
        e[0].record()
        tracer.ray_directions(ray_range[0], ray_range[1], out=dirs)
        e[1].record()
        job.enqueue(w["tx"], 1.0, ray_range=ray_range, dirs=dirs)      # environment trace + literal replay
        e[2].record()
        _, summary = job.collect(seg_cap[0], ir=ir)                     # pack + all-gather + sort + ordered responses
        e[3].record()
        if timed:
            marks.append(e)
        return summary

    def read_summary(summary):
        c = summary.cpu().numpy()
        assert not c[_lib.SUM_COUNTERS + _lib.CTR_QUEUE_OVERFLOW]
        return c

    # size the exchange segment from a first pass (every rank sees the same summary, so all agree), then warm up
    c = read_summary(device_step(False))
    assert c[_lib.SUM_MAX_CANDIDATES] <= job.cand_capacity and c[_lib.SUM_MAX_RECORDS] <= job.rec_capacity
    seg_cap[0] = 1 << int(np.ceil(np.log2(max(1.25 * c[_lib.SUM_MAX_RECORDS], 4096))))
    for _ in range(max(args.warmup - 1, 0) + 1):
        device_step(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = job.kernel_launches
    summary = None
    for _ in range(args.steps):
        flush.fill_(1)  # evict L2 between timed steps (not timed)
        summary = device_step(True)
    torch.cuda.synchronize(dev)
    c = read_summary(summary)
    barrier()
    clocks = sampler.stop()
    assert c[_lib.SUM_OVERFLOWED] == 0 and c[_lib.SUM_MAX_CANDIDATES] <= job.cand_capacity
    step_ms = [e[0].elapsed_time(e[3]) for e in marks]
    total_ms = float(sum(step_ms))
    launches = (job.kernel_launches - launches0) + args.steps  # + rfrt_ray_directions per step
    seg_step_all = int(c[_lib.SUM_COUNTERS + _lib.CTR_SEGMENTS])   # all ranks, one step (summed by the exchange itself)
    records_all = int(c[_lib.SUM_RECORDS])
    g_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in marks]))
    tr_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in marks]))
    x_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in marks]))

    # the dominant kernel alone, same inputs (roofline): bracket ONLY rfrt_trace with events
    def trace_only_ms():
        from rf_ray_tracing_warp_b200._lib import check, float3
        a, b = ev(), ev()
        job.counters_t.zero_()
        a.record()
        check(tracer._lib.rfrt_trace(tracer._env, job.rxset, float3(w["tx"]), w["bounces"], ray_range[0], ray_range[1], 1,
                                     dirs.data_ptr(), R, job.counters_t.data_ptr(), job.cands.data_ptr(),
                                     job.cand_capacity, None, None, torch.cuda.current_stream().cuda_stream), "rfrt_trace")
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b), int(job.counters_t[_lib.CTR_SEGMENTS].item())

    t_ms = [trace_only_ms() for _ in range(3)]
    k_ms = float(np.mean([t for t, _ in t_ms[1:]]))
    seg_step = t_ms[0][1]  # this rank's segments per step

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = seg_step_all * args.steps / (total_ms * 1e-3)

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

        for _ in range(2):
            api_step()
        barrier()
        t0 = time.perf_counter()
        e_segs = 0
        for _ in range(args.steps):
            s, ir_host = api_step()
            e_segs += s  # already summed over ranks by the record exchange
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": e_segs / float(dt.item()), "unit": "segments/s", "ms_per_step": 1e3 * float(dt.item()) / args.steps,
               "h2d_bytes_per_step": int(tx_host.nbytes // 2 + rx_host.nbytes + 42 * 3 * 8 + 80 * 3 * 4),
               "d2h_bytes_per_step": int(ir_host.nbytes + 8 * _lib.SUM_COUNT), "steps": args.steps,
               "api": "Tracer.compute_cir_multi (receiver-set build + trace + replay + record exchange + ordered binning + "
                      "D2H of the impulse responses)"}

    if rank == 0:
        peak, peak_src = load_peaks()
        ncu = load_ncu("k_trace_small")
        scale = seg_step / ncu["segments_per_launch"] if ncu.get("segments_per_launch") else None
        sm_hz = 1e6 * (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        sms = 148
        try:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
        except Exception:
            pass
        # Instruction-issue roofline: every SM sub-partition issues at most one warp instruction per clock.
        issue_peak = sms * 4 * sm_hz / 1e9                                   # G warp-instructions / s at the sampled clock
        warp_inst = ncu.get("warp_instructions_per_launch")
        issue_achieved = warp_inst * scale / (k_ms * 1e-3) / 1e9 if warp_inst and scale else None
        dram = ncu.get("dram_bytes_per_launch")
        line = {"metric": "traced ray-segments/s", "value": value, "unit": "segments/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(R), "segments_per_step_per_gpu": seg_step,
                           "received_records_per_step": records_all,
                           "l2": "256 MiB flush between timed steps; the 4 GiB direction buffer streamed by every step exceeds the 126 MB L2",
                           "parallelism": f"ray-range sharding x{world}, BVH replicated, one all-gather of record segments "
                                          f"({seg_cap[0]} records/rank) inside the timed step"},
                "roofline": {"bound": "issue", "achieved": issue_achieved, "peak": issue_peak, "unit": "Gwarp-inst/s",
                             "frac": issue_achieved / issue_peak if issue_achieved else None,
                             "kernel": "k_trace_small", "kernel_ms": k_ms,
                             "how": "the 44-triangle scene and the ray states live in shared memory, so instruction issue bounds "
                                    "this kernel, not HBM: achieved = warp instructions per launch (smsp__inst_executed.sum of the "
                                    "committed ncu capture of this kernel on this scene and receiver set, profiles/ncu_summary.json, scaled by this run's segment "
                                    "count) / the kernel's CUDA-event time in THIS run; peak = SMs x 4 schedulers x the SM clock "
                                    "sampled in this run",
                             "active_lanes_per_instruction": ncu.get("active_lanes"), "issue_active_pct_ncu": ncu.get("issue_active_pct"),
                             "traffic": int(dram * scale) if dram and scale else None,
                             "hbm": {"achieved": (dram * scale / (k_ms * 1e-3) / 1e9) if dram and scale else None, "peak": peak,
                                     "unit": "GB/s", "frac": (dram * scale / (k_ms * 1e-3) / 1e9 / peak) if dram and scale else None,
                                     "peak_source": peak_src,
                                     "note": "measured DRAM bytes (ncu dram__bytes_read+write, the direction buffer) over this "
                                             "run's kernel time"},
                             "algorithmic": {"bytes_per_segment": BYTES_PER_SEGMENT,
                                             "gbs": seg_step * BYTES_PER_SEGMENT / (k_ms * 1e-3) / 1e9,
                                             "note": "SURVEY 8(d)'s figure for a BVH walk through memory; NOT a bound here (the "
                                                     "scene never leaves shared memory), kept for comparison with round 1"}},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "kernel_ms": {"k_gen_dirs": g_ms, "trace+replay": tr_ms, "exchange+sort+binning": x_ms, "k_trace_small_alone": k_ms,
                              "step": total_ms / args.steps}}
        if invariance:
            line["invariance"] = invariance
        if world == 1:
            if args.cpu_sample > 0:
                line["cpu_baseline"] = cpu_baseline(args.cpu_sample)
            if not args.no_coverage:
                job.close()
                del dirs, flush
                torch.cuda.empty_cache()
                line["coverage"] = coverage_record(torch, dev, (1 << 20) if args.cpu_sample > 0 else 0)
        print(json.dumps(line), flush=True)
    job.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
