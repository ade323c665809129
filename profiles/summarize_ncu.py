#!/usr/bin/env python
"""Summarises an .ncu-rep (raw + source pages) into the few numbers the roofline discussion needs.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [--lines 25]
"""
import csv
import io
import json
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
    summary = {}
    for k in keys:
        if k in m:
            summary[k] = " ".join(x for x in m[k] if x)
            print(f"{k:88s} {summary[k]}")
    src = page(rep, "source")
    h = src[1]
    body = src[2:]
    iS, iI, iAvg = h.index("Source"), h.index("Instructions Executed"), h.index("Avg. Threads Executed")
    tot = sum(float(r[iI] or 0) for r in body)
    w = sum(float(r[iI] or 0) * float(r[iAvg] or 0) for r in body)
    print(f"\nSASS instructions: {len(body)}; warp-level executed {tot:.4g}; mean active threads {w / tot:.2f}")
    top = sorted(body, key=lambda r: -float(r[iI] or 0))[:nlines]
    for r in top:
        print(f"  {100 * float(r[iI]) / tot:5.2f}%  active={float(r[iAvg] or 0):4.1f}  {r[iS][:90]}")
    summary["mean_active_threads"] = w / tot
    print("\nJSON:", json.dumps(summary))


if __name__ == "__main__":
    main()
