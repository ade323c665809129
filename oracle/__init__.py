"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A restatement of the reference hot path (kernel.py + tracer.py + main.py power) used to check the
CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; rf_ray_tracing_warp_b200 never does.

Parity status: pinned by KAT-1 (tests/golden/kat1.json, decoded from the reference's web/scene.html)
for RNG, direction sampling and the received-ray set; "parity unpinned" against real Warp for
closest-hit last bits / tie order (warp-lang is not installable here) — see rfrt_oracle.c header.
"""
