#!/usr/bin/env python
"""Monte-Carlo model of k_trace_small's trip schedule — CPU only, no GPU needed.

A warp of 32 lanes, K rays per lane.  A ray needs a full sweep when it is fresh or after a failed self-re-hit test;
after an environment hit it stands on a triangle and is eligible for the (3x cheaper) re-hit trip.  Per trip the warp
runs ONE kind and each lane contributes at most one of its rays.  The model reproduces the measured lane counts of the
one-ray-per-lane kernel (re-hit trips at ~18 of 32 lanes) and predicted the gain of several rays per lane before the
kernel was written (measured afterwards on room.stl: 31.0e9 -> 35.0e9 segments/s with K = 5).

    python scripts/trip_schedule_sim.py [n_rays]

Model constants (room.stl, 8 bounces): P(segment misses) = 0.07, P(re-hit test fails) = 0.22, warp instructions per
re-hit trip / full sweep = 350 / 1150 (profiles/ncu_trace_r01l_summary.txt).
"""
import random
import sys


def simulate(slots, policy, n_rays=60000, c_rehit=350, c_sweep=1150, p_miss=0.07, p_fail=0.22, bounces=8, seed=1):
    rnd = random.Random(seed)
    lanes = [[None] * slots for _ in range(32)]  # per slot: None or [kind the ray waits for, bounce]
    remaining = n_rays
    trips = {"R": 0, "S": 0}
    lanes_in = {"R": 0, "S": 0}
    while True:
        for lane in lanes:
            for k in range(slots):
                if lane[k] is None and remaining > 0:
                    lane[k] = ["S", 0]
                    remaining -= 1
        n_r = sum(1 for lane in lanes if any(s and s[0] == "R" for s in lane))
        n_s = sum(1 for lane in lanes if any(s and s[0] == "S" for s in lane))
        if n_r == 0 and n_s == 0:
            break
        kind = policy(n_r, n_s)
        if kind == "R" and n_r == 0:
            kind = "S"
        if kind == "S" and n_s == 0:
            kind = "R"
        trips[kind] += 1
        lanes_in[kind] += n_r if kind == "R" else n_s
        for lane in lanes:
            for k in range(slots):
                s = lane[k]
                if s and s[0] == kind:
                    if kind == "R" and rnd.random() < p_fail:
                        s[0] = "S"  # not a self re-hit: goes through a full sweep
                    elif rnd.random() < p_miss:
                        lane[k] = None  # the segment misses the environment
                    else:
                        s[1] += 1
                        if s[1] >= bounces:
                            lane[k] = None
                        else:
                            s[0] = "R"
                    break
    cost = (trips["R"] * c_rehit + trips["S"] * c_sweep) / n_rays
    return cost, lanes_in["R"] / max(trips["R"], 1), lanes_in["S"] / max(trips["S"], 1)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    print("slots policy                          warp-instr/ray  lanes per re-hit trip  lanes per sweep")
    rows = [(1, "re-hit trip when >= 8 lanes eligible", lambda r, s: "R" if r >= 8 else "S")]
    for k in (2, 3, 4, 5):
        rows.append((k, "the kind more lanes can join", lambda r, s: "R" if r >= s else "S"))
        rows.append((k, "re-hit trip when >= 24 lanes eligible", lambda r, s: "R" if r >= 24 else "S"))
    for k, name, pol in rows:
        c, lr, ls = simulate(k, pol, n)
        print(f"{k:5d} {name:38s} {c:12.1f} {lr:20.1f} {ls:18.1f}")


if __name__ == "__main__":
    main()
