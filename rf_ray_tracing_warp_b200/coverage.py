"""Headless equivalent of the reference's coverage.py (coverage map).

    python -m rf_ray_tracing_warp_b200.coverage [--model models/room.stl] [--tx 10 0 5] [--grid reference|NxM] ...

coverage.py:38-43 re-runs the full trace once per receiver position; here ONE trace tests every segment
against all receivers (identical per-receiver results, see tests), then coverage.py:45-55 (power per receiver)
runs as one kernel over the sparse arrivals.  The viridis colouring / 3-D viewer (coverage.py:31-36,57-59) are
replaced by result files (power grid as .npy).
"""
import argparse
import json
import os

import numpy as np

LIGHT_SPEED_MPS = 2.998e8   # coverage.py:12
SAMPLE_RATE_HZ = 100e9      # coverage.py:13
SAMPLE_WINDOW_S = 100.0e-9  # coverage.py:14
MAX_BOUNCES = 2             # coverage.py:15
TX_NUM_RAYS = 1_000_000     # coverage.py:16


def reference_lattice():
    """coverage.py:38-40: x, y in range(-15, 16, 2), z in range(0, 16, 2) -> 16*16*8 = 2048 receivers."""
    return np.array([[x, y, z] for x in range(-15, 16, 2) for y in range(-15, 16, 2) for z in range(0, 16, 2)],
                    dtype=np.float64)


def plane_lattice(nx, ny, extent=15.0, z=2.0):
    """Cell-centred nx x ny lattice over [-extent, extent]^2 at height z."""
    xs = -extent + 2 * extent * (np.arange(nx) + 0.5) / nx
    ys = -extent + 2 * extent * (np.arange(ny) + 0.5) / ny
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    return np.stack([X, Y, np.full_like(X, z)], axis=-1).reshape(-1, 3)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default=os.path.join(os.path.dirname(__file__), "..", "models", "room.stl"))
    ap.add_argument("--tx", type=float, nargs=3, default=[10, 0, 5])     # coverage.py:19 (commented-out default)
    ap.add_argument("--grid", default="reference", help="'reference' (coverage.py:38-40) or NxM plane lattice")
    ap.add_argument("--z", type=float, default=2.0)
    ap.add_argument("--tx-power", type=float, default=1)                 # coverage.py:25
    ap.add_argument("--rx-radius", type=float, default=0.1)              # coverage.py:26
    ap.add_argument("--rays", type=int, default=TX_NUM_RAYS)
    ap.add_argument("--bounces", type=int, default=MAX_BOUNCES)
    ap.add_argument("--out", default=None)
    ap.add_argument("--mode", default="reference", choices=["reference", "physical"],
                    help="reference = coverage.py:43-55 per receiver; physical = Tracer.trace_physical (free-space "
                         "loss, Fresnel amplitude, carrier phase; analytic-sphere receivers)")
    ap.add_argument("--carrier", type=float, default=2.4e9)              # coverage.py:46
    args = ap.parse_args(argv)

    from . import Tracer, load_mesh
    if args.grid == "reference":
        rx, shape = reference_lattice(), (16, 16, 8)
    else:
        nx, ny = (int(v) for v in args.grid.lower().split("x"))
        rx, shape = plane_lattice(nx, ny, z=args.z), (nx, ny)
    tracer = Tracer(load_mesh(args.model), LIGHT_SPEED_MPS, SAMPLE_RATE_HZ, SAMPLE_WINDOW_S, args.bounces, args.rays,
                    max_candidates=1 << 22, max_records=1 << 22)
    if args.mode == "physical":
        cov = tracer.trace_physical(np.array(args.tx), args.tx_power, rx, args.rx_radius, carrier_hz=args.carrier)
        cov["dbm"] = np.where(cov["power"] > 0, cov["dbm"], np.nan)
    else:
        cov = tracer.coverage(np.array(args.tx), args.tx_power, rx, args.rx_radius, carrier_hz=args.carrier)
    dbm = cov["dbm"].reshape(shape)
    covered = int(np.isfinite(dbm).sum())
    print(f"coverage: {rx.shape[0]} receivers, {covered} with signal, "
          f"max {np.nanmax(dbm) if covered else float('nan'):.2f} dBm; stats {cov['stats']}")
    if args.out:
        os.makedirs(args.out, exist_ok=True)
        np.save(os.path.join(args.out, "coverage_dbm.npy"), dbm)
        np.save(os.path.join(args.out, "coverage_power.npy"), cov["power"].reshape(shape))
        np.save(os.path.join(args.out, "receivers.npy"), rx)
        with open(os.path.join(args.out, "result.json"), "w") as f:
            json.dump(dict(model=args.model, tx=args.tx, receivers=int(rx.shape[0]), covered=covered, **cov["stats"]), f)
    return cov


if __name__ == "__main__":
    main()
