"""Headless equivalent of the reference's main.py (single-link channel impulse response).

    python -m rf_ray_tracing_warp_b200.main [--model models/room.stl] [--tx 10 0 5] [--rx -10 0 5] ...

Follows main.py:15-19 (constants), :29-37 (scene, Tracer, compute_cir), :39,46-55 (RX power) of the reference.
The blocking matplotlib plot (main.py:41-43) and the HTTP viewer (main.py:67) are replaced by result files.
"""
import argparse
import json
import os

import numpy as np

LIGHT_SPEED_MPS = 2.998e8   # main.py:15
SAMPLE_RATE_HZ = 100e9      # main.py:16
SAMPLE_WINDOW_S = 200.0e-9  # main.py:17
MAX_BOUNCES = 4             # main.py:18
TX_NUM_RAYS = 5_000_000     # main.py:19


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default=os.path.join(os.path.dirname(__file__), "..", "models", "room.stl"))
    ap.add_argument("--tx", type=float, nargs=3, default=[10, 0, 5])     # main.py:30
    ap.add_argument("--rx", type=float, nargs=3, default=[-10, 0, 5])    # main.py:31
    ap.add_argument("--tx-power", type=float, default=1)                 # main.py:33
    ap.add_argument("--rx-radius", type=float, default=0.1)              # main.py:34
    ap.add_argument("--rays", type=int, default=TX_NUM_RAYS)
    ap.add_argument("--bounces", type=int, default=MAX_BOUNCES)
    ap.add_argument("--out", default=None, help="directory for impulse_response.npy / paths.npz / result.json / scene.glb")
    ap.add_argument("--mode", default="reference", choices=["reference", "physical"],
                    help="reference = main.py:36-55; physical = Tracer.trace_physical (free-space loss, Fresnel amplitude, "
                         "carrier phase, analytic-sphere receiver; complex impulse response)")
    ap.add_argument("--carrier", type=float, default=2.4e9)              # main.py:46
    ap.add_argument("--scene", default="glb", choices=["glb", "html", "none"],
                    help="scene export written next to the results (replaces viz/visualization.py)")
    args = ap.parse_args(argv)

    from . import Tracer, load_mesh, to_dbm
    mesh = load_mesh(args.model)
    tracer = Tracer(mesh, LIGHT_SPEED_MPS, SAMPLE_RATE_HZ, SAMPLE_WINDOW_S, args.bounces, args.rays, verbose=True)
    if args.mode == "physical":
        out = tracer.trace_physical(np.array(args.tx), args.tx_power, [args.rx], args.rx_radius, carrier_hz=args.carrier,
                                    want_ir=True)
        power, dbm = float(out["power"][0]), float(out["dbm"][0])
        print(f"Signal RX power: {dbm} dBm (physical mode, {out['stats']['arrivals']} arrivals)")
        result = dict(model=args.model, tx=args.tx, rx=args.rx, rays=args.rays, bounces=args.bounces, mode="physical",
                      rx_power=power, rx_power_dbm=dbm, **out["stats"])
        if args.out:
            os.makedirs(args.out, exist_ok=True)
            np.save(os.path.join(args.out, "impulse_response_complex.npy"), out["impulse_response"].cpu().numpy()[0])
            with open(os.path.join(args.out, "result.json"), "w") as f:
                json.dump(result, f)
        return result
    paths, impulse_response = tracer.compute_cir(np.array(args.tx), args.tx_power, np.array(args.rx), args.rx_radius)
    out = tracer.compute_cir_multi(np.array(args.tx), args.tx_power, [args.rx], args.rx_radius, dense=False)
    power = float(tracer.rx_power(out["records"], 1).cpu().numpy()[0])   # main.py:46-55
    dbm = float(to_dbm(power))
    print(f"Signal RX power: {dbm} dBm")                                 # main.py:55
    result = dict(model=args.model, tx=args.tx, rx=args.rx, rays=args.rays, bounces=args.bounces,
                  received_paths=len(paths), rx_power=power, rx_power_dbm=dbm, **tracer.last_stats)
    if args.out:
        os.makedirs(args.out, exist_ok=True)
        np.save(os.path.join(args.out, "impulse_response.npy"), impulse_response)
        np.savez(os.path.join(args.out, "paths.npz"), *paths)
        with open(os.path.join(args.out, "result.json"), "w") as f:
            json.dump(result, f)
        if args.scene != "none":  # main.py:67 visualize(mesh, tx_pos, rx_pos, paths)
            from .scene_export import export_scene
            export_scene(os.path.join(args.out, "scene." + args.scene), mesh, args.tx, args.rx, paths)
    return result


if __name__ == "__main__":
    main()
