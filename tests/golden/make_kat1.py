#!/usr/bin/env python
"""Regenerates tests/golden/kat1.json from the reference's only golden artefact.

    python tests/golden/make_kat1.py            (needs /root/reference; run in the build container)

Source: /root/reference/web/scene.html — a trimesh `scene_to_html` page saved by an earlier run of
main.py + viz/visualization.py.  It embeds a base64 GLB holding the almost_empty mesh, two marker
spheres (TX at (20,0,4.5), RX at (-20,0,4.8)) and 119 GL_LINES polylines = the received ray paths.
This script decodes the polylines and stores them verbatim (fp32 values as JSON doubles).  It then
uses the ORACLE's direction generator to find, for every polyline, the ray id < 80 000 000 whose
direction passes closest to the recorded receiver entry point, and the set of ray ids whose
direction hits the analytic r=0.1 receiver sphere; both are stored so the tests can check the
oracle (and the GPU) against the file without /root/reference being present.
"""
import base64
import hashlib
import json
import os
import re
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
SRC = "/root/reference/web/scene.html"


def decode_glb(html):
    m = re.search(r'base64_data\s*=\s*"([A-Za-z0-9+/=]+)"', html)
    glb = base64.b64decode(m.group(1))
    magic, _ver, length = struct.unpack_from("<III", glb, 0)
    assert magic == 0x46546C67 and length == len(glb)
    off = 12
    clen, _ = struct.unpack_from("<II", glb, off)
    js = json.loads(glb[off + 8: off + 8 + clen])
    off += 8 + clen
    blen, _ = struct.unpack_from("<II", glb, off)
    return js, glb[off + 8: off + 8 + blen]


def accessor(js, binc, idx):
    acc = js["accessors"][idx]
    bv = js["bufferViews"][acc["bufferView"]]
    assert acc["componentType"] == 5126 and acc["type"] == "VEC3"
    start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
    return np.frombuffer(binc, dtype="<f4", count=3 * acc["count"], offset=start).reshape(-1, 3)


def main():
    from oracle import cpu

    raw = open(SRC, "rb").read()
    js, binc = decode_glb(raw.decode("utf-8", errors="replace"))
    tx = np.array([20.0, 0.0, 4.5])
    rx = np.array([-20.0, 0.0, 4.8])
    radius = 0.1
    n_rays = 80_000_000

    polylines = []
    for mesh in js["meshes"]:
        prim = mesh["primitives"][0]
        if prim.get("mode", 4) != 1:
            continue
        pts = accessor(js, binc, prim["attributes"]["POSITION"])
        # GL_LINES: consecutive vertex pairs; rebuild the polyline
        poly = [pts[0]]
        for k in range(0, len(pts), 2):
            assert np.array_equal(pts[k], poly[-1])
            poly.append(pts[k + 1])
        polylines.append(np.asarray(poly, dtype=np.float32))
    assert len(polylines) == 119

    hits = cpu.sphere_hits(0, n_rays, tx, rx, radius)
    dirs = cpu.ray_directions_list(hits).astype(np.float64)
    matched, resid = [], []
    for poly in polylines:
        assert np.array_equal(poly[0], tx.astype(np.float32))
        e = poly[1].astype(np.float64) - tx
        perp = e[None, :] - (dirs @ e)[:, None] * dirs / (dirs * dirs).sum(1)[:, None]
        r = np.sqrt((perp * perp).sum(1))
        j = int(np.argmin(r))
        matched.append(int(hits[j]))
        resid.append(float(r[j]))
    next_hit = cpu.sphere_hits(n_rays, 200_000, tx, rx, radius)
    out = {
        "source": "reference web/scene.html",
        "source_sha256": hashlib.sha256(raw).hexdigest(),
        "mesh": "almost_empty.stl",
        "tx_pos": tx.tolist(), "rx_pos": rx.tolist(), "rx_radius": radius, "n_rays": n_rays,
        "paths": [[[float(x) for x in v] for v in p] for p in polylines],
        "matched_tids": matched,
        "match_residual_m": resid,
        "analytic_sphere_hit_tids": [int(t) for t in hits],
        "next_hit_tid_after_n_rays": int(next_hit[0]) if len(next_hit) else None,
    }
    with open(os.path.join(HERE, "kat1.json"), "w") as f:
        json.dump(out, f)
    print("polylines", len(polylines), "analytic hits", len(hits), "max residual", max(resid),
          "median", float(np.median(resid)), "sets equal", sorted(matched) == sorted(int(t) for t in hits),
          "next", out["next_hit_tid_after_n_rays"])


if __name__ == "__main__":
    main()
