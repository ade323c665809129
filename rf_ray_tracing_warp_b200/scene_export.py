"""Scene export: mesh + TX/RX markers + received ray paths as a binary glTF (.glb) and a small HTML page.

Headless replacement of the reference's viz/visualization.py:8-41 (a ``trimesh.Scene`` with the environment mesh,
two marker spheres of radius 0.25 and one polyline per received path, written to viz/scene.html) — without trimesh
and without the blocking HTTP server (viz/visualization.py:43-50).  The GLB has the same structure as the one embedded
in the reference's web/scene.html: mesh 0 = environment (TRIANGLES), meshes 1-2 = markers, then one GL_LINES
primitive per path with its vertices stored as segment pairs, so ``read_glb_paths`` reads both.
"""
import base64
import json
import struct

import numpy as np

from .mesh_io import unit_icosphere

_HTML = """<!DOCTYPE html><html><head><meta charset="utf-8"><title>rf ray tracing scene</title>
<script type="module" src="https://ajax.googleapis.com/ajax/libs/model-viewer/3.4.0/model-viewer.min.js"></script></head>
<body style="margin:0"><model-viewer style="width:100vw;height:100vh" camera-controls src="data:model/gltf-binary;base64,%s">
</model-viewer><script>base64_data = "%s";</script></body></html>
"""


class _Builder:
    def __init__(self):
        self.bin = bytearray()
        self.views, self.accessors, self.meshes, self.nodes = [], [], [], []

    def _accessor(self, arr, target, ctype, atype):
        arr = np.ascontiguousarray(arr)
        while len(self.bin) % 4:
            self.bin.append(0)
        self.views.append(dict(buffer=0, byteOffset=len(self.bin), byteLength=arr.nbytes, target=target))
        self.bin += arr.tobytes()
        acc = dict(bufferView=len(self.views) - 1, componentType=ctype, count=int(arr.shape[0]), type=atype)
        if atype == "VEC3":
            acc["min"], acc["max"] = arr.min(0).astype(float).tolist(), arr.max(0).astype(float).tolist()
        self.accessors.append(acc)
        return len(self.accessors) - 1

    def add(self, name, vertices, indices=None, mode=4, color=(0.8, 0.8, 0.8, 1.0)):
        prim = dict(attributes=dict(POSITION=self._accessor(np.asarray(vertices, dtype="<f4"), 34962, 5126, "VEC3")),
                    mode=mode, material=len(self.meshes))
        if indices is not None:
            prim["indices"] = self._accessor(np.asarray(indices, dtype="<u4").reshape(-1), 34963, 5125, "SCALAR")
        self.meshes.append((dict(name=name, primitives=[prim]), color))
        self.nodes.append(dict(name=name, mesh=len(self.meshes) - 1))

    def glb(self):
        js = dict(asset=dict(version="2.0", generator="rf_ray_tracing_warp_b200"), scene=0,
                  scenes=[dict(nodes=list(range(len(self.nodes))))], nodes=self.nodes,
                  meshes=[m for m, _ in self.meshes],
                  materials=[dict(pbrMetallicRoughness=dict(baseColorFactor=list(c), metallicFactor=0.0),
                                  doubleSided=True) for _, c in self.meshes],
                  accessors=self.accessors, bufferViews=self.views, buffers=[dict(byteLength=len(self.bin))])
        jb = json.dumps(js, separators=(",", ":")).encode()
        jb += b" " * (-len(jb) % 4)
        bb = bytes(self.bin) + b"\0" * (-len(self.bin) % 4)
        total = 12 + 8 + len(jb) + 8 + len(bb)
        return struct.pack("<III", 0x46546C67, 2, total) + struct.pack("<II", len(jb), 0x4E4F534A) + jb + \
            struct.pack("<II", len(bb), 0x004E4942) + bb


def export_scene(path, mesh, tx_pos=None, rx_pos=None, paths=(), marker_radius=0.25):
    """Writes ``path`` (.glb, or .html with the GLB embedded).  ``mesh`` has .vertices/.faces (tracer.py:22-23's duck
    type); ``paths`` is the list of (k,3) arrays returned by ``Tracer.compute_cir``."""
    b = _Builder()
    b.add("environment", np.asarray(mesh.vertices, dtype=np.float32), np.asarray(mesh.faces))
    sv, sf = unit_icosphere(2)
    for name, p, col in (("tx", tx_pos, (1.0, 0.2, 0.2, 1.0)), ("rx", rx_pos, (0.2, 0.4, 1.0, 1.0))):
        if p is not None:  # viz/visualization.py:12-24 (spheres of radius 0.25)
            b.add(name, (np.asarray(p, dtype=np.float64) + marker_radius * sv).astype(np.float32), sf, color=col)
    for i, p in enumerate(paths):  # viz/visualization.py:27-31: one polyline per path, stored as GL_LINES pairs
        p = np.asarray(p, dtype=np.float32)
        if p.shape[0] >= 2:
            b.add(f"path_{i}", np.stack([p[:-1], p[1:]], axis=1).reshape(-1, 3), mode=1, color=(0.1, 0.8, 0.1, 1.0))
    glb = b.glb()
    if str(path).lower().endswith(".html"):
        enc = base64.b64encode(glb).decode()
        with open(path, "w") as f:
            f.write(_HTML % (enc, enc))
    else:
        with open(path, "wb") as f:
            f.write(glb)
    return len(glb)


def read_glb_paths(path_or_bytes):
    """Polylines (list of (k,3) float32 arrays) of a GLB written by ``export_scene`` — or of the GLB embedded in a
    trimesh ``scene_to_html`` page such as the reference's web/scene.html (``base64_data = "..."``)."""
    import re
    raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    if raw[:4] != b"glTF":
        raw = base64.b64decode(re.search(rb'base64_data\s*=\s*"([A-Za-z0-9+/=]+)"', raw).group(1))
    clen, _ = struct.unpack_from("<II", raw, 12)
    js = json.loads(raw[20:20 + clen])
    binc = raw[20 + clen + 8:]
    out = []
    for m in js["meshes"]:
        for prim in m["primitives"]:
            if prim.get("mode", 4) != 1:
                continue
            acc = js["accessors"][prim["attributes"]["POSITION"]]
            bv = js["bufferViews"][acc["bufferView"]]
            start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
            seg = np.frombuffer(binc, dtype="<f4", count=3 * acc["count"], offset=start).reshape(-1, 2, 3)
            out.append(np.concatenate([seg[:1, 0], seg[:, 1]]).astype(np.float32))
    return out
