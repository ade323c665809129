"""Mesh inputs without trimesh: STL loader, receiver icosphere, synthetic terrain.

The reference gets these from trimesh (absent here; /root/reference README.md:8):
  * ``tm.load_mesh(path)``                      main.py:21,25,29      -> :func:`load_mesh`
  * ``tm.primitives.Sphere(..., subdivisions=1)`` tracer.py:27        -> :func:`unit_icosphere`
``apollo_17_landing_site.stl`` is missing from the reference (.MISSING_LARGE_BLOBS:1), so the terrain
configs use :func:`synthetic_terrain`, a seeded height field that is fully specified here.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Mesh:
    """Duck-type of the two trimesh attributes the reference reads (tracer.py:22-23)."""

    vertices: np.ndarray  # (V, 3) float64
    faces: np.ndarray  # (F, 3) int64

    @property
    def triangles(self):
        return self.vertices[self.faces]


def _parse_ascii_stl(text):
    verts = []
    for line in text.splitlines():
        parts = line.split()
        if len(parts) == 4 and parts[0] == "vertex":
            verts.append([np.float32(parts[1]), np.float32(parts[2]), np.float32(parts[3])])
    if len(verts) % 3:
        raise ValueError("ASCII STL: vertex count is not a multiple of 3")
    return np.asarray(verts, dtype=np.float32).reshape(-1, 3, 3)


def load_stl_triangles(path):
    """(F, 3, 3) float32 facet corners in file order (binary or ASCII STL)."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) >= 84:
        count = int(np.frombuffer(raw, dtype="<u4", count=1, offset=80)[0])
        if len(raw) == 84 + 50 * count:  # the binary-STL size identity (trimesh uses the same test)
            rec = np.dtype([("n", "<f4", (3,)), ("v", "<f4", (3, 3)), ("attr", "<u2")])
            return np.ascontiguousarray(np.frombuffer(raw, dtype=rec, count=count, offset=84)["v"])
    return _parse_ascii_stl(raw.decode("ascii", errors="replace"))


def load_stl_attributes(path):
    """(F,) uint16 "attribute byte count" word of every facet of a binary STL (zeros for ASCII).  room.stl carries
    20083 on every facet; CAD exporters use the word for colour / material ids."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) >= 84:
        count = int(np.frombuffer(raw, dtype="<u4", count=1, offset=80)[0])
        if len(raw) == 84 + 50 * count:
            rec = np.dtype([("n", "<f4", (3,)), ("v", "<f4", (3, 3)), ("attr", "<u2")])
            return np.ascontiguousarray(np.frombuffer(raw, dtype=rec, count=count, offset=84)["attr"])
    return np.zeros(load_stl_triangles(path).shape[0], dtype=np.uint16)


def materials_from_attributes(attributes, table, default=5.0):
    """Per-triangle refractive index from the STL attribute words: ``table`` maps word -> index; anything else gets
    ``default`` (the reference's hard-coded 5.0, tracer.py:43)."""
    out = np.full(len(attributes), float(default), dtype=np.float32)
    for word, n in table.items():
        out[np.asarray(attributes) == int(word)] = float(n)
    return out


def load_mesh(path):
    """Equivalent of ``trimesh.load_mesh`` for STL: bit-identical corners are merged into shared vertices;
    face order and winding are the file's, no face is dropped.  Triangle i == facet i of the file."""
    tris = load_stl_triangles(path)
    flat = tris.reshape(-1, 3)
    uniq, inverse = np.unique(flat.view(np.uint32).reshape(-1, 3), axis=0, return_inverse=True)
    vertices = uniq.view(np.float32).reshape(-1, 3).astype(np.float64)
    faces = np.asarray(inverse, dtype=np.int64).reshape(-1, 3)
    return Mesh(vertices=vertices, faces=faces)


def mesh_from_triangles(tris):
    tris = np.asarray(tris, dtype=np.float32).reshape(-1, 3, 3)
    return Mesh(vertices=tris.reshape(-1, 3).astype(np.float64),
                faces=np.arange(tris.shape[0] * 3, dtype=np.int64).reshape(-1, 3))


# --- receiver icosphere (tracer.py:27: Sphere(subdivisions=1) -> 42 vertices / 80 faces) --------------

def unit_icosphere(subdivisions=1):
    """Unit-sphere vertices (float64) and faces (int32) of an icosahedron subdivided `subdivisions` times,
    every vertex projected onto the sphere.  Receiver k's mesh is float32(center_k + radius * vertices)."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(int(subdivisions)):
        edges = np.sort(np.stack([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=1).reshape(-1, 2), axis=1)
        uniq, inverse = np.unique(edges, axis=0, return_inverse=True)
        # keep first-appearance order of the edges so the vertex numbering is stable
        first = np.full(len(uniq), len(edges), dtype=np.int64)
        np.minimum.at(first, inverse.reshape(-1), np.arange(len(edges)))
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
        mid_index = (rank[inverse.reshape(-1)] + len(v)).reshape(-1, 3)  # columns: ab, bc, ca
        ordered = uniq[np.argsort(first, kind="stable")]
        mids = 0.5 * (v[ordered[:, 0]] + v[ordered[:, 1]])
        v = np.vstack([v, mids])
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        mab, mbc, mca = mid_index[:, 0], mid_index[:, 1], mid_index[:, 2]
        f = np.stack([np.stack([a, mab, mca], 1), np.stack([mab, b, mbc], 1), np.stack([mca, mbc, c], 1),
                      np.stack([mab, mbc, mca], 1)], axis=1).reshape(-1, 3)
        v = v / np.sqrt((v * v).sum(axis=1))[:, None]
    return np.ascontiguousarray(v), np.ascontiguousarray(f.astype(np.int32))


# --- synthetic terrain (stands in for the missing apollo STL) ----------------------------------------

def _pcg(v):
    v = v.astype(np.uint32)
    b = v * np.uint32(747796405) + np.uint32(2891336453)
    c = ((b >> ((b >> np.uint32(28)) + np.uint32(4))) ^ b) * np.uint32(277803737)
    return (c >> np.uint32(22)) ^ c


def _lattice_value(ix, iy, seed):
    with np.errstate(over="ignore"):
        h = _pcg(ix.astype(np.uint32) * np.uint32(73856093) ^ _pcg(iy.astype(np.uint32) * np.uint32(19349663) ^ np.uint32(seed)))
    return (h >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0) * 2.0 - 1.0


def terrain_height(x, y, seed=17, octaves=5, base_freq=0.125):
    """fBm value noise on a hashed integer lattice; |height| <= 1."""
    z = np.zeros_like(x, dtype=np.float64)
    amp, freq, norm = 1.0, base_freq, 0.0
    for o in range(octaves):
        fx, fy = x * freq, y * freq
        ix, iy = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
        tx, ty = fx - ix, fy - iy
        sx, sy = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty)
        s = seed + 1013 * o
        v00 = _lattice_value(ix, iy, s); v10 = _lattice_value(ix + 1, iy, s)
        v01 = _lattice_value(ix, iy + 1, s); v11 = _lattice_value(ix + 1, iy + 1, s)
        z += amp * ((v00 * (1 - sx) + v10 * sx) * (1 - sy) + (v01 * (1 - sx) + v11 * sx) * sy)
        norm += amp
        amp *= 0.5
        freq *= 2.0
    return z / norm


def synthetic_terrain(n=1024, extent=20.0, seed=17, height=1.5):
    """n x n quads (2 n^2 triangles) over [-extent, extent]^2, z = height * fBm(x, y); row-major quads,
    two triangles per quad ((i,j),(i+1,j),(i+1,j+1)) and ((i,j),(i+1,j+1),(i,j+1))."""
    g = np.linspace(-extent, extent, n + 1)
    X, Y = np.meshgrid(g, g, indexing="ij")
    Z = height * terrain_height(X, Y, seed=seed)
    vertices = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(np.float32).astype(np.float64)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    v00 = (i * (n + 1) + j).reshape(-1); v10 = ((i + 1) * (n + 1) + j).reshape(-1)
    v11 = ((i + 1) * (n + 1) + j + 1).reshape(-1); v01 = (i * (n + 1) + j + 1).reshape(-1)
    faces = np.stack([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], axis=1).reshape(-1, 3)
    return Mesh(vertices=vertices, faces=faces.astype(np.int64))
