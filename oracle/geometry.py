"""Oracle-side geometry helpers (independent of the product's loaders).

Restates what the reference gets from trimesh (absent here, version unpinned, README.md:8):
  * tm.load_mesh(<binary stl>)  (main.py:21,25,29): facet corners in file order, fp32 exact.
  * tm.primitives.Sphere(center, radius, subdivisions=1)  (tracer.py:27): icosahedron -> one midpoint
    subdivision -> project to the sphere -> *radius + center in fp64; cast to fp32 at tracer.py:28.
"""
import struct

import numpy as np


def load_stl_soup(path):
    """(F,3,3) float32 triangle soup in file order."""
    raw = open(path, "rb").read()
    (count,) = struct.unpack_from("<I", raw, 80)
    if len(raw) != 84 + 50 * count:
        raise ValueError("oracle STL reader handles binary STL only")
    out = np.empty((count, 3, 3), dtype=np.float32)
    for i in range(count):
        vals = struct.unpack_from("<12f", raw, 84 + 50 * i)
        out[i] = np.asarray(vals[3:], dtype=np.float32).reshape(3, 3)
    return out


_T = (1.0 + 5.0 ** 0.5) / 2.0
_ICO_V = [[-1, _T, 0], [1, _T, 0], [-1, -_T, 0], [1, -_T, 0], [0, -1, _T], [0, 1, _T], [0, -1, -_T], [0, 1, -_T],
          [_T, 0, -1], [_T, 0, 1], [-_T, 0, -1], [-_T, 0, 1]]
_ICO_F = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
          [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
          [8, 6, 7], [9, 8, 1]]


def unit_icosphere_subdiv1():
    """42 unit vertices (float64) and 80 faces (int), trimesh-style subdivision order."""
    verts = [np.asarray(v, dtype=np.float64) for v in _ICO_V]
    verts = [v / np.sqrt(np.dot(v, v)) for v in verts]
    cache = {}

    def midpoint(i, j):
        key = (min(i, j), max(i, j))
        if key not in cache:
            cache[key] = len(verts)
            verts.append((verts[i] + verts[j]) * 0.5)
        return cache[key]

    faces = []
    # trimesh.remesh.subdivide ordering: all edge midpoints first, then per face
    # [a, m_ab, m_ca], [m_ab, b, m_bc], [m_ca, m_bc, c], [m_ab, m_bc, m_ca]
    for a, b, c in _ICO_F:
        midpoint(a, b); midpoint(b, c); midpoint(c, a)
    for a, b, c in _ICO_F:
        mab, mbc, mca = midpoint(a, b), midpoint(b, c), midpoint(c, a)
        faces += [[a, mab, mca], [mab, b, mbc], [mca, mbc, c], [mab, mbc, mca]]
    v = np.asarray(verts, dtype=np.float64)
    v = v / np.sqrt((v * v).sum(axis=1))[:, None]
    return v, np.asarray(faces, dtype=np.int64)


def rx_soup(center, radius):
    """(80,3,3) float32 triangle soup of the receiver mesh (tracer.py:26-30)."""
    u, f = unit_icosphere_subdiv1()
    v = (np.asarray(center, dtype=np.float64)[None, :] + float(radius) * u).astype(np.float32)
    return np.ascontiguousarray(v[f])
