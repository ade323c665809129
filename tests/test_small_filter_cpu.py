"""CPU test of the small-scene candidate filter (csrc/rfrt_small.cu + closest_hit_small, phase 1).

The filter must keep a SUPERSET of what the exact watertight test accepts: for every segment of real ray
trajectories (oracle, kernel.py:57-98) the triangle the oracle hits has to be among the candidates.  The tables
come from the product's own host code through the C ABI (rfrt_small_scene_tables, no device work); the per-ray
filter arithmetic of sweep_pairs is restated here in NumPy float32."""
import ctypes

import numpy as np
import pytest

from oracle import cpu, geometry


def _tables(soup):
    from rf_ray_tracing_warp_b200 import _lib
    lib = _lib.load()
    soup = np.ascontiguousarray(soup, dtype=np.float32).reshape(-1, 9)
    recs = np.zeros((32, 28), dtype=np.float32)
    slot_tri = np.zeros(64, dtype=np.int32)
    n_pairs, extent = ctypes.c_int32(), ctypes.c_float()
    nbr = np.zeros((max(soup.shape[0], 1), 4), dtype=np.uint32)
    cls = np.zeros(5, dtype=np.int32)
    rc = lib.rfrt_small_scene_tables(soup.ctypes.data, soup.shape[0], recs.ctypes.data, slot_tri.ctypes.data,
                                     ctypes.byref(n_pairs), ctypes.byref(extent), nbr.ctypes.data, cls.ctypes.data)
    _tables.nbr = nbr[:soup.shape[0]]
    _tables.class_begin = cls
    return rc, recs[:n_pairs.value], slot_tri[:2 * n_pairs.value], extent.value


def _candidates(recs, slot_tri, extent, pos, dirs):
    """sweep_pairs for many rays at once: bool (n_rays, n_slots)."""
    f32 = np.float32
    pos, dirs = pos.astype(f32), dirs.astype(f32)
    dl = ((f32(extent) + np.abs(pos).sum(axis=1, dtype=f32)) * f32(1.0 / 65536.0)).astype(f32)
    dl_h = (dl * (np.sqrt((dirs * dirs).sum(axis=1, dtype=f32)) * f32(1.001))).astype(f32)
    keep = np.zeros((pos.shape[0], 2 * recs.shape[0]), dtype=bool)
    with np.errstate(all="ignore"):
        for k, R in enumerate(recs):
            n, d = R[:3], R[3]
            nd = (dirs @ n).astype(f32)
            npd = ((pos @ n) - d).astype(f32)
            r = (f32(1.0) / nd).astype(f32)
            t = (-npd * r).astype(f32)
            ar = np.abs(r)
            thr = np.where(t < -(dl * ar), f32(np.inf), -(dl_h * ar)).astype(f32)
            h = (pos + t[:, None] * dirs).astype(f32)
            for j in range(2):
                E = R[4 + 12 * j:16 + 12 * j].reshape(3, 4)
                dist = (h @ E[:, :3].T + E[:, 3]).astype(f32)
                m = np.fmin.reduce(dist, axis=1)          # fmin ignores NaN like FMNMX3
                x = m - thr
                keep[:, 2 * k + j] = ~(np.signbit(x) & ~np.isnan(x))
    return keep


def _segments(soup, tx, bounces, n_rays, tid0=0):
    """(pos, dir, hit triangle) of every alive segment, replaying reflect in float32 as kernel.py:6-8 does."""
    out = cpu.trace_paths(soup, None, tx, bounces, tid0, n_rays)
    traced, hit_tri = out["traced"], out["hit_tri"]
    d = cpu.ray_directions(tid0, n_rays)
    normals = np.stack([cpu.tri_normal(t) for t in np.asarray(soup, dtype=np.float32).reshape(-1, 9)])
    alive = np.ones(n_rays, dtype=bool)
    P, D, F = [], [], []
    f32 = np.float32
    for b in range(bounces):
        P.append(traced[alive, b]); D.append(d[alive]); F.append(hit_tri[alive, b])
        hit = alive & (hit_tri[:, b] >= 0)
        n = normals[np.where(hit, hit_tri[:, b], 0)]
        dot = ((d[:, 0] * n[:, 0]).astype(f32) + (d[:, 1] * n[:, 1]).astype(f32)).astype(f32)
        dot = (dot + (d[:, 2] * n[:, 2]).astype(f32)).astype(f32)
        s = (f32(2.0) * dot).astype(f32)
        refl = (d - (s[:, None] * n).astype(f32)).astype(f32)
        d = np.where(hit[:, None], refl, d)
        alive = hit
    return np.concatenate(P), np.concatenate(D), np.concatenate(F)


@pytest.mark.parametrize("name,tx", [("room", [10.0, 0.0, 5.0]), ("almost_empty", [1.0, 0.0, 1.0]),
                                     ("almost_empty", [0.02, 0.01, 0.0])])
def test_filter_keeps_every_oracle_hit(name, tx, repo_root):
    soup = geometry.load_stl_soup(f"{repo_root}/models/{name}.stl")
    rc, recs, slot_tri, extent = _tables(soup)
    assert rc == 0 and sorted(set(slot_tri.tolist())) == list(range(soup.reshape(-1, 9).shape[0]))
    cls = _tables.class_begin  # both shipped scenes are all axis-aligned walls: no general-class pair
    assert cls[0] == 0 and cls[1] == 0 and cls[4] == recs.shape[0] and np.all(np.diff(cls) >= 0)
    for c, axis in ((1, 0), (2, 1), (3, 2)):
        n = recs[cls[c]:cls[c + 1], :3]
        assert np.all(n[:, axis] == 1.0) and np.all(np.delete(n, axis, axis=1) == 0.0)
    pos, dirs, face = _segments(soup, tx, 8, 200_000)
    keep = _candidates(recs, slot_tri, extent, pos, dirs)
    hit = face >= 0
    # the slot(s) of the oracle's triangle must be kept
    is_face = slot_tri[None, :] == face[:, None]
    assert np.all((keep & is_face).any(axis=1)[hit]), "filter dropped a triangle the exact test hits"
    # and the filter must actually filter (else phase 2 degenerates to the full exact sweep)
    assert keep.sum(axis=1).mean() < 6.0


def test_filter_on_random_soups():
    rng = np.random.default_rng(5)
    for trial in range(6):
        n = int(rng.integers(1, 33))
        soup = rng.uniform(-3, 3, size=(n, 9)).astype(np.float32)
        if trial == 0:
            soup[0, 3:6] = soup[0, 0:3]                       # degenerate: two equal vertices
        if trial == 1:
            soup[0, 6:9] = 0.5 * (soup[0, 0:3] + soup[0, 3:6])  # degenerate: collinear
        rc, recs, slot_tri, extent = _tables(soup)
        assert rc == 0
        pos, dirs, face = _segments(soup, [0.1, -0.2, 0.3], 4, 20_000, tid0=1000 * trial)
        keep = _candidates(recs, slot_tri, extent, pos, dirs)
        hit = face >= 0
        assert hit.sum() > 1000
        assert np.all((keep & (slot_tri[None, :] == face[:, None])).any(axis=1)[hit])


def test_tables_refuse_scenes_that_do_not_fit():
    rng = np.random.default_rng(6)
    soup = rng.uniform(-3, 3, size=(40, 9)).astype(np.float32)  # 40 distinct planes -> 40 pairs > 32
    rc, _, _, _ = _tables(soup)
    assert rc == -1
    rc, recs, slot_tri, _ = _tables(np.zeros((0, 9), dtype=np.float32))
    assert rc == 0 and recs.shape[0] == 0


def test_self_rehit_neighbour_masks(repo_root):
    """The self-re-hit shortcut (small_self_rehit): when the ray's own triangle f is hit again within tau, the
    overall closest hit must be f or a triangle of a neighbour pair of f (rfrt_small_scene_tables' h_nbr)."""
    soup = geometry.load_stl_soup(f"{repo_root}/models/room.stl").reshape(-1, 9)
    rc, recs, slot_tri, extent = _tables(soup)
    nbr = _tables.nbr[:, 0] | _tables.nbr[:, 1]  # interior | boundary neighbours
    assert np.all((_tables.nbr[:, 2] | _tables.nbr[:, 3]) & ~nbr == 0)
    n, B, tx = 20_000, 8, [10.0, 0.0, 5.0]
    out = cpu.trace_paths(soup, None, tx, B, 0, n)
    pos, dirs, face = _segments(soup, tx, B, n)
    # previous face of every alive segment (same ordering as _segments: bounce-major over alive rays)
    hit_tri = out["hit_tri"]
    alive = np.ones(n, dtype=bool)
    prev = []
    last = np.full(n, -1)
    for b in range(B):
        prev.append(last[alive])
        last = np.where(alive, hit_tri[:, b], last)
        alive = alive & (hit_tri[:, b] >= 0)
    prev = np.concatenate(prev)
    tau = 2.5e-5 * extent
    checked = 0
    for i in np.nonzero(prev >= 0)[0][:60_000]:
        f = int(prev[i])
        hit, t_f, _ = cpu.query(soup[f:f + 1], pos[i], dirs[i])
        if not hit or t_f * float(np.linalg.norm(dirs[i])) * 1.001 > tau:
            continue
        g = int(face[i])
        assert g >= 0
        if g != f:
            pairs = {k for k in range(len(slot_tri) // 2) if g in slot_tri[2 * k:2 * k + 2]}
            assert any(nbr[f] >> k & 1 for k in pairs), (i, f, g)
        checked += 1
    assert checked > 20_000
