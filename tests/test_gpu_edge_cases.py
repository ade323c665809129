"""GPU edge cases: empty / tiny inputs, ragged sizes, overflow handling, random meshes (BVH == brute force)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
C = 2.998e8


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device — no CPU fallback exists")
    torch.cuda.set_device(0)
    return torch


def _mesh(tris):
    from rf_ray_tracing_warp_b200 import mesh_from_triangles
    return mesh_from_triangles(np.asarray(tris, dtype=np.float32))


def test_empty_mesh_and_zero_rays(torch_cuda, room_stl):
    from rf_ray_tracing_warp_b200 import Mesh, Tracer, load_mesh
    empty = Mesh(vertices=np.zeros((0, 3)), faces=np.zeros((0, 3), dtype=np.int64))
    tr = Tracer(empty, C, 100e9, 200e-9, 3, 1000)
    out = tr.trace_segments([0, 0, 0], dump=True)
    assert out["segments"] == 1000 and out["env_hits"] == 0          # every ray: one segment, then dead
    assert (out["hit_tri"].cpu().numpy() == -1).all()
    # a receiver in empty space still captures line-of-sight rays (kernel.py:85: `not maybe_hit_env`)
    paths, ir = tr.compute_cir([0, 0, 0], 1, [1.0, 0, 0], 0.5)
    assert len(paths) > 10 and all(2 <= len(p) <= 4 for p in paths)
    # zero rays / zero bounces
    tr0 = Tracer(load_mesh(room_stl), C, 100e9, 200e-9, 3, 0)
    paths, ir = tr0.compute_cir([10, 0, 5], 1, [-10, 0, 5], 0.1)
    assert paths == [] and ir.shape == (20000,) and not ir.any()
    trb = Tracer(load_mesh(room_stl), C, 100e9, 200e-9, 0, 1000)
    paths, ir = trb.compute_cir([10, 0, 5], 1, [-10, 0, 5], 0.1)
    assert paths == [] and not ir.any()


@pytest.mark.parametrize("ntri", [1, 2, 3, 5])
def test_tiny_meshes_match_oracle(torch_cuda, ntri):
    from oracle import cpu
    rng = np.random.default_rng(ntri)
    tris = rng.uniform(-2, 2, size=(ntri, 3, 3)).astype(np.float32)
    from rf_ray_tracing_warp_b200 import Tracer
    n, B = 50000, 4
    tr = Tracer(_mesh(tris), C, 100e9, 200e-9, B, n)
    out = tr.trace_segments([0.1, 0.2, 0.3], dump=True)
    seg, tri, t = cpu.trace_env(tris, [0.1, 0.2, 0.3], B, 0, n)
    assert out["segments"] == seg and (tri >= 0).sum() > 100
    assert np.array_equal(out["hit_tri"].cpu().numpy(), tri)
    assert np.array_equal(out["hit_t"].cpu().numpy().view(np.uint32), t.view(np.uint32))


def test_random_soups_bvh_equals_brute_force(torch_cuda):
    """Property test over random triangle soups incl. degenerate and duplicated triangles (ties -> lowest index)."""
    from oracle import cpu
    from rf_ray_tracing_warp_b200 import Tracer
    rng = np.random.default_rng(99)
    for trial in range(6):
        ntri = int(rng.integers(8, 400))
        tris = rng.uniform(-5, 5, size=(ntri, 3, 3)).astype(np.float32)
        tris[: ntri // 8] = tris[ntri // 8: 2 * (ntri // 8)]          # exact duplicates: equal-t ties
        tris[-1, 2] = tris[-1, 1]                                      # a degenerate (zero-area) triangle
        tris[-2] = np.round(tris[-2])                                  # axis-snapped vertices
        n, B = 20000, 5
        tr = Tracer(_mesh(tris), C, 100e9, 200e-9, B, n)
        out = tr.trace_segments([0.0, 0.0, 0.0], dump=True)
        seg, tri, t = cpu.trace_env(tris, [0.0, 0.0, 0.0], B, 0, n)
        assert out["segments"] == seg
        assert np.array_equal(out["hit_tri"].cpu().numpy(), tri), trial
        assert np.array_equal(out["hit_t"].cpu().numpy().view(np.uint32), t.view(np.uint32)), trial


def test_ragged_ray_ranges_and_chunks(torch_cuda, room_stl):
    """Results are independent of how ray ids are split into ranges / chunks (the multi-GPU invariant)."""
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    mesh = load_mesh(room_stl)
    n, B, tx, rx, r = 100003, 4, [10, 0, 5], [3.0, 6.0, 5.0], 0.7
    whole = Tracer(mesh, C, 100e9, 200e-9, B, n)
    paths, ir = whole.compute_cir(tx, 1, rx, r)
    seg = whole.last_stats["segments"]
    parts, seg_parts, ir_sum = [], 0, np.zeros_like(ir)
    for begin, end in [(0, 1), (1, 33333), (33333, 33334), (33334, 100003)]:
        t = Tracer(mesh, C, 100e9, 200e-9, B, n, ray_range=(begin, end), chunk_rays=7777)
        p, i = t.compute_cir(tx, 1, rx, r)
        parts += p
        seg_parts += t.last_stats["segments"]
        ir_sum += i
    assert seg_parts == seg and len(parts) == len(paths) > 50
    assert all(np.array_equal(a, b) for a, b in zip(parts, paths))
    np.testing.assert_allclose(ir_sum, ir, rtol=1e-12, atol=0)


def test_work_list_overflow_is_retried(torch_cuda, room_stl):
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    mesh = load_mesh(room_stl)
    n, B, tx, rx, r = 1 << 16, 3, [10, 0, 5], [3.0, 6.0, 5.0], 1.5
    big = Tracer(mesh, C, 100e9, 200e-9, B, n)
    small = Tracer(mesh, C, 100e9, 200e-9, B, n, max_candidates=16, max_records=16)
    p1, ir1 = big.compute_cir(tx, 1, rx, r)
    p2, ir2 = small.compute_cir(tx, 1, rx, r)
    assert len(p1) == len(p2) > 100 and small.max_records >= len(p2)
    assert all(np.array_equal(a, b) for a, b in zip(p1, p2)) and np.array_equal(ir1, ir2)


def test_bad_arguments_are_rejected(torch_cuda, room_stl):
    from rf_ray_tracing_warp_b200 import Mesh, RfrtError, Tracer, load_mesh
    bad = Mesh(vertices=np.zeros((3, 3)), faces=np.array([[0, 1, 7]]))
    with pytest.raises(RfrtError):
        Tracer(bad, C, 100e9, 200e-9, 3, 10)
    tr = Tracer(load_mesh(room_stl), C, 100e9, 200e-9, 40, 10)   # replay path supports <= 32 bounces
    with pytest.raises(RfrtError):
        tr.compute_cir([10, 0, 5], 1, [-10, 0, 5], 0.1)
    with pytest.raises(RfrtError):
        Tracer(load_mesh(room_stl), C, 100e9, 200e-9, 3, 10).compute_cir([10, 0, 5], 1, [-10, 0, 5], -1.0)


def test_far_receiver_uses_exact_fallback(torch_cuda, almost_empty_stl):
    """A receiver > 8192 radii away takes the all-faces branch of the receiver query.  Hits are ~1e-9 per ray, so
    the oracle first finds the ray ids that hit the receiver's bounding sphere and replays only those."""
    from oracle import cpu, geometry, post
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    n, B, tx, rx, r = 1 << 29, 2, [0, 0, 1.0], [0, 0, 101.0], 0.012
    tr = Tracer(load_mesh(almost_empty_stl), C, 100e9, 200e-9, B, n)
    out = tr.compute_cir_multi(tx, 1, [rx], r, return_paths=True, dense=False)
    rec = {k: v.cpu().numpy() for k, v in out["records"].items()}
    soup, rxs = geometry.load_stl_soup(almost_empty_stl), geometry.rx_soup(rx, r)
    expect = {}
    for tid in cpu.sphere_hits(0, n, tx, rx, r * 1.0001):
        o = cpu.trace_paths(soup, rxs, tx, B, int(tid), 1, instrument=False)
        if o["mask"][0]:
            expect[int(tid)] = post.clean_paths(o["received"], o["mask"])[0]
    assert sorted(expect) == sorted(int(t) & 0xFFFFFFFF for t in rec["ray"])
    for t, row, nv in zip(rec["ray"], rec["paths"], rec["nverts"]):
        assert np.array_equal(row[:nv], expect[int(t) & 0xFFFFFFFF])


@pytest.mark.parametrize("case", ["far_from_origin", "many_planes_fallback", "degenerates_and_duplicates", "box_rooms"])
def test_small_scene_sweep_equals_bvh_and_oracle(torch_cuda, case):
    """The shared-memory sweep (candidate filter + nearest-first exact tests + self-re-hit shortcut) on awkward small
    scenes: hit triangle / distance per (ray, bounce) == oracle, and the order-independent checksum == the BVH walk."""
    from oracle import cpu
    from rf_ray_tracing_warp_b200 import Tracer
    rng = np.random.default_rng(11)
    if case == "far_from_origin":      # large coordinates: the filter tolerance scales with |coordinate|
        tris = rng.uniform(-3, 3, size=(24, 3, 3)).astype(np.float32) + np.array([1000.0, -2000.0, 500.0], dtype=np.float32)
        tx = [1000.2, -2000.1, 500.3]
    elif case == "many_planes_fallback":  # 40 distinct planes need 80 filter slots -> the BVH path takes over
        tris = rng.uniform(-3, 3, size=(40, 3, 3)).astype(np.float32)
        tx = [0.1, 0.2, 0.3]
    elif case == "degenerates_and_duplicates":
        tris = rng.uniform(-3, 3, size=(20, 3, 3)).astype(np.float32)
        tris[3] = tris[2]                      # exact duplicate: equal-t tie -> lowest index
        tris[5, 2] = tris[5, 1]                # two equal vertices
        tris[7, 2] = 0.5 * (tris[7, 0] + tris[7, 1])  # collinear
        tris[9] = np.float32(1.25)             # a point
        tx = [0.0, 0.1, -0.2]
    else:                                    # nested axis-aligned boxes: coplanar pairs, T-junction-free rooms, rays stuck at t == 0
        def box(lo, hi):
            x0, y0, z0 = lo; x1, y1, z1 = hi
            v = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], np.float32)
            f = [[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 5, 1], [0, 4, 5], [1, 6, 2], [1, 5, 6], [2, 7, 3], [2, 6, 7], [3, 4, 0], [3, 7, 4]]
            return v[np.array(f)]
        tris = np.concatenate([box((-4, -3, 0), (4, 3, 3)), box((-1, -1, 0), (1, 1, 3)), box((2, -3, 0), (4, -1, 1))]).astype(np.float32)
        tx = [-2.5, 0.5, 1.5]
    n, B = 1 << 17, 7
    res = {}
    for force_bvh in (False, True):
        tr = Tracer(_mesh(tris), C, 100e9, 200e-9, B, n, force_bvh=force_bvh)
        res[force_bvh] = tr.trace_segments(tx, dump=True)
    seg, tri, t = cpu.trace_env(tris, tx, B, 0, n)
    for k in (False, True):
        assert res[k]["segments"] == seg and (tri >= 0).sum() > 1000
        assert np.array_equal(res[k]["hit_tri"].cpu().numpy(), tri), (case, k)
        assert np.array_equal(res[k]["hit_t"].cpu().numpy().view(np.uint32), t.view(np.uint32)), (case, k)
    assert res[False]["checksum"] == res[True]["checksum"]


def test_multi_scale_mesh_keeps_the_hierarchy_within_the_walk_stack(torch_cuda, monkeypatch):
    """Triangles clustered at many scales (each cluster 2x closer to a corner than the last) give a hierarchy over
    63-bit Morton codes one level per scale; past the walk's 64-entry stack the builder falls back to coarser cells.
    Here the limit is lowered (RFRT_BVH_MAX_DEPTH, a test aid) so that 22 scales trigger the fallback; the answers stay
    those of the brute-force oracle (any valid BVH returns them)."""
    from oracle import cpu
    from rf_ray_tracing_warp_b200 import Tracer
    rng = np.random.default_rng(3)
    tris = []
    for k in range(22):
        c = 10.0 * 2.0 ** -k * np.ones(3)
        for _ in range(6):
            tris.append(c + 0.2 * 2.0 ** -k * rng.normal(size=(3, 3)))
    soup = np.asarray(tris, dtype=np.float32).reshape(-1, 9)
    mesh = _mesh(soup.reshape(-1, 3, 3))
    deep = Tracer(mesh, C, 100e9, 200e-9, 1, 1).mesh_info()["max_depth"]
    monkeypatch.setenv("RFRT_BVH_MAX_DEPTH", "18")
    tr = Tracer(mesh, C, 100e9, 200e-9, 1, 1)
    monkeypatch.delenv("RFRT_BVH_MAX_DEPTH")
    info = tr.mesh_info()
    assert deep > 20 and info["n_triangles"] == len(soup) and info["max_depth"] < deep, (deep, info)
    n = 4000
    o = rng.uniform(-1, 11, size=(n, 3)).astype(np.float32)
    target = soup[rng.integers(0, len(soup), n)].reshape(n, 3, 3).mean(axis=1)
    d = (target - o + rng.normal(scale=0.01, size=(n, 3))).astype(np.float32)
    t, f = tr.query_closest(o, d)
    t, f = t.cpu().numpy(), f.cpu().numpy()
    hits = 0
    for i in range(n):
        hit, t_o, f_o = cpu.query(soup, o[i], d[i])
        assert (f[i] >= 0) == hit
        if hit:
            hits += 1
            assert f[i] == f_o and np.float32(t[i]) == np.float32(t_o)
    assert hits > 300


def test_receiver_enumeration_paths_agree(torch_cuda, room_stl, monkeypatch):
    """64 heavily overlapping receivers on a line: a segment through them overlaps far more receiver boxes than the
    per-lane enumeration queues at once (12), so its drain-and-resume path runs; the cooperative enumeration of the
    same set must produce the same records, and both the oracle's for sampled receivers."""
    from oracle import cpu, geometry, post
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    n, B, tx, r = 1 << 16, 4, [10, 0, 5], 0.5
    rxs = np.array([[2.0 + 0.05 * k, 6.0, 5.0] for k in range(64)])
    results = []
    for coop in ("0", "1"):
        monkeypatch.setenv("RFRT_RX_COOP", coop)
        tr = Tracer(load_mesh(room_stl), C, 100e9, 200e-9, B, n)
        out = tr.compute_cir_multi(tx, 1, rxs, r, return_paths=True)
        rec = {k: v.cpu().numpy() for k, v in out["records"].items()}
        results.append(rec)
    monkeypatch.delenv("RFRT_RX_COOP")
    a, b = results
    assert a["ray"].shape[0] > 2000
    for name in a:
        assert np.array_equal(a[name].view(np.uint8), b[name].view(np.uint8)), name
    soup = geometry.load_stl_soup(room_stl)
    for k in (0, 31, 63):
        o = cpu.trace_paths(soup, geometry.rx_soup(rxs[k], r), tx, B, 0, n, instrument=False)
        o_paths = post.clean_paths(o["received"], o["mask"])
        sel = a["rx"] == k
        assert np.array_equal(a["ray"][sel].astype(np.uint32), np.nonzero(o["mask"])[0].astype(np.uint32))
        for row, nv, op in zip(a["paths"][sel], a["nverts"][sel], o_paths):
            assert np.array_equal(row[:nv].view(np.uint32), op.view(np.uint32))


def test_plain_and_general_slab_tests_agree(torch_cuda, room_stl, monkeypatch):
    """The walks use the plain slab test when the transmitter stands within 8 x the mesh's largest coordinate and the
    general one (per-axis origin-dependent offsets) otherwise.  Same hits either way and both equal to brute force:
    a fine terrain (triangles of 0.4 m in boxes padded by 2e-4), near transmitter with the general test forced, a
    transmitter ten times the scene's size away (general test picked by the library), and the replay's records."""
    from oracle import cpu
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh, synthetic_terrain
    mesh = synthetic_terrain(96, 20.0, 5)
    soup = mesh.vertices[mesh.faces].astype(np.float32)
    n, B = 1 << 16, 5
    for tx, n in (([10.0, 0.0, 4.5], 1 << 16), ([200.0, -30.0, 60.0], 1 << 18)):   # 8 x 20 m = 160 m is the switch
        outs = []
        for far in (None, "1"):
            if far:
                monkeypatch.setenv("RFRT_SLAB_FAR", far)
            tr = Tracer(mesh, C, 100e9, 200e-9, B, n)
            out = tr.trace_segments(tx, dump=True)
            outs.append((out["segments"], out["hit_tri"].cpu().numpy(), out["hit_t"].cpu().numpy().view(np.uint32)))
            if far:
                monkeypatch.delenv("RFRT_SLAB_FAR")
        seg, tri, t = cpu.trace_env(soup, tx, B, 0, n)
        assert (tri >= 0).sum() > 50
        for o in outs:
            assert o[0] == seg and np.array_equal(o[1], tri) and np.array_equal(o[2], t.view(np.uint32)), tx
    recs, n = [], 1 << 18
    for far in (None, "1"):
        if far:
            monkeypatch.setenv("RFRT_SLAB_FAR", far)
        tr = Tracer(load_mesh(room_stl), C, 100e9, 200e-9, 4, n, force_bvh=True)
        out = tr.compute_cir_multi([10, 0, 5], 1, np.array([[-10.0, 0, 5], [0.0, 3.0, 4.0]]), 1.0, return_paths=True)
        recs.append({k: v.cpu().numpy() for k, v in out["records"].items()})
        if far:
            monkeypatch.delenv("RFRT_SLAB_FAR")
    assert recs[0]["ray"].shape[0] > 50
    for name in recs[0]:
        assert np.array_equal(recs[0][name].view(np.uint8), recs[1][name].view(np.uint8)), name


def test_ray_order_by_cells_equals_radix_order(torch_cuda, monkeypatch):
    """Waves of a BVH scene are traced in direction-coherent order: a counting sort over the 2^24 direction cells for
    big waves, a radix sort for small ones.  Both are permutations of the wave's rays, so every ray's trajectory must be
    the same either way (and the same as with no ordering at all)."""
    from rf_ray_tracing_warp_b200 import Tracer, _lib, synthetic_terrain
    mesh = synthetic_terrain(96, 20.0, 5)
    n, B, tx = (1 << 18) + 777, 4, [10.0, 0.0, 4.5]
    outs = []
    for order in ("radix", "cells", None):
        if order:
            monkeypatch.setenv("RFRT_RAY_ORDER", order)
        tr = Tracer(mesh, C, 100e9, 200e-9, B, n)
        if order is None:
            tr.trace_flags |= _lib.FLAG_NO_RAY_SORT
        out = tr.trace_segments(tx, dump=True, checksum=True)
        outs.append((out["segments"], out["env_hits"], out["checksum"], out["hit_tri"].cpu().numpy(),
                     out["hit_t"].cpu().numpy().view(np.uint32)))
        if order:
            monkeypatch.delenv("RFRT_RAY_ORDER")
    assert outs[0][0] > n and outs[0][1] > 1000
    for o in outs[1:]:
        assert o[:3] == outs[0][:3]
        assert np.array_equal(o[3], outs[0][3]) and np.array_equal(o[4], outs[0][4])
