"""GPU parity on the BASELINE configs as stated (C1 exact, C2-shaped), the receiver mesh, far receivers, the record
pipeline (library sort + ordered arrival sums, privatised / atomic binning) and the two headless drivers.

Bars: ray ids, path vertices, bins, counts -> bit-exact; impulse responses 1e-5 relative (libm vs libdevice in the
Fresnel chain); received power per coverage cell 1e-4 relative (BASELINE.json north_star).
"""
import json
import os

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


def test_config_c1_exact(torch_cuda, room_stl):
    """BASELINE config C1 = the reference's own CPU-runnable case (main.py:15-19,29-37 with room.stl): 1 M rays,
    3 bounces, tx (10,0,5), rx (-10,0,5), r 0.1, 100 GS/s x 200 ns = 20 000 bins."""
    from oracle import cpu, geometry, post
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    n, B, tx, rx, r = 1_000_000, 3, [10, 0, 5], [-10, 0, 5], 0.1
    tr = Tracer(load_mesh(room_stl), C, 100e9, 200e-9, B, n)
    paths, ir = tr.compute_cir(tx, 1, rx, r)
    soup = geometry.load_stl_soup(room_stl)
    seg, _, _ = cpu.trace_env(soup, tx, B, 0, n, instrument=False)
    tids, rows = cpu.trace_received(soup, geometry.rx_soup(rx, r), tx, B, 0, n)
    o_paths = post.clean_paths(rows, np.ones(len(rows), dtype=np.uint32))
    o_ir = post.impulse_response(o_paths, 1, n, C, 100e9, 200e-9)
    assert tr.last_stats["segments"] == seg and 2_500_000 < seg < 2_700_000   # ~2 590 9xx
    assert len(paths) == len(o_paths) >= 1
    for a, b in zip(paths, o_paths):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert ir.shape == (20000,) and np.array_equal(ir != 0, o_ir != 0)
    np.testing.assert_allclose(ir, o_ir, rtol=1e-5, atol=0)
    # and the RX power observable of main.py:46-55
    out = tr.compute_cir_multi(tx, 1, [rx], r, dense=False)
    assert np.array_equal(out["records"]["ray"].cpu().numpy().astype(np.int64), tids)
    p = float(tr.rx_power(out["records"], 1).cpu().numpy()[0])
    np.testing.assert_allclose(p, post.rx_power(o_ir, 200e-9), rtol=1e-4)


def test_config_c2_lattice_cells(torch_cuda, almost_empty_stl):
    """C2-shaped coverage map (coverage.py:38-57 on almost_empty.stl, 256 x 256 lattice at z = 2, r = 0.1, 4 bounces,
    10 000 bins) at 2^20 rays: ONE trace for the 65 536 receivers; 64 sampled cells against the reference procedure
    (a full per-receiver trace + np.convolve power) within 1e-4 relative per cell."""
    from oracle import cpu, geometry, post
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    from rf_ray_tracing_warp_b200.coverage import plane_lattice
    n, B, tx, r = 1 << 20, 4, [1.0, 0.0, 1.0], 0.1
    lattice = plane_lattice(256, 256, 15.0, 2.0)
    tr = Tracer(load_mesh(almost_empty_stl), C, 100e9, 100e-9, B, n, max_candidates=1 << 23, max_records=1 << 23)
    cov = tr.coverage(tx, 1, lattice, r)
    assert cov["power"].shape == (65536,) and cov["stats"]["records"] > 1_000_000
    rng = np.random.default_rng(2)
    # cells around the transmitter's foot point and the cube (most arrivals) + a uniform sample of the grid
    d = np.linalg.norm(lattice - np.asarray(tx), axis=1)
    cells = np.concatenate([np.argsort(d)[:8], rng.choice(65536, 56, replace=False)])
    soup = geometry.load_stl_soup(almost_empty_stl)
    checked = 0
    for k in cells:
        tids, rows = cpu.trace_received(soup, geometry.rx_soup(lattice[k], r), tx, B, 0, n)
        o_ir = post.impulse_response(post.clean_paths(rows, np.ones(len(rows), dtype=np.uint32)), 1, n, C, 100e9, 100e-9)
        p = post.rx_power(o_ir, 100e-9)
        if np.isnan(p):
            assert np.isnan(cov["power"][k]), k
        else:
            np.testing.assert_allclose(cov["power"][k], p, rtol=1e-4, err_msg=f"cell {k}")
            np.testing.assert_allclose(cov["dbm"][k], post.to_dbm(p), rtol=1e-4, atol=1e-3)
            checked += 1
    assert checked >= 48
    # the sparse path (sorted records + ordered arrival sums) gives the same map
    cov2 = tr.coverage(tx, 1, lattice[cells], r, dense_budget_bytes=0)
    np.testing.assert_allclose(cov2["power"], cov["power"][cells], rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("center,radius", [([0.0, 0.0, 0.0], 1.0), ([-10.0, 0.0, 5.0], 0.1), ([1234.5678, -987.654, 321.0], 0.1),
                                           ([1.0e5, 2.0e5, -3.0e5], 2.5), ([0.3, -0.7, 0.2], 1.0e-3)])
def test_receiver_mesh_matches_rx_soup(torch_cuda, room_stl, center, radius):
    """tracer.py:26-30 (_generate_rx_mesh): the 80 triangles of the generated receiver == the independent restatement
    (set equality, bit-exact fp32), incl. large coordinates where the fp64 -> fp32 rounding matters."""
    from oracle import geometry
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    tr = Tracer(load_mesh(room_stl), C, 100e9, 200e-9, 1, 1)
    verts, faces = tr.receiver_mesh([center, [0.0, 1.0, 2.0]], radius)
    assert verts.shape == (2, 42, 3) and faces.shape == (80, 3)
    got = verts[0][faces].view(np.uint32).reshape(80, 9)
    want = geometry.rx_soup(center, radius).view(np.uint32).reshape(80, 9)
    assert sorted(map(bytes, got)) == sorted(map(bytes, want))
    # every vertex within an ulp-scale band of the sphere, receiver 1 is an independent translate
    rad = np.linalg.norm(verts[0].astype(np.float64) - np.asarray(center), axis=1)
    assert np.all(np.abs(rad - radius) <= 4e-7 * (np.abs(center).max() + radius) + 1e-7 * radius)
    got1 = verts[1][faces].view(np.uint32).reshape(80, 9)
    want1 = geometry.rx_soup([0.0, 1.0, 2.0], radius).view(np.uint32).reshape(80, 9)
    assert sorted(map(bytes, got1)) == sorted(map(bytes, want1))


def test_far_receiver_parity(torch_cuda, almost_empty_stl):
    """A small receiver 150 m from the transmitter (coordinates ~100 m): every line-of-sight ray the reference
    receives is received here (the candidate filter in front of the exact receiver query must stay conservative when
    |c - p|^2 ~ 2e4 m^2 dwarfs r^2 = 1e-2 m^2)."""
    from oracle import cpu, geometry, post
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    n, B, tx, rx, r = 1 << 27, 2, [100.0, 3.0, 1.0], [-50.0, -2.0, 1.5], 0.1
    tr = Tracer(load_mesh(almost_empty_stl), C, 100e9, 1000e-9, B, n)
    out = tr.compute_cir_multi(tx, 1, [rx], r, return_paths=True, dense=False)
    rays = out["records"]["ray"].cpu().numpy().astype(np.int64)
    # candidates: rays whose direction passes the analytic sphere (a superset of the inscribed icosphere's hits)
    cand = cpu.sphere_hits(0, n, tx, rx, r)
    soup = geometry.load_stl_soup(almost_empty_stl)
    rxs = geometry.rx_soup(rx, r)
    want, want_paths = [], []
    for t in cand:
        o = cpu.trace_paths(soup, rxs, tx, B, int(t), 1, instrument=False)
        if o["mask"][0]:
            want.append(int(t))
            want_paths += post.clean_paths(o["received"], o["mask"])
    assert len(want) >= 6
    assert rays.tolist() == want
    nv = out["records"]["nverts"].cpu().numpy()
    paths = out["records"]["paths"].cpu().numpy()
    for row, k, p in zip(paths, nv, want_paths):
        assert np.array_equal(row[:k].view(np.uint32), p.view(np.uint32))


def _random_records(rng, n, n_rx, n_bins, clustered):
    rx = np.sort(rng.integers(0, n_rx, n)).astype(np.int32)
    ray = np.empty(n, dtype=np.uint32)
    for k in range(n_rx):   # ascending ray ids inside every receiver: the (receiver, ray id) order
        sel = np.nonzero(rx == k)[0]
        ray[sel] = np.sort(rng.choice(1 << 31, sel.size, replace=False)).astype(np.uint32)
    if clustered:           # arrivals pile up in a few bins (line-of-sight delays) + a uniform tail + out-of-window bins
        b = np.where(rng.random(n) < 0.7, rng.integers(0, 8, n) * 97 + 13, rng.integers(-5, n_bins + 50, n))
    else:
        b = rng.integers(0, n_bins, n)
    amp = rng.random(n) * 10.0 ** rng.integers(-12, -6, n)
    return rx, ray, b.astype(np.int64), amp


def _ordered_ir(rx, b, amp, n_rx, n_bins):
    ir = np.zeros((n_rx, n_bins))
    for k, bb, a in zip(rx, b, amp):   # tracer.py:116-117 in record order
        if 0 <= bb < n_bins:
            ir[k, bb] += a
    return ir


@pytest.mark.parametrize("n,n_rx,n_bins,clustered", [(70_000, 3, 10_000, True), (200_000, 16, 200_000, False),
                                                       (5_000, 700, 100, True), (1, 1, 1, False)])
def test_arrivals_build_is_the_reference_order(torch_cuda, n, n_rx, n_bins, clustered):
    """rfrt_arrivals_build: per (receiver, bin) the amplitudes are added in ray-id order — bit-identical to the
    sequential `impulse_response[bin] += amp` loop of tracer.py:116-117 — and the CSR feeds rfrt_rx_power."""
    import torch
    from rf_ray_tracing_warp_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n)
    rx, ray, b, amp = _random_records(rng, n, n_rx, n_bins, clustered)
    want = _ordered_ir(rx, b, amp, n_rx, n_bins)
    dev = torch.device("cuda", 0)
    slots = n + 1000                                            # spare slots beyond the device-side count
    pad = lambda a, fill: torch.from_numpy(np.concatenate([a, np.full(1000, fill, dtype=a.dtype)])).to(dev)  # noqa: E731
    d_rx, d_bin, d_amp = pad(rx, 0), pad(b, 0), pad(amp, 123.0)
    d_n = torch.tensor([n], dtype=torch.int64, device=dev)
    need = _lib.c_i64(0)
    assert lib.rfrt_records_workspace_bytes(slots, need) == 0
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    ir = torch.zeros((n_rx, n_bins), dtype=torch.float64, device=dev)
    off = torch.zeros(n_rx + 1, dtype=torch.int64, device=dev)
    abin = torch.empty(slots, dtype=torch.int32, device=dev)
    aamp = torch.empty(slots, dtype=torch.float64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.rfrt_arrivals_build(d_rx.data_ptr(), d_bin.data_ptr(), d_amp.data_ptr(), slots, d_n.data_ptr(), n_rx, n_bins,
                                       off.data_ptr(), abin.data_ptr(), aamp.data_ptr(), ir.data_ptr(), ws.data_ptr(), ws.numel(), s),
               "rfrt_arrivals_build")
    got = ir.cpu().numpy()
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    off = off.cpu().numpy()
    nnz = np.count_nonzero(want)
    assert off[0] == 0 and off[-1] == nnz and np.array_equal(np.diff(off), np.count_nonzero(want, axis=1))
    kk, bb = np.nonzero(want)
    assert np.array_equal(abin.cpu().numpy()[:nnz], bb) and np.array_equal(aamp.cpu().numpy()[:nnz], want[kk, bb])
    # the order-free paths agree up to fp64 summation order: atomics, and the shared-memory privatised histogram
    # (taken when <= 4 receivers' rows fit in shared memory and there are >= 65 536 records)
    for det in (0, 1):
        ir2 = torch.zeros_like(ir)
        _lib.check(lib.rfrt_bin_ir(d_rx.data_ptr(), d_bin.data_ptr(), d_amp.data_ptr(), slots, d_n.data_ptr(), n_rx, n_bins, det,
                                   ir2.data_ptr(), s), "rfrt_bin_ir")
        if det:
            assert np.array_equal(ir2.cpu().numpy().view(np.uint64), want.view(np.uint64))
        else:
            np.testing.assert_allclose(ir2.cpu().numpy(), want, rtol=1e-12, atol=0)


def test_privatised_histogram_path(torch_cuda):
    """rfrt_bin_ir(deterministic=0) with 4 receivers x 10 000 bins and 2^20 records takes k_bin_privatised
    (shared-memory windows, warp-aggregated adds): equal to the ordered sums to 1e-12, incl. a device-side count
    that cuts the record list in the middle of a CTA's slice."""
    import torch
    from rf_ray_tracing_warp_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(7)
    n, n_rx, n_bins = 1 << 20, 4, 10_000
    rx, ray, b, amp = _random_records(rng, n, n_rx, n_bins, True)
    perm = rng.permutation(n)                                   # order-free path: any record order
    rx, b, amp = rx[perm], b[perm], amp[perm]
    dev = torch.device("cuda", 0)
    d_rx, d_bin, d_amp = (torch.from_numpy(a).to(dev) for a in (rx, b, amp))
    s = torch.cuda.current_stream().cuda_stream
    for count in (n, n - 12345):
        want = np.zeros((n_rx, n_bins))
        ok = (b[:count] >= 0) & (b[:count] < n_bins)
        np.add.at(want, (rx[:count][ok], b[:count][ok]), amp[:count][ok])
        d_n = torch.tensor([count], dtype=torch.int64, device=dev)
        ir = torch.zeros((n_rx, n_bins), dtype=torch.float64, device=dev)
        _lib.check(lib.rfrt_bin_ir(d_rx.data_ptr(), d_bin.data_ptr(), d_amp.data_ptr(), n, d_n.data_ptr(), n_rx, n_bins, 0,
                                   ir.data_ptr(), s), "rfrt_bin_ir")
        np.testing.assert_allclose(ir.cpu().numpy(), want, rtol=1e-12, atol=0)


def test_records_sort_merges_segments(torch_cuda):
    """rfrt_records_pack + rfrt_records_sort: three 'ranks' with ragged record counts (one empty) merge into
    (receiver, ray id) order with every field carried along; an undersized segment reports its overflow."""
    import torch
    from rf_ray_tracing_warp_b200 import _lib, sharding
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(3)
    n_rx, B = 300, 3
    row = (B + 1) * 3
    cap = 5000
    nbytes = _lib.c_i64(0)
    assert lib.rfrt_record_segment_bytes(cap, row, nbytes) == 0
    segs = torch.zeros(3 * nbytes.value, dtype=torch.uint8, device=dev)
    recs, keep = [], []
    for g, cnt in enumerate([4000, 0, 5000]):
        ray = rng.choice(1 << 32, cnt, replace=False).astype(np.uint32)
        r = dict(ray=ray, rx=rng.integers(0, n_rx, cnt).astype(np.int32), nverts=rng.integers(2, B + 2, cnt).astype(np.int32),
                 bin=rng.integers(-3, 1 << 40, cnt).astype(np.int64), amp=rng.random(cnt), dist=rng.random(cnt) * 100,
                 paths=rng.random((cnt, row)).astype(np.float32))
        recs.append(r)
        d = {k: torch.from_numpy(np.concatenate([v, np.zeros((1,) + v.shape[1:], dtype=v.dtype)])).to(dev) for k, v in r.items()}
        keep.append(d)
        counters = torch.zeros(_lib.CTR_COUNT, dtype=torch.int64, device=dev)
        counters[_lib.CTR_RECORDS] = cnt
        counters[_lib.CTR_SEGMENTS] = 1000 * (g + 1)
        counters[_lib.CTR_CANDIDATES] = 7 * cnt
        _lib.check(lib.rfrt_records_pack(counters.data_ptr(), d["ray"].data_ptr(), d["rx"].data_ptr(), d["nverts"].data_ptr(),
                                         d["bin"].data_ptr(), d["amp"].data_ptr(), d["dist"].data_ptr(), d["paths"].data_ptr(),
                                         max(cnt, 1), row, segs.data_ptr() + g * nbytes.value, cap, s), "rfrt_records_pack")
    host = sharding.read_segments(segs, 3, cap, row)            # the host mirror of the layout reads what the device packed
    assert [h["produced"] for h in host] == [4000, 0, 5000]
    assert np.array_equal(host[2]["ray"], recs[2]["ray"]) and np.array_equal(host[0]["paths"], recs[0]["paths"])
    n = 3 * cap
    out = dict(ray=torch.empty(n, dtype=torch.int32, device=dev), rx=torch.empty(n, dtype=torch.int32, device=dev),
               nverts=torch.empty(n, dtype=torch.int32, device=dev), bin=torch.empty(n, dtype=torch.int64, device=dev),
               amp=torch.empty(n, dtype=torch.float64, device=dev), dist=torch.empty(n, dtype=torch.float64, device=dev),
               paths=torch.empty((n, row), dtype=torch.float32, device=dev))
    summary = torch.zeros(_lib.SUM_COUNT, dtype=torch.int64, device=dev)
    need = _lib.c_i64(0)
    lib.rfrt_records_workspace_bytes(n, need)
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    _lib.check(lib.rfrt_records_sort(segs.data_ptr(), 3, cap, row, n_rx, *(out[k].data_ptr() for k in
                                                                              ("ray", "rx", "nverts", "bin", "amp", "dist", "paths")),
                                     summary.data_ptr(), ws.data_ptr(), ws.numel(), s), "rfrt_records_sort")
    c = summary.cpu().numpy()
    assert c[_lib.SUM_RECORDS] == 9000 and c[_lib.SUM_OVERFLOWED] == 0 and c[_lib.SUM_MAX_RECORDS] == 5000
    assert c[_lib.SUM_COUNTERS + _lib.CTR_SEGMENTS] == 6000 and c[_lib.SUM_MAX_CANDIDATES] == 35000
    cat = {k: np.concatenate([r[k] for r in recs]) for k in recs[0]}
    order = np.argsort((cat["rx"].astype(np.int64) << 32) | cat["ray"].astype(np.int64))
    for k in cat:
        got = out[k].cpu().numpy()[:9000]
        assert np.array_equal(got.view(np.uint32) if k == "ray" else got, cat[k][order]), k
    # overflow: a segment smaller than the record count keeps the true count in its header
    small = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
    d = keep[2]
    counters = torch.zeros(_lib.CTR_COUNT, dtype=torch.int64, device=dev)
    counters[_lib.CTR_RECORDS] = 5000
    _lib.check(lib.rfrt_records_pack(counters.data_ptr(), d["ray"].data_ptr(), d["rx"].data_ptr(), d["nverts"].data_ptr(),
                                     d["bin"].data_ptr(), d["amp"].data_ptr(), d["dist"].data_ptr(), None, 5000, row,
                                     small.data_ptr(), 1000, s), "rfrt_records_pack")
    _lib.check(lib.rfrt_records_sort(small.data_ptr(), 1, 1000, row, n_rx, *(out[k].data_ptr() for k in
                                                                              ("ray", "rx", "nverts", "bin", "amp", "dist")), None,
                                     summary.data_ptr(), ws.data_ptr(), ws.numel(), s), "rfrt_records_sort")
    c = summary.cpu().numpy()
    assert c[_lib.SUM_RECORDS] == 1000 and c[_lib.SUM_OVERFLOWED] == 1 and c[_lib.SUM_MAX_RECORDS] == 5000


def test_exchange_capacity_overflow_is_retried(torch_cuda, room_stl):
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    mesh = load_mesh(room_stl)
    n, B, tx, rx, r = 1 << 16, 3, [10, 0, 5], [3.0, 6.0, 5.0], 1.5
    a = Tracer(mesh, C, 100e9, 200e-9, B, n)
    b = Tracer(mesh, C, 100e9, 200e-9, B, n, exchange_records=8)
    p1, ir1 = a.compute_cir(tx, 1, rx, r)
    p2, ir2 = b.compute_cir(tx, 1, rx, r)
    assert len(p1) == len(p2) > 100 and b.exchange_records >= len(p2)
    assert all(np.array_equal(x, y) for x, y in zip(p1, p2)) and np.array_equal(ir1, ir2)


def test_headless_main_driver(torch_cuda, room_stl, tmp_path, capsys):
    """python -m rf_ray_tracing_warp_b200.main (main.py:36-55 without the blocking plot / viewer) on a small input:
    files written, received paths / impulse response / RX power == the reference procedure."""
    from oracle import cpu, geometry, post
    from rf_ray_tracing_warp_b200 import main as drv
    n, B, tx, rx, r = 200_000, 4, [10, 0, 5], [3.0, 6.0, 5.0], 0.5
    res = drv.main(["--model", room_stl, "--tx", *map(str, tx), "--rx", *map(str, rx), "--rx-radius", str(r), "--rays", str(n),
                    "--bounces", str(B), "--out", str(tmp_path), "--scene", "glb"])
    o = cpu.trace_paths(geometry.load_stl_soup(room_stl), geometry.rx_soup(rx, r), tx, B, 0, n, instrument=False)
    o_paths = post.clean_paths(o["received"], o["mask"])
    o_ir = post.impulse_response(o_paths, 1, n, C, drv.SAMPLE_RATE_HZ, drv.SAMPLE_WINDOW_S)
    ir = np.load(tmp_path / "impulse_response.npy")
    assert ir.shape == (20000,) and np.array_equal(ir != 0, o_ir != 0)
    np.testing.assert_allclose(ir, o_ir, rtol=1e-5, atol=0)
    saved = np.load(tmp_path / "paths.npz")
    assert len(saved.files) == len(o_paths) == res["received_paths"] > 20
    for k, p in enumerate(o_paths):
        assert np.array_equal(saved[f"arr_{k}"], p)
    want = post.rx_power(o_ir, drv.SAMPLE_WINDOW_S)
    np.testing.assert_allclose(res["rx_power"], want, rtol=1e-4)
    np.testing.assert_allclose(res["rx_power_dbm"], post.to_dbm(want), rtol=1e-4)
    assert json.load(open(tmp_path / "result.json"))["received_paths"] == len(o_paths)
    assert (tmp_path / "scene.glb").stat().st_size > 1000
    assert "Signal RX power" in capsys.readouterr().out                 # main.py:55


def test_headless_coverage_driver(torch_cuda, room_stl, tmp_path):
    """python -m rf_ray_tracing_warp_b200.coverage with the reference lattice (coverage.py:38-40: 16 x 16 x 8 receivers)
    at 200 000 rays x 2 bounces: the dBm grid of sampled cells == the reference procedure per receiver."""
    from oracle import cpu, geometry, post
    from rf_ray_tracing_warp_b200 import coverage as drv
    n, tx = 200_000, [10, 0, 5]
    cov = drv.main(["--model", room_stl, "--tx", *map(str, tx), "--rays", str(n), "--out", str(tmp_path)])
    dbm = np.load(tmp_path / "coverage_dbm.npy")
    rxs = np.load(tmp_path / "receivers.npy")
    assert dbm.shape == (16, 16, 8) and rxs.shape == (2048, 3) and np.array_equal(rxs, drv.reference_lattice())
    soup = geometry.load_stl_soup(room_stl)
    flat = dbm.reshape(-1)
    lit = np.nonzero(np.isfinite(flat))[0]
    assert lit.size > 200
    rng = np.random.default_rng(0)
    for k in np.concatenate([rng.choice(lit, 10, replace=False), rng.choice(2048, 6, replace=False)]):
        o = cpu.trace_paths(soup, geometry.rx_soup(rxs[k], drv.__dict__.get("RX_RADIUS", 0.1)), tx, drv.MAX_BOUNCES, 0, n,
                            instrument=False)
        o_ir = post.impulse_response(post.clean_paths(o["received"], o["mask"]), 1, n, C, drv.SAMPLE_RATE_HZ, drv.SAMPLE_WINDOW_S)
        p = post.rx_power(o_ir, drv.SAMPLE_WINDOW_S)
        if np.isnan(p):
            assert not np.isfinite(flat[k])
        else:
            np.testing.assert_allclose(cov["power"][k], p, rtol=1e-4)
            np.testing.assert_allclose(flat[k], post.to_dbm(p), rtol=1e-4, atol=1e-3)


def test_dense_power_kernel_matches_np_convolve(torch_cuda):
    """rfrt_rx_power_dense (closed-form sums per run of samples with a constant arrival window) against the literal
    main.py:39,46-55 evaluation (np.convolve "same", np.nonzero selection) of the same impulse-response rows: sparse
    rows with arrivals at the window's ends, at the centre bin, single arrivals, an empty row, and even / odd lengths."""
    import torch
    from oracle import post
    from rf_ray_tracing_warp_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(23)
    for L, window in ((1000, 10e-9), (1001, 10e-9), (10_000, 100e-9)):
        half = (L - 1) // 2
        rows = np.zeros((48, L))
        for k in range(48):
            nnz = [0, 1, 1, 2, 3][k] if k < 5 else int(rng.integers(1, 60))
            bins = rng.choice(L, size=nnz, replace=False)
            if k == 1:
                bins = np.array([half])            # the one arrival sits on the centre bin
            if k == 2:
                bins = np.array([0])
            if k == 3:
                bins = np.array([0, L - 1])
            if k == 4:
                bins = np.array([half - 1, half, half + 1])
            rows[k, bins] = rng.uniform(0.1, 2.0, size=len(bins)) * 10.0 ** rng.integers(-9, -3)
        d_ir = torch.from_numpy(rows).to(dev)
        power = torch.empty(48, dtype=torch.float64, device=dev)
        _lib.check(lib.rfrt_rx_power_dense(d_ir.data_ptr(), 48, L, window, 2.4e9, power.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "rfrt_rx_power_dense")
        got = power.cpu().numpy()
        want = np.array([post.rx_power(rows[k], window) for k in range(48)])
        assert np.isnan(got[0]) and np.isnan(want[0])
        np.testing.assert_allclose(got[1:], want[1:], rtol=1e-9, atol=0)


@pytest.mark.parametrize("scene", ["room", "terrain"])
def test_records_do_not_depend_on_keeping_the_paths(torch_cuda, room_stl, scene):
    """compute_cir_multi with and without return_paths: ray, receiver, vertex count, bin, amplitude and distance of the
    records are bit-identical (the path-less variants of the replay may post-process differently — a streamed
    evaluation was tried and measured slower — but never to another result), also with a per-triangle material table."""
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh, synthetic_terrain
    if scene == "room":
        mesh, tx, B = load_mesh(room_stl), [10, 0, 5], 6
        rxs, r = np.array([[3.0, 6.0, 5.0], [-8.0, 8.0, 3.0], [5.0, -3.0, 2.0], [5.2, -3.1, 2.1]]), 0.6
    else:
        mesh, tx, B = synthetic_terrain(96, 20.0, 17), [10, 0, 4.5], 6
        rxs, r = np.array([[0.0, 0.0, 3.0], [4.0, -3.0, 2.5], [-6.0, 5.0, 3.5]]), 0.9
    n = 1 << 17
    tr = Tracer(mesh, C, 100e9, 200e-9, B, n)
    rng = np.random.default_rng(1)
    for materials in (None, rng.uniform(1.5, 9.0, size=tr.mesh_info()["n_triangles"])):
        tr.set_materials(materials)
        with_paths = {k: v.cpu().numpy() for k, v in tr.compute_cir_multi(tx, 1, rxs, r, return_paths=True)["records"].items() if v is not None}
        without = {k: v.cpu().numpy() for k, v in tr.compute_cir_multi(tx, 1, rxs, r, return_paths=False)["records"].items() if v is not None}
        assert with_paths["ray"].shape[0] > 300 and (with_paths["nverts"] > 3).sum() > 20
        for name in ("ray", "rx", "nverts", "bin", "amp", "dist"):
            assert np.array_equal(with_paths[name].view(np.uint8), without[name].view(np.uint8)), (name, materials is not None)
