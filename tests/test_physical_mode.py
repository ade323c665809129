"""Physical mode (include/rfrt.h rfrt_trace_physical): no reference to compare with — the CPU
restatement (oracle_trace_physical) defines it, Friis' formula validates it, the CUDA path must agree with it."""
import numpy as np
import pytest

C = 2.998e8
F = 2.4e9


def test_oracle_line_of_sight_matches_friis():
    """One receiver in free space: |field|^2 must approach (lambda / (4 pi d))^2."""
    from oracle import cpu
    n = 1 << 22
    d, r = 10.0, 0.2
    out = cpu.trace_physical(np.zeros((0, 9), np.float32), [[d, 0.0, 0.0]], r, [0.0, 0.0, 0.0], 3, 0, n, n, F, C)
    lam = C / F
    friis = (lam / (4 * np.pi * d)) ** 2
    expected_rays = n * r * r / (4 * d * d)
    assert abs(out["arrivals"] - expected_rays) < 5 * np.sqrt(expected_rays)
    assert out["segments"] == n
    np.testing.assert_allclose(abs(out["field"][0]) ** 2, friis, rtol=0.15)


def test_oracle_mirror_reflection_adds_second_path(repo_root):
    """A large perfect-ish mirror (n = 1e6 -> |Gamma| -> 1) below a TX/RX pair: two-ray model."""
    from oracle import cpu
    s = 200.0
    soup = np.array([[-s, -s, 0, s, -s, 0, s, s, 0], [-s, -s, 0, s, s, 0, -s, s, 0]], dtype=np.float32)
    n, r, h, d = 1 << 23, 0.25, 2.0, 12.0
    out = cpu.trace_physical(soup, [[d, 0.0, h]], r, [0.0, 0.0, h], 2, 0, n, n, F, C, materials=[1.0e6, 1.0e6])
    lam = C / F
    d2 = np.hypot(d, 2 * h)
    # Gamma -> (ct - n ci)/(ct + n ci) -> -1 for n -> inf
    two_ray = lam / (4 * np.pi) * (np.exp(-2j * np.pi * d / lam) / d - np.exp(-2j * np.pi * d2 / lam) / d2)
    assert abs(abs(out["field"][0]) - abs(two_ray)) < 0.2 * abs(two_ray) + 0.1 * lam / (4 * np.pi * d)


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["room", "terrain"])
def test_gpu_physical_matches_oracle(room_stl, scene):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device — no CPU fallback exists"
    from oracle import cpu, geometry
    from rf_ray_tracing_warp_b200 import Tracer, load_mesh, synthetic_terrain
    rng = np.random.default_rng(3)
    if scene == "room":
        mesh, soup, tx, B, n = load_mesh(room_stl), geometry.load_stl_soup(room_stl), [10.0, 0.0, 5.0], 5, 1 << 18
        rxs = np.array([[3.0, 6.0, 5.0], [-8.0, 8.0, 3.0], [12.0, -12.0, 14.0], [10.3, 0.1, 5.0], [0.0, -5.0, 50.0]])
        radius, bvh = 0.6, None
    else:
        mesh = synthetic_terrain(96, 20.0, 17)
        soup, tx, B, n = mesh.triangles.astype(np.float32), [10.0, 0.0, 4.5], 4, 1 << 17
        rxs = np.array([[0.0, 0.0, 3.0], [5.0, 5.0, 1.5], [-10.0, 3.0, 2.0]])
        radius, bvh = 1.0, cpu.Bvh(soup)
    mats = rng.uniform(1.5, 8.0, soup.reshape(-1, 9).shape[0]).astype(np.float32)
    tr = Tracer(mesh, C, 100e9, 100e-9, B, n, chunk_rays=50_000)
    tr.set_materials(mats)
    out = tr.trace_physical(tx, 2.0, rxs, radius, carrier_hz=F, want_ir=True)
    o = cpu.trace_physical(soup, rxs, radius, tx, B, 0, n, n, F, C, materials=mats, sample_rate=100e9, n_bins=10000, bvh=bvh)
    assert out["stats"]["segments"] == o["segments"] and out["stats"]["arrivals"] == o["arrivals"] > 1000
    scale = np.abs(o["field"]).max()
    np.testing.assert_allclose(out["field"], o["field"], rtol=1e-9, atol=1e-9 * scale)
    np.testing.assert_allclose(out["power"], 2.0 * np.abs(o["field"]) ** 2, rtol=1e-8, atol=1e-16)
    ir = out["impulse_response"].cpu().numpy()
    assert np.array_equal(ir != 0, o["ir"] != 0)
    np.testing.assert_allclose(ir, o["ir"], rtol=1e-9, atol=1e-9 * scale)
    # a receiver far outside everything hears nothing; default materials path
    tr.set_materials(None)
    out5 = tr.trace_physical(tx, 1.0, rxs, radius, carrier_hz=F)
    o5 = cpu.trace_physical(soup, rxs, radius, tx, B, 0, n, n, F, C, bvh=bvh)
    np.testing.assert_allclose(out5["field"], o5["field"], rtol=1e-9, atol=1e-9 * scale)
