"""CPU tests: the C-ABI library loads and exports every symbol include/rfrt.h declares (no compute calls),
the host-side loaders agree with the oracle's independent ones, and the product never touches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest


def _declared(repo_root):
    text = open(os.path.join(repo_root, "include", "rfrt.h")).read()
    return sorted(set(re.findall(r"RFRT_API\s+(?:const\s+char\s*\*|int)\s*(rfrt_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(repo_root):
    names = _declared(repo_root)
    assert len(names) >= 17
    path = os.path.join(repo_root, "rf_ray_tracing_warp_b200", "csrc", "librfrt.so")
    assert os.path.exists(path), "librfrt.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(path)
    for n in names:
        assert hasattr(lib, n), n
    lib.rfrt_version.restype = ctypes.c_int
    assert lib.rfrt_version() == 100


def test_binding_table_matches_header(repo_root):
    from rf_ray_tracing_warp_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared(repo_root)
    _lib.load()
    # argument counts agree with the header
    text = open(os.path.join(repo_root, "include", "rfrt.h")).read()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"RFRT_API[^;]*?\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        args = m.group(1).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert n == len(argtypes), name


def test_binding_constants_match_header(repo_root):
    """Every RFRT_FLAG_* / RFRT_CTR_* / RFRT_SMALL_MAX_TRIS of include/rfrt.h has the same value in the ctypes mirror."""
    from rf_ray_tracing_warp_b200 import _lib
    text = open(os.path.join(repo_root, "include", "rfrt.h")).read()
    defines = dict(re.findall(r"#define\s+RFRT_((?:FLAG|CTR)_\w+|SMALL_MAX_TRIS)\s+(\d+)u?\b", text))
    assert len(defines) >= 14
    for name, value in defines.items():
        assert getattr(_lib, name) == int(value), name


def test_error_path_without_gpu(repo_root):
    """Error behaviour: bad arguments return a negative status and set a message; nothing throws."""
    from rf_ray_tracing_warp_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_uint64(0)
    assert lib.rfrt_mesh_create(None, -1, None, 5, None, h) == -1
    assert b"rfrt_mesh_create" in lib.rfrt_last_error()
    assert lib.rfrt_mesh_destroy(12345) == -3
    assert lib.rfrt_trace(999, 0, None, 3, 0, 10, 0, None, 0, None, None, 0, None, None, None) == -3


def test_tracer_fails_loudly_without_cuda(room_stl):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from rf_ray_tracing_warp_b200 import RfrtError, Tracer, load_mesh
    with pytest.raises(RfrtError):
        Tracer(load_mesh(room_stl), 2.998e8, 100e9, 200e-9, 3, 1000)


def test_product_does_not_import_oracle(repo_root):
    pkg = os.path.join(repo_root, "rf_ray_tracing_warp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle's answer", ""), os.path.join(dirpath, f)


def test_stl_loader_matches_oracle_reader(room_stl, almost_empty_stl):
    from oracle import geometry
    from rf_ray_tracing_warp_b200 import load_mesh
    for path, nv in [(room_stl, 24), (almost_empty_stl, 8)]:
        m = load_mesh(path)
        assert m.vertices.dtype == np.float64 and m.faces.dtype == np.int64
        assert m.vertices.shape == (nv, 3)     # trimesh-style merge of identical corners
        assert np.array_equal(m.triangles.astype(np.float32), geometry.load_stl_soup(path))


def test_ascii_stl(tmp_path):
    from rf_ray_tracing_warp_b200 import load_mesh
    p = tmp_path / "t.stl"
    p.write_text("solid t\nfacet normal 0 0 1\nouter loop\nvertex 0 0 0\nvertex 1 0 0\nvertex 0 1 0\nendloop\nendfacet\n"
                 "facet normal 0 0 1\nouter loop\nvertex 1 0 0\nvertex 1 1 0\nvertex 0 1 0\nendloop\nendfacet\nendsolid t\n")
    m = load_mesh(str(p))
    assert m.faces.shape == (2, 3) and m.vertices.shape == (4, 3)


def test_unit_icosphere_matches_oracle():
    from oracle import geometry
    from rf_ray_tracing_warp_b200 import unit_icosphere
    v, f = unit_icosphere(1)
    u, g = geometry.unit_icosphere_subdiv1()
    assert np.array_equal(v, u) and np.array_equal(f, g)
    v2, f2 = unit_icosphere(2)
    assert v2.shape == (162, 3) and f2.shape == (320, 3)


def test_synthetic_terrain_is_deterministic():
    from rf_ray_tracing_warp_b200 import synthetic_terrain
    a, b = synthetic_terrain(32, seed=17), synthetic_terrain(32, seed=17)
    assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.faces, b.faces)
    assert a.faces.shape == (2048, 3) and np.abs(a.vertices[:, 2]).max() <= 1.5
    assert not np.array_equal(a.vertices, synthetic_terrain(32, seed=18).vertices)


def test_scene_export_round_trip_and_reads_trimesh_pages(tmp_path, room_stl):
    """viz/visualization.py replacement: GLB/HTML written by export_scene parse back to the same polylines, and the
    reader also understands a trimesh scene_to_html page (structure of the reference's web/scene.html)."""
    from rf_ray_tracing_warp_b200 import load_mesh
    from rf_ray_tracing_warp_b200.scene_export import export_scene, read_glb_paths
    rng = np.random.default_rng(0)
    paths = [rng.normal(size=(k, 3)).astype(np.float32) for k in (2, 3, 5, 4)]
    for name in ("scene.glb", "scene.html"):
        p = tmp_path / name
        assert export_scene(str(p), load_mesh(room_stl), [10, 0, 5], [-10, 0, 5], paths) > 2000
        back = read_glb_paths(str(p))
        assert len(back) == len(paths) and all(np.array_equal(a, b) for a, b in zip(back, paths))


def test_stl_attribute_words_and_material_table(room_stl, almost_empty_stl):
    from rf_ray_tracing_warp_b200 import load_stl_attributes, materials_from_attributes
    a = load_stl_attributes(room_stl)
    assert a.dtype == np.uint16 and a.shape == (44,) and np.all(a == 20083)
    assert np.all(load_stl_attributes(almost_empty_stl) == 0)
    m = materials_from_attributes(np.array([1, 2, 2, 7]), {2: 3.5, 7: 1.2})
    assert m.dtype == np.float32 and np.allclose(m, [5.0, 3.5, 3.5, 1.2])
