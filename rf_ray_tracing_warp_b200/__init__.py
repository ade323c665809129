"""rf_ray_tracing_warp_b200 — B200-native (sm_100a) RF ray-tracing hot path behind the reference's API.

    from rf_ray_tracing_warp_b200 import Tracer, load_mesh
    tracer = Tracer(load_mesh("models/room.stl"), 2.998e8, 100e9, 200e-9, 3, 1_000_000)
    paths, impulse_response = tracer.compute_cir([10, 0, 5], 1, [-10, 0, 5], 0.1)

The compute path is librfrt.so (rf_ray_tracing_warp_b200/csrc, C ABI in include/rfrt.h).  No CPU fallback.
"""
from .mesh_io import (Mesh, load_mesh, load_stl_attributes, load_stl_triangles, materials_from_attributes,
                      mesh_from_triangles, synthetic_terrain, unit_icosphere)
from ._lib import RfrtError


def __getattr__(name):
    # Tracer needs torch + the CUDA library; import lazily so mesh utilities work without them
    if name in ("Tracer", "to_dbm"):
        from . import tracer
        return getattr(tracer, name)
    raise AttributeError(name)


__all__ = ["Tracer", "to_dbm", "Mesh", "load_mesh", "load_stl_attributes", "load_stl_triangles", "materials_from_attributes",
           "mesh_from_triangles",
           "synthetic_terrain", "unit_icosphere", "RfrtError"]
