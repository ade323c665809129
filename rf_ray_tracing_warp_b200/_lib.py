"""ctypes binding of librfrt.so (the C ABI declared in include/rfrt.h).

There is NO fallback: if the CUDA library is missing or fails to load, importing the product's compute
entry points raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C rf_ray_tracing_warp_b200/csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librfrt.so")

c_void_p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_u32 = ctypes.c_uint32
c_u64 = ctypes.c_uint64
c_f = ctypes.c_float
c_d = ctypes.c_double

CTR_SEGMENTS, CTR_CANDIDATES, CTR_RECORDS, CTR_ENV_HITS, CTR_NEXT_RAY, CTR_NEXT_CAND, CTR_CHECKSUM, CTR_QUEUE_OVERFLOW, \
    CTR_NODE_VISITS, CTR_TRI_TESTS, CTR_COUNT = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
FLAG_NONE, FLAG_DIRS_READY, FLAG_FORCE_BVH, FLAG_CHECKSUM, FLAG_NO_RAY_SORT = 0, 1, 2, 8, 16
SMALL_MAX_TRIS = 64
TRI_TEST_WOOP, TRI_TEST_MT = 0, 1
# d_summary of rfrt_records_sort (u64[16])
SUM_RECORDS, SUM_OVERFLOWED, SUM_COUNTERS, SUM_MAX_RECORDS, SUM_MAX_CANDIDATES, SUM_COUNT = 0, 1, 2, 12, 13, 16

# name -> (restype, argtypes); mirrors include/rfrt.h one to one
SIGNATURES = {
    "rfrt_version": (ctypes.c_int, []),
    "rfrt_last_error": (ctypes.c_char_p, []),
    "rfrt_device_info": (ctypes.c_int, [ctypes.POINTER(c_i32)] * 3),
    "rfrt_mesh_create": (ctypes.c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, ctypes.POINTER(c_u64)]),
    "rfrt_mesh_destroy": (ctypes.c_int, [c_u64]),
    "rfrt_mesh_set_materials": (ctypes.c_int, [c_u64, c_void_p, c_void_p]),
    "rfrt_mesh_set_triangle_test": (ctypes.c_int, [c_u64, c_i32]),
    "rfrt_mesh_info": (ctypes.c_int, [c_u64, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_f),
                                      ctypes.POINTER(c_i32), ctypes.POINTER(c_f)]),
    "rfrt_mesh_export": (ctypes.c_int, [c_u64, c_void_p, c_void_p, c_void_p]),
    "rfrt_small_scene_tables": (ctypes.c_int, [c_void_p, c_i32, c_void_p, c_void_p, ctypes.POINTER(c_i32),
                                               ctypes.POINTER(c_f), c_void_p, c_void_p]),
    "rfrt_rxset_create": (ctypes.c_int, [c_void_p, c_i64, c_d, ctypes.POINTER(c_d), c_i32, ctypes.POINTER(c_i32),
                                         c_i32, c_void_p, ctypes.POINTER(c_u64)]),
    "rfrt_rxset_destroy": (ctypes.c_int, [c_u64]),
    "rfrt_rxset_export": (ctypes.c_int, [c_u64, c_void_p, c_void_p]),
    "rfrt_ray_directions": (ctypes.c_int, [c_i64, c_i64, c_void_p, c_void_p]),
    "rfrt_trace": (ctypes.c_int, [c_u64, c_u64, ctypes.POINTER(c_f), c_i32, c_i64, c_i64, c_u32, c_void_p, c_i64,
                                  c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p]),
    "rfrt_trace_receive": (ctypes.c_int, [c_u64, c_u64, ctypes.POINTER(c_f), c_i32, c_void_p, c_i64, c_void_p, c_d,
                                          c_d, c_d, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_i64, c_void_p, c_i64, c_void_p]),
    "rfrt_bin_ir": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i64, c_i32, c_void_p,
                                   c_void_p]),
    "rfrt_mesh_reserve_rays": (ctypes.c_int, [c_u64, c_i64]),
    "rfrt_record_segment_bytes": (ctypes.c_int, [c_i64, c_i32, ctypes.POINTER(c_i64)]),
    "rfrt_records_workspace_bytes": (ctypes.c_int, [c_i64, ctypes.POINTER(c_i64)]),
    "rfrt_records_pack": (ctypes.c_int, [c_void_p] * 8 + [c_i64, c_i32, c_void_p, c_i64, c_void_p]),
    "rfrt_records_sort": (ctypes.c_int, [c_void_p, c_i64, c_i64, c_i32, c_i64] + [c_void_p] * 9 + [c_i64, c_void_p]),
    "rfrt_arrivals_build": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i64, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "rfrt_rx_power": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_d, c_d, c_void_p, c_void_p,
                                     c_void_p]),
    "rfrt_rx_power_dense": (ctypes.c_int, [c_void_p, c_i64, c_i64, c_d, c_d, c_void_p, c_void_p]),
    "rfrt_trace_physical": (ctypes.c_int, [c_u64, c_u64, ctypes.POINTER(c_f), c_i32, c_i64, c_i64, c_i64, c_d, c_d, c_d,
                                           c_i64, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rfrt_trace_paths_compat": (ctypes.c_int, [c_u64, ctypes.POINTER(c_f), c_u64, c_i64, c_i32, c_i64, c_i64,
                                               c_void_p, c_void_p, c_void_p, c_void_p]),
    "rfrt_query_closest": (ctypes.c_int, [c_u64, c_void_p, c_void_p, c_i64, c_f, c_void_p, c_void_p, c_void_p]),
}

_LIB = None


class RfrtError(RuntimeError):
    pass


def load():
    """Loads librfrt.so and declares every entry point; raises if the library is absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RfrtError(f"{LIB_PATH} not found — the CUDA extension is not built and there is no CPU fallback "
                            "(run `make -C rf_ray_tracing_warp_b200/csrc`)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _LIB = lib
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = load().rfrt_last_error()
        raise RfrtError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def float3(v):
    return (c_f * 3)(float(v[0]), float(v[1]), float(v[2]))
