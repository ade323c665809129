"""ctypes wrapper over oracle/liboracle.so (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i8p = ctypes.POINTER(ctypes.c_int8)
_f64p = ctypes.POINTER(ctypes.c_double)


def build():
    subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.oracle_bvh_create.restype = ctypes.c_void_p
        L.oracle_bvh_create.argtypes = [_f32p, ctypes.c_int64]
        L.oracle_bvh_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_trace_paths.restype = None
        L.oracle_trace_paths.argtypes = [_f32p, ctypes.c_int64, ctypes.c_void_p, _f32p, ctypes.c_int64, _f32p,
                                         ctypes.c_int, ctypes.c_int64, ctypes.c_int64, _f32p, _f32p, _u32p, _i32p,
                                         _f32p, _i8p, ctypes.c_int]
        L.oracle_trace_received.restype = ctypes.c_int64
        L.oracle_trace_received.argtypes = [_f32p, ctypes.c_int64, ctypes.c_void_p, _f32p, ctypes.c_int64, _f32p,
                                            ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _i64p,
                                            _f32p, ctypes.c_int]
        L.oracle_trace_env.restype = ctypes.c_uint64
        L.oracle_trace_env.argtypes = [_f32p, ctypes.c_int64, ctypes.c_void_p, _f32p, ctypes.c_int, ctypes.c_int64,
                                       ctypes.c_int64, _i32p, _f32p, ctypes.c_int]
        L.oracle_trace_physical.restype = ctypes.c_uint64
        L.oracle_trace_physical.argtypes = [_f32p, ctypes.c_int64, ctypes.c_void_p, _f32p, _f64p, ctypes.c_int64,
                                            ctypes.c_double, _f32p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_int64, _f64p, _f64p, ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_ray_directions.argtypes = [ctypes.c_int64, ctypes.c_int64, _f32p]
        L.oracle_ray_directions_list.argtypes = [_i64p, ctypes.c_int64, _f32p]
        L.oracle_pcg.restype = ctypes.c_uint32
        L.oracle_pcg.argtypes = [ctypes.c_uint32]
        L.oracle_det_math.argtypes = [_f64p, ctypes.c_int64, _f64p, _f64p, _f64p]
        L.oracle_query.restype = ctypes.c_int
        L.oracle_query.argtypes = [_f32p, ctypes.c_int64, ctypes.c_void_p, _f32p, _f32p, ctypes.c_float, _f32p,
                                   _i32p]
        L.oracle_tri_normal.argtypes = [_f32p, _f32p]
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_sphere_hits.restype = ctypes.c_int64
        L.oracle_sphere_hits.argtypes = [ctypes.c_int64, ctypes.c_int64, _f64p, _f64p, ctypes.c_double, _i64p,
                                         ctypes.c_int64, ctypes.c_int]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _soup(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, 9))
    return a


class Bvh:
    """Optional CPU BVH over a soup (same answers as brute force; only faster)."""

    def __init__(self, soup):
        self.soup = _soup(soup)
        self.handle = lib().oracle_bvh_create(_p(self.soup, _f32p), self.soup.shape[0])

    def __del__(self):
        if getattr(self, "handle", None):
            lib().oracle_bvh_destroy(self.handle)
            self.handle = None


def ray_directions(tid_begin, n):
    out = np.empty((n, 3), dtype=np.float32)
    lib().oracle_ray_directions(int(tid_begin), int(n), _p(out, _f32p))
    return out


def ray_directions_list(tids):
    tids = np.ascontiguousarray(np.asarray(tids, dtype=np.int64))
    out = np.empty((tids.size, 3), dtype=np.float32)
    lib().oracle_ray_directions_list(_p(tids, _i64p), tids.size, _p(out, _f32p))
    return out


def pcg(s):
    return int(lib().oracle_pcg(ctypes.c_uint32(int(s) & 0xFFFFFFFF)))


def det_math(x):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    s = np.empty_like(x); c = np.empty_like(x); ac = np.empty_like(x)
    lib().oracle_det_math(_p(x, _f64p), x.size, _p(s, _f64p), _p(c, _f64p), _p(ac, _f64p))
    return s, c, ac


def query(soup, p, d, max_t=1.0e6, bvh=None):
    soup = _soup(soup if bvh is None else bvh.soup)
    p = np.ascontiguousarray(np.asarray(p, dtype=np.float32)); d = np.ascontiguousarray(np.asarray(d, dtype=np.float32))
    t = ctypes.c_float(0.0); f = ctypes.c_int32(-1)
    hit = lib().oracle_query(_p(soup, _f32p), soup.shape[0], bvh.handle if bvh else None, _p(p, _f32p), _p(d, _f32p),
                             ctypes.c_float(max_t), ctypes.byref(t), ctypes.byref(f))
    return bool(hit), float(t.value), int(f.value)


def tri_normal(tri):
    tri = np.ascontiguousarray(np.asarray(tri, dtype=np.float32).reshape(9))
    n = np.empty(3, dtype=np.float32)
    lib().oracle_tri_normal(_p(tri, _f32p), _p(n, _f32p))
    return n


def trace_paths(env_soup, rx_soup, tx_pos, max_bounces, tid_begin, n, instrument=True, bvh=None, nthreads=0):
    """Dense reference-shaped outputs (kernel.py:38-98 launched as tracer.py:75-79).

    returns dict(traced (n,B+1,3) f32, received (n,B+1,3) f32, mask (n,) u32[, hit_tri, hit_t, event (n,B)])
    """
    env = _soup(env_soup if bvh is None else bvh.soup)
    rx = _soup(rx_soup) if rx_soup is not None else np.zeros((0, 9), dtype=np.float32)
    tx = np.ascontiguousarray(np.asarray(tx_pos, dtype=np.float32))
    B = int(max_bounces)
    traced = np.empty((n, B + 1, 3), dtype=np.float32)
    received = np.empty((n, B + 1, 3), dtype=np.float32)
    mask = np.empty(n, dtype=np.uint32)
    hit_tri = np.empty((n, B), dtype=np.int32) if instrument else None
    hit_t = np.empty((n, B), dtype=np.float32) if instrument else None
    event = np.empty((n, B), dtype=np.int8) if instrument else None
    lib().oracle_trace_paths(_p(env, _f32p), env.shape[0], bvh.handle if bvh else None, _p(rx, _f32p), rx.shape[0],
                             _p(tx, _f32p), B, int(tid_begin), int(n), _p(traced, _f32p), _p(received, _f32p),
                             _p(mask, _u32p), _p(hit_tri, _i32p), _p(hit_t, _f32p), _p(event, _i8p), int(nthreads))
    out = dict(traced=traced, received=received, mask=mask)
    if instrument:
        out.update(hit_tri=hit_tri, hit_t=hit_t, event=event)
    return out


def trace_received(env_soup, rx_soup, tx_pos, max_bounces, tid_begin, n, cap=1 << 20, bvh=None, nthreads=0):
    """Sparse variant for big n: (tids (m,), rows (m,B+1,3)) of received rays, ascending tid."""
    env = _soup(env_soup if bvh is None else bvh.soup)
    rx = _soup(rx_soup)
    tx = np.ascontiguousarray(np.asarray(tx_pos, dtype=np.float32))
    B = int(max_bounces)
    tids = np.empty(cap, dtype=np.int64)
    rows = np.empty((cap, B + 1, 3), dtype=np.float32)
    m = lib().oracle_trace_received(_p(env, _f32p), env.shape[0], bvh.handle if bvh else None, _p(rx, _f32p),
                                    rx.shape[0], _p(tx, _f32p), B, int(tid_begin), int(n), int(cap), _p(tids, _i64p),
                                    _p(rows, _f32p), int(nthreads))
    if m > cap:
        raise RuntimeError(f"oracle: {m} received rays exceed cap {cap}")
    return tids[:m].copy(), rows[:m].copy()


def trace_env(env_soup, tx_pos, max_bounces, tid_begin, n, instrument=True, bvh=None, nthreads=0):
    """Environment-only trajectory: (segments, hit_tri (n,B) i32, hit_t (n,B) f32)."""
    env = _soup(env_soup if bvh is None else bvh.soup)
    tx = np.ascontiguousarray(np.asarray(tx_pos, dtype=np.float32))
    B = int(max_bounces)
    hit_tri = np.empty((n, B), dtype=np.int32) if instrument else None
    hit_t = np.empty((n, B), dtype=np.float32) if instrument else None
    seg = lib().oracle_trace_env(_p(env, _f32p), env.shape[0], bvh.handle if bvh else None, _p(tx, _f32p), B,
                                 int(tid_begin), int(n), _p(hit_tri, _i32p), _p(hit_t, _f32p), int(nthreads))
    return int(seg), hit_tri, hit_t


def trace_physical(env_soup, rx_centers, rx_radius, tx_pos, max_bounces, tid_begin, n, n_total, carrier_hz, light_speed,
                   materials=None, sample_rate=0.0, n_bins=0, bvh=None):
    """Physical mode (oracle_trace_physical): returns dict(field (R,) complex128, ir (R,L) complex128 or None,
    segments, arrivals)."""
    env = _soup(env_soup if bvh is None else bvh.soup)
    c = np.ascontiguousarray(np.asarray(rx_centers, dtype=np.float64).reshape(-1, 3))
    tx = np.ascontiguousarray(np.asarray(tx_pos, dtype=np.float32))
    mat = np.ascontiguousarray(np.asarray(materials, dtype=np.float32)) if materials is not None else None
    field = np.zeros((c.shape[0], 2), dtype=np.float64)
    ir = np.zeros((c.shape[0], int(n_bins), 2), dtype=np.float64) if n_bins else None
    arr = ctypes.c_uint64(0)
    seg = lib().oracle_trace_physical(_p(env, _f32p), env.shape[0], bvh.handle if bvh else None, _p(mat, _f32p),
                                      _p(c, _f64p), c.shape[0], float(rx_radius), _p(tx, _f32p), int(max_bounces),
                                      int(tid_begin), int(n), int(n_total), float(carrier_hz), float(light_speed),
                                      float(sample_rate), int(n_bins), _p(field, _f64p), _p(ir, _f64p), ctypes.byref(arr))
    return dict(field=field[:, 0] + 1j * field[:, 1], ir=None if ir is None else ir[..., 0] + 1j * ir[..., 1],
                segments=int(seg), arrivals=int(arr.value))


def max_threads():
    return int(lib().oracle_max_threads())


class triangle_test:
    """Context manager: which functor the restated mesh_query_ray runs — "woop" (the reference's watertight test,
    default) or "mt" (Moeller-Trumbore, rfrt_oracle.c mt_tri)."""
    KINDS = {"woop": 0, "mt": 1}

    def __init__(self, kind):
        self.kind = self.KINDS[kind]

    def __enter__(self):
        self.prev = int(lib().oracle_get_triangle_test())
        lib().oracle_set_triangle_test(self.kind)
        return self

    def __exit__(self, *exc):
        lib().oracle_set_triangle_test(self.prev)
        return False


def sphere_hits(tid_begin, n, tx, center, radius, cap=1 << 20, nthreads=0):
    """Ray ids whose generated direction hits the analytic sphere (KAT-1 helper)."""
    tx = np.ascontiguousarray(np.asarray(tx, dtype=np.float64)); center = np.ascontiguousarray(np.asarray(center, dtype=np.float64))
    out = np.empty(cap, dtype=np.int64)
    m = lib().oracle_sphere_hits(int(tid_begin), int(n), _p(tx, _f64p), _p(center, _f64p), float(radius),
                                 _p(out, _i64p), int(cap), int(nthreads))
    if m > cap:
        raise RuntimeError("sphere_hits: cap too small")
    return out[:m].copy()
