"""Oracle post-processing: tracer.py:84-117 and main.py:12-13,39,46-55 restated in NumPy.

TEST INFRASTRUCTURE ONLY.  dtype semantics are written out explicitly and follow the NumPy 1.x
promotion rules of the reference's era (python float (+) np.float32 scalar -> float64), so the
result does not depend on the NumPy version installed:
  * segment vectors, their lengths, the dot product and the arccos are fp32 (tracer.py:107-110);
    the 3-term sums are evaluated left to right without FMA: (x*x + y*y) + z*z; the fp32 arccos is the
    correctly rounded one (fp64 acos rounded to fp32)
  * the Fresnel chain is fp64 `math` (tracer.py:34-61), fed with the fp32 angle
  * distance accumulates fp32 lengths in fp64 (tracer.py:104,112-113)
The debug prints (tracer.py:36,41,46,56,59) are omitted.
"""
import math

import numpy as np

_f32 = np.float32


def clean_paths(received, mask):
    """tracer.py:84-97 — mask gather (ascending ray id) + strip at the first vertex containing a NaN."""
    rows = received[np.asarray(mask) != 0]
    cleaned = []
    for row in rows:
        k = 0
        while k < row.shape[0] and not np.isnan(row[k]).any():
            k += 1
        cleaned.append(np.array(row[:k], dtype=np.float32))
    return cleaned


def _norm32(v):
    return _f32(np.sqrt(_f32(_f32(_f32(v[0] * v[0]) + _f32(v[1] * v[1])) + _f32(v[2] * v[2]))))


def _dot32(a, b):
    return _f32(_f32(_f32(a[0] * b[0]) + _f32(a[1] * b[1])) + _f32(a[2] * b[2]))


def bounce_amplitude(angle_between, n_1=5.0):
    """tracer.py:34-61 (p-polarised Fresnel power reflectance, n1=5 -> n2=1, clamp, NaN -> 0).
    n_1: material-table extension; the reference hard-codes 5.0."""
    angle_between = float(angle_between)
    if math.isnan(angle_between):
        return 0.0
    theta = (math.pi / 2) - (angle_between / 2)
    n_1 = float(n_1)
    n_2 = 1.0
    theta_i = math.asin((n_2 * math.sin(theta)) / n_1)
    num = n_2 * math.cos(theta_i) - n_1 * math.cos(theta)
    denom = n_2 * math.cos(theta_i) + n_1 * math.cos(theta)
    amp = -(num / denom) ** 2
    if amp < -1:
        amp = -1
    if math.isnan(amp):
        return 0.0
    return -amp


def path_amplitude_delay(path, tx_power, tx_num_rays, light_speed_mps, sample_rate_hz, vertex_n=None):
    """tracer.py:103-115 for one cleaned path -> (amplitude fp64, distance fp64, delay_samples int).
    vertex_n: optional refractive index per path vertex (material table; default 5.0 = tracer.py:43)."""
    path = np.asarray(path, dtype=np.float32)
    amplitude = tx_power / tx_num_rays
    distance = 0.0
    for k, (p1, p2, p3) in enumerate(zip(path[:-2], path[1:-1], path[2:])):
        seg1 = (p2 - p1).astype(np.float32)
        seg2 = (p3 - p2).astype(np.float32)
        seg1_len = _norm32(seg1)
        with np.errstate(invalid="ignore", divide="ignore"):
            q = _f32(_dot32(seg1, seg2) / _f32(seg1_len * _norm32(seg2)))
        # np.arccos on a float32 scalar (tracer.py:110): NaN if q > 1, q < -1 or 0/0.  NumPy's fp32 arccos is
        # SIMD-dispatch dependent (up to a few ulp), so the oracle pins it to the correctly rounded fp32 value.
        if np.isnan(q) or q > 1 or q < -1:
            angle_between = _f32(np.nan)
        else:
            angle_between = _f32(math.acos(float(q)))
        amplitude *= bounce_amplitude(angle_between, 5.0 if vertex_n is None else vertex_n[k + 1])
        distance += float(seg1_len)
    distance += float(_norm32((path[-2] - path[-1]).astype(np.float32)))
    delay_samples = int((distance / light_speed_mps) * sample_rate_hz)
    return amplitude, distance, delay_samples


def impulse_response(cleaned_paths, tx_power, tx_num_rays, light_speed_mps, sample_rate_hz, sample_window_s,
                     vertex_n=None):
    """tracer.py:101-117.  vertex_n: optional list (one array per path) of per-vertex refractive indices."""
    ir = np.zeros(int(sample_window_s * sample_rate_hz))
    for i, path in enumerate(cleaned_paths):
        amp, _, d = path_amplitude_delay(path, tx_power, tx_num_rays, light_speed_mps, sample_rate_hz,
                                         None if vertex_n is None else vertex_n[i])
        if d < ir.shape[0]:
            ir[d] += amp
    return ir


def to_dbm(power):
    """main.py:12-13"""
    with np.errstate(divide="ignore", invalid="ignore"):
        return 10 * np.log10(power / 1e-3)


def rx_power(ir, sample_window_s, carrier_hz=2.4e9):
    """main.py:39,46-55 / coverage.py:45-55 -> mean-square power (linear); NaN when nothing is non-zero."""
    time = np.linspace(0, sample_window_s, ir.shape[0])
    signal_tx = np.sin(2 * np.pi * carrier_hz * time)
    signal_rx = np.convolve(ir, signal_tx, mode="same")
    r = np.nonzero(signal_rx)[:10000]
    signal_rx = signal_rx[r]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sum(signal_rx ** 2) / np.float64(signal_rx.shape[0])
