// rfrt_math.cuh — device arithmetic of the hot path (sm_100a).
//
// Everything here that decides WHICH triangle a ray hits is written as an explicit sequence of IEEE
// fp32 operations through the __f*_rn intrinsics, which nvcc never contracts (the library is built with the default
// -fmad=true: pruning code such as the slab tests and the candidate filters may fuse freely, parity-critical code
// may not, so it never uses plain * and +); the only fused parity-critical operations are the __fmaf_rn calls
// written out below.  The sequence restates:
//   kernel.py:51-52  wp.rand_init / wp.sample_unit_sphere_surface   (PCG hash, 24-bit randf)
//   kernel.py:71,82  wp.mesh_query_ray -> intersect_ray_tri_woop    (watertight test)
//   kernel.py:6-8    reflect
// of /root/reference (warp-lang itself is third-party and not vendored there).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rfrt {

// ---------------------------------------------------------------------------------------------
// RNG
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pcg_hash(uint32_t s)
{
    uint32_t b = s * 747796405u + 2891336453u;
    uint32_t c = ((b >> ((b >> 28) + 4u)) ^ b) * 277803737u;
    return (c >> 22) ^ c;
}

__device__ __forceinline__ float randf(uint32_t &state)
{
    state = pcg_hash(state);
    return __uint2float_rn(state >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------------------------
// Deterministic fp64 sin / cos / acos: IEEE + - * / sqrt / floor / fma only, Horner in a fixed order.
// Identical bits on any IEEE machine (no libm / libdevice involved).
// ---------------------------------------------------------------------------------------------
// polynomial coefficients live in the constant bank so that each Horner step is ONE DFMA with a constant operand
// (as immediates every fp64 constant costs two extra UMOV issue slots)
__constant__ double c_det_sin[8] = {-0.16666666666666666, 0.008333333333333333, -0.0001984126984126984, 2.7557319223985893e-06,
                                    -2.505210838544172e-08, 1.6059043836821613e-10, -7.647163731819816e-13,
                                    2.8114572543455206e-15};
__constant__ double c_det_cos[9] = {-0.5, 0.041666666666666664, -0.001388888888888889, 2.48015873015873e-05,
                                    -2.755731922398589e-07, 2.08767569878681e-09, -1.1470745597729725e-11,
                                    4.779477332387385e-14, -1.5619206968586225e-16};
__constant__ double c_det_asin[29] = {
    1.0, 0.16666666666666666, 0.075, 0.044642857142857144, 0.030381944444444444, 0.022372159090909092,
    0.017352764423076924, 0.01396484375, 0.011551800896139705, 0.009761609529194078, 0.008390335809616815,
    0.0073125258735988454, 0.006447210311889649, 0.005740037670841924, 0.005153309682319905,
    0.004660143486915096, 0.004240907093679363, 0.003880964558837669, 0.0035692053938259347,
    0.003297059503473485, 0.0030578216492580306, 0.002846178401108942, 0.00265787063820729,
    0.0024894486782468836, 0.002338091892111975, 0.0022014739737101384, 0.0020776610325181676,
    0.0019650336162772837, 0.0018622264064031275};

__device__ __forceinline__ void det_sincos(double x, double &s_out, double &c_out)
{
    const double TWO_OVER_PI = 0.6366197723675814;
    const double PIO2_1 = 1.5707963267341256;     // first 33 bits of pi/2
    const double PIO2_1T = 6.077100506506192e-11; // pi/2 - PIO2_1
    double kf = floor(__dadd_rn(__dmul_rn(x, TWO_OVER_PI), 0.5));
    int k = (int)kf;
    double r = __fma_rn(-kf, PIO2_1T, __fma_rn(-kf, PIO2_1, x));
    double w = __dmul_rn(r, r);
    double ps = c_det_sin[7];
#pragma unroll
    for (int j = 6; j >= 0; --j) ps = __fma_rn(w, ps, c_det_sin[j]);
    double sn = __fma_rn(__dmul_rn(r, w), ps, r);
    double pc = c_det_cos[8];
#pragma unroll
    for (int j = 7; j >= 0; --j) pc = __fma_rn(w, pc, c_det_cos[j]);
    double cs = __fma_rn(w, pc, 1.0);
    switch (k & 3) {
    case 0: s_out = sn; c_out = cs; break;
    case 1: s_out = cs; c_out = -sn; break;
    case 2: s_out = -sn; c_out = -cs; break;
    default: s_out = -cs; c_out = sn; break;
    }
}

__device__ __forceinline__ double det_asin_small(double x)
{
    double w = __dmul_rn(x, x);
    double p = c_det_asin[28];
#pragma unroll
    for (int j = 27; j >= 0; --j) p = __fma_rn(w, p, c_det_asin[j]);
    return __dmul_rn(x, p);
}

// One series evaluation for all three ranges: written as three branches the 28-step Horner chain was executed three
// times by every warp (the lanes of a warp fall into all ranges).  Per lane the operations and their operands are the
// ones of the branchy form — 1 + z for z < -0.5 and 1 - z for z > 0.5 are both 1 - |z|, the same IEEE subtraction.
__device__ __forceinline__ double det_acos(double z)
{
    const double PI = 3.141592653589793;
    const double PIO2 = 1.5707963267948966;
    const double az = fabs(z);
    const double x = az > 0.5 ? __dsqrt_rn(__dmul_rn(__dsub_rn(1.0, az), 0.5)) : z;
    const double p = det_asin_small(x);
    if (z > 0.5) return __dmul_rn(2.0, p);
    if (z < -0.5) return __dsub_rn(PI, __dmul_rn(2.0, p));
    return __dsub_rn(PIO2, p);
}

// kernel.py:51-52
__device__ __forceinline__ float3 ray_direction(uint32_t tid)
{
    uint32_t state = pcg_hash(tid);
    float u1 = randf(state);
    float u2 = randf(state);
    float phi = __double2float_rn(det_acos(__dsub_rn(1.0, __dmul_rn(2.0, (double)u1))));
    float theta = __fmul_rn(6.2831854820251465f, u2);
    double s, c;
    det_sincos((double)phi, s, c);
    float sp = __double2float_rn(s), cp = __double2float_rn(c);
    det_sincos((double)theta, s, c);
    float st = __double2float_rn(s), ct = __double2float_rn(c);
    return make_float3(__fmul_rn(ct, sp), __fmul_rn(st, sp), cp);
}

// ---------------------------------------------------------------------------------------------
// Watertight ray/triangle test (Woop, Benthin, Wald 2013) in the operation order of warp's
// intersect_ray_tri_woop.  The per-ray part (dominant axis, shear) is hoisted out of the
// per-triangle part; the arithmetic performed per triangle is unchanged.
// ---------------------------------------------------------------------------------------------
struct WoopRay {
    int kx, ky, kz;
    float Sx, Sy, Sz;
    float px, py, pz;    // ray origin
    float pkx, pky, pkz; // ... permuted (p[kx], p[ky], p[kz])
};

__device__ __forceinline__ float sel3(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

__device__ __forceinline__ WoopRay woop_setup(float3 p, float3 d)
{
    WoopRay r;
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    // kz = dominant axis (x only if strictly greater than both, else y if |y| > |z|, else z); kx, ky follow cyclically
    const bool k0 = ax > ay && ax > az;
    const bool k1 = !k0 && ay > az;
    const float dkz = k0 ? d.x : (k1 ? d.y : d.z);
    float dkx = k0 ? d.y : (k1 ? d.z : d.x), dky = k0 ? d.z : (k1 ? d.x : d.y);
    float pkx = k0 ? p.y : (k1 ? p.z : p.x), pky = k0 ? p.z : (k1 ? p.x : p.y);
    int kx = k0 ? 1 : (k1 ? 2 : 0), ky = k0 ? 2 : (k1 ? 0 : 1);
    if (dkz < 0.0f) { // swap kx and ky to preserve the winding
        float tf = dkx; dkx = dky; dky = tf;
        tf = pkx; pkx = pky; pky = tf;
        int ti = kx; kx = ky; ky = ti;
    }
    r.kx = kx; r.ky = ky; r.kz = k0 ? 0 : (k1 ? 1 : 2);
    r.Sx = __fdiv_rn(dkx, dkz);
    r.Sy = __fdiv_rn(dky, dkz);
    r.Sz = __fdiv_rn(1.0f, dkz);
    r.px = p.x; r.py = p.y; r.pz = p.z;
    r.pkx = pkx; r.pky = pky; r.pkz = k0 ? p.x : (k1 ? p.y : p.z);
    return r;
}

__device__ __forceinline__ float diff_product(float a, float b, float c, float d)
{
    float cd = __fmul_rn(c, d);
    float diff = __fmaf_rn(a, b, -cd);
    float err = __fmaf_rn(-c, d, cd);
    return __fadd_rn(diff, err);
}

// The test proper, on the vertex offsets already permuted to (kx, ky, kz).  Returns true and t when the triangle is
// hit (any sign of t that the reference accepts here: callers apply  t >= 0 && t < best).
__device__ __forceinline__ bool woop_hit_permuted(const WoopRay &r, float Akx, float Aky, float Akz, float Bkx, float Bky,
                                                  float Bkz, float Ckx, float Cky, float Ckz, float &t_out)
{
    float Ax = __fsub_rn(Akx, __fmul_rn(r.Sx, Akz));
    float Ay = __fsub_rn(Aky, __fmul_rn(r.Sy, Akz));
    float Bx = __fsub_rn(Bkx, __fmul_rn(r.Sx, Bkz));
    float By = __fsub_rn(Bky, __fmul_rn(r.Sy, Bkz));
    float Cx = __fsub_rn(Ckx, __fmul_rn(r.Sx, Ckz));
    float Cy = __fsub_rn(Cky, __fmul_rn(r.Sy, Ckz));

    float U = diff_product(Cx, By, Cy, Bx);
    float V = diff_product(Ax, Cy, Ay, Cx);
    float W = diff_product(Bx, Ay, By, Ax);

    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        // products of two floats are exact in fp64; one rounding for the difference, one to fp32
        double CxBy = __dmul_rn((double)Cx, (double)By);
        double CyBx = __dmul_rn((double)Cy, (double)Bx);
        U = __double2float_rn(__dsub_rn(CxBy, CyBx));
        double AxCy = __dmul_rn((double)Ax, (double)Cy);
        double AyCx = __dmul_rn((double)Ay, (double)Cx);
        V = __double2float_rn(__dsub_rn(AxCy, AyCx));
        double BxAy = __dmul_rn((double)Bx, (double)Ay);
        double ByAx = __dmul_rn((double)By, (double)Ax);
        W = __double2float_rn(__dsub_rn(BxAy, ByAx));
    }

    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;

    float det = __fadd_rn(__fadd_rn(U, V), W);
    if (det == 0.0f) return false;

    float Az = __fmul_rn(r.Sz, Akz);
    float Bz = __fmul_rn(r.Sz, Bkz);
    float Cz = __fmul_rn(r.Sz, Ckz);
    float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));

    // xorf(T, sign_mask(det)) < 0 -> reject (so T == +-0 and NaN pass)
    float x = __uint_as_float(__float_as_uint(T) ^ (__float_as_uint(det) & 0x80000000u));
    if (x < 0.0f) return false;

    float rcp_det = __fdiv_rn(1.0f, det);
    t_out = __fmul_rn(T, rcp_det);
    return true;
}

__device__ __forceinline__ bool woop_hit(const WoopRay &r, float3 a, float3 b, float3 c, float &t_out)
{
    float A0 = __fsub_rn(a.x, r.px), A1 = __fsub_rn(a.y, r.py), A2 = __fsub_rn(a.z, r.pz);
    float B0 = __fsub_rn(b.x, r.px), B1 = __fsub_rn(b.y, r.py), B2 = __fsub_rn(b.z, r.pz);
    float C0 = __fsub_rn(c.x, r.px), C1 = __fsub_rn(c.y, r.py), C2 = __fsub_rn(c.z, r.pz);
    return woop_hit_permuted(r, sel3(A0, A1, A2, r.kx), sel3(A0, A1, A2, r.ky), sel3(A0, A1, A2, r.kz),
                             sel3(B0, B1, B2, r.kx), sel3(B0, B1, B2, r.ky), sel3(B0, B1, B2, r.kz),
                             sel3(C0, C1, C2, r.kx), sel3(C0, C1, C2, r.ky), sel3(C0, C1, C2, r.kz), t_out);
}

// same test with the vertices in memory (3 consecutive floats each): the permuted components are fetched directly
// (v[k] - p[k] is the same operation whether the permutation is applied before or after the subtraction)
__device__ __forceinline__ bool woop_hit_mem(const WoopRay &r, const float *a, const float *b, const float *c, float &t_out)
{
    return woop_hit_permuted(r, __fsub_rn(__ldg(a + r.kx), r.pkx), __fsub_rn(__ldg(a + r.ky), r.pky), __fsub_rn(__ldg(a + r.kz), r.pkz),
                             __fsub_rn(__ldg(b + r.kx), r.pkx), __fsub_rn(__ldg(b + r.ky), r.pky), __fsub_rn(__ldg(b + r.kz), r.pkz),
                             __fsub_rn(__ldg(c + r.kx), r.pkx), __fsub_rn(__ldg(c + r.ky), r.pky), __fsub_rn(__ldg(c + r.kz), r.pkz), t_out);
}

// ---------------------------------------------------------------------------------------------
// Moeller-Trumbore functor (BASELINE.json north_star (2); SURVEY.md 7.3 "ship both functors"; selected per mesh with
// rfrt_mesh_set_triangle_test).  NOT the reference's arithmetic — Warp's mesh_query_ray runs the watertight test
// above — but the classic two-sided test without an epsilon (det == 0 rejects), every operation a separate fp32
// rounding, dot products summed left to right (the CPU restatement under tests/ runs the same sequence).
// ---------------------------------------------------------------------------------------------
struct MtRay {
    float px, py, pz, dx, dy, dz;
};

__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

__device__ __forceinline__ bool mt_hit(const MtRay &r, float3 a, float3 b, float3 c, float &t_out)
{
    const float e1x = __fsub_rn(b.x, a.x), e1y = __fsub_rn(b.y, a.y), e1z = __fsub_rn(b.z, a.z);
    const float e2x = __fsub_rn(c.x, a.x), e2y = __fsub_rn(c.y, a.y), e2z = __fsub_rn(c.z, a.z);
    const float pvx = __fsub_rn(__fmul_rn(r.dy, e2z), __fmul_rn(r.dz, e2y));
    const float pvy = __fsub_rn(__fmul_rn(r.dz, e2x), __fmul_rn(r.dx, e2z));
    const float pvz = __fsub_rn(__fmul_rn(r.dx, e2y), __fmul_rn(r.dy, e2x));
    const float det = dot3_rn(e1x, e1y, e1z, pvx, pvy, pvz);
    if (det == 0.0f || det != det) return false;
    const float inv_det = __fdiv_rn(1.0f, det);
    const float tvx = __fsub_rn(r.px, a.x), tvy = __fsub_rn(r.py, a.y), tvz = __fsub_rn(r.pz, a.z);
    const float u = __fmul_rn(dot3_rn(tvx, tvy, tvz, pvx, pvy, pvz), inv_det);
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const float qvx = __fsub_rn(__fmul_rn(tvy, e1z), __fmul_rn(tvz, e1y));
    const float qvy = __fsub_rn(__fmul_rn(tvz, e1x), __fmul_rn(tvx, e1z));
    const float qvz = __fsub_rn(__fmul_rn(tvx, e1y), __fmul_rn(tvy, e1x));
    const float v = __fmul_rn(dot3_rn(r.dx, r.dy, r.dz, qvx, qvy, qvz), inv_det);
    if (!(v >= 0.0f && __fadd_rn(u, v) <= 1.0f)) return false;
    t_out = __fmul_rn(dot3_rn(e2x, e2y, e2z, qvx, qvy, qvz), inv_det);
    return true;
}

// The functor is chosen by the type of the per-ray constants: WoopRay (watertight, reference-faithful) or MtRay.
template <class RAY> __device__ __forceinline__ RAY tri_ray_setup(float3 p, float3 d);
template <> __device__ __forceinline__ WoopRay tri_ray_setup<WoopRay>(float3 p, float3 d) { return woop_setup(p, d); }
template <> __device__ __forceinline__ MtRay tri_ray_setup<MtRay>(float3 p, float3 d)
{
    MtRay r;
    r.px = p.x; r.py = p.y; r.pz = p.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
    return r;
}
__device__ __forceinline__ bool tri_hit(const WoopRay &r, float3 a, float3 b, float3 c, float &t) { return woop_hit(r, a, b, c, t); }
__device__ __forceinline__ bool tri_hit(const MtRay &r, float3 a, float3 b, float3 c, float &t) { return mt_hit(r, a, b, c, t); }

// normalize(cross(b-a, c-a)); zero vector when degenerate  (normal returned by mesh_query_ray)
__device__ __forceinline__ float3 tri_normal(float3 a, float3 b, float3 c)
{
    float abx = __fsub_rn(b.x, a.x), aby = __fsub_rn(b.y, a.y), abz = __fsub_rn(b.z, a.z);
    float acx = __fsub_rn(c.x, a.x), acy = __fsub_rn(c.y, a.y), acz = __fsub_rn(c.z, a.z);
    float cx = __fsub_rn(__fmul_rn(aby, acz), __fmul_rn(abz, acy));
    float cy = __fsub_rn(__fmul_rn(abz, acx), __fmul_rn(abx, acz));
    float cz = __fsub_rn(__fmul_rn(abx, acy), __fmul_rn(aby, acx));
    float l = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
    if (l > 0.0f) return make_float3(__fdiv_rn(cx, l), __fdiv_rn(cy, l), __fdiv_rn(cz, l));
    return make_float3(0.0f, 0.0f, 0.0f);
}

// kernel.py:6-8   v - 2*dot(v,n)*n
__device__ __forceinline__ float3 reflect(float3 v, float3 n)
{
    float d = __fadd_rn(__fadd_rn(__fmul_rn(v.x, n.x), __fmul_rn(v.y, n.y)), __fmul_rn(v.z, n.z));
    float s = __fmul_rn(2.0f, d);
    return make_float3(__fsub_rn(v.x, __fmul_rn(s, n.x)), __fsub_rn(v.y, __fmul_rn(s, n.y)),
                       __fsub_rn(v.z, __fmul_rn(s, n.z)));
}

// pos + dir*t   (kernel.py:87,94)
__device__ __forceinline__ float3 advance(float3 p, float3 d, float t)
{
    return make_float3(__fadd_rn(p.x, __fmul_rn(d.x, t)), __fadd_rn(p.y, __fmul_rn(d.y, t)),
                       __fadd_rn(p.z, __fmul_rn(d.z, t)));
}

} // namespace rfrt
