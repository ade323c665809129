/*
 * rfrt_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's hot path so that the CUDA path can be
 * checked bit-for-bit.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (rf_ray_tracing_warp_b200/) never links, imports or calls it.
 *
 * What it restates (all citations are into /root/reference):
 *   kernel.py:6-8     reflect(v, n) = v - 2*dot(v,n)*n
 *   kernel.py:38-98   trace_paths_kernel: per ray seed RNG from the thread id, pick a
 *                     direction, loop max_bounces times { RX query, env query, arbitrate,
 *                     record vertex, reflect }, incl. the `ray_finished` reset quirk (:58)
 *   tracer.py:26-30   RX mesh = icosphere(subdivisions=1) (vertex table built in oracle.py)
 *   tracer.py:67-72   NaN-filled (N,B+1,3) path buffers, zeroed row mask
 * and the third-party pieces those lines call, which are NOT in /root/reference
 * (warp-lang, version unpinned, README.md:8).  Their published algorithms are restated:
 *   wp.rand_init / randf / sample_unit_sphere_surface  -> PCG hash, 24-bit randf
 *                     (constants + draw order confirmed by the KAT-1 golden vector,
 *                     web/scene.html, see tests/golden/)
 *   wp.mesh_query_ray -> closest hit, 0 <= t < max_t, double sided, strict '<'
 *   intersect_ray_tri_woop -> Woop/Benthin/Wald watertight test, fp32, with the
 *                     fmaf-based diff_product and the fp64 fallback when U,V or W == 0
 *
 * PARITY STATUS: direction sampling and the received-ray set are pinned by KAT-1.
 * Everything beyond it (closest-hit t, tie order, FMA policy, libm last bits) is
 * "parity unpinned" against real Warp — warp-lang cannot be installed here — so this
 * file DEFINES those bits:
 *   - IEEE fp32, no FMA contraction (compile with -ffp-contract=off); the only fused
 *     operations are the explicit fmaf() calls inside diff_product.
 *   - Equal-t ties resolve to the LOWEST triangle index (brute force in index order
 *     with strict '<'), which makes the answer independent of any BVH topology.
 *   - sin/cos/acos are evaluated by the deterministic fp64 routines below (IEEE
 *     + - * / sqrt only) and rounded to fp32, so CPU and GPU agree bit-for-bit; they
 *     are within 1 ulp of libm's sinf/cosf/acosf.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -mfma).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * RNG — warp rand.h [restated; constants confirmed by KAT-1], called at kernel.py:51-52
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t pcg_hash(uint32_t s)
{
    uint32_t b = s * 747796405u + 2891336453u;
    uint32_t c = ((b >> ((b >> 28) + 4u)) ^ b) * 277803737u;
    return (c >> 22) ^ c;
}

static inline float randf(uint32_t *state)
{
    *state = pcg_hash(*state);
    return (float)(*state >> 8) * (1.0f / 16777216.0f);
}

/* ------------------------------------------------------------------------------------------
 * Deterministic fp64 sin / cos / acos (IEEE add/sub/mul/div/sqrt/floor only).
 * Taylor coefficients are exact rationals rounded once to double.
 * ---------------------------------------------------------------------------------------- */
static const double DM_TWO_OVER_PI = 0.6366197723675814;
static const double DM_PIO2_1 = 1.5707963267341256;     /* first 33 bits of pi/2 */
static const double DM_PIO2_1T = 6.077100506506192e-11; /* pi/2 - DM_PIO2_1 */
static const double DM_PI = 3.141592653589793;
static const double DM_PIO2 = 1.5707963267948966;

static const double DM_S[8] = {
    -0.16666666666666666, 0.008333333333333333, -0.0001984126984126984, 2.7557319223985893e-06,
    -2.505210838544172e-08, 1.6059043836821613e-10, -7.647163731819816e-13, 2.8114572543455206e-15};
static const double DM_C[9] = {
    -0.5, 0.041666666666666664, -0.001388888888888889, 2.48015873015873e-05, -2.755731922398589e-07,
    2.08767569878681e-09, -1.1470745597729725e-11, 4.779477332387385e-14, -1.5619206968586225e-16};
static const double DM_A[29] = {
    1.0, 0.16666666666666666, 0.075, 0.044642857142857144, 0.030381944444444444, 0.022372159090909092,
    0.017352764423076924, 0.01396484375, 0.011551800896139705, 0.009761609529194078, 0.008390335809616815,
    0.0073125258735988454, 0.006447210311889649, 0.005740037670841924, 0.005153309682319905,
    0.004660143486915096, 0.004240907093679363, 0.003880964558837669, 0.0035692053938259347,
    0.003297059503473485, 0.0030578216492580306, 0.002846178401108942, 0.00265787063820729,
    0.0024894486782468836, 0.002338091892111975, 0.0022014739737101384, 0.0020776610325181676,
    0.0019650336162772837, 0.0018622264064031275};

/* x >= 0, x < ~8.  Writes sin(x), cos(x). */
static void det_sincos(double x, double *s_out, double *c_out)
{
    double kf = floor(x * DM_TWO_OVER_PI + 0.5);
    int k = (int)kf;
    /* every step is one IEEE fused multiply-add (fma() here, DFMA on the GPU): identical bits on both sides */
    double r = fma(-kf, DM_PIO2_1T, fma(-kf, DM_PIO2_1, x));
    double w = r * r;
    double ps = DM_S[7];
    for (int j = 6; j >= 0; --j) ps = fma(w, ps, DM_S[j]);
    double sn = fma(r * w, ps, r);
    double pc = DM_C[8];
    for (int j = 7; j >= 0; --j) pc = fma(w, pc, DM_C[j]);
    double cs = fma(w, pc, 1.0);
    switch (k & 3) {
    case 0: *s_out = sn; *c_out = cs; break;
    case 1: *s_out = cs; *c_out = -sn; break;
    case 2: *s_out = -sn; *c_out = -cs; break;
    default: *s_out = -cs; *c_out = sn; break;
    }
}

/* asin(x) for |x| <= 0.5 */
static double det_asin_small(double x)
{
    double w = x * x;
    double p = DM_A[28];
    for (int j = 27; j >= 0; --j) p = fma(w, p, DM_A[j]);
    return x * p;
}

/* acos(z), -1 <= z <= 1 */
static double det_acos(double z)
{
    if (z > 0.5) return 2.0 * det_asin_small(sqrt((1.0 - z) * 0.5));
    if (z < -0.5) return DM_PI - 2.0 * det_asin_small(sqrt((1.0 + z) * 0.5));
    return DM_PIO2 - det_asin_small(z);
}

/* fp32 wrappers: evaluate in fp64, round once */
static inline float det_sinf(float x) { double s, c; det_sincos((double)x, &s, &c); return (float)s; }
static inline float det_cosf(float x) { double s, c; det_sincos((double)x, &s, &c); return (float)c; }

/* warp sample_unit_sphere_surface [restated], called at kernel.py:52:
 *   phi   = acos(1.0 - 2.0*randf)   (double literals -> fp64, rounded to float)
 *   theta = randf(state, 0, 2*pi)   = float(2*pi) * randf  (fp32)
 *   dir   = (cos(theta)*sin(phi), sin(theta)*sin(phi), cos(phi))          */
static void ray_direction(uint32_t tid, float dir[3])
{
    uint32_t state = pcg_hash(tid); /* rand_init(tid), kernel.py:51 */
    float u1 = randf(&state);
    float u2 = randf(&state);
    float phi = (float)det_acos(1.0 - 2.0 * (double)u1);
    float theta = 6.2831854820251465f * u2;
    float sp = det_sinf(phi), cp = det_cosf(phi);
    float st = det_sinf(theta), ct = det_cosf(theta);
    dir[0] = ct * sp;
    dir[1] = st * sp;
    dir[2] = cp;
}

/* ------------------------------------------------------------------------------------------
 * intersect_ray_tri_woop [restated from warp intersect.h]; called via mesh_query_ray
 * at kernel.py:71 and kernel.py:82.
 * ---------------------------------------------------------------------------------------- */
static inline float diff_product(float a, float b, float c, float d)
{
    float cd = c * d;
    float diff = fmaf(a, b, -cd);
    float err = fmaf(-c, d, cd);
    return diff + err;
}

static inline int sign_bit(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    return (int)(u >> 31);
}

/* The per-ray part of the test (dominant axis, shear constants) is hoisted out of the per-triangle
 * part — the arithmetic and its order are exactly those of intersect_ray_tri_woop, evaluated once per
 * query instead of once per triangle (what an optimising compiler does after inlining). */
typedef struct {
    int kx, ky, kz;
    float Sx, Sy, Sz;
    float p[3];
    float d[3]; /* the direction itself (Moeller-Trumbore functor) */
} WoopRay;

static inline void woop_setup(const float p[3], const float dir[3], WoopRay *r)
{
    float ax = fabsf(dir[0]), ay = fabsf(dir[1]), az = fabsf(dir[2]);
    int kz;
    if (ax > ay && ax > az) kz = 0;
    else if (ay > az) kz = 1;
    else kz = 2;
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    if (dir[kz] < 0.0f) { int tmp = kx; kx = ky; ky = tmp; }
    r->kx = kx; r->ky = ky; r->kz = kz;
    r->Sx = dir[kx] / dir[kz];
    r->Sy = dir[ky] / dir[kz];
    r->Sz = 1.0f / dir[kz];
    r->p[0] = p[0]; r->p[1] = p[1]; r->p[2] = p[2];
    r->d[0] = dir[0]; r->d[1] = dir[1]; r->d[2] = dir[2];
}

static inline int woop_tri_exact(const WoopRay *r, const float a[3], const float b[3], const float c[3], float *t_out)
{
    const int kx = r->kx, ky = r->ky, kz = r->kz;
    const float Sx = r->Sx, Sy = r->Sy, Sz = r->Sz;
    const float *p = r->p;

    float A[3] = {a[0] - p[0], a[1] - p[1], a[2] - p[2]};
    float B[3] = {b[0] - p[0], b[1] - p[1], b[2] - p[2]};
    float C[3] = {c[0] - p[0], c[1] - p[1], c[2] - p[2]};

    float Ax = A[kx] - Sx * A[kz];
    float Ay = A[ky] - Sy * A[kz];
    float Bx = B[kx] - Sx * B[kz];
    float By = B[ky] - Sy * B[kz];
    float Cx = C[kx] - Sx * C[kz];
    float Cy = C[ky] - Sy * C[kz];

    float U = diff_product(Cx, By, Cy, Bx);
    float V = diff_product(Ax, Cy, Ay, Cx);
    float W = diff_product(Bx, Ay, By, Ax);

    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By;
        double CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy;
        double AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay;
        double ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }

    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;

    float det = U + V + W;
    if (det == 0.0f) return 0;

    float Az = Sz * A[kz];
    float Bz = Sz * B[kz];
    float Cz = Sz * C[kz];
    float T = U * Az + V * Bz + W * Cz;

    /* xorf(T, sign_mask(det)) < 0  -> reject;  +-0 and NaN fall through */
    {
        uint32_t tu, du;
        float x;
        memcpy(&tu, &T, 4);
        memcpy(&du, &det, 4);
        tu ^= (du & 0x80000000u);
        memcpy(&x, &tu, 4);
        if (x < 0.0f) return 0;
    }
    float rcp_det = 1.0f / det;
    *t_out = T * rcp_det;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Moeller-Trumbore functor (BASELINE.json north_star (2); SURVEY.md 7.3 "ship both functors").
 * NOT the reference's arithmetic (Warp's mesh_query_ray runs the watertight test above): the
 * classic two-sided test, no epsilon (det == 0 rejects), every operation a separate fp32
 * rounding, dot products summed left to right.  Callers apply 0 <= t < best as for the
 * watertight test.  The CUDA side (rfrt_math.cuh mt_hit) runs the same sequence.
 * ---------------------------------------------------------------------------------------- */
static inline int mt_tri(const WoopRay *r, const float a[3], const float b[3], const float c[3], float *t_out)
{
    const float *p = r->p, *d = r->d;
    float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    float e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    float pv[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
    float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
    if (det == 0.0f || det != det) return 0;
    float inv_det = 1.0f / det;
    float tv[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
    float u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv_det;
    if (!(u >= 0.0f && u <= 1.0f)) return 0;
    float qv[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
    float v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv_det;
    if (!(v >= 0.0f && u + v <= 1.0f)) return 0;
    *t_out = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv_det;
    return 1;
}

/* which functor mesh_query_ray runs: 0 = watertight (reference-faithful, default), 1 = Moeller-Trumbore */
static int g_tri_test = 0;
ORACLE_API void oracle_set_triangle_test(int kind) { g_tri_test = kind; }
ORACLE_API int oracle_get_triangle_test(void) { return g_tri_test; }

static inline int woop_tri(const WoopRay *r, const float a[3], const float b[3], const float c[3], float *t_out)
{
    return g_tri_test ? mt_tri(r, a, b, c, t_out) : woop_tri_exact(r, a, b, c, t_out);
}

static inline int woop(const float p[3], const float dir[3], const float a[3], const float b[3], const float c[3],
                       float *t_out)
{
    WoopRay r;
    woop_setup(p, dir, &r);
    return woop_tri(&r, a, b, c, t_out);
}

/* mesh_query_ray [restated]: closest accepted hit with 0 <= t < max_t.  Brute force in
 * index order with strict '<'  ==>  equal-t ties go to the lowest triangle index. */
static int query_closest_skip(const float *tris, int64_t ntris, const float p[3], const float dir[3], float max_t,
                              int skip, float *t_out, int *face_out)
{
    float min_t = max_t;
    int min_face = -1;
    WoopRay wr;
    woop_setup(p, dir, &wr);
    for (int64_t i = 0; i < ntris; ++i) {
        const float *a = tris + 9 * i;
        float t;
        if ((int)i != skip && woop_tri(&wr, a, a + 3, a + 6, &t)) {
            if (t < min_t && t >= 0.0f) {
                min_t = t;
                min_face = (int)i;
            }
        }
    }
    if (min_t < max_t) {
        *t_out = min_t;
        *face_out = min_face;
        return 1;
    }
    return 0;
}

/* normal returned by mesh_query_ray: normalize(cross(b-a, c-a)); normalize = v / length(v),
 * zero vector if the length is 0. */
static void tri_normal(const float *a, const float *b, const float *c, float n[3])
{
    float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    float ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    float cx = ab[1] * ac[2] - ab[2] * ac[1];
    float cy = ab[2] * ac[0] - ab[0] * ac[2];
    float cz = ab[0] * ac[1] - ab[1] * ac[0];
    float l = sqrtf(cx * cx + cy * cy + cz * cz);
    if (l > 0.0f) {
        n[0] = cx / l;
        n[1] = cy / l;
        n[2] = cz / l;
    } else {
        n[0] = n[1] = n[2] = 0.0f;
    }
}

/* kernel.py:6-8 */
static inline void reflect(float v[3], const float n[3])
{
    float d = v[0] * n[0] + v[1] * n[1] + v[2] * n[2];
    float s = 2.0f * d;
    v[0] = v[0] - s * n[0];
    v[1] = v[1] - s * n[1];
    v[2] = v[2] - s * n[2];
}

/* ------------------------------------------------------------------------------------------
 * A simple median-split BVH used ONLY to make the oracle usable on large meshes and to
 * make the CPU baseline a fair one (Warp's CPU device also traverses a BVH).  Its result
 * is defined to be identical to query_closest (tests check this): boxes are padded and
 * culling uses '<=' so equal-t candidates are never dropped; the hit rule is the same
 * lexicographic (t, index) minimum.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float lo[3], hi[3];
    int left, right; /* internal: child node ids; leaf: left = first prim slot, right = -count */
} ONode;

typedef struct {
    int64_t ntris;
    const float *tris; /* borrowed */
    ONode *nodes;
    int nnodes;
    int *prim; /* permutation */
} OBvh;

static void obvh_bounds(const OBvh *b, int first, int count, float lo[3], float hi[3])
{
    for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
    for (int i = first; i < first + count; ++i) {
        const float *t = b->tris + 9 * (int64_t)b->prim[i];
        for (int v = 0; v < 3; ++v)
            for (int k = 0; k < 3; ++k) {
                float x = t[3 * v + k];
                if (x < lo[k]) lo[k] = x;
                if (x > hi[k]) hi[k] = x;
            }
    }
}

static int cmp_axis;
static const float *cmp_tris;
static int cmp_centroid(const void *pa, const void *pb)
{
    int ia = *(const int *)pa, ib = *(const int *)pb;
    const float *ta = cmp_tris + 9 * (int64_t)ia, *tb = cmp_tris + 9 * (int64_t)ib;
    float ca = ta[cmp_axis] + ta[3 + cmp_axis] + ta[6 + cmp_axis];
    float cb = tb[cmp_axis] + tb[3 + cmp_axis] + tb[6 + cmp_axis];
    if (ca < cb) return -1;
    if (ca > cb) return 1;
    return (ia > ib) - (ia < ib);
}

static int obvh_build_rec(OBvh *b, int first, int count)
{
    int id = b->nnodes++;
    ONode *n = &b->nodes[id];
    obvh_bounds(b, first, count, n->lo, n->hi);
    if (count <= 2) {
        n->left = first;
        n->right = -count;
        return id;
    }
    int axis = 0;
    float ext = n->hi[0] - n->lo[0];
    for (int k = 1; k < 3; ++k)
        if (n->hi[k] - n->lo[k] > ext) { ext = n->hi[k] - n->lo[k]; axis = k; }
    cmp_axis = axis;
    cmp_tris = b->tris;
    qsort(b->prim + first, (size_t)count, sizeof(int), cmp_centroid);
    int half = count / 2;
    int l = obvh_build_rec(b, first, half);
    int r = obvh_build_rec(b, first + half, count - half);
    b->nodes[id].left = l;
    b->nodes[id].right = r;
    return id;
}

ORACLE_API void *oracle_bvh_create(const float *tris, int64_t ntris)
{
    OBvh *b = (OBvh *)calloc(1, sizeof(OBvh));
    b->ntris = ntris;
    b->tris = tris;
    b->nodes = (ONode *)malloc(sizeof(ONode) * (size_t)(2 * ntris + 1));
    b->prim = (int *)malloc(sizeof(int) * (size_t)(ntris > 0 ? ntris : 1));
    for (int64_t i = 0; i < ntris; ++i) b->prim[i] = (int)i;
    if (ntris > 0) obvh_build_rec(b, 0, (int)ntris);
    return b;
}

ORACLE_API void oracle_bvh_destroy(void *h)
{
    OBvh *b = (OBvh *)h;
    if (!b) return;
    free(b->nodes);
    free(b->prim);
    free(b);
}

static int obvh_query_skip(const OBvh *b, const float p[3], const float dir[3], float max_t, int skip, float *t_out,
                           int *face_out)
{
    if (b->ntris == 0) return 0;
    float pad = 1.0e-3f;
    {
        const ONode *r = &b->nodes[0];
        float m = 0.0f;
        for (int k = 0; k < 3; ++k) {
            if (fabsf(r->lo[k]) > m) m = fabsf(r->lo[k]);
            if (fabsf(r->hi[k]) > m) m = fabsf(r->hi[k]);
        }
        if (m * 1.0e-5f > pad) pad = m * 1.0e-5f;
    }
    /* zero direction components are tilted by 1e-18 so the slab arithmetic stays NaN free (pruning only) */
    float inv[3];
    for (int k = 0; k < 3; ++k) {
        float d = dir[k];
        if (fabsf(d) < 1.0e-18f) d = copysignf(1.0e-18f, d);
        inv[k] = 1.0f / d;
    }
    WoopRay wr;
    woop_setup(p, dir, &wr);
    float min_t = max_t;
    int min_face = -1;
    int stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const ONode *n = &b->nodes[stack[--sp]];
        float tmin = 0.0f, tmax = INFINITY;
        for (int k = 0; k < 3; ++k) {
            float t0 = ((n->lo[k] - pad) - p[k]) * inv[k];
            float t1 = ((n->hi[k] + pad) - p[k]) * inv[k];
            float tn = t0 < t1 ? t0 : t1, tf = t0 < t1 ? t1 : t0;
            tmin = tn > tmin ? tn : tmin;
            tmax = tf < tmax ? tf : tmax;
        }
        if (!(tmax * 1.0000004f >= tmin) || !(tmin <= min_t)) continue;
        if (n->right < 0) {
            for (int i = n->left; i < n->left - n->right; ++i) {
                int f = b->prim[i];
                const float *a = b->tris + 9 * (int64_t)f;
                float t;
                if (f != skip && woop_tri(&wr, a, a + 3, a + 6, &t)) {
                    if (t >= 0.0f && (t < min_t || (t == min_t && min_face >= 0 && f < min_face))) {
                        min_t = t;
                        min_face = f;
                    }
                }
            }
        } else {
            stack[sp++] = n->left;
            stack[sp++] = n->right;
        }
    }
    if (min_t < max_t) {
        *t_out = min_t;
        *face_out = min_face;
        return 1;
    }
    return 0;
}

static inline int query_skip(const float *tris, int64_t ntris, const OBvh *bvh, const float p[3], const float dir[3],
                             float max_t, int skip, float *t, int *face)
{
    if (bvh) return obvh_query_skip(bvh, p, dir, max_t, skip, t, face);
    return query_closest_skip(tris, ntris, p, dir, max_t, skip, t, face);
}

static inline int query(const float *tris, int64_t ntris, const OBvh *bvh, const float p[3], const float dir[3],
                        float max_t, float *t, int *face)
{
    return query_skip(tris, ntris, bvh, p, dir, max_t, -1, t, face);
}

/* ------------------------------------------------------------------------------------------
 * kernel.py:38-98 — one ray, one RX mesh.  traced/received rows are (B+1)*3 floats
 * (pre-filled with NaN by the caller, tracer.py:67-71).  hit_tri/hit_t/event rows are B
 * entries of extra instrumentation (not in the reference): event 0 = miss, 1 = env hit,
 * 2 = rx hit.
 * ---------------------------------------------------------------------------------------- */
static void trace_one(const float *env, int64_t nenv, const OBvh *env_bvh, const float *rx, int64_t nrx,
                      const OBvh *rx_bvh,
                      const float tx[3], int max_bounces, uint32_t tid, float *traced, float *received,
                      uint32_t *mask, int32_t *hit_tri, float *hit_t, int8_t *event)
{
    float dir[3], pos[3];
    ray_direction(tid, dir);
    pos[0] = tx[0]; pos[1] = tx[1]; pos[2] = tx[2];
    traced[0] = pos[0]; traced[1] = pos[1]; traced[2] = pos[2];

    for (int bounce = 0; bounce < max_bounces; ++bounce) {
        /* kernel.py:58 — ray_finished is reset every iteration, so nothing ever stops early */
        float t_rx = 0.0f, t_env = 0.0f;
        int f_rx = 0, f_env = 0;
        int maybe_hit_rx = nrx > 0 ? query(rx, nrx, rx_bvh, pos, dir, 1.0e6f, &t_rx, &f_rx) : 0; /* :71 */
        int maybe_hit_env = query(env, nenv, env_bvh, pos, dir, 1.0e6f, &t_env, &f_env);       /* :82 */
        int hit_recv = maybe_hit_rx && (!maybe_hit_env || (maybe_hit_env && t_env > t_rx));    /* :85 */
        if (hit_recv) {
            pos[0] = pos[0] + dir[0] * t_rx; /* :87 */
            pos[1] = pos[1] + dir[1] * t_rx;
            pos[2] = pos[2] + dir[2] * t_rx;
            float *v = traced + 3 * (bounce + 1);
            v[0] = pos[0]; v[1] = pos[1]; v[2] = pos[2]; /* :88 */
            memcpy(received, traced, sizeof(float) * 3 * (size_t)(bounce + 2)); /* :89-90 */
            *mask = 1u;                                                         /* :91 */
            if (event) { event[bounce] = 2; hit_tri[bounce] = -1; hit_t[bounce] = t_rx; }
        } else if (maybe_hit_env) {
            pos[0] = pos[0] + dir[0] * t_env; /* :94 */
            pos[1] = pos[1] + dir[1] * t_env;
            pos[2] = pos[2] + dir[2] * t_env;
            float *v = traced + 3 * (bounce + 1);
            v[0] = pos[0]; v[1] = pos[1]; v[2] = pos[2]; /* :95 */
            float n[3];
            const float *a = env + 9 * (int64_t)f_env;
            tri_normal(a, a + 3, a + 6, n);
            reflect(dir, n); /* :96 */
            if (event) { event[bounce] = 1; hit_tri[bounce] = f_env; hit_t[bounce] = t_env; }
        } else {
            if (event) { event[bounce] = 0; hit_tri[bounce] = -1; hit_t[bounce] = 0.0f; }
        }
    }
}

/* Dense, reference-shaped outputs for ray ids [tid_begin, tid_begin+n).
 *   env:  nenv*9 floats (triangle soup a,b,c);  rx: nrx*9 floats
 *   traced/received: n*(B+1)*3 floats — overwritten with NaN first (tracer.py:67-71)
 *   mask: n uint32 — zeroed first (tracer.py:72)
 *   hit_tri (n*B int32), hit_t (n*B float), event (n*B int8): optional (may be NULL)
 *   env_bvh: optional handle from oracle_bvh_create (NULL = brute force)            */
ORACLE_API void oracle_trace_paths(const float *env, int64_t nenv, void *env_bvh, const float *rx, int64_t nrx,
                                   const float *tx, int max_bounces, int64_t tid_begin, int64_t n, float *traced,
                                   float *received, uint32_t *mask, int32_t *hit_tri, float *hit_t, int8_t *event,
                                   int nthreads)
{
    const int64_t row = 3 * (int64_t)(max_bounces + 1);
    /* with an environment BVH the receiver mesh gets one too (Warp traverses a BVH for both meshes) */
    OBvh *rx_bvh = (env_bvh && nrx > 0) ? (OBvh *)oracle_bvh_create(rx, nrx) : NULL;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads)
#endif
    for (int64_t i = 0; i < n; ++i) {
        float *tr = traced + i * row, *rc = received + i * row;
        for (int64_t k = 0; k < row; ++k) { tr[k] = NAN; rc[k] = NAN; }
        mask[i] = 0u;
        trace_one(env, nenv, (const OBvh *)env_bvh, rx, nrx, rx_bvh, tx, max_bounces, (uint32_t)(tid_begin + i), tr, rc,
                  mask + i, hit_tri ? hit_tri + i * max_bounces : NULL, hit_t ? hit_t + i * max_bounces : NULL,
                  event ? event + i * max_bounces : NULL);
    }
    oracle_bvh_destroy(rx_bvh);
}

/* Sparse variant for big N: same loop, but only received rows are kept.
 *   recv_tid (cap), recv_paths (cap*(B+1)*3): rows of received rays in ascending tid order
 *   returns the number of received rays (may exceed cap; only the first cap are stored). */
ORACLE_API int64_t oracle_trace_received(const float *env, int64_t nenv, void *env_bvh, const float *rx,
                                         int64_t nrx, const float *tx, int max_bounces, int64_t tid_begin,
                                         int64_t n, int64_t cap, int64_t *recv_tid, float *recv_paths,
                                         int nthreads)
{
    const int64_t row = 3 * (int64_t)(max_bounces + 1);
    int64_t count = 0;
    const int64_t chunk = 1 << 16;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
    uint8_t *flag = (uint8_t *)malloc((size_t)chunk);
    float *rows = (float *)malloc(sizeof(float) * (size_t)(chunk * row));
    OBvh *rx_bvh = (env_bvh && nrx > 0) ? (OBvh *)oracle_bvh_create(rx, nrx) : NULL;
    for (int64_t c0 = 0; c0 < n; c0 += chunk) {
        int64_t m = n - c0 < chunk ? n - c0 : chunk;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024) num_threads(nthreads)
#endif
        for (int64_t i = 0; i < m; ++i) {
            float tr[3 * 65];
            float *rc = rows + i * row;
            uint32_t msk = 0;
            for (int64_t k = 0; k < row; ++k) { tr[k] = NAN; rc[k] = NAN; }
            trace_one(env, nenv, (const OBvh *)env_bvh, rx, nrx, rx_bvh, tx, max_bounces, (uint32_t)(tid_begin + c0 + i),
                      tr, rc, &msk, NULL, NULL, NULL);
            flag[i] = (uint8_t)msk;
        }
        for (int64_t i = 0; i < m; ++i)
            if (flag[i]) {
                if (count < cap) {
                    recv_tid[count] = tid_begin + c0 + i;
                    memcpy(recv_paths + count * row, rows + i * row, sizeof(float) * (size_t)row);
                }
                ++count;
            }
    }
    free(flag);
    free(rows);
    oracle_bvh_destroy(rx_bvh);
    return count;
}

/* Environment-only trajectory (the loop of kernel.py:57-98 with no receiver): this is the
 * trajectory every receiver's run shares until its first RX hit.  Counts "segments"
 * (SURVEY.md §8d): bounce iterations of rays that have not yet missed the environment.
 *   hit_tri (n*B int32, -1 = miss), hit_t (n*B float): optional
 * returns the segment count. */
ORACLE_API uint64_t oracle_trace_env(const float *env, int64_t nenv, void *env_bvh, const float *tx, int max_bounces,
                                     int64_t tid_begin, int64_t n, int32_t *hit_tri, float *hit_t, int nthreads)
{
    uint64_t segments = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads) reduction(+ : segments)
#endif
    for (int64_t i = 0; i < n; ++i) {
        float dir[3], pos[3] = {tx[0], tx[1], tx[2]};
        ray_direction((uint32_t)(tid_begin + i), dir);
        int alive = 1;
        for (int b = 0; b < max_bounces; ++b) {
            int32_t f = -1;
            float t = 0.0f;
            if (alive) {
                ++segments;
                int fe = 0;
                if (query(env, nenv, (const OBvh *)env_bvh, pos, dir, 1.0e6f, &t, &fe)) {
                    f = fe;
                    pos[0] = pos[0] + dir[0] * t;
                    pos[1] = pos[1] + dir[1] * t;
                    pos[2] = pos[2] + dir[2] * t;
                    float nrm[3];
                    const float *a = env + 9 * (int64_t)fe;
                    tri_normal(a, a + 3, a + 6, nrm);
                    reflect(dir, nrm);
                } else {
                    alive = 0;
                    t = 0.0f;
                }
            }
            if (hit_tri) hit_tri[i * max_bounces + b] = f;
            if (hit_t) hit_t[i * max_bounces + b] = t;
        }
    }
    return segments;
}

/* ---- small exported probes used by unit tests ------------------------------------------ */
ORACLE_API void oracle_ray_directions(int64_t tid_begin, int64_t n, float *dirs)
{
    for (int64_t i = 0; i < n; ++i) ray_direction((uint32_t)(tid_begin + i), dirs + 3 * i);
}

ORACLE_API void oracle_ray_directions_list(const int64_t *tids, int64_t n, float *dirs)
{
    for (int64_t i = 0; i < n; ++i) ray_direction((uint32_t)tids[i], dirs + 3 * i);
}

ORACLE_API uint32_t oracle_pcg(uint32_t s) { return pcg_hash(s); }

ORACLE_API void oracle_det_math(const double *x, int64_t n, double *s, double *c, double *ac)
{
    for (int64_t i = 0; i < n; ++i) {
        det_sincos(fabs(x[i]), s + i, c + i);
        double z = x[i];
        if (z > 1.0) z = 1.0;
        if (z < -1.0) z = -1.0;
        ac[i] = det_acos(z);
    }
}

/* single closest-hit query; returns 1 on hit.  use_bvh: handle or NULL */
ORACLE_API int oracle_query(const float *tris, int64_t ntris, void *bvh, const float *p, const float *dir,
                            float max_t, float *t, int *face)
{
    return query(tris, ntris, (const OBvh *)bvh, p, dir, max_t, t, face);
}

ORACLE_API void oracle_tri_normal(const float *tri9, float *n) { tri_normal(tri9, tri9 + 3, tri9 + 6, n); }

/* ------------------------------------------------------------------------------------------
 * "Physical" mode (NOT reference behaviour — it removes quirks Q1-Q6):
 *   - the triangle a ray has just left is excluded from its next closest-hit query;
 *   - receivers are analytic spheres; a segment [0, t_env] whose closest approach to a centre lies inside the
 *     sphere (origin outside it) is one arrival with unfolded path length L = L_prev + t_c * |dir|;
 *   - field (unit transmit power, isotropic antennas):
 *         E = (L * lambda / (pi * N * r^2)) * prod_i Gamma_i * exp(-j 2 pi L / lambda)
 *     i.e. free-space loss lambda / (4 pi L) times the reception-sphere weight 4 L^2 / (N r^2);
 *     Gamma = (cos(theta_t) - n cos(theta_i)) / (cos(theta_t) + n cos(theta_i)), sin(theta_t) = sin(theta_i) / n
 *     (p-polarised Fresnel amplitude coefficient, air -> index n; tracer.py:43-53 uses its square with n = 5);
 *   - field[k] (re, im) = sum of E over all arrivals at receiver k; optional complex impulse response
 *     ir[k][bin] += E with bin = int(L / c * rate) (tracer.py:115).
 * Geometry is fp32 with the trace's own operations; the sphere test and the field are fp64.
 * materials: per-triangle index n, or NULL for 5.0.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API uint64_t oracle_trace_physical(const float *env, int64_t nenv, void *env_bvh, const float *materials,
                                          const double *rx_centers, int64_t nrx, double rx_radius, const float *tx,
                                          int max_bounces, int64_t tid_begin, int64_t n, int64_t n_total,
                                          double carrier_hz, double light_speed, double sample_rate, int64_t n_bins,
                                          double *field, double *ir, uint64_t *arrivals_out)
{
    const double lambda = light_speed / carrier_hz;
    const double wk = lambda / (3.141592653589793 * (double)n_total * (rx_radius * rx_radius));
    const double two_pi_over_lambda = (2.0 * 3.141592653589793) / lambda;
    const double r2 = rx_radius * rx_radius;
    uint64_t segments = 0, arrivals = 0;
    for (int64_t i = 0; i < n; ++i) {
        float dir[3], pos[3] = {tx[0], tx[1], tx[2]};
        ray_direction((uint32_t)(tid_begin + i), dir);
        int prev = -1;
        double L = 0.0, gamma = 1.0;
        for (int b = 0; b < max_bounces; ++b) {
            float t_env = 0.0f;
            int face = -1;
            int hit = query_skip(env, nenv, (const OBvh *)env_bvh, pos, dir, 1.0e6f, prev, &t_env, &face);
            ++segments;
            const float dlen = sqrtf((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]);
            const double t_lim = hit ? (double)t_env : 1.0e6;
            const double dd = ((double)dir[0] * (double)dir[0] + (double)dir[1] * (double)dir[1]) + (double)dir[2] * (double)dir[2];
            for (int64_t k = 0; k < nrx; ++k) {
                const double ox = rx_centers[3 * k] - (double)pos[0], oy = rx_centers[3 * k + 1] - (double)pos[1],
                             oz = rx_centers[3 * k + 2] - (double)pos[2];
                const double oo = (ox * ox + oy * oy) + oz * oz;
                if (oo <= r2) continue; /* the segment starts inside the sphere: no new arrival */
                const double od = (ox * (double)dir[0] + oy * (double)dir[1]) + oz * (double)dir[2];
                const double tc = od / dd;
                if (!(tc >= 0.0) || !(tc <= t_lim)) continue;
                const double perp2 = oo - tc * od;
                if (!(perp2 <= r2)) continue;
                const double Lk = L + tc * (double)dlen;
                const double a = (Lk * wk) * gamma;
                const double ph = two_pi_over_lambda * Lk;
                const double re = a * cos(ph), im = -(a * sin(ph));
                field[2 * k] += re;
                field[2 * k + 1] += im;
                if (ir) {
                    const int64_t bin = (int64_t)((Lk / light_speed) * sample_rate);
                    if (bin >= 0 && bin < n_bins) {
                        ir[2 * (k * n_bins + bin)] += re;
                        ir[2 * (k * n_bins + bin) + 1] += im;
                    }
                }
                ++arrivals;
            }
            if (!hit) break;
            const float *a = env + 9 * (int64_t)face;
            float nrm[3];
            tri_normal(a, a + 3, a + 6, nrm);
            const float dn = (dir[0] * nrm[0] + dir[1] * nrm[1]) + dir[2] * nrm[2];
            const double nmat = materials ? (double)materials[face] : 5.0;
            double ci = fabs((double)dn) / (double)dlen;
            if (ci > 1.0) ci = 1.0;
            const double si2 = 1.0 - ci * ci;
            const double ct = sqrt(1.0 - si2 / (nmat * nmat));
            gamma = gamma * ((ct - nmat * ci) / (ct + nmat * ci));
            L = L + (double)t_env * (double)dlen;
            pos[0] = pos[0] + dir[0] * t_env;
            pos[1] = pos[1] + dir[1] * t_env;
            pos[2] = pos[2] + dir[2] * t_env;
            reflect(dir, nrm);
            prev = face;
        }
    }
    if (arrivals_out) *arrivals_out = arrivals;
    return segments;
}

ORACLE_API int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* KAT-1 helper: ray ids in [tid_begin, tid_begin+n) whose direction (fp32, as generated above)
 * hits the ANALYTIC sphere (center, radius) from tx; geometry evaluated in fp64.
 * Returns the number found (stores up to cap, ascending). */
ORACLE_API int64_t oracle_sphere_hits(int64_t tid_begin, int64_t n, const double *tx, const double *center,
                                      double radius, int64_t *out_tids, int64_t cap, int nthreads)
{
    int64_t count = 0;
    const int64_t chunk = 1 << 22;
    uint8_t *flag = (uint8_t *)malloc((size_t)chunk);
    double oc[3] = {center[0] - tx[0], center[1] - tx[1], center[2] - tx[2]};
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
    for (int64_t c0 = 0; c0 < n; c0 += chunk) {
        int64_t m = n - c0 < chunk ? n - c0 : chunk;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
        for (int64_t i = 0; i < m; ++i) {
            float d[3];
            ray_direction((uint32_t)(tid_begin + c0 + i), d);
            double dd = (double)d[0] * d[0] + (double)d[1] * d[1] + (double)d[2] * d[2];
            double tc = ((double)d[0] * oc[0] + (double)d[1] * oc[1] + (double)d[2] * oc[2]) / dd;
            double px = oc[0] - tc * d[0], py = oc[1] - tc * d[1], pz = oc[2] - tc * d[2];
            flag[i] = (tc > 0.0 && px * px + py * py + pz * pz <= radius * radius) ? 1 : 0;
        }
        for (int64_t i = 0; i < m; ++i)
            if (flag[i]) {
                if (count < cap) out_tids[count] = tid_begin + c0 + i;
                ++count;
            }
    }
    free(flag);
    return count;
}
