// rfrt_trace.cuh — device-side BVH traversal shared by the trace kernels.
#pragma once
#include "rfrt_internal.h"
#include "rfrt_math.cuh"

namespace rfrt {

// Ray constants for the (conservative) slab test.  FMA is fine here: boxes are padded and the slab
// test never decides a hit, it only prunes.
struct SlabRay {
    float ix, iy, iz;    // 1/d   (|d| clamped to >= 1e-18 so it stays finite)
    float nx, ny, nz;    // o * (1/d) + delta * |1/d| : subtracted on the NEAR face of a slab (entry parameter rounded down)
    float fx, fy, fz;    // o * (1/d) - delta * |1/d| : subtracted on the FAR face (exit parameter rounded up)
    bool px, py, pz;     // d >= 0 along the axis: the near face is the lower one
};
// delta = 2^-21 * max |o_k| covers what depends on the ORIGIN of the ray, per axis and in position space: the rounding of
// o_k * (1/d_k) (half an ulp) and the reach of the exact triangle test around a triangle (a few ulps of |o| + |vertex|).
// What depends on the MESH's own coordinates is covered by the padding of the boxes (rfrt_bvh.cu: 1e-5 * max
// |coordinate|) — so the padding need not anticipate how far away a transmitter may stand (round 1 padded by >= 1e-3 like
// Warp does: 8 % of a 1.3 cm triangle's box on the 20 M-triangle terrain, 2.12 instead of 1.54 triangle tests per
// segment).  Near / far faces are picked by the sign of the direction (ray-uniform selects instead of min / max: the
// same instruction count), which is what lets the two faces carry different offsets.
__device__ __forceinline__ void slab_finish(SlabRay &s, float3 p)
{
    const float ox = p.x * s.ix, oy = p.y * s.iy, oz = p.z * s.iz;
    const float delta = fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z)) * (1.0f / 2097152.0f);
    const float ex = delta * fabsf(s.ix), ey = delta * fabsf(s.iy), ez = delta * fabsf(s.iz);
    s.nx = ox + ex; s.ny = oy + ey; s.nz = oz + ez;
    s.fx = ox - ex; s.fy = oy - ey; s.fz = oz - ez;
    s.px = s.ix >= 0.0f; s.py = s.iy >= 0.0f; s.pz = s.iz >= 0.0f;
}

__device__ __forceinline__ SlabRay slab_setup(float3 p, float3 d)
{
    SlabRay s;
    // a zero (or denormal-small) component would give inf * 0 = NaN in the fused slab arithmetic; tilt the
    // ray by 1e-18 instead: a parallel ray then sees (-huge, +huge) inside a slab and (huge, huge) outside
    const float tiny = 1.0e-18f;
    float dx = fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x;
    float dy = fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y;
    float dz = fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z;
    s.ix = __fdiv_rn(1.0f, dx); s.iy = __fdiv_rn(1.0f, dy); s.iz = __fdiv_rn(1.0f, dz);
    slab_finish(s, p);
    return s;
}

// rcp.approx instead of IEEE division (2 ulp: a relative error of the parameters along one axis against another of
// 2.4e-7 — the relative slack of slab_hit below)
__device__ __forceinline__ SlabRay slab_setup_fast(float3 p, float3 d)
{
    SlabRay s;
    const float tiny = 1.0e-18f;
    float dx = fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x;
    float dy = fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y;
    float dz = fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s.ix) : "f"(dx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s.iy) : "f"(dy));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s.iz) : "f"(dz));
    slab_finish(s, p);
    return s;
}

// entry distance of the padded box; hit iff  max(tnear,0) <= min(tfar, t_max).
__device__ __forceinline__ bool slab_hit(const SlabRay &s, float lx, float ly, float lz, float hx, float hy, float hz,
                                         float t_max, float &t_near)
{
    const float tnx = fmaf(s.px ? lx : hx, s.ix, -s.nx), tfx = fmaf(s.px ? hx : lx, s.ix, -s.fx);
    const float tny = fmaf(s.py ? ly : hy, s.iy, -s.ny), tfy = fmaf(s.py ? hy : ly, s.iy, -s.fy);
    const float tnz = fmaf(s.pz ? lz : hz, s.iz, -s.nz), tfz = fmaf(s.pz ? hz : lz, s.iz, -s.fz);
    float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
    float tf = fminf(fminf(tfx, tfy), fminf(tfz, t_max));
    t_near = tn;
    // a small relative slack keeps the FMA rounding of the parameters (and of an approximate 1/d) from ever culling a
    // box the exact test would enter
    return tn <= tf * 1.0000008f + 1.0e-30f;
}

// Receiver boxes (rfrt_rxset_create) keep Warp's padding of >= 1e-3 — a hundredth of the usual 0.1 m receiver, nothing
// to gain from less — and the plain test: six floats of ray constants, min / max per axis.
struct RxSlabRay {
    float ix, iy, iz, ox, oy, oz;
};
__device__ __forceinline__ RxSlabRay rx_slab_setup(float3 p, float3 d)
{
    RxSlabRay s;
    const float tiny = 1.0e-18f;
    float dx = fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x;
    float dy = fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y;
    float dz = fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s.ix) : "f"(dx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s.iy) : "f"(dy));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s.iz) : "f"(dz));
    s.ox = p.x * s.ix; s.oy = p.y * s.iy; s.oz = p.z * s.iz;
    return s;
}
__device__ __forceinline__ bool rx_slab_hit(const RxSlabRay &s, float lx, float ly, float lz, float hx, float hy, float hz,
                                            float t_max, float &t_near)
{
    float t0x = fmaf(lx, s.ix, -s.ox), t1x = fmaf(hx, s.ix, -s.ox);
    float t0y = fmaf(ly, s.iy, -s.oy), t1y = fmaf(hy, s.iy, -s.oy);
    float t0z = fmaf(lz, s.iz, -s.oz), t1z = fmaf(hz, s.iz, -s.oz);
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), t_max));
    t_near = tn;
    return tn <= tf * 1.0000004f + 1.0e-30f;
}
// The plain test is also enough for an ENVIRONMENT walk whose origins all lie within 8 x the mesh's largest coordinate
// m of the coordinate origin: delta <= 2^-21 * 8 m = 3.8e-6 m then sits inside the boxes' padding of 1e-5 m with more
// than half of it to spare.  The replay (k_trace_receive) knows that per launch — its origins are the transmitter and
// hit points — and picks this cheaper ray (three registers fewer through its walk) unless the transmitter stands far
// outside the scene.
__device__ __forceinline__ bool slab_hit(const RxSlabRay &s, float lx, float ly, float lz, float hx, float hy, float hz,
                                         float t_max, float &t_near)
{
    return rx_slab_hit(s, lx, ly, lz, hx, hy, hz, t_max, t_near);
}
template <class SLAB>
__device__ __forceinline__ SLAB slab_make(float3 p, float3 d)
{
    if constexpr (std::is_same<SLAB, RxSlabRay>::value) return rx_slab_setup(p, d);
    else return slab_setup(p, d);
}

__device__ __forceinline__ void tri_vertices(const BvhTri *__restrict__ tris, int slot, float3 &a, float3 &b,
                                             float3 &c, int &index)
{
    const float4 *p = reinterpret_cast<const float4 *>(tris + slot);
    float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
    a = make_float3(v0.x, v0.y, v0.z);
    b = make_float3(v0.w, v1.x, v1.y);
    c = make_float3(v1.z, v1.w, v2.x);
    index = __float_as_int(v2.y);
}

struct Hit {
    float t;   // best distance so far (starts at max_t)
    int face;  // original triangle index (-1 = none)
    int slot;  // sorted slot of that triangle
};

// Traversal state codes: node >= 0 internal node, node < 0 leaf (~slot), TRAV_DONE = walk finished.
constexpr int TRAV_DONE = (int)0x80000000;

// pop the next stack entry that the current best does not already rule out (TRAV_DONE if none)
__device__ __forceinline__ int stack_pop(int *stack, float *stack_t, int stride, int &sp, float best_t)
{
    int next = TRAV_DONE;
    while (sp > 0) {
        --sp;
        if (stack_t[sp * stride] <= best_t) { next = stack[sp * stride]; break; }
    }
    return next;
}

// One 64-byte node = two 256-bit loads (LDG.E.ENL2.256 on sm_100a; nodes come from cudaMalloc and are 64-byte aligned):
// half the load instructions of four 128-bit loads for lanes that all sit on different nodes (+2 % on the 20 M-triangle
// terrain).  Only the environment walks use it: in the receiver walk of k_trace_small the eight-register alignment of
// the wide load costs the kernel its register allocation (C4 kernel 10.36 -> 10.97 ms).
__device__ __forceinline__ void load_node(const BvhNode *__restrict__ nd, float4 &q0, float4 &q1, float4 &q2, int4 &q3)
{
#ifdef RFRT_NODE_LDG128
    const float4 *np = reinterpret_cast<const float4 *>(nd);
    q0 = __ldg(np); q1 = __ldg(np + 1); q2 = __ldg(np + 2);
    q3 = __ldg(reinterpret_cast<const int4 *>(np + 3));
#else
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w), "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w)
        : "l"(nd));
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(q2.x), "=f"(q2.y), "=f"(q2.z), "=f"(q2.w), "=r"(q3.x), "=r"(q3.y), "=r"(q3.z), "=r"(q3.w)
        : "l"(reinterpret_cast<const char *>(nd) + 32));
#endif
}

// visit one internal node: slab-test both child boxes (4 x 128-bit loads), go to the nearer hit child and
// defer the other (with its entry distance) on the stack.  Returns the next state code.
template <class SLAB>
__device__ __forceinline__ int node_step(const BvhNode *__restrict__ nodes, int node, const SLAB &sr, float best_t,
                                         int *stack, float *stack_t, int stride, int &sp)
{
    float4 q0, q1, q2;
    int4 q3;
    load_node(nodes + node, q0, q1, q2, q3);
    float tn0, tn1;
    bool h0 = slab_hit(sr, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, best_t, tn0);
    bool h1 = slab_hit(sr, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, best_t, tn1);
    const int c0 = q3.x, c1 = q3.y;
    if (c1 == c0) h1 = false; // single-primitive tree
    int next;
    if (h0 && h1) {
        const bool swap = tn1 < tn0;
        next = swap ? c1 : c0;
        stack[sp * stride] = swap ? c0 : c1;
        stack_t[sp * stride] = swap ? tn0 : tn1;
        ++sp;
    } else if (h0) {
        next = c0;
    } else if (h1) {
        next = c1;
    } else {
        next = stack_pop(stack, stack_t, stride, sp, best_t);
    }
    return next;
}

// test the triangle of one leaf (closest t, 0 <= t < best; equal t -> lowest triangle index), then pop
// (skip: a triangle index the query ignores — physical mode's "the triangle just left"; -1 = none)
// (RAY: WoopRay = the reference's watertight functor, MtRay = Moeller-Trumbore; rfrt_math.cuh)
template <class RAY>
__device__ __forceinline__ int leaf_step(const BvhTri *__restrict__ tris, int node, const RAY &wr, Hit &h,
                                         int *stack, float *stack_t, int stride, int &sp, int skip = -1)
{
    const int slot = ~node;
    float3 a, b, c; int idx; float t;
    tri_vertices(tris, slot, a, b, c, idx);
    if (idx != skip && tri_hit(wr, a, b, c, t) && t >= 0.0f && (t < h.t || (t == h.t && h.face >= 0 && idx < h.face))) {
        h.t = t; h.face = idx; h.slot = slot;
    }
    return stack_pop(stack, stack_t, stride, sp, h.t);
}

// mesh_query_ray (kernel.py:71,82): closest accepted hit, 0 <= t < max_t; equal t -> lowest index.
// `stack` / `stack_t` point at this thread's column of the shared-memory stack (stride = blockDim.x).
// "while-while" (Aila & Laine): the inner loop walks internal nodes until THIS lane holds a leaf, then the
// lanes holding one test their triangles together.  Each loop has ONE back-edge so the warp re-converges on
// every trip (several `continue` back-edges let sub-groups of lanes run the loop separately: measured 6/32
// lanes active; testing leaves inside the node loop: 3/32 active in the triangle test).
// COUNT: also count the internal nodes fetched and the triangles tested (RFRT_CTR_NODE_VISITS / RFRT_CTR_TRI_TESTS)
template <bool COUNT = false, class RAY = WoopRay, class SLAB = SlabRay>
__device__ __forceinline__ void closest_hit(const BvhNode *__restrict__ nodes, const BvhTri *__restrict__ tris,
                                            int64_t n_prims, const RAY &wr, const SLAB &sr, int *stack,
                                            float *stack_t, int stride, Hit &h, int skip = -1,
                                            unsigned *n_nodes = nullptr, unsigned *n_tests = nullptr)
{
    int sp = 0;
    int node = n_prims > 0 ? 0 : TRAV_DONE;
    while (node != TRAV_DONE) {
        while (node >= 0) {
            node = node_step(nodes, node, sr, h.t, stack, stack_t, stride, sp);
            if (COUNT) ++*n_nodes;
        }
        if (node != TRAV_DONE) {
            node = leaf_step(tris, node, wr, h, stack, stack_t, stride, sp, skip);
            if (COUNT) ++*n_tests;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Small scenes (<= RFRT_SMALL_MAX_TRIS triangles): closest hit by a warp-lockstep sweep over the whole scene,
// staged in shared memory, instead of a BVH walk.  Every lane executes the same instruction stream (no traversal
// divergence), which on a 44-triangle room beats the BVH although it looks at every triangle.
//   phase 1 (lockstep): a conservative candidate filter (tables of rfrt_small.cu).  Per supporting plane: ray
//            parameter t and hit point h; per triangle: the three in-plane edge distances of h.  A triangle is
//            dropped only when h is farther outside one of its edges than the tolerance, or its plane lies behind
//            the origin by more than the tolerance.  All loads are warp-uniform (shared-memory broadcasts); NaN /
//            inf (ray parallel to the plane, degenerate triangle) always KEEP the candidate.
//   phase 2 (per lane): the exact watertight test — the fp32 operation sequence of woop_hit — on the candidates,
//            nearest first, until the rest provably lies behind the best hit; closest t in [0, best), equal t ->
//            lowest triangle index (same rule as the BVH path).
// The answer is therefore bit-identical to testing all triangles exactly (tests/test_gpu_parity.py compares it
// with the CPU restatement and with the BVH kernel; tests/test_small_filter_cpu.py checks the superset property).
// ---------------------------------------------------------------------------------------------------------
struct SmallScene {
    const float4 *recs;   // 7 x float4 per pair of coplanar triangles: (n.xyz, d), 2 x 3 x (m_i.xyz, c_i)
    const int *slot_tri;  // slot -> triangle index (slots 2k, 2k+1 = pair k)
    const float *soup;    // [9*n] original order
    const float *normals; // [3*n] original order
    const uint4 *nbr;     // [n] neighbour pair masks of triangle f: interior, boundary, interior & lower index, boundary & lower
    const int *tri_slot;  // [n] a filter slot of triangle f
    int n_pairs;
    int cls[5];           // pair ranges of the plane classes (general, x-, y-, z-aligned)
    float extent;         // max |coordinate| of the scene
    float tau;            // self-re-hit shortcut: path-length bound
    float erode;          // ... clearance that rules out the boundary neighbours (2 * reach)
};

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// minimum that ignores NaN operands (NaN only when all three are NaN): one FMNMX3 on sm_100a
__device__ __forceinline__ float min3f(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// U, V, W of one shared-memory triangle (offset o = 9*f) with the permuted components fetched through the
// per-lane base pointers tkx/tky/tkz — the same fp32 operations, in the same order, as woop_hit.
struct WoopUVW {
    float U, V, W, Akz, Bkz, Ckz;
};
__device__ __forceinline__ WoopUVW woop_uvw_smem(const float *tkx, const float *tky, const float *tkz, int o, float pkx,
                                                 float pky, float pkz, float Sx, float Sy)
{
    WoopUVW r;
    const float Akx = __fsub_rn(tkx[o], pkx), Aky = __fsub_rn(tky[o], pky);
    r.Akz = __fsub_rn(tkz[o], pkz);
    const float Bkx = __fsub_rn(tkx[o + 3], pkx), Bky = __fsub_rn(tky[o + 3], pky);
    r.Bkz = __fsub_rn(tkz[o + 3], pkz);
    const float Ckx = __fsub_rn(tkx[o + 6], pkx), Cky = __fsub_rn(tky[o + 6], pky);
    r.Ckz = __fsub_rn(tkz[o + 6], pkz);
    const float Ax = __fsub_rn(Akx, __fmul_rn(Sx, r.Akz));
    const float Ay = __fsub_rn(Aky, __fmul_rn(Sy, r.Akz));
    const float Bx = __fsub_rn(Bkx, __fmul_rn(Sx, r.Bkz));
    const float By = __fsub_rn(Bky, __fmul_rn(Sy, r.Bkz));
    const float Cx = __fsub_rn(Ckx, __fmul_rn(Sx, r.Ckz));
    const float Cy = __fsub_rn(Cky, __fmul_rn(Sy, r.Ckz));
    r.U = diff_product(Cx, By, Cy, Bx);
    r.V = diff_product(Ax, Cy, Ay, Cx);
    r.W = diff_product(Bx, Ay, By, Ax);
    if (r.U == 0.0f || r.V == 0.0f || r.W == 0.0f) {
        // fp64 fallback of intersect_ray_tri_woop (rare: edge / vertex hits)
        r.U = __double2float_rn(__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx)));
        r.V = __double2float_rn(__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx)));
        r.W = __double2float_rn(__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax)));
    }
    return r;
}

// filter of one pair record: shifts the two "definitely outside" bits (slot 2k+1, then slot 2k) into `dropped`
__device__ __forceinline__ unsigned pair_shift_in(const float4 *rec, float3 pos, float3 dir, float dl, float dl_h, unsigned dropped)
{
    const float4 P = rec[0];
    const float nd = fmaf(P.x, dir.x, fmaf(P.y, dir.y, P.z * dir.z));
    const float np = fmaf(P.x, pos.x, fmaf(P.y, pos.y, fmaf(P.z, pos.z, -P.w)));
    const float r = rcp_approx(nd);
    const float t = -np * r;
    const float ar = fabsf(r);
    // plane behind the origin by more than the tolerance: thr = +inf drops both triangles
    const float thr = (t < -(dl * ar)) ? __int_as_float(0x7f800000) : -(dl_h * ar);
    const float hx = fmaf(t, dir.x, pos.x), hy = fmaf(t, dir.y, pos.y), hz = fmaf(t, dir.z, pos.z);
#pragma unroll
    for (int j = 1; j >= 0; --j) {
        const float4 e0 = rec[1 + 3 * j], e1 = rec[2 + 3 * j], e2 = rec[3 + 3 * j];
        const float d0 = fmaf(e0.x, hx, fmaf(e0.y, hy, fmaf(e0.z, hz, e0.w)));
        const float d1 = fmaf(e1.x, hx, fmaf(e1.y, hy, fmaf(e1.z, hz, e1.w)));
        const float d2 = fmaf(e2.x, hx, fmaf(e2.y, hy, fmaf(e2.z, hz, e2.w)));
        // shift the sign of (min_i d_i - thr) into the mask: 1 = definitely outside.  inf - inf and NaN inputs
        // give the canonical NaN 0x7fffffff (sign clear), i.e. the triangle is kept.
        dropped = __funnelshift_l(__float_as_uint(min3f(d0, d1, d2) - thr), dropped, 1);
    }
    return dropped;
}

// The same filter for pairs whose plane is normal to axis K (n = +e_K exactly, so the edge normals have no K
// component): the zero terms of pair_shift_in are left out — identical values, 32 instead of 46 instructions per pair.
// r, kt, thr0: 1 / dir[K], dl * |r|, -dl_h * |r| (constant for the whole class).
template <int K>
__device__ __forceinline__ unsigned pair_shift_in_axis(const float4 *rec, float3 pos, float3 dir, float r, float kt, float thr0,
                                                       unsigned dropped)
{
    const float pk = K == 0 ? pos.x : (K == 1 ? pos.y : pos.z);
    const float pi = K == 0 ? pos.y : pos.x, pj = K == 2 ? pos.y : pos.z; // the two in-plane axes, in x < y < z order
    const float di = K == 0 ? dir.y : dir.x, dj = K == 2 ? dir.y : dir.z;
    const float np = pk - rec[0].w;
    const float t = -np * r;
    const float thr = (t < -kt) ? __int_as_float(0x7f800000) : thr0;
    const float hi = fmaf(t, di, pi), hj = fmaf(t, dj, pj);
#pragma unroll
    for (int j = 1; j >= 0; --j) {
        const float4 e0 = rec[1 + 3 * j], e1 = rec[2 + 3 * j], e2 = rec[3 + 3 * j];
        const float d0 = fmaf(K == 0 ? e0.y : e0.x, hi, fmaf(K == 2 ? e0.y : e0.z, hj, e0.w));
        const float d1 = fmaf(K == 0 ? e1.y : e1.x, hi, fmaf(K == 2 ? e1.y : e1.z, hj, e1.w));
        const float d2 = fmaf(K == 0 ? e2.y : e2.x, hi, fmaf(K == 2 ? e2.y : e2.z, hj, e2.w));
        dropped = __funnelshift_l(__float_as_uint(min3f(d0, d1, d2) - thr), dropped, 1);
    }
    return dropped;
}

template <int K>
__device__ __forceinline__ unsigned sweep_axis_range(const float4 *recs, int b, int e, float3 pos, float3 dir, float dl, float dl_h,
                                                     unsigned dropped)
{
    if (e <= b) return dropped;
    const float r = rcp_approx(K == 0 ? dir.x : (K == 1 ? dir.y : dir.z));
    const float ar = fabsf(r), kt = dl * ar, thr0 = -(dl_h * ar);
    const float4 *rec = recs + 7 * (e - 1);
#pragma unroll 2
    for (int k = e - 1; k >= b; --k, rec -= 7) dropped = pair_shift_in_axis<K>(rec, pos, dir, r, kt, thr0, dropped);
    return dropped;
}

// phase 1 over the pairs [first, first + count) (walked backwards, so that slot s ends up at mask bit s - 2 * first):
// the candidate mask.  cls[0..4]: pair ranges of the plane classes general / x / y / z (rfrt_small.cu).
__device__ __forceinline__ unsigned sweep_pairs(const float4 *recs, const int *cls, int first, int count, float3 pos, float3 dir,
                                                float dl, float dl_h)
{
    const int last = first + count;
    unsigned dropped = 0u;
    dropped = sweep_axis_range<2>(recs, max(first, cls[3]), min(last, cls[4]), pos, dir, dl, dl_h, dropped);
    dropped = sweep_axis_range<1>(recs, max(first, cls[2]), min(last, cls[3]), pos, dir, dl, dl_h, dropped);
    dropped = sweep_axis_range<0>(recs, max(first, cls[1]), min(last, cls[2]), pos, dir, dl, dl_h, dropped);
    {
        const int b = max(first, cls[0]), e = min(last, cls[1]);
        const float4 *rec = recs + 7 * (e - 1);
#pragma unroll 2
        for (int k = e - 1; k >= b; --k, rec -= 7) dropped = pair_shift_in(rec, pos, dir, dl, dl_h, dropped);
    }
    return ~dropped & (count >= 16 ? 0xffffffffu : (1u << (2 * count)) - 1u);
}

// per-ray constants of the exact test against the shared-memory soup
struct SmallExact {
    const float *tkx, *tky, *tkz; // soup base pointers offset by the permuted component
    float pkx, pky, pkz;
};
__device__ __forceinline__ SmallExact small_exact_setup(const SmallScene &S, const WoopRay &wr)
{
    SmallExact x;
    x.tkx = S.soup + wr.kx; x.tky = S.soup + wr.ky; x.tkz = S.soup + wr.kz;
    x.pkx = wr.pkx; x.pky = wr.pky; x.pkz = wr.pkz;
    return x;
}

// the exact test of triangle f: closest t in [0, best], equal t -> lowest triangle index
__device__ __forceinline__ void small_exact(const SmallExact &X, int f, const WoopRay &wr, Hit &h)
{
    const WoopUVW q = woop_uvw_smem(X.tkx, X.tky, X.tkz, 9 * f, X.pkx, X.pky, X.pkz, wr.Sx, wr.Sy);
    const float U = q.U, V = q.V, W = q.W;
    const bool mixed = (U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f);
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    const float Az = __fmul_rn(wr.Sz, q.Akz), Bz = __fmul_rn(wr.Sz, q.Bkz), Cz = __fmul_rn(wr.Sz, q.Ckz);
    const float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));
    const float x = __uint_as_float(__float_as_uint(T) ^ (__float_as_uint(det) & 0x80000000u));
    if (!mixed && det != 0.0f && !(x < 0.0f)) {
        const float t = __fmul_rn(T, __fdiv_rn(1.0f, det));
        if (t >= 0.0f && (t < h.t || (t == h.t && h.face >= 0 && f < h.face))) { h.t = t; h.face = f; h.slot = f; }
    }
}

// lower bound of the exact t of the triangles of slot's pair (NaN-free: -inf when undecidable)
__device__ __forceinline__ float small_t_lower(const SmallScene &S, int slot, float3 pos, float3 dir, float dl)
{
    const float4 P = S.recs[7 * (slot >> 1)];
    const float nd = fmaf(P.x, dir.x, fmaf(P.y, dir.y, P.z * dir.z));
    const float np = fmaf(P.x, pos.x, fmaf(P.y, pos.y, fmaf(P.z, pos.z, -P.w)));
    const float r = rcp_approx(nd);
    return fmaxf(fmaf(-dl, fabsf(r), -np * r), -__int_as_float(0x7f800000)); // t - kt; NaN -> -inf
}

__device__ __forceinline__ int pop_slot(unsigned &lo, unsigned &hi)
{
    if (lo) { const int b = 31 - __clz((int)lo); lo ^= 1u << b; return b; }
    const int b = 31 - __clz((int)hi); hi ^= 1u << b; return 32 + b;
}

// phase 2: the exact test on the candidate slots (lo, hi), nearest first.  The plane parameter minus the tolerance
// (t_lo) is a lower bound of a candidate's exact t, so once the best exact hit lies below the t_lo of everything that
// is left, the rest cannot win (on room.stl this cuts the exact tests per segment from 2.3 to 1.0: the ray's own
// surface at t ~ 0 and the shell behind the hit drop out).  SKIP: ignore the slots of triangle `skip` (already in h).
template <bool SKIP>
__device__ __forceinline__ void small_resolve(const SmallScene &S, unsigned lo, unsigned hi, float3 pos, float3 dir, float dl,
                                              const SmallExact &X, const WoopRay &wr, int skip, Hit &h)
{
    const float INF = __int_as_float(0x7f800000);
    float k1 = INF, k2 = INF; // the two smallest t_lo ...
    int s1 = -1, s2 = -1;     // ... and their slots
    {
        unsigned mlo = lo, mhi = hi;
        while (mlo | mhi) {
            const int slot = pop_slot(mlo, mhi);
            float key = small_t_lower(S, slot, pos, dir, dl);
            if (SKIP && S.slot_tri[slot] == skip) key = INF;
            const bool lt1 = key < k1, lt2 = key < k2;
            s2 = lt1 ? s1 : (lt2 ? slot : s2);
            k2 = lt1 ? k1 : (lt2 ? key : k2);
            s1 = lt1 ? slot : s1;
            k1 = lt1 ? key : k1;
        }
    }
    int cur = (s1 >= 0 && k1 <= h.t) ? s1 : -1;
    int stage = 0;
    while (cur >= 0) {
        small_exact(X, S.slot_tri[cur], wr, h);
        if (cur < 32) lo &= ~(1u << cur); else hi &= ~(1u << (cur - 32));
        int next = -1;
        if (stage == 0) {
            if (s2 >= 0 && k2 <= h.t) next = s2;
        } else if (k2 <= h.t) {
            // rare: more than two candidates may reach below the best hit -> walk the rest with the bound test
            while (lo | hi) {
                const int slot = pop_slot(lo, hi);
                if (SKIP && S.slot_tri[slot] == skip) continue;
                if (small_t_lower(S, slot, pos, dir, dl) <= h.t) { next = slot; break; }
            }
        }
        stage = 1;
        cur = next;
    }
}

// WIDE: more than 16 pairs (the candidate mask needs a second word)
// (skip: a triangle the query ignores — physical mode's "the triangle just left"; -1 = none)
template <bool WIDE>
__device__ __forceinline__ void closest_hit_small(const SmallScene &S, float3 pos, float3 dir, const WoopRay &wr, Hit &h,
                                                  int skip = -1)
{
    // ---- phase 1: lockstep candidate filter ------------------------------------------------------------------
    // tolerance (metres) = 2^-16 * (extent + |p|_1) / sin(angle between ray and plane); see rfrt_small.cu
    const float dl = (S.extent + fabsf(pos.x) + fabsf(pos.y) + fabsf(pos.z)) * (1.0f / 65536.0f);
    const float dl_h = dl * (sqrt_approx(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z) * 1.001f);
    const int n_lo = WIDE ? 16 : S.n_pairs, n_hi = WIDE ? S.n_pairs - 16 : 0;
    const unsigned lo = sweep_pairs(S.recs, S.cls, 0, n_lo, pos, dir, dl, dl_h);
    const unsigned hi = WIDE ? sweep_pairs(S.recs, S.cls, 16, n_hi, pos, dir, dl, dl_h) : 0u;
    // ---- phase 2 ---------------------------------------------------------------------------------------------
    const SmallExact X = small_exact_setup(S, wr);
    if (skip >= 0) small_resolve<true>(S, lo, hi, pos, dir, dl, X, wr, skip, h);
    else small_resolve<false>(S, lo, hi, pos, dir, dl, X, wr, -1, h);
}

// pointers into the small-scene image staged at `img` (layout: rfrt_internal.h)
__device__ __forceinline__ SmallScene small_scene_view(const float *img, int n_pairs, int n_tris, const int *cls, float extent,
                                                       float tau, float erode)
{
    SmallScene S;
#pragma unroll
    for (int c = 0; c < 5; ++c) S.cls[c] = cls[c];
    S.recs = reinterpret_cast<const float4 *>(img);
    S.nbr = reinterpret_cast<const uint4 *>(img + 28 * n_pairs);
    S.slot_tri = reinterpret_cast<const int *>(img + 28 * n_pairs + 4 * n_tris);
    S.soup = img + 30 * n_pairs + 4 * n_tris;
    S.normals = img + 30 * n_pairs + 13 * n_tris;
    S.tri_slot = reinterpret_cast<const int *>(img + 30 * n_pairs + 16 * n_tris);
    S.n_pairs = n_pairs; S.extent = extent; S.tau = tau; S.erode = erode;
    return S;
}

// Self-re-hit shortcut.  The reference never offsets a reflected ray (kernel.py:94-96), so most segments that start on
// a surface hit that very triangle again at t ~ 0 (78 % of all room.stl segments).  For a ray standing on triangle f:
// run the exact test on f alone; if it is hit within tau (path length), only triangles that come within `reach` of f
// can beat or tie it, and only if they pass within rho = tau + tolerance of the ray origin.  The host lists their
// pairs per triangle (rfrt_small.cu: interior / boundary neighbours; boundary neighbours are skipped when the origin
// keeps a clearance from f's edges; for t == 0 only lower indices matter, since ties go to the lowest index).  The
// few triangles that survive the distance filter get the same exact test as in the full sweep, so the result is again
// bit-identical.  Returns false (h untouched) when f is not hit that close: the caller then runs the full sweep.
__device__ __forceinline__ bool small_self_rehit(const SmallScene &S, int f, float3 pos, float3 dir, const WoopRay &wr, Hit &h)
{
    const SmallExact X = small_exact_setup(S, wr);
    Hit hf;
    hf.t = 1.0e6f; hf.face = -1; hf.slot = -1;
    small_exact(X, f, wr, hf);
    const float dlen = sqrt_approx(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z) * 1.001f;
    if (!(hf.face >= 0 && hf.t * dlen <= S.tau)) return false;
    h = hf;
    const float dl = (S.extent + fabsf(pos.x) + fabsf(pos.y) + fabsf(pos.z)) * (1.0f / 65536.0f);
    const float rho = fmaf(S.tau, 1.01f, 4.0f * dl); // a winner passes this close to pos
    // clearance of pos from f's own edges (in-plane distances, positive inside)
    const int fs = S.tri_slot[f];
    const float4 *fe = S.recs + 7 * (fs >> 1) + 1 + 3 * (fs & 1);
    const float4 e0 = fe[0], e1 = fe[1], e2 = fe[2];
    const float clear = min3f(fmaf(e0.x, pos.x, fmaf(e0.y, pos.y, fmaf(e0.z, pos.z, e0.w))),
                              fmaf(e1.x, pos.x, fmaf(e1.y, pos.y, fmaf(e1.z, pos.z, e1.w))),
                              fmaf(e2.x, pos.x, fmaf(e2.y, pos.y, fmaf(e2.z, pos.z, e2.w))));
    const uint4 nb = S.nbr[f];
    // (inside f eroded by `erode`, with a slack of tau + dl for the fp32 clearance — its own rounding error is ~1e-6 *
    // extent, a twentieth of dl; the full rho of the distance filters below is not needed here)
    const bool inner = clear >= S.erode + fmaf(S.tau, 1.01f, dl); // false for NaN
    unsigned m = (hf.t == 0.0f) ? (inner ? nb.z : (nb.z | nb.w)) : (inner ? nb.x : (nb.x | nb.y));
    while (m) {
        const int k = 31 - __clz((int)m);
        m ^= 1u << k;
        const float4 *rec = S.recs + 7 * k;
        const float4 P = rec[0];
        const float np = fmaf(P.x, pos.x, fmaf(P.y, pos.y, fmaf(P.z, pos.z, -P.w)));
        if (fabsf(np) > rho) continue; // the pair's plane is too far (NaN stays)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float4 g0 = rec[1 + 3 * j], g1 = rec[2 + 3 * j], g2 = rec[3 + 3 * j];
            const float d = min3f(fmaf(g0.x, pos.x, fmaf(g0.y, pos.y, fmaf(g0.z, pos.z, g0.w))),
                                  fmaf(g1.x, pos.x, fmaf(g1.y, pos.y, fmaf(g1.z, pos.z, g1.w))),
                                  fmaf(g2.x, pos.x, fmaf(g2.y, pos.y, fmaf(g2.z, pos.z, g2.w))));
            if (d < -rho) continue; // pos lies farther than rho outside one edge (NaN stays)
            const int g = S.slot_tri[2 * k + j];
            if (g != f) small_exact(X, g, wr, h);
        }
    }
    return true;
}

// One receiver as seen by the kernels: its fp32 world-space vertices plus the shared unit-icosphere BVH.
struct RxView {
    const float *verts;          // [n_unit*3] world-space vertices of THIS receiver
    const BvhNode *unit_nodes;   // BVH over the unit icosphere's faces (shared by all receivers)
    const int32_t *unit_order;   // sorted slot -> face index
    float cx, cy, cz, inv_r;     // centre and 1/radius: world -> unit space (approximate, pruning only)
};

// Exact receiver query (kernel.py:71 against the 80-triangle icosphere of one receiver): closest t in
// [0, max_t) over ALL faces.  The faces actually tested are pruned with the unit-space BVH (boxes inflated by
// 4e-3 of the radius; t is the same parameter in both spaces); every triangle test itself is the exact
// world-space Woop test, so the result equals the brute-force minimum.
template <class RAY>
__device__ __forceinline__ bool rx_query(const RxView &rx, const uint8_t *faces, int n_faces, const RAY &wr,
                                         float3 pos, float3 dir, float max_t, int *stack, float *stack_t, int stride,
                                         float &t_out)
{
    const float3 ou = make_float3((pos.x - rx.cx) * rx.inv_r, (pos.y - rx.cy) * rx.inv_r, (pos.z - rx.cz) * rx.inv_r);
    const float3 du = make_float3(dir.x * rx.inv_r, dir.y * rx.inv_r, dir.z * rx.inv_r);
    const SlabRay su = slab_setup(ou, du);
    float best = max_t;
    int sp = 0;
    int node = 0;
    if (fmaxf(fmaxf(fabsf(ou.x), fabsf(ou.y)), fabsf(ou.z)) > 8192.0f) {
        // origin farther than 8192 radii: fp32 unit-space coordinates get too coarse for the 4e-3 box inflation,
        // so test every face (n_faces == number of unit-BVH leaves)
        node = TRAV_DONE;
        for (int f = 0; f < n_faces; ++f) {
            const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
            const float *v = rx.verts;
            float3 a = make_float3(__ldg(v + 3 * i0), __ldg(v + 3 * i0 + 1), __ldg(v + 3 * i0 + 2));
            float3 b = make_float3(__ldg(v + 3 * i1), __ldg(v + 3 * i1 + 1), __ldg(v + 3 * i1 + 2));
            float3 c = make_float3(__ldg(v + 3 * i2), __ldg(v + 3 * i2 + 1), __ldg(v + 3 * i2 + 2));
            float t;
            if (tri_hit(wr, a, b, c, t) && t < best && t >= 0.0f) best = t;
        }
    }
    while (node != TRAV_DONE) {
        while (node >= 0) node = node_step(rx.unit_nodes, node, su, best, stack, stack_t, stride, sp);
        if (node != TRAV_DONE) {
            const int f = __ldg(rx.unit_order + (~node));
            const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
            const float *v = rx.verts;
            float3 a = make_float3(__ldg(v + 3 * i0), __ldg(v + 3 * i0 + 1), __ldg(v + 3 * i0 + 2));
            float3 b = make_float3(__ldg(v + 3 * i1), __ldg(v + 3 * i1 + 1), __ldg(v + 3 * i1 + 2));
            float3 c = make_float3(__ldg(v + 3 * i2), __ldg(v + 3 * i2 + 1), __ldg(v + 3 * i2 + 2));
            float t;
            if (tri_hit(wr, a, b, c, t) && t < best && t >= 0.0f) best = t;
            node = stack_pop(stack, stack_t, stride, sp, best);
        }
    }
    t_out = best;
    return best < max_t;
}

// ---------------------------------------------------------------------------------------------------------
// Warp-cooperative enumeration of the receivers whose (padded) box a segment overlaps.
// Per-lane enumeration of a dense receiver lattice leaves 5 of 32 lanes busy (each crossing segment walks ~300 nodes
// of the receiver BVH on its own).  Here the warp takes its lanes' segments one at a time and walks the receiver BVH
// together: a per-warp node queue in shared memory, every lane pops one node per step (last-in first-out batches of 32,
// which bounds the queue by ~32 x depth), tests its two child boxes, pushes the internal children (ballot + popc
// offsets) and hands leaf receivers to `leaf`.  Must be called by all 32 lanes of a converged warp.
//   want        this lane has a segment (pos, dir, t_limit) to enumerate
//   leaf(k, src) is called by SOME lane for every receiver k whose box the segment of lane `src` overlaps; it has to
//               fetch the segment's data with __shfl_sync-free means, so callers broadcast what they need through
//               `bcast` (called by all lanes with the source lane before the walk of each segment)
// Returns false if the queue overflowed (results incomplete: the caller counts that in RFRT_CTR_QUEUE_OVERFLOW).
// ---------------------------------------------------------------------------------------------------------
constexpr int RX_QUEUE_CAP = 1024; // ints per warp

//   flush(final) is called by all lanes at converged points: after every step of a walk (final = false) and after the
//               last step of a segment (final = true) — the place to drain what `leaf` has staged for this segment
template <class Bcast, class Leaf, class Flush>
__device__ __forceinline__ bool rx_enumerate_coop(const BvhNode *__restrict__ rx_nodes, const int32_t *__restrict__ rx_order,
                                                  bool want, int *queue, Bcast &&bcast, Leaf &&leaf, Flush &&flush)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    bool ok = true;
    unsigned todo = __ballot_sync(FULL, want);
    while (todo) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        float3 pos, dir; float t_limit;
        bcast(src, pos, dir, t_limit); // every lane now holds lane src's segment
        const RxSlabRay sr = rx_slab_setup(pos, dir);
        int n = 1;
        if (lane == 0) queue[0] = 0;
        __syncwarp();
        while (n > 0) {
            const int batch = n < 32 ? n : 32;
            int node = -1;
            if (lane < batch) node = queue[n - 1 - lane];
            n -= batch;
            __syncwarp();
            bool p0 = false, p1 = false;
            int c0 = 0, c1 = 0;
            if (node >= 0) {
                const float4 *np = reinterpret_cast<const float4 *>(rx_nodes + node);
                const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
                const int4 q3 = __ldg(reinterpret_cast<const int4 *>(np + 3));
                float tn0, tn1;
                bool h0 = rx_slab_hit(sr, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, t_limit, tn0);
                bool h1 = rx_slab_hit(sr, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, t_limit, tn1);
                c0 = q3.x; c1 = q3.y;
                if (c1 == c0) h1 = false;
                p0 = h0 && c0 >= 0; p1 = h1 && c1 >= 0;
                if (h0 && c0 < 0) leaf(__ldg(rx_order + (~c0)), pos, dir, t_limit);
                if (h1 && c1 < 0) leaf(__ldg(rx_order + (~c1)), pos, dir, t_limit);
            }
            const unsigned b0 = __ballot_sync(FULL, p0), b1 = __ballot_sync(FULL, p1);
            const int off0 = n + __popc(b0 & lt), off1 = n + __popc(b0) + __popc(b1 & lt);
            if (p0 && off0 < RX_QUEUE_CAP) queue[off0] = c0;
            if (p1 && off1 < RX_QUEUE_CAP) queue[off1] = c1;
            n += __popc(b0) + __popc(b1);
            if (n > RX_QUEUE_CAP) { n = RX_QUEUE_CAP; ok = false; }
            __syncwarp();
            flush(n == 0);
        }
    }
    return ok;
}

// Receiver query of the replay kernel, lockstep: the same answer as rx_query (closest exact t in [0, max_t) over all
// faces of one receiver), but the faces to test are found by the plane / edge-distance filter (see rfrt_small.cu) over
// the unit shape's face records in shared memory — every lane runs the same n_faces plane steps (stage A), then checks
// the edge distances of the handful of faces whose plane is crossed inside the receiver's sphere (stage B), where the
// unit-BVH walk left 7.7 of 32 lanes active.  The ray is mapped into unit space in fp32, so the tolerance also carries the rounding of
// that mapping (2^-21 * (|c|_1 + |p|_1) / r).
//
// RxFaceCache: the candidate faces are a property of the LINE.  A receiver hit moves the origin along the ray and keeps
// the direction (kernel.py:87), so the queries that follow it — the exit through the far side, the t ~ 0 repeats — run
// on the same line up to the rounding of `advance` (one ulp of the position, an eighth of the mapping tolerance above):
// a face the filter dropped for the first origin (plane crossed outside the chord of the inflated sphere, behind the
// origin, or outside the face by more than the tolerance) stays dropped for every later origin on that line, because
// the later origins lie farther along the ray (what was behind stays behind) and no farther from the receiver than the
// first origin or the receiver's own sphere (their tolerance is no larger than the one taken here).  The first query
// of a line therefore stores its candidate masks — computed with TWICE that tolerance — and the later ones run only
// the exact tests on them (4-5 queries per received pair on a dense lattice:
// the filter stages were 2/3 of the replay kernel's instructions).  The caller invalidates the cache whenever the
// direction changes.
struct RxFaceCache {
    unsigned m[4]; // candidate faces, 32 per word (n_faces <= 128)
    bool valid;
};

__device__ __forceinline__ bool rx_query_sweep(const RxView &rx, const float4 *s_recs, const uint8_t *faces, int n_faces,
                                               const WoopRay &wr, float3 pos, float3 dir, float max_t, float &t_out,
                                               RxFaceCache *cache = nullptr)
{
    unsigned masks[4] = {0u, 0u, 0u, 0u};
    if (cache && cache->valid) {
#pragma unroll
        for (int w = 0; w < 4; ++w) masks[w] = cache->m[w];
    } else {
        const float slack = cache ? 2.0f : 1.0f;
        const float3 ou = make_float3((pos.x - rx.cx) * rx.inv_r, (pos.y - rx.cy) * rx.inv_r, (pos.z - rx.cz) * rx.inv_r);
        const float3 du = make_float3(dir.x * rx.inv_r, dir.y * rx.inv_r, dir.z * rx.inv_r);
        const float far = fmaxf(fmaxf(fabsf(ou.x), fabsf(ou.y)), fabsf(ou.z));
        const bool all_faces = !(far <= 8192.0f); // origin farther than 8192 radii (or NaN): unit space too coarse
        // (with a cache: later origins of the line lie on or inside the inflated sphere, |ou|_1 <= 1.75, or between
        // the first origin and the sphere)
        const float ou1 = fabsf(ou.x) + fabsf(ou.y) + fabsf(ou.z);
        const float dl = slack * ((1.0f + (cache ? fmaxf(ou1, 1.75f) : ou1)) * (1.0f / 65536.0f) +
                                  (fabsf(rx.cx) + fabsf(rx.cy) + fabsf(rx.cz) + fabsf(pos.x) + fabsf(pos.y) + fabsf(pos.z)) * rx.inv_r * (1.0f / 2097152.0f));
        const float dd = du.x * du.x + du.y * du.y + du.z * du.z;
        const float dl_h = dl * (sqrt_approx(dd) * 1.001f);
        // Stage A works on the chord of the unit sphere inflated by 1 % (+ tolerance): the shape lies inside the unit
        // sphere, so a face can only be hit where its plane is crossed inside that sphere, and not behind the origin.
        // (Chord through the closest-approach point: no cancellation for far origins.)
        const float inv_dd = rcp_approx(dd);
        const float tc = -(ou.x * du.x + ou.y * du.y + ou.z * du.z) * inv_dd;
        const float mx = fmaf(tc, du.x, ou.x), my = fmaf(tc, du.y, ou.y), mz = fmaf(tc, du.z, ou.z);
        const float rho2 = 1.0201f + 4.0f * dl + 1.0e-3f * far * (1.0f / 8192.0f);
        const float half2 = (rho2 - (mx * mx + my * my + mz * mz)) * inv_dd;
        const float half = sqrt_approx(half2);
        // no hit is possible when the line misses the inflated sphere, or when the whole chord lies more than 0.05 radii
        // behind the origin (hits need t >= 0); NaN falls through to the sweep
        const bool can_hit = all_faces || !(half2 < 0.0f || (tc + half) * sqrt_approx(dd) < -0.05f);
        if (can_hit) {
            const float lo0 = all_faces ? -3.0e38f : fmaxf(tc - half, 0.0f), hi0 = all_faces ? 3.0e38f : tc + half;
#pragma unroll 1
            for (int f0 = 0; f0 < n_faces; f0 += 32) {
                const int cnt = n_faces - f0 < 32 ? n_faces - f0 : 32;
                // Stage A (lockstep, plane parameter only) for faces f0 .. f0 + cnt - 1
                unsigned dropped = 0u;
                const float4 *rec = s_recs + 4 * (f0 + cnt - 1);
#pragma unroll 4
                for (int k = 0; k < cnt; ++k, rec -= 4) { // backwards: face f0 + j ends up at bit j
                    const float4 P = rec[0];
                    const float nd = fmaf(P.x, du.x, fmaf(P.y, du.y, P.z * du.z));
                    const float np = fmaf(P.x, ou.x, fmaf(P.y, ou.y, fmaf(P.z, ou.z, -P.w)));
                    const float r = rcp_approx(nd);
                    const float t = -np * r;
                    const float kt = dl * fabsf(r);
                    // sign set <=> t + kt < lo0 or t - kt > hi0 (NaN operands are ignored by min: kept)
                    dropped = __funnelshift_l(__float_as_uint(fminf((t + kt) - lo0, hi0 - (t - kt))), dropped, 1);
                }
                unsigned m = ~dropped & (cnt >= 32 ? 0xffffffffu : (1u << cnt) - 1u);
                // Stage B (per lane, the few faces left): in-plane edge distances of the plane hit point
                if (!all_faces) {
                    unsigned mb = m;
                    while (mb) {
                        const int b = __ffs((int)mb) - 1;
                        mb &= mb - 1u;
                        const float4 *fr = s_recs + 4 * (f0 + b);
                        const float4 P = fr[0], e0 = fr[1], e1 = fr[2], e2 = fr[3];
                        const float nd = fmaf(P.x, du.x, fmaf(P.y, du.y, P.z * du.z));
                        const float np = fmaf(P.x, ou.x, fmaf(P.y, ou.y, fmaf(P.z, ou.z, -P.w)));
                        const float r = rcp_approx(nd);
                        const float t = -np * r;
                        const float thr = -(dl_h * fabsf(r));
                        const float hx = fmaf(t, du.x, ou.x), hy = fmaf(t, du.y, ou.y), hz = fmaf(t, du.z, ou.z);
                        const float d0 = fmaf(e0.x, hx, fmaf(e0.y, hy, fmaf(e0.z, hz, e0.w)));
                        const float d1 = fmaf(e1.x, hx, fmaf(e1.y, hy, fmaf(e1.z, hz, e1.w)));
                        const float d2 = fmaf(e2.x, hx, fmaf(e2.y, hy, fmaf(e2.z, hz, e2.w)));
                        if (min3f(d0, d1, d2) < thr) m &= ~(1u << b); // definitely outside the face (false for NaN: kept)
                    }
                }
                masks[f0 >> 5] = m;
            }
        }
        if (cache) {
#pragma unroll
            for (int w = 0; w < 4; ++w) cache->m[w] = masks[w];
            cache->valid = true;
        }
    }
    // exact test on the candidates (a separate loop: the lanes that still hold a face run it together; ONE copy of the
    // test in the instruction stream — unrolled per mask word the kernel outgrew the instruction cache)
    float best = max_t;
    unsigned m0 = masks[0], m1 = masks[1], m2 = masks[2], m3 = masks[3];
#pragma unroll 1
    while (m0 | m1 | m2 | m3) {
        int f;
        if (m0) { f = __ffs((int)m0) - 1; m0 &= m0 - 1u; }
        else if (m1) { f = 31 + __ffs((int)m1); m1 &= m1 - 1u; }
        else if (m2) { f = 63 + __ffs((int)m2); m2 &= m2 - 1u; }
        else { f = 95 + __ffs((int)m3); m3 &= m3 - 1u; }
        const float *v = rx.verts;
        float t;
        if (woop_hit_mem(wr, v + 3 * faces[3 * f], v + 3 * faces[3 * f + 1], v + 3 * faces[3 * f + 2], t) && t < best && t >= 0.0f)
            best = t;
    }
    t_out = best;
    return best < max_t;
}

} // namespace rfrt
