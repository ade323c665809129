// rfrt_phys.cu — "physical" mode of the trace: the same deterministic rays and exact
// closest-hit arithmetic as the reference path, without its quirks and with explicit path loss, reflection
// coefficients and carrier phase:
//   * the triangle a ray has just left is excluded from its next query (no t ~ 0 re-hits: kernel.py:94-96 quirk Q3);
//   * receivers are analytic spheres; a segment whose closest approach to a centre lies inside the sphere (and that
//     starts outside it) is ONE arrival, the ray itself goes on unchanged (quirks Q1, Q2, Q4, Q12);
//   * field of an arrival with unfolded path length L (unit transmit power, isotropic antennas):
//         E = (L * lambda / (pi * N * r^2)) * prod_i Gamma_i * exp(-j 2 pi L / lambda)
//     = free-space loss lambda / (4 pi L) x reception-sphere weight 4 L^2 / (N r^2) (a wavefront of N rays puts
//     N r^2 / (4 L^2) of them through a sphere of radius r at range L); Gamma is the p-polarised Fresnel amplitude
//     coefficient of the hit triangle's refractive index (tracer.py:43-53 uses its square with n = 5 for everything);
//   * outputs: per receiver the coherent sum of E (grid power = P_tx * |sum|^2), optionally a complex impulse
//     response ir[k][bin] += E, bin = int(L / c * rate) (tracer.py:115).
// No replay, no 10 000-bin rows, no convolution: one pass of k_trace_phys fills the whole coverage grid.
#include "rfrt_trace.cuh"

namespace rfrt {
namespace {

constexpr int PHYS_THREADS = 128;

struct PhysParams {
    const BvhNode *nodes;
    const BvhTri *tris;
    const float4 *normals; // sorted order
    int64_t n_tris;
    const float *materials; // [n_tris] original order, or NULL (5.0)
    const BvhNode *rx_nodes; // BVH over the receivers' bounding cubes (centre +- radius)
    const int32_t *rx_order;
    const double *rx_centers;
    int64_t n_rx;
    double r2, wk, two_pi_over_lambda, light_speed, sample_rate;
    float3 tx;
    int32_t max_bounces;
    int64_t chunk_begin, chunk_n;
    const float4 *dirs;
    unsigned long long *counters;
    double *field; // [n_rx*2]
    double *ir;    // [n_rx*n_bins*2] or NULL
    int64_t n_bins;
    int32_t stack_depth;
    int32_t rx_coop; // dense receiver sets: warp-cooperative enumeration (see rfrt_trace.cu)
    const float *small;  // small scenes: shared-memory image of the lockstep sweep (rfrt_small.cu), else NULL
    const float *face_normals; // [n_tris*3] original order
    int32_t small_pairs;
    int32_t small_class[5];
    float small_extent;
};

// one receiver against one segment: fp64 closest-approach test + field accumulation
__device__ __forceinline__ void phys_arrival(const PhysParams &P, int64_t k, float3 pos, float3 dir, double dd, double dlen,
                                             double t_lim, double L, double gamma, unsigned &n_arr)
{
    const double ox = __dsub_rn(__ldg(P.rx_centers + 3 * k), (double)pos.x), oy = __dsub_rn(__ldg(P.rx_centers + 3 * k + 1), (double)pos.y),
                 oz = __dsub_rn(__ldg(P.rx_centers + 3 * k + 2), (double)pos.z);
    const double oo = __dadd_rn(__dadd_rn(__dmul_rn(ox, ox), __dmul_rn(oy, oy)), __dmul_rn(oz, oz));
    if (oo <= P.r2) return; // the segment starts inside the sphere: no new arrival
    const double od = __dadd_rn(__dadd_rn(__dmul_rn(ox, (double)dir.x), __dmul_rn(oy, (double)dir.y)), __dmul_rn(oz, (double)dir.z));
    const double tc = __ddiv_rn(od, dd);
    if (!(tc >= 0.0) || !(tc <= t_lim)) return;
    const double perp2 = __dsub_rn(oo, __dmul_rn(tc, od));
    if (!(perp2 <= P.r2)) return;
    const double Lk = __dadd_rn(L, __dmul_rn(tc, dlen));
    const double a = __dmul_rn(__dmul_rn(Lk, P.wk), gamma);
    double sn, cs;
    sincos(__dmul_rn(P.two_pi_over_lambda, Lk), &sn, &cs);
    const double re = a * cs, im = -(a * sn);
    atomicAdd(P.field + 2 * k, re);
    atomicAdd(P.field + 2 * k + 1, im);
    if (P.ir) {
        const long long bin = (long long)__dmul_rn(__ddiv_rn(Lk, P.light_speed), P.sample_rate);
        if (bin >= 0 && bin < P.n_bins) {
            atomicAdd(P.ir + 2 * (k * P.n_bins + bin), re);
            atomicAdd(P.ir + 2 * (k * P.n_bins + bin) + 1, im);
        }
    }
    ++n_arr;
}

// SMALL: 0 = BVH walk, 1 = lockstep sweep of <= 16 triangle pairs, 2 = of <= 32 pairs (closest_hit_small with the
//        triangle just left excluded from the exact tests)
template <bool LSTACK, int SMALL>
__global__ void __launch_bounds__(PHYS_THREADS) k_trace_phys(const PhysParams P)
{
    extern __shared__ __align__(16) int s_stack_raw[];
    int l_stack[LSTACK ? 64 : 1];
    float l_stack_t[LSTACK ? 64 : 1];
    int *stack = LSTACK ? l_stack : s_stack_raw + threadIdx.x;
    float *stack_t = LSTACK ? l_stack_t : reinterpret_cast<float *>(s_stack_raw + P.stack_depth * PHYS_THREADS) + threadIdx.x;
    constexpr int STRIDE = LSTACK ? 1 : PHYS_THREADS;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    int *rx_queue = s_stack_raw + 2 * P.stack_depth * PHYS_THREADS + RX_QUEUE_CAP * (threadIdx.x >> 5); // this warp's node queue
    SmallScene S;
    if (SMALL) { // the scene image lives behind the stacks and (dense receiver sets) the node queues
        float *img = reinterpret_cast<float *>(s_stack_raw + 2 * P.stack_depth * PHYS_THREADS + (P.rx_coop ? RX_QUEUE_CAP * (PHYS_THREADS / 32) : 0));
        const int n = (int)P.n_tris, np = P.small_pairs;
        for (int i = threadIdx.x; i < 30 * np + 17 * n; i += PHYS_THREADS) img[i] = __ldg(P.small + i);
        __syncthreads();
        S = small_scene_view(img, np, n, P.small_class, P.small_extent, 0.0f, 0.0f);
    }

    bool has_ray = false, exhausted = false;
    float3 pos = make_float3(0.f, 0.f, 0.f), dir = make_float3(0.f, 0.f, 1.f);
    int bounce = 0, prev = -1;
    double L = 0.0, gamma = 1.0;
    unsigned n_seg = 0, n_hit = 0, n_arr = 0;

    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !has_ray);
        if (idle != 0u && !exhausted) {
            const int cnt = __popc(idle);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&P.counters[RFRT_CTR_NEXT_RAY], (unsigned long long)cnt);
            base = __shfl_sync(FULL, base, 0);
            if (!has_ray) {
                const int64_t r = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
                if (r < P.chunk_n) {
                    const float4 d4 = __ldg(P.dirs + r);
                    dir = make_float3(d4.x, d4.y, d4.z);
                    pos = P.tx;
                    bounce = 0; prev = -1; L = 0.0; gamma = 1.0;
                    has_ray = true;
                }
            }
            if ((int64_t)base + cnt >= P.chunk_n) exhausted = true;
        }
        if (!__any_sync(FULL, has_ray)) break;
        Hit h;
        h.t = 1.0e6f; h.face = -1; h.slot = -1;
        float dlen_f = 1.0f;
        if (has_ray) {
            const WoopRay wr = woop_setup(pos, dir);
            if (SMALL) {
                closest_hit_small<SMALL == 2>(S, pos, dir, wr, h, prev);
            } else {
                const SlabRay sr = slab_setup(pos, dir);
                closest_hit(P.nodes, P.tris, P.n_tris, wr, sr, stack, stack_t, STRIDE, h, prev);
            }
            ++n_seg;
            dlen_f = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dir.x, dir.x), __fmul_rn(dir.y, dir.y)), __fmul_rn(dir.z, dir.z)));
        }
        const bool hit = h.face >= 0;
        const double dlen = (double)dlen_f;
        if (P.n_rx > 0 && P.rx_coop) {
            // dense receiver sets: the warp enumerates, segment by segment, the receivers whose bounding cube overlaps
            // [0, t_limit] (rx_enumerate_coop)
            const float t_limit = hit ? h.t : 1.0e6f;
            double L_b = 0.0, gamma_b = 1.0, dlen_b = 1.0;
            const bool ok = rx_enumerate_coop(
                P.rx_nodes, P.rx_order, has_ray, rx_queue,
                [&](int src, float3 &bp, float3 &bd, float &bt) {
                    bp.x = __shfl_sync(FULL, pos.x, src); bp.y = __shfl_sync(FULL, pos.y, src); bp.z = __shfl_sync(FULL, pos.z, src);
                    bd.x = __shfl_sync(FULL, dir.x, src); bd.y = __shfl_sync(FULL, dir.y, src); bd.z = __shfl_sync(FULL, dir.z, src);
                    bt = __shfl_sync(FULL, t_limit, src);
                    L_b = __shfl_sync(FULL, L, src); gamma_b = __shfl_sync(FULL, gamma, src); dlen_b = __shfl_sync(FULL, dlen, src);
                },
                [&](int k, float3 bp, float3 bd, float bt) {
                    const double dd = __dadd_rn(__dadd_rn(__dmul_rn((double)bd.x, (double)bd.x), __dmul_rn((double)bd.y, (double)bd.y)),
                                                __dmul_rn((double)bd.z, (double)bd.z));
                    phys_arrival(P, k, bp, bd, dd, dlen_b, bt < 1.0e6f ? (double)bt : 1.0e6, L_b, gamma_b, n_arr);
                },
                [](bool) {});
            if (!ok && lane == 0) atomicAdd(&P.counters[RFRT_CTR_QUEUE_OVERFLOW], 1ull);
        } else if (P.n_rx > 0 && has_ray) {
            // sparse receiver sets: every lane walks the receiver BVH for its own segment
            const double dd = __dadd_rn(__dadd_rn(__dmul_rn((double)dir.x, (double)dir.x), __dmul_rn((double)dir.y, (double)dir.y)),
                                        __dmul_rn((double)dir.z, (double)dir.z));
            const double t_lim = hit ? (double)h.t : 1.0e6;
            const float t_limit = hit ? h.t : 1.0e6f;
            const RxSlabRay sr = rx_slab_setup(pos, dir);
            int sp = 0;
            int node = 0;
            while (node >= 0) {
                const float4 *np = reinterpret_cast<const float4 *>(P.rx_nodes + node);
                const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
                const int4 q3 = __ldg(reinterpret_cast<const int4 *>(np + 3));
                float tn0, tn1;
                bool h0 = rx_slab_hit(sr, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, t_limit, tn0);
                bool h1 = rx_slab_hit(sr, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, t_limit, tn1);
                const int c0 = q3.x, c1 = q3.y;
                if (c1 == c0) h1 = false;
                if (h0) {
                    if (c0 < 0) phys_arrival(P, __ldg(P.rx_order + (~c0)), pos, dir, dd, dlen, t_lim, L, gamma, n_arr);
                    else { stack[sp * STRIDE] = c0; ++sp; }
                }
                if (h1) {
                    if (c1 < 0) phys_arrival(P, __ldg(P.rx_order + (~c1)), pos, dir, dd, dlen, t_lim, L, gamma, n_arr);
                    else { stack[sp * STRIDE] = c1; ++sp; }
                }
                node = -1;
                if (sp > 0) { --sp; node = stack[sp * STRIDE]; }
            }
        }
        if (has_ray) {
            if (hit) {
                ++n_hit;
                float3 nrm; // normalize(cross(b-a, c-a)) of the hit triangle, precomputed at build time
                if (SMALL) {
                    nrm = make_float3(S.normals[3 * h.face], S.normals[3 * h.face + 1], S.normals[3 * h.face + 2]);
                } else {
                    const float4 n4 = __ldg(P.normals + h.slot);
                    nrm = make_float3(n4.x, n4.y, n4.z);
                }
                const float dn = __fadd_rn(__fadd_rn(__fmul_rn(dir.x, nrm.x), __fmul_rn(dir.y, nrm.y)), __fmul_rn(dir.z, nrm.z));
                const double nmat = P.materials ? (double)__ldg(P.materials + h.face) : 5.0;
                double ci = __ddiv_rn(fabs((double)dn), dlen);
                if (ci > 1.0) ci = 1.0;
                const double si2 = __dsub_rn(1.0, __dmul_rn(ci, ci));
                const double ct = __dsqrt_rn(__dsub_rn(1.0, __ddiv_rn(si2, __dmul_rn(nmat, nmat))));
                gamma = __dmul_rn(gamma, __ddiv_rn(__dsub_rn(ct, __dmul_rn(nmat, ci)), __dadd_rn(ct, __dmul_rn(nmat, ci))));
                L = __dadd_rn(L, __dmul_rn((double)h.t, dlen));
                pos = advance(pos, dir, h.t);
                dir = reflect(dir, nrm);
                prev = h.face;
                ++bounce;
                if (bounce >= P.max_bounces) has_ray = false;
            } else {
                has_ray = false;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_seg += __shfl_xor_sync(FULL, n_seg, o);
        n_hit += __shfl_xor_sync(FULL, n_hit, o);
        n_arr += __shfl_xor_sync(FULL, n_arr, o);
    }
    if (lane == 0) {
        atomicAdd(&P.counters[RFRT_CTR_SEGMENTS], (unsigned long long)n_seg);
        atomicAdd(&P.counters[RFRT_CTR_ENV_HITS], (unsigned long long)n_hit);
        atomicAdd(&P.counters[RFRT_CTR_RECORDS], (unsigned long long)n_arr);
    }
}

__global__ void k_phys_gen_dirs(int64_t ray_begin, int64_t n, float4 *__restrict__ dirs)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 d = ray_direction((uint32_t)(ray_begin + i));
    dirs[i] = make_float4(d.x, d.y, d.z, 0.0f);
}

} // namespace
} // namespace rfrt

using namespace rfrt;

extern "C" int rfrt_trace_physical(rfrt_handle env_mesh, rfrt_handle rxset, const float *h_tx_pos, int32_t max_bounces,
                                   int64_t ray_begin, int64_t ray_end, int64_t n_rays_total, double carrier_hz,
                                   double light_speed_mps, double sample_rate_hz, int64_t n_bins,
                                   const float *d_materials, float *d_dir_scratch, int64_t chunk_rays,
                                   uint64_t *d_counters, double *d_field, double *d_ir, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Mesh *m = get_mesh(env_mesh);
    if (!m) { set_error("rfrt_trace_physical: unknown environment mesh handle"); return RFRT_ERR_HANDLE; }
    if (m->tri_test != RFRT_TRI_TEST_WOOP) { set_error("rfrt_trace_physical: only the watertight triangle test is supported"); return RFRT_ERR_INVALID; }
    RxSet *r = nullptr;
    if (rxset) {
        r = get_rxset(rxset);
        if (!r) { set_error("rfrt_trace_physical: unknown receiver set handle"); return RFRT_ERR_HANDLE; }
        if (!d_field) { set_error("rfrt_trace_physical: d_field required with receivers"); return RFRT_ERR_INVALID; }
    }
    const int64_t n = ray_end - ray_begin;
    if (!h_tx_pos || !d_counters || max_bounces < 0 || n < 0 || ray_begin < 0 || ray_end > (1ll << 32) || n_rays_total <= 0 ||
        !(carrier_hz > 0.0) || !(light_speed_mps > 0.0) || (d_ir && (n_bins <= 0 || !(sample_rate_hz > 0.0)))) {
        set_error("rfrt_trace_physical: bad arguments");
        return RFRT_ERR_INVALID;
    }
    if (n == 0 || max_bounces == 0) return RFRT_OK;
    if (!d_dir_scratch) { set_error("rfrt_trace_physical: d_dir_scratch required"); return RFRT_ERR_INVALID; }
    if (chunk_rays <= 0) chunk_rays = 1ll << 24;

    PhysParams P;
    P.nodes = m->bvh.nodes; P.tris = m->tris; P.normals = m->normals; P.n_tris = m->bvh.n_prims;
    P.small = (m->small && m->bvh.n_prims > 0) ? m->small : nullptr; P.face_normals = m->face_normals;
    P.small_pairs = m->small_pairs; P.small_extent = m->small_extent;
    for (int c = 0; c < 5; ++c) P.small_class[c] = m->small_class[c];
    P.materials = d_materials;
    P.rx_nodes = r ? r->bvh.nodes : nullptr; P.rx_order = r ? r->bvh.prim_order : nullptr;
    P.rx_centers = r ? r->centers : nullptr; P.n_rx = r ? r->n_receivers : 0;
    const double lambda = light_speed_mps / carrier_hz;
    const double radius = r ? r->radius : 1.0;
    P.r2 = radius * radius;
    P.wk = lambda / (3.141592653589793 * (double)n_rays_total * (radius * radius));
    P.two_pi_over_lambda = (2.0 * 3.141592653589793) / lambda;
    P.light_speed = light_speed_mps; P.sample_rate = sample_rate_hz;
    P.tx = make_float3(h_tx_pos[0], h_tx_pos[1], h_tx_pos[2]);
    P.max_bounces = max_bounces;
    P.dirs = (const float4 *)d_dir_scratch;
    P.counters = (unsigned long long *)d_counters;
    P.field = d_field; P.ir = d_ir; P.n_bins = n_bins;
    // receiver enumeration: per lane (stack) for sparse sets, warp-cooperative (queue) for dense ones (rfrt_trace.cu)
    P.rx_coop = 0;
    if (r && r->n_receivers > 1) {
        const double ex = r->bvh.bounds[3] - r->bvh.bounds[0], ey = r->bvh.bounds[4] - r->bvh.bounds[1], ez = r->bvh.bounds[5] - r->bvh.bounds[2];
        double face = ex * ey > ey * ez ? ex * ey : ey * ez;
        if (ex * ez > face) face = ex * ez;
        const double d2 = 4.0 * r->radius * r->radius;
        P.rx_coop = (double)r->n_receivers * d2 >= 8.0 * (face > d2 ? face : d2) ? 1 : 0;
    }
    int depth = P.small ? 0 : m->bvh.max_depth; // (small scenes: lockstep sweep, the stack only serves the receiver walk)
    if (r && !P.rx_coop && r->bvh.max_depth > depth) depth = r->bvh.max_depth;
    depth += 2;
    if (depth < 8) depth = 8;
    const bool lstack = depth > 16;
    if (lstack && depth > 64) { set_error("rfrt_trace_physical: BVH deeper than 64 levels"); return RFRT_ERR_INVALID; }
    P.stack_depth = lstack ? 0 : depth;
    const size_t smem = (size_t)P.stack_depth * PHYS_THREADS * 2 * sizeof(int) +
                        (P.rx_coop ? sizeof(int) * RX_QUEUE_CAP * (PHYS_THREADS / 32) : 0) +
                        (P.small ? sizeof(float) * small_image_floats(m->small_pairs, (int)P.n_tris) : 0);
    typedef void (*kern_t)(const PhysParams);
    const int sv = P.small ? (m->small_pairs > 16 ? 2 : 1) : 0;
    static const kern_t kerns[2][3] = {{k_trace_phys<false, 0>, k_trace_phys<false, 1>, k_trace_phys<false, 2>},
                                       {k_trace_phys<true, 0>, k_trace_phys<true, 1>, k_trace_phys<true, 2>}};
    const kern_t kfn = kerns[lstack ? 1 : 0][sv];
    const void *kern = (const void *)kfn;
    int dev = 0, sms = 0, per_sm = 0;
    RFRT_CUDA(cudaGetDevice(&dev));
    RFRT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (smem > 48 * 1024) RFRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RFRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PHYS_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const int grid = sms * per_sm;
    for (int64_t c0 = ray_begin; c0 < ray_end; c0 += chunk_rays) {
        const int64_t cn = ray_end - c0 < chunk_rays ? ray_end - c0 : chunk_rays;
        k_phys_gen_dirs<<<(unsigned)((cn + 255) / 256), 256, 0, stream>>>(c0, cn, (float4 *)d_dir_scratch);
        RFRT_CUDA(cudaMemsetAsync(d_counters + RFRT_CTR_NEXT_RAY, 0, sizeof(uint64_t), stream));
        P.chunk_begin = c0; P.chunk_n = cn;
        int g = grid;
        const int64_t need = (cn + PHYS_THREADS - 1) / PHYS_THREADS;
        if (need < g) g = (int)need;
        kfn<<<g, PHYS_THREADS, smem, stream>>>(P);
    }
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}
