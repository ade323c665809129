// rfrt_trace.cu — the trace kernels (sm_100a) and their C-ABI launchers.
//
//   k_gen_dirs      kernel.py:51-52  directions of a chunk of rays (fp64 deterministic math, no divergence)
//   k_trace_small   kernel.py:57-98  scenes of <= 64 filter slots, staged in shared memory: persistent kernel, every
//                                    lane owns several rays and runs one bounce iteration of one of them per loop
//                                    trip (lockstep sweep or self-re-hit test, whichever more lanes can join)
//   k_trace_walk    kernel.py:57-98  BVH scenes: persistent kernel, every lane owns one ray and keeps its traversal state
//                                    across the phases of the loop: lanes whose walk has ended finish their segment,
//                                    start the next one or take a fresh ray (warp-aggregated fetch: ballot + popc + one
//                                    atomic per block of rays) while the other lanes stay in the middle of theirs
//   k_trace_receive kernel.py:38-98  literal replay for the rare (ray, receiver) candidates + the per-path
//                                    post-processing of tracer.py:102-115
//   k_trace_compat  kernel.py:38-98  the reference kernel's dense contract (tracer.py:75-79)
//   k_query         test probe for closest_hit
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "rfrt_trace.cuh"

namespace rfrt {

__constant__ uint8_t c_rx_faces[3 * 128];

namespace {

constexpr int TRACE_THREADS = 128;
constexpr int RX_CAND_BUF = 256; // receivers staged per warp between two appends (cooperative enumeration)
constexpr int RX_COOP_INTS = RX_QUEUE_CAP + RX_CAND_BUF + 4; // per-warp shared memory of the cooperative enumeration
constexpr int MAX_RECV_BOUNCES = 32;
#ifndef RECV_MIN_CTAS
#define RECV_MIN_CTAS 8
#endif
#ifndef WALK_MIN_CTAS
#define WALK_MIN_CTAS 9
#endif

struct TraceParams {
    const BvhNode *nodes;
    const BvhTri *tris;
    const float4 *normals;      // [n_tris] sorted order
    const float *small;         // small scenes: shared-memory image (rfrt_internal.h), else NULL
    int32_t small_pairs;
    int32_t small_class[5];
    float small_extent;
    int64_t n_tris;
    // receivers
    const BvhNode *rx_nodes;
    const int32_t *rx_order;
    const float *rx_verts;
    const double *rx_centers;
    int64_t n_rx;
    float rx_lo[3], rx_hi[3];   // padded bounds of all receivers (cheap pre-test of a segment's box)
    float env_lo[3], env_hi[3]; // padded bounds of the environment (BVH scenes: tested before every walk)
    int32_t n_unit;
    int32_t n_faces;
    float rx_radius;
    // rays
    float3 tx;
    int32_t max_bounces;
    int64_t chunk_begin; // global id of the first ray of this chunk
    int64_t chunk_n;
    const float4 *dirs;
    const uint64_t *order; // BVH scenes: (direction cell << 32 | ray index in chunk), sorted -> coherent warps; or NULL
    const uint32_t *order32; // ... or as 32-bit ray numbers (counting sort by direction cell)
    // outputs
    unsigned long long *counters;
    uint4 *candidates;
    int64_t cand_capacity;
    int32_t *hit_tri; // dense dumps, row = ray - dump_begin
    float *hit_t;
    int64_t dump_begin;
    int32_t stack_depth;
    int32_t fetch_block; // rays a warp claims per atomic (32..256, sized so that every warp sees >= 64 blocks)
    int32_t walk_refill;   // k_trace_walk: lanes waiting for phase A that end phase B
    int32_t walk_node_min; // k_trace_walk: fewest lanes worth another node step while other lanes hold a triangle
    int32_t rx_coop;   // dense receiver sets: the warp enumerates its lanes' segments together (rx_enumerate_coop)
};

__global__ void k_gen_dirs(int64_t ray_begin, int64_t n, float4 *__restrict__ dirs)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 d = ray_direction((uint32_t)(ray_begin + i));
    dirs[i] = make_float4(d.x, d.y, d.z, 0.0f);
}

// Sort key of a ray: 24-bit Morton code of its direction in the octahedral map, above the ray's index in the chunk.
// Rays are independent (kernel.py:48-55), so the trace may visit them in any order; neighbouring lanes that share a
// direction cell walk the same BVH nodes (primary rays all start at the transmitter), which is what turns the
// scattered node fetches of a big scene into cache hits and equalises the walk lengths inside a warp.
__device__ __forceinline__ uint32_t spread12(uint32_t v)
{
    v &= 0xfffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
// The top bit of the 24-bit code says whether the primary ray misses the scene's (padded) bounding box: those rays are
// one trivial segment each and are traced LAST, where they fill the end-game of the persistent kernel (the longest rays
// of a terrain — grazing, 6 bounces of ~200 dependent node fetches each — take about a millisecond on their own, and
// the warps that fetched them late used to finish alone: 7.3 -> 7.0 ms at 2^24 rays; ordering the entering rays by
// their chord through the box, longest first, brought nothing more).  What remains of the fixed ~1.1 ms per wave is
// 0.5 ms of drain after the last fetch and 0.6 ms before it (measured with %globaltimer probes); waves of 2^26 rays
// amortise it (4.7e9 segments/s on 20 M triangles against 4.05e9 at 2^24 and 2.6e9 at 2^22).
// CELLS: also count the ray into its direction cell (the 24-bit code) and keep its arrival rank there instead of its
// id: the input of the counting sort below.
// GEN: the directions are generated here (and stored for the walk) instead of being read back from k_gen_dirs' output:
// the fp64 series then run under the latency of the cell atomics.
template <bool CELLS, bool GEN = false>
__global__ void k_dir_keys(float4 *__restrict__ dirs, int64_t n, uint64_t *__restrict__ keys, float3 tx, float3 env_lo,
                           float3 env_hi, uint32_t *__restrict__ cells, int cell_shift, int64_t ray_begin = 0)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 d;
    if (GEN) {
        const float3 g = ray_direction((uint32_t)(ray_begin + i));
        d = make_float4(g.x, g.y, g.z, 0.0f);
        dirs[i] = d;
    } else {
        d = dirs[i];
    }
    float tn;
    const bool enters = slab_hit(slab_setup_fast(tx, make_float3(d.x, d.y, d.z)), env_lo.x, env_lo.y, env_lo.z, env_hi.x, env_hi.y,
                                 env_hi.z, 1.0e6f, tn);
    const float s = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + 1.0e-30f);
    float u = d.x * s, v = d.y * s;
    if (d.z < 0.0f) { // fold the lower hemisphere over the diagonals
        const float uu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f), vv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f);
        u = uu; v = vv;
    }
    const uint32_t qu = (uint32_t)fminf(fmaxf((u * 0.5f + 0.5f) * 4096.0f, 0.0f), 4095.0f);
    const uint32_t qv = (uint32_t)fminf(fmaxf((v * 0.5f + 0.5f) * 4096.0f, 0.0f), 4095.0f);
    const uint32_t code = ((spread12(qu) | (spread12(qv) << 1)) >> 1) | (enters ? 0u : 0x800000u);
    if (CELLS) keys[i] = ((uint64_t)code << 32) | (uint64_t)atomicAdd(cells + (code >> cell_shift), 1u);
    else keys[i] = ((uint64_t)code << 32) | (uint64_t)(uint32_t)i;
}

// ---- counting sort of a wave's rays by direction cell ---------------------------------------------------------------
// The key is a 24-bit cell number and the rays spread over the cells about evenly (one per cell at 2^24 rays), so the
// order needs no general sort: count the rays of every cell with one atomic each (k_dir_keys<true>, which keeps the
// arrival rank), scan the 2^24 counters (three small kernels over 67 MB), place every ray at its cell's offset + rank.
// Measured per 16.8 M rays (profiles/launches_terrain_r02_final.csv): directions + keys + atomics 227 us, scan 19 us,
// placement 180 us, against 163 (k_gen_dirs) + 86 + 3 x 265 us with the radix sort (wave of the 20 M-triangle terrain
// 6.0 -> 5.6 ms).
// The order inside a cell is the arrival order of the atomics — the hits of a ray do not depend on the order the rays
// are traced in.
constexpr int RAY_CELLS = 1 << 24;   // most cells a wave is ordered by (the whole 24-bit code)
constexpr int CELL_BLOCK = 4096;     // cells per CTA of the scan: 256 threads x 16 cells; <= 4096 block sums, scanned by one CTA

__device__ __forceinline__ uint32_t cell_thread_load(const uint32_t *cells, uint32_t (&v)[16])
{
    const uint4 *p = reinterpret_cast<const uint4 *>(cells + (size_t)blockIdx.x * CELL_BLOCK + threadIdx.x * 16);
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint4 q = p[j];
        v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        sum += q.x + q.y + q.z + q.w;
    }
    return sum;
}

// exclusive prefix of one value per thread over the CTA (blockDim.x <= 1024); total to every thread
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t x, uint32_t &total)
{
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t prefix = 0, tot = 0;
    for (int w = 0; w < nwarps; ++w) {
        const uint32_t ws = warp_sums[w];
        if (w < warp) prefix += ws;
        tot += ws;
    }
    __syncthreads();
    total = tot;
    return prefix + incl - x;
}

__global__ void __launch_bounds__(256) k_cells_reduce(const uint32_t *__restrict__ cells, uint32_t *__restrict__ block_sums)
{
    uint32_t v[16], total;
    const uint32_t sum = cell_thread_load(cells, v);
    cta_exclusive_scan(sum, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_cells_scan_sums(uint32_t *block_sums)
{
    uint4 q = reinterpret_cast<uint4 *>(block_sums)[threadIdx.x];
    uint32_t total;
    uint32_t base = cta_exclusive_scan(q.x + q.y + q.z + q.w, total);
    uint4 o;
    o.x = base; o.y = base + q.x; o.z = o.y + q.y; o.w = o.z + q.z;
    reinterpret_cast<uint4 *>(block_sums)[threadIdx.x] = o;
}

__global__ void __launch_bounds__(256) k_cells_apply(uint32_t *__restrict__ cells, const uint32_t *__restrict__ block_sums)
{
    uint32_t v[16], total;
    const uint32_t sum = cell_thread_load(cells, v);
    uint32_t run = cta_exclusive_scan(sum, total) + block_sums[blockIdx.x];
    uint4 *p = reinterpret_cast<uint4 *>(cells + (size_t)blockIdx.x * CELL_BLOCK + threadIdx.x * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = run; run += v[4 * j];
        o.y = run; run += v[4 * j + 1];
        o.z = run; run += v[4 * j + 2];
        o.w = run; run += v[4 * j + 3];
        p[j] = o;
    }
}

// keys[i] = (code << 32 | rank in the cell)  ->  order[offset of the cell + rank] = ray.  32-bit entries: the order of
// a 2^24-ray wave is 67 MB and its scattered 4-byte stores merge in L2 (64-bit entries = 134 MB did not: 563 us)
__global__ void k_cells_place(const uint64_t *__restrict__ keys, int64_t n, const uint32_t *__restrict__ cell_offset,
                              int cell_shift, uint32_t *__restrict__ order)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    const uint32_t cell = (uint32_t)(key >> 32) >> cell_shift;
    order[(size_t)__ldg(cell_offset + cell) + (uint32_t)key] = (uint32_t)i;
}

// Conservative sphere filter in front of the exact 80-triangle receiver query: can the segment
// [0, t_limit] of the ray come within the receiver's bounding sphere?  (Approximate reciprocal / square root: the
// 1 % inflation of the radius is five orders of magnitude above their 2-ulp error.)  The perpendicular distance is
// taken from the closest-approach VECTOR m = (c - p) - tc * d: |c - p|^2 - tc * ((c - p) . d) cancels two terms of
// size D^2 and loses the 1e-2 m^2 that matter once the receiver is D > ~50 m away (every rounding of m is ~D * 2^-24,
// far inside the 1e-5 * D slack on the radius).
__device__ __forceinline__ bool rx_sphere_filter(float3 p, float3 d, float cx, float cy, float cz, float radius,
                                                 float t_limit)
{
    float ox = cx - p.x, oy = cy - p.y, oz = cz - p.z;
    float dd = d.x * d.x + d.y * d.y + d.z * d.z;
    float od = ox * d.x + oy * d.y + oz * d.z;
    float oo = ox * ox + oy * oy + oz * oz;
    float r = radius * 1.01f + 1.0e-5f * (sqrt_approx(oo) + 1.0f);
    float r2 = r * r;
    if (oo <= r2) return true; // origin inside the (inflated) sphere
    float inv_dd = rcp_approx(dd);
    float tc = od * inv_dd;
    if (tc < 0.0f) return false;
    float mx = fmaf(-tc, d.x, ox), my = fmaf(-tc, d.y, oy), mz = fmaf(-tc, d.z, oz);
    float perp2 = mx * mx + my * my + mz * mz;
    if (perp2 > r2 * 1.01f + 1.0e-12f) return false;
    float half = sqrt_approx(fmaxf(r2 * 1.01f - perp2, 0.0f) * inv_dd);
    return tc - half * 1.0001f <= t_limit * 1.0001f + 1.0e-6f;
}

// Receiver candidates for one segment (kernel.py:71,85): a conservative bounding-sphere filter only.  The exact
// 80-triangle test is NOT done here — the literal replay (k_trace_receive) performs it anyway and drops the
// candidates whose first receiver hit is not at this bounce (false positives are ~10 % of the candidates).
__device__ __forceinline__ void rx_filter_and_emit(const TraceParams &P, int k, float3 pos, float3 dir, float t_limit,
                                                   uint32_t gid, int bounce)
{
    float cx = (float)__ldg(P.rx_centers + 3 * k), cy = (float)__ldg(P.rx_centers + 3 * k + 1),
          cz = (float)__ldg(P.rx_centers + 3 * k + 2);
    if (!rx_sphere_filter(pos, dir, cx, cy, cz, P.rx_radius, t_limit)) return;
    // warp-aggregated append: one atomic for all lanes that arrive here together (dense receiver lattices emit
    // several candidates per segment; 56 M single-address atomics were a quarter of C2's trace time)
    const unsigned am = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs((int)am) - 1;
    unsigned long long slot = 0;
    if (lane == leader) slot = atomicAdd(&P.counters[RFRT_CTR_CANDIDATES], (unsigned long long)__popc(am));
    slot = __shfl_sync(am, slot, leader) + __popc(am & ((1u << lane) - 1u));
    if ((int64_t)slot < P.cand_capacity) P.candidates[slot] = make_uint4(gid, (uint32_t)k, (uint32_t)bounce, 0u);
}

// Receivers of one finished segment, per lane (sparse receiver sets): the segment's own box against the bounds of all
// receivers (no division: most segments of a sparse set stop here), then this lane walks the receiver BVH with its
// column of the shared-memory stack.  The receivers whose box the segment overlaps are only QUEUED during the walk (a
// few entries of local memory) and filtered + appended afterwards, in a loop that the lanes with queued receivers run
// together: handled right where the walk finds them, the sphere filter and the append ran at 2 of 32 lanes and were
// half of the trace kernel's instructions on a 256 x 256 lattice (C2).
// RUNS: append a segment's candidates as one contiguous block (sets of many receivers: the replay's neighbouring lanes
//       then share their ray); otherwise one candidate per lane and step (a dozen instructions instead of a hundred:
//       the small-scene kernel of sparse sets is sensitive to the size of its loop body — 43.6 vs 46.5 ms on C4)
constexpr int RX_LANE_QUEUE = 12;
template <bool RUNS>
__device__ __forceinline__ void receivers_lane(const TraceParams &P, float3 pos, float3 dir, float t_limit, uint32_t gid,
                                               int bounce, int *stack, int stride)
{
    if (P.n_rx == 1) {
        rx_filter_and_emit(P, 0, pos, dir, t_limit, gid, bounce);
        return;
    }
    const float ex = fmaf(dir.x, t_limit, pos.x), ey = fmaf(dir.y, t_limit, pos.y), ez = fmaf(dir.z, t_limit, pos.z);
    const bool near_rx = fminf(pos.x, ex) <= P.rx_hi[0] && fmaxf(pos.x, ex) >= P.rx_lo[0] &&
                         fminf(pos.y, ey) <= P.rx_hi[1] && fmaxf(pos.y, ey) >= P.rx_lo[1] &&
                         fminf(pos.z, ez) <= P.rx_hi[2] && fmaxf(pos.z, ez) >= P.rx_lo[2];
    if (!near_rx) return;
    const RxSlabRay sr = rx_slab_setup(pos, dir);
    {   // the segment itself against the set's bounds (the box overlap above is loose for a diagonal segment and a thin
        // set — C4's 16 receivers on a line: a quarter of the trips had a lane in here, almost all to miss the root)
        float tn;
        if (!rx_slab_hit(sr, P.rx_lo[0], P.rx_lo[1], P.rx_lo[2], P.rx_hi[0], P.rx_hi[1], P.rx_hi[2], t_limit, tn)) return;
    }
    int queue[RX_LANE_QUEUE];
    int qn = 0, sp = 0;
    // next internal node from the stack; leaf entries (deferred while the queue was full) move to the queue on the way.
    // TRAV_DONE when the stack is empty or the queue is full (the caller drains it and asks again)
    auto pop = [&]() -> int {
        while (sp > 0 && qn < RX_LANE_QUEUE) {
            const int c = stack[(--sp) * stride];
            if (c >= 0) return c;
            queue[qn++] = ~c;
        }
        return TRAV_DONE;
    };
    int node = 0;
    for (;;) {
        while (node >= 0) {
            const float4 *np = reinterpret_cast<const float4 *>(P.rx_nodes + node);
            float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
            int4 q3 = __ldg(reinterpret_cast<const int4 *>(np + 3));
            float tn0, tn1;
            bool h0 = rx_slab_hit(sr, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, t_limit, tn0);
            bool h1 = rx_slab_hit(sr, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, t_limit, tn1);
            const int c0 = q3.x, c1 = q3.y;
            if (c1 == c0) h1 = false;
            if (h0) { if (c0 < 0 && qn < RX_LANE_QUEUE) queue[qn++] = ~c0; else { stack[sp * stride] = c0; ++sp; } }
            if (h1) { if (c1 < 0 && qn < RX_LANE_QUEUE) queue[qn++] = ~c1; else { stack[sp * stride] = c1; ++sp; } }
            node = pop();
        }
        // drain: sphere filter of the queued receivers, then ONE block of candidate slots per lane (one atomic per group
        // of lanes that drain together, a shuffle prefix over them): a segment's candidates stay next to each other, so
        // the replay's neighbouring lanes share their ray (interleaved lane by lane, every lane of a replay warp had
        // another ray)
        if (!RUNS) {
            while (qn > 0) rx_filter_and_emit(P, __ldg(P.rx_order + queue[--qn]), pos, dir, t_limit, gid, bounce);
            if (sp == 0) break;
            node = pop();
            continue;
        }
        int m = 0;
        for (int i = 0; i < qn; ++i) {
            const int k = __ldg(P.rx_order + queue[i]);
            const float cx = (float)__ldg(P.rx_centers + 3 * k), cy = (float)__ldg(P.rx_centers + 3 * k + 1),
                        cz = (float)__ldg(P.rx_centers + 3 * k + 2);
            if (rx_sphere_filter(pos, dir, cx, cy, cz, P.rx_radius, t_limit)) queue[m++] = k;
        }
        qn = 0;
        if (m > 0) {
            const unsigned am = __activemask();
            const int lane = threadIdx.x & 31, leader = __ffs((int)am) - 1;
            int before = 0, total = 0;
            for (unsigned rest = am; rest; rest &= rest - 1u) {
                const int src = __ffs((int)rest) - 1;
                const int v = __shfl_sync(am, m, src);
                if (src < lane) before += v;
                total += v;
            }
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(&P.counters[RFRT_CTR_CANDIDATES], (unsigned long long)total);
            base = __shfl_sync(am, base, leader) + (unsigned long long)before;
            for (int i = 0; i < m; ++i)
                if ((int64_t)(base + i) < P.cand_capacity) P.candidates[base + i] = make_uint4(gid, (uint32_t)queue[i], (uint32_t)bounce, 0u);
        }
        if (sp == 0) break;
        node = pop();
    }
}

// Dense receiver sets: the warp enumerates the receivers of its lanes' finished segments together (converged call).
__device__ __forceinline__ void receivers_coop(const TraceParams &P, bool seg_done, float3 pos, float3 dir, float t_limit,
                                               uint32_t gid, int bounce, int *rx_queue, int lane)
{
    const unsigned FULL = 0xffffffffu;
    // first the segment's own box against the bounds of all receivers, then the warp enumerates together
    bool near_rx = false;
    if (seg_done) {
        const float ex = fmaf(dir.x, t_limit, pos.x), ey = fmaf(dir.y, t_limit, pos.y), ez = fmaf(dir.z, t_limit, pos.z);
        near_rx = fminf(pos.x, ex) <= P.rx_hi[0] && fmaxf(pos.x, ex) >= P.rx_lo[0] &&
                  fminf(pos.y, ey) <= P.rx_hi[1] && fmaxf(pos.y, ey) >= P.rx_lo[1] &&
                  fminf(pos.z, ez) <= P.rx_hi[2] && fmaxf(pos.z, ez) >= P.rx_lo[2];
    }
    if (!__any_sync(FULL, near_rx)) return;
    // The receivers that pass the sphere filter are staged per segment in a small shared-memory buffer and
    // appended in runs (one global atomic per run): consecutive candidates then belong to the same ray, so
    // the replay kernel's warps work on one ray at a time instead of a mixture of several warps' rays.
    uint32_t gid_b = 0; int bounce_b = 0;
    int *cbuf = rx_queue + RX_QUEUE_CAP;        // [RX_CAND_BUF] receiver ids, then the fill count
    int *ccount = cbuf + RX_CAND_BUF;
    if (lane == 0) *ccount = 0;
    __syncwarp();
    const bool ok = rx_enumerate_coop(
        P.rx_nodes, P.rx_order, near_rx, rx_queue,
        [&](int src, float3 &bp, float3 &bd, float &bt) {
            bp.x = __shfl_sync(FULL, pos.x, src); bp.y = __shfl_sync(FULL, pos.y, src); bp.z = __shfl_sync(FULL, pos.z, src);
            bd.x = __shfl_sync(FULL, dir.x, src); bd.y = __shfl_sync(FULL, dir.y, src); bd.z = __shfl_sync(FULL, dir.z, src);
            bt = __shfl_sync(FULL, t_limit, src);
            gid_b = __shfl_sync(FULL, gid, src);
            bounce_b = __shfl_sync(FULL, bounce, src);
        },
        [&](int k, float3 bp, float3 bd, float bt) {
            const float cx = (float)__ldg(P.rx_centers + 3 * k), cy = (float)__ldg(P.rx_centers + 3 * k + 1),
                        cz = (float)__ldg(P.rx_centers + 3 * k + 2);
            if (rx_sphere_filter(bp, bd, cx, cy, cz, P.rx_radius, bt)) cbuf[atomicAdd(ccount, 1)] = k;
        },
        [&](bool final) {
            const int cnt = *ccount; // (every step adds at most 64 entries: the buffer never overflows)
            if (cnt >= RX_CAND_BUF - 64 || (final && cnt > 0)) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&P.counters[RFRT_CTR_CANDIDATES], (unsigned long long)cnt);
                base = __shfl_sync(FULL, base, 0);
                for (int j = lane; j < cnt; j += 32)
                    if ((int64_t)(base + j) < P.cand_capacity)
                        P.candidates[base + j] = make_uint4(gid_b, (uint32_t)cbuf[j], (uint32_t)bounce_b, 0u);
                __syncwarp();
                if (lane == 0) *ccount = 0;
                __syncwarp();
            }
        });
    if (!ok && lane == 0) atomicAdd(&P.counters[RFRT_CTR_QUEUE_OVERFLOW], 1ull);
}

__device__ __forceinline__ unsigned long long segment_hash(uint32_t gid, int bounce, int face, float t)
{
    // splitmix64 finaliser of (ray id, bounce | triangle, bits of t): RFRT_CTR_CHECKSUM sums it over all segments
    unsigned long long z = (((unsigned long long)gid << 8) | (unsigned)bounce) * 0x9E3779B97F4A7C15ull +
                           (((unsigned long long)(uint32_t)face << 32) | __float_as_uint(t));
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// BVH scenes: the walk is decoupled from the trip.  With one segment per lane and trip (round 1's k_trace_env) a trip
// ends when the LONGEST of the warp's 32 walks ends (measured on the 20 M-triangle terrain: 10-12 of 32 lanes active —
// the walk lengths of secondary rays vary from 3 to 200 nodes; 3.1e9 segments/s against 3.9e9 with this schedule on
// the same tree).  Here every lane keeps its traversal state (node, stack pointer, best hit, slab constants) across
// the phases of one loop:
//   phase A  lanes whose walk has ended finish their segment (counters, receivers, advance + reflect, kernel.py:85-96),
//            take a fresh ray if theirs is dead, and start the next segment (scene-box test, per-ray constants);
//   phase B  the warp walks — a node loop for the lanes that hold an internal node, a leaf step for the lanes that hold
//            a triangle — until `refill` lanes are waiting for phase A again.
// The node loop stops early for a leaf step only when fewer than `node_min` lanes would still take part in it.
// The per-segment functions (closest hit, ties, reflect) do not depend on the schedule: the segments are checked by the
// segment checksum against the small-scene kernel and the CPU restatement.
#ifndef WALK_UNROLL_N
#define WALK_UNROLL_N 2
#endif
constexpr int WALK_UNROLL = WALK_UNROLL_N;
// LSTACK: deep trees (big meshes) keep the traversal stack in per-thread local memory (L1-cached) instead of
//         shared memory, whose depth x 1 KiB per CTA would otherwise cap the occupancy of this latency-bound case
// COOP:   dense receiver sets — the warp enumerates its lanes' segments together (rx_enumerate_coop) at the converged
//         point of phase A; otherwise every lane handles its own receivers right where its segment is finished
// MT:     the Moeller-Trumbore functor instead of the reference's watertight test (rfrt_mesh_set_triangle_test) // node steps per vote of the node loop
// NEAR:   the transmitter stands within 8 x the mesh's largest coordinate of the coordinate origin, so every origin of
//         a segment does and the plain slab test is enough (rfrt_trace.cuh; +2 % on the 20 M-triangle terrain)
template <bool DUMP, bool LSTACK, bool COOP, bool MT, bool NEAR = false>
__global__ void __launch_bounds__(TRACE_THREADS, COOP ? WALK_MIN_CTAS - 1 : WALK_MIN_CTAS) k_trace_walk(const TraceParams P)
{
    using Ray = typename std::conditional<MT, MtRay, WoopRay>::type;
    extern __shared__ __align__(16) int s_stack_raw[];
    int l_stack[LSTACK ? 64 : 1];
    float l_stack_t[LSTACK ? 64 : 1];
    int *stack = LSTACK ? l_stack : s_stack_raw + threadIdx.x;
    float *stack_t = LSTACK ? l_stack_t : reinterpret_cast<float *>(s_stack_raw + P.stack_depth * TRACE_THREADS) + threadIdx.x;
    constexpr int STRIDE = LSTACK ? 1 : TRACE_THREADS;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    // per-ray constants of the watertight test, written once per segment in phase A (where the starting lanes compute
    // them together) and read back at every leaf: (kx | ky << 2 | kz << 4, Sx, Sy, Sz); the permuted origin follows from pos
    float4 *s_wr = reinterpret_cast<float4 *>(s_stack_raw + 2 * P.stack_depth * TRACE_THREADS) + threadIdx.x;
    int *rx_queue = s_stack_raw + 2 * P.stack_depth * TRACE_THREADS + 4 * TRACE_THREADS + (COOP ? RX_COOP_INTS * (threadIdx.x >> 5) : 0);
    const int REFILL = P.walk_refill, NODE_MIN = P.walk_node_min;

    bool has_ray = false;   // this lane owns a ray
    bool walking = false;   // ... and a segment of it is under way (node == TRAV_DONE: finished, waits for phase A)
    bool exhausted = false; // warp-uniform: no rays left to fetch
    int node = TRAV_DONE, sp = 0;
    float3 pos = make_float3(0.f, 0.f, 0.f), dir = make_float3(0.f, 0.f, 1.f);
    int bounce = 0;
    int64_t ray = 0;
    Hit h;
    h.t = 1.0e6f; h.face = -1; h.slot = -1;
    using WSlab = typename std::conditional<NEAR, RxSlabRay, SlabRay>::type;
    WSlab sr = slab_make<WSlab>(pos, dir);
    unsigned int n_seg = 0, n_hit = 0, n_nodes = 0, n_tests = 0;
    unsigned long long csum = 0ull;
    const int FETCH_BLOCK = P.fetch_block;
    int64_t blk_next = 0, blk_end = 0; // warp-uniform

    for (;;) {
        // ================= phase A: lanes that are not in the middle of a walk =================
        const bool seg_done = has_ray && walking && node == TRAV_DONE;
        // ---- receivers hit strictly before the environment, or at all if it is missed (kernel.py:71,85) ----
        if (COOP)
            receivers_coop(P, seg_done, pos, dir, h.face >= 0 ? h.t : 1.0e6f, (uint32_t)(P.chunk_begin + ray), bounce, rx_queue, lane);
        if (seg_done) {
            const bool hit_env = h.face >= 0;
            ++n_seg;
            if (!COOP && P.n_rx > 0)
                receivers_lane<true>(P, pos, dir, hit_env ? h.t : 1.0e6f, (uint32_t)(P.chunk_begin + ray), bounce, stack, STRIDE);
            if (DUMP) {
                int64_t row = (P.chunk_begin + ray - P.dump_begin) * P.max_bounces + bounce;
                if (P.hit_tri) P.hit_tri[row] = hit_env ? h.face : -1;
                if (P.hit_t) P.hit_t[row] = hit_env ? h.t : 0.0f;
                csum += segment_hash((uint32_t)(P.chunk_begin + ray), bounce, hit_env ? h.face : -1, hit_env ? h.t : 0.0f);
            }
            walking = false;
            if (hit_env) {
                ++n_hit;
                pos = advance(pos, dir, h.t);                 // kernel.py:94
                const float4 n4 = __ldg(P.normals + h.slot);  // normalize(cross(b-a, c-a)), precomputed at build time
                dir = reflect(dir, make_float3(n4.x, n4.y, n4.z)); // kernel.py:96
                ++bounce;
                if (bounce >= P.max_bounces) has_ray = false;
            } else {
                has_ray = false; // a miss repeats forever in the reference (kernel.py:97-98): nothing more to do
            }
        }
        // ---- refill (rays are handed out in blocks per warp: one atomic per block; second pass: the block ran out) ----
        {
            unsigned idle = __ballot_sync(FULL, !has_ray);
#pragma unroll 1
            for (int pass = 0; pass < 2 && idle != 0u && !exhausted; ++pass) {
                if (blk_next == blk_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(&P.counters[RFRT_CTR_NEXT_RAY], (unsigned long long)FETCH_BLOCK);
                    base = __shfl_sync(FULL, base, 0);
                    blk_next = (int64_t)base;
                    blk_end = blk_next + FETCH_BLOCK < P.chunk_n ? blk_next + FETCH_BLOCK : P.chunk_n;
                    if (blk_next >= P.chunk_n) { exhausted = true; blk_end = blk_next; }
                }
                if (!has_ray) {
                    int64_t r = blk_next + __popc(idle & ((1u << lane) - 1u));
                    if (r < blk_end) {
                        if (P.order32) r = (int64_t)__ldg(P.order32 + r);             // direction-coherent order
                        else if (P.order) r = (int64_t)(uint32_t)__ldg(P.order + r);
                        float4 d4 = __ldg(P.dirs + r);
                        dir = make_float3(d4.x, d4.y, d4.z);
                        pos = P.tx;
                        bounce = 0;
                        ray = r;
                        has_ray = true;
                    }
                }
                blk_next = blk_next + __popc(idle) < blk_end ? blk_next + __popc(idle) : blk_end;
                idle = __ballot_sync(FULL, !has_ray);
            }
        }
        if (!__any_sync(FULL, has_ray)) break;
        // ---- start the next segment: scene-box test first (a miss is a finished segment with no hit) ----
        if (has_ray && !walking) {
            // (approximate reciprocals: the slab test only prunes, and a hit point inside a box padded by 1e-5 of the
            // mesh's largest coordinate leaves a chord through it some thirty times the 2.4e-7 relative error per axis)
            if constexpr (NEAR) sr = rx_slab_setup(pos, dir);
            else sr = slab_setup_fast(pos, dir);
            float tn;
            const bool entered = P.n_tris > 0 &&
                                 slab_hit(sr, P.env_lo[0], P.env_lo[1], P.env_lo[2], P.env_hi[0], P.env_hi[1], P.env_hi[2], 1.0e6f, tn);
            h.t = 1.0e6f; h.face = -1; h.slot = -1;
            walking = true;
            sp = 0;
            node = entered ? 0 : TRAV_DONE;
            if constexpr (!MT) {
                if (entered) {
                    const WoopRay w0 = woop_setup(pos, dir);
                    *s_wr = make_float4(__int_as_float(w0.kx | (w0.ky << 2) | (w0.kz << 4)), w0.Sx, w0.Sy, w0.Sz);
                }
            }
        }
        // ================= phase B: walk until enough lanes wait for phase A =================
        // (a lane waits for phase A when its walk has ended and phase A has something for it: amask)
        const unsigned amask = __ballot_sync(FULL, has_ray || !exhausted);
        for (;;) {
            for (;;) {
                const unsigned internal = __ballot_sync(FULL, node >= 0), ended = __ballot_sync(FULL, node == TRAV_DONE);
                if (internal == 0u || __popc(ended & amask) >= REFILL ||
                    ((internal | ended) != FULL && __popc(internal) < NODE_MIN)) break;
#pragma unroll
                for (int u = 0; u < WALK_UNROLL; ++u) {
                    if (node >= 0) {
                        node = node_step(P.nodes, node, sr, h.t, stack, stack_t, STRIDE, sp);
                        if (DUMP) ++n_nodes;
                    }
                }
            }
            if (node < 0 && node != TRAV_DONE) {
                // (the per-ray constants of the triangle test come back from shared memory instead of living in 9
                // registers through the node loop)
                Ray wr;
                if constexpr (MT) {
                    wr = tri_ray_setup<Ray>(pos, dir);
                } else {
                    const float4 w4 = *s_wr;
                    const int kk = __float_as_int(w4.x);
                    wr.kx = kk & 3; wr.ky = (kk >> 2) & 3; wr.kz = (kk >> 4) & 3;
                    wr.Sx = w4.y; wr.Sy = w4.z; wr.Sz = w4.w;
                    wr.px = pos.x; wr.py = pos.y; wr.pz = pos.z;
                    wr.pkx = sel3(pos.x, pos.y, pos.z, wr.kx); wr.pky = sel3(pos.x, pos.y, pos.z, wr.ky);
                    wr.pkz = sel3(pos.x, pos.y, pos.z, wr.kz);
                }
                node = leaf_step(P.tris, node, wr, h, stack, stack_t, STRIDE, sp);
                if (DUMP) ++n_tests;
            }
            const unsigned ended = __ballot_sync(FULL, node == TRAV_DONE);
            if (__popc(ended & amask) >= REFILL || ended == FULL) break;
        }
    }

    for (int o = 16; o > 0; o >>= 1) {
        n_seg += __shfl_xor_sync(FULL, n_seg, o);
        n_hit += __shfl_xor_sync(FULL, n_hit, o);
    }
    if (lane == 0) {
        atomicAdd(&P.counters[RFRT_CTR_SEGMENTS], (unsigned long long)n_seg);
        atomicAdd(&P.counters[RFRT_CTR_ENV_HITS], (unsigned long long)n_hit);
    }
    if (DUMP) {
        unsigned long long nn = n_nodes, nt = n_tests;
        for (int o = 16; o > 0; o >>= 1) {
            csum += __shfl_xor_sync(FULL, csum, o);
            nn += __shfl_xor_sync(FULL, nn, o);
            nt += __shfl_xor_sync(FULL, nt, o);
        }
        if (lane == 0) {
            atomicAdd(&P.counters[RFRT_CTR_CHECKSUM], csum);
            atomicAdd(&P.counters[RFRT_CTR_NODE_VISITS], nn);
            atomicAdd(&P.counters[RFRT_CTR_TRI_TESTS], nt);
        }
    }
}

// ---- small scenes (<= 64 filter slots): lockstep sweep, several rays per lane ------------------------------------
// The scene image (rfrt_small.cu) lives in shared memory.  A trip of the warp is one of two kinds: a self-re-hit trip
// (cheap: the lanes whose ray stands on a triangle test whether it hits that triangle again at t ~ 0 —
// small_self_rehit; 3/4 of room.stl's segments) or a full sweep (closest_hit_small, ~3x the instructions).  With one
// ray per lane a self-re-hit trip finds only about half the lanes eligible (measured 17 of 32).  So every lane owns
// SMALL_SLOTS rays whose states live in its own column of shared memory (two 128-bit words per ray, conflict-free):
// the trip is the kind that more lanes can take part in, and each lane contributes whichever of its rays fits
// (28 of 32 lanes in both kinds; -10 % warp instructions, +13 % segments/s on room.stl).  The per-segment functions
// are those of the BVH path's parity tests, so the segments are the same; only the schedule differs.
//   word A = (pos.xyz, dir.x)   word B = (dir.y, dir.z, ray index in chunk, bounce << 8 | (triangle stood on + 1))
//   SMALL: 1 = <= 16 triangle pairs (one 32-bit candidate mask), 2 = <= 32 pairs
constexpr int SMALL_SLOTS = 5;       // rays per lane     (measured on room.stl: 2: 28e9, 3: 33.6e9, 4: 34.5e9, 5: 35.1e9
constexpr int SMALL_REFILL_MIN = 16; // lanes with an empty slot that make a refill pass worth its instructions (8 / 16 / 24: 34.6 / 35.1 / 35.2e9)

template <bool DUMP, int SMALL, bool COOP, bool RUNS>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_small(const TraceParams P)
{
    extern __shared__ __align__(16) int s_raw[];
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    // shared memory: ray slots | scene image | receiver stacks (one int column per thread) | cooperative queues
    float4 *slots = reinterpret_cast<float4 *>(s_raw) + threadIdx.x; // [2 * SMALL_SLOTS][TRACE_THREADS]
    float *img = reinterpret_cast<float *>(s_raw) + 8 * SMALL_SLOTS * TRACE_THREADS;
    const int n = (int)P.n_tris, np = P.small_pairs;
    const int img_floats = 30 * np + 17 * n;
    for (int i = threadIdx.x; i < img_floats; i += TRACE_THREADS) img[i] = __ldg(P.small + i);
    __syncthreads();
    const SmallScene S = small_scene_view(img, np, n, P.small_class, P.small_extent, (float)SMALL_TAU_REL * P.small_extent,
                                          (float)(2.0 * SMALL_REACH_REL) * P.small_extent);
    int *stack = reinterpret_cast<int *>(img + img_floats) + threadIdx.x;
    int *rx_queue = reinterpret_cast<int *>(img + img_floats) + P.stack_depth * TRACE_THREADS + (COOP ? RX_COOP_INTS * (threadIdx.x >> 5) : 0);

    unsigned valid = 0u;  // bit k: slot k holds a ray
    unsigned onface = 0u; // bit k: that ray stands on a triangle and has not failed the self-re-hit test there
    bool exhausted = false; // warp-uniform
    unsigned int n_seg = 0, n_hit = 0;
    unsigned long long csum = 0ull;
    const int FETCH_BLOCK = P.fetch_block;
    uint32_t blk_next = 0u; // warp-uniform: next ray of the warp's block (index in the chunk) and how many are left
    int blk_left = 0;
    constexpr unsigned ALL_SLOTS = (1u << SMALL_SLOTS) - 1u;

    for (;;) {
        // ---- refill: one empty slot per lane and trip (rays are handed out in blocks per warp: one atomic per block;
        //      second pass: the block ran out in the middle of the refill) ----
        {
            const int ke = __ffs((int)(~valid & ALL_SLOTS)) - 1; // -1: no empty slot
            unsigned idle = __ballot_sync(FULL, ke >= 0);
            bool filled = false;
            // (with several slots per lane a few empty ones can wait: refill when it pays for the pass)
#pragma unroll 1
            for (int pass = 0; pass < 2 && __popc(idle) >= (pass == 0 ? SMALL_REFILL_MIN : 1) && !exhausted; ++pass) {
                if (blk_left == 0) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(&P.counters[RFRT_CTR_NEXT_RAY], (unsigned long long)FETCH_BLOCK);
                    base = __shfl_sync(FULL, base, 0);
                    if ((int64_t)base >= P.chunk_n) {
                        exhausted = true;
                    } else {
                        blk_next = (uint32_t)base;
                        blk_left = P.chunk_n - (int64_t)base < FETCH_BLOCK ? (int)(P.chunk_n - (int64_t)base) : FETCH_BLOCK;
                    }
                }
                const int rank = __popc(idle & ((1u << lane) - 1u));
                if (ke >= 0 && !filled && rank < blk_left) {
                    const uint32_t r = blk_next + (uint32_t)rank;
                    const float4 d4 = __ldg(P.dirs + r);
                    slots[(2 * ke) * TRACE_THREADS] = make_float4(P.tx.x, P.tx.y, P.tx.z, d4.x);
                    slots[(2 * ke + 1) * TRACE_THREADS] = make_float4(d4.y, d4.z, __uint_as_float(r), __int_as_float(0));
                    valid |= 1u << ke;
                    filled = true;
                }
                const int took = __popc(idle) < blk_left ? __popc(idle) : blk_left;
                blk_next += (uint32_t)took;
                blk_left -= took;
                idle = __ballot_sync(FULL, ke >= 0 && !filled);
            }
        }
        if (!__any_sync(FULL, valid != 0u)) break;

        // ---- pick the trip: the kind that more lanes can take part in (a full sweep is always valid) ----
        const unsigned wants_sweep = valid & ~onface;
        const int n_r = __popc(__ballot_sync(FULL, onface != 0u));
        const int n_s = __popc(__ballot_sync(FULL, wants_sweep != 0u));
        const bool shortcut = n_r > 0 && n_r >= n_s; // (a bias of +-2..4 lanes either way: no difference)
        const unsigned pick = shortcut ? onface : wants_sweep;
        const bool act = pick != 0u;
        const int k = __ffs((int)pick) - 1;

        float3 pos = make_float3(0.f, 0.f, 0.f), dir = make_float3(0.f, 0.f, 1.f);
        Hit h;
        h.t = 1.0e6f; h.face = -1; h.slot = -1;
        bool seg_done = false;
        if (act) {
            const float4 A = slots[(2 * k) * TRACE_THREADS], B = slots[(2 * k + 1) * TRACE_THREADS];
            pos = make_float3(A.x, A.y, A.z);
            dir = make_float3(A.w, B.x, B.y);
            const WoopRay wr = woop_setup(pos, dir);
            bool resolved = true;
            if (shortcut) resolved = small_self_rehit(S, (__float_as_int(B.w) & 0xff) - 1, pos, dir, wr, h);
            else closest_hit_small<SMALL == 2>(S, pos, dir, wr, h);
            if (!resolved) onface &= ~(1u << k); // not a self re-hit: this segment goes through a full sweep
            seg_done = resolved;
        }
        // (ray id and bounce are read back from the slot only now: they need not live through the sweep)
        uint32_t ray = 0u;
        int bounce = 0;
        if (seg_done) {
            const float4 B = slots[(2 * k + 1) * TRACE_THREADS];
            ray = __float_as_uint(B.z);
            bounce = __float_as_int(B.w) >> 8;
        }

        // ---- receivers hit strictly before the environment, or at all if it is missed (kernel.py:71,85) ----
        const float t_limit = h.face >= 0 ? h.t : 1.0e6f;
        const uint32_t gid = (uint32_t)(P.chunk_begin + ray);
        if (COOP) receivers_coop(P, seg_done, pos, dir, t_limit, gid, bounce, rx_queue, lane);

        if (seg_done) {
            const bool hit_env = h.face >= 0;
            ++n_seg;
            if (!COOP && P.n_rx > 0) receivers_lane<RUNS>(P, pos, dir, t_limit, gid, bounce, stack, TRACE_THREADS);
            if (DUMP) {
                int64_t row = (P.chunk_begin + ray - P.dump_begin) * P.max_bounces + bounce;
                if (P.hit_tri) P.hit_tri[row] = hit_env ? h.face : -1;
                if (P.hit_t) P.hit_t[row] = hit_env ? h.t : 0.0f;
                csum += segment_hash(gid, bounce, hit_env ? h.face : -1, hit_env ? h.t : 0.0f);
            }
            bool alive = false;
            if (hit_env) {
                ++n_hit;
                pos = advance(pos, dir, h.t);                 // kernel.py:94
                const float *v = S.normals + 3 * h.face;
                dir = reflect(dir, make_float3(v[0], v[1], v[2])); // kernel.py:96
                ++bounce;
                alive = bounce < P.max_bounces;
            } // (a miss repeats forever in the reference, kernel.py:97-98: nothing more to do)
            if (alive) {
                slots[(2 * k) * TRACE_THREADS] = make_float4(pos.x, pos.y, pos.z, dir.x);
                slots[(2 * k + 1) * TRACE_THREADS] = make_float4(dir.y, dir.z, __uint_as_float(ray), __int_as_float((bounce << 8) | (h.face + 1)));
                onface |= 1u << k;
            } else {
                valid &= ~(1u << k);
                onface &= ~(1u << k);
            }
        }
    }

    for (int o = 16; o > 0; o >>= 1) {
        n_seg += __shfl_xor_sync(FULL, n_seg, o);
        n_hit += __shfl_xor_sync(FULL, n_hit, o);
    }
    if (lane == 0) {
        atomicAdd(&P.counters[RFRT_CTR_SEGMENTS], (unsigned long long)n_seg);
        atomicAdd(&P.counters[RFRT_CTR_ENV_HITS], (unsigned long long)n_hit);
    }
    if (DUMP) {
        for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(FULL, csum, o);
        if (lane == 0) atomicAdd(&P.counters[RFRT_CTR_CHECKSUM], csum);
    }
}

// ---- literal replay ------------------------------------------------------------------------------
struct LiteralEnv {
    const BvhNode *nodes;
    const BvhTri *tris;
    int64_t n_tris;
};

// kernel.py:38-98 for one ray and one receiver.  Sink receives the vertex writes and RX-hit events.
// s_recs: unit-space face records in shared memory (lockstep receiver query) or NULL (unit-BVH walk)
// NEAR: every origin of the walk lies within 8 x the mesh's largest coordinate (rfrt_trace.cuh: the plain slab test)
template <bool MT, bool NEAR = false, class Sink>
__device__ __forceinline__ void literal_trace(const LiteralEnv &E, const RxView *rx, int n_faces, float3 tx,
                                              int max_bounces, uint32_t tid, int *stack, float *stack_t, int stride,
                                              Sink &sink, const float4 *s_recs = nullptr)
{
    float3 dir = ray_direction(tid); // kernel.py:51-52
    float3 pos = tx;                 // :53
    sink.vertex(0, pos);             // :55
    // A receiver hit that leaves the ray where it is — t == 0, or a t so small that pos + dir * t rounds back to pos —
    // with the direction it had (kernel.py:87, no reflection) makes every later iteration evaluate the same two queries
    // on the same values and take the same branch: the "stuck" paths of the reference's golden scene (SURVEY.md
    // Appendix C).  They are replayed without queries.
    bool stuck = false;
    float t_stuck = 0.0f;
    RxFaceCache rx_cache; // candidate faces of the current line (valid from a receiver query until the direction changes)
    rx_cache.valid = false;
    for (int bounce = 0; bounce < max_bounces; ++bounce) {
        if (stuck) {
            pos = advance(pos, dir, t_stuck); // :87
            sink.vertex(bounce + 1, pos);     // :88
            sink.received(bounce);            // :89-91
            continue;
        }
        using Ray = typename std::conditional<MT, MtRay, WoopRay>::type;
        const Ray wr = tri_ray_setup<Ray>(pos, dir);
        using Slab = typename std::conditional<NEAR, RxSlabRay, SlabRay>::type;
        const Slab sr = slab_make<Slab>(pos, dir);
        float t_rx = 0.0f;
        bool maybe_hit_rx = false; // :71
        if (rx) {
            // (the lockstep receiver query's candidate filter is derived for the watertight test only)
            if constexpr (!MT) {
                maybe_hit_rx = s_recs ? rx_query_sweep(*rx, s_recs, c_rx_faces, n_faces, wr, pos, dir, 1.0e6f, t_rx, &rx_cache)
                                      : rx_query(*rx, c_rx_faces, n_faces, wr, pos, dir, 1.0e6f, stack, stack_t, stride, t_rx);
            } else {
                maybe_hit_rx = rx_query(*rx, c_rx_faces, n_faces, wr, pos, dir, 1.0e6f, stack, stack_t, stride, t_rx);
            }
        }
        Hit h;
        h.t = 1.0e6f; h.face = -1; h.slot = -1;
        closest_hit(E.nodes, E.tris, E.n_tris, wr, sr, stack, stack_t, stride, h);                          // :82
        bool maybe_hit_env = h.face >= 0;
        bool hit_recv = maybe_hit_rx && (!maybe_hit_env || h.t > t_rx);                                     // :85
        if (hit_recv) {
            const float3 moved = advance(pos, dir, t_rx);  // :87
            if (moved.x == pos.x && moved.y == pos.y && moved.z == pos.z) { stuck = true; t_stuck = t_rx; }
            pos = moved;
            sink.vertex(bounce + 1, pos);   // :88
            sink.received(bounce);          // :89-91
        } else if (maybe_hit_env) {
            pos = advance(pos, dir, h.t);   // :94
            sink.vertex(bounce + 1, pos);   // :95
            sink.env_face(bounce + 1, h.face);
            float3 a, b, c; int idx;
            tri_vertices(E.tris, h.slot, a, b, c, idx);
            dir = reflect(dir, tri_normal(a, b, c)); // :96
            rx_cache.valid = false;
        } else {
            break; // :97-98: nothing changes, so every later iteration repeats the same two misses
        }
    }
}

struct CompatSink {
    float *traced;
    float *recv_row;
    uint32_t *mask;
    __device__ __forceinline__ void env_face(int, int) {}
    __device__ __forceinline__ void vertex(int i, float3 p)
    {
        traced[3 * i] = p.x; traced[3 * i + 1] = p.y; traced[3 * i + 2] = p.z;
    }
    __device__ __forceinline__ void received(int bounce)
    {
        for (int i = 0; i < 3 * (bounce + 2); ++i) recv_row[i] = traced[i]; // kernel.py:89-90
        *mask = 1u;                                                         // :91
    }
};

template <bool MT>
__global__ void __launch_bounds__(TRACE_THREADS)
k_trace_compat(LiteralEnv E, RxView rx, const double *rx_center, int n_faces, float3 tx, int max_bounces, int64_t ray_begin,
               int64_t n_rays, float *traced, float *received, uint32_t *mask, int stack_depth)
{
    const int has_rx = rx_center != nullptr;
    if (has_rx) { rx.cx = (float)__ldg(rx_center); rx.cy = (float)__ldg(rx_center + 1); rx.cz = (float)__ldg(rx_center + 2); }
    extern __shared__ int s_stack_raw[];
    int *stack = s_stack_raw + threadIdx.x;
    float *stack_t = reinterpret_cast<float *>(s_stack_raw + stack_depth * TRACE_THREADS) + threadIdx.x;
    const int64_t row = 3 * (int64_t)(max_bounces + 1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rays; i += (int64_t)gridDim.x * blockDim.x) {
        CompatSink sink{traced + i * row, received + i * row, mask + i};
        literal_trace<MT>(E, has_rx ? &rx : nullptr, n_faces, tx, max_bounces, (uint32_t)(ray_begin + i), stack, stack_t,
                          TRACE_THREADS, sink);
    }
}

struct RecordSink {
    float path[3 * (MAX_RECV_BOUNCES + 1)];
    int face[MAX_RECV_BOUNCES + 1]; // triangle of an environment vertex, -1 for tx / receiver vertices (materials)
    int last_rx_bounce;
    int first_rx_bounce;
    bool want_faces;
    __device__ __forceinline__ void env_face(int i, int f) { if (want_faces) face[i] = f; }
    __device__ __forceinline__ void vertex(int i, float3 p)
    {
        path[3 * i] = p.x; path[3 * i + 1] = p.y; path[3 * i + 2] = p.z;
    }
    __device__ __forceinline__ void received(int bounce)
    {
        if (first_rx_bounce < 0) first_rx_bounce = bounce;
        last_rx_bounce = bounce;
    }
};

__device__ __forceinline__ float norm3_f32(float x, float y, float z)
{
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// tracer.py:34-61.  Explicit _rn operations: nvcc would otherwise contract n_2*cos(theta_i) -/+ n_1*cos(theta) into
// FMAs, which CPython's float arithmetic never does (asin / sin / cos themselves are libdevice vs libm: <= 1e-15).
__device__ __forceinline__ double bounce_amplitude(double angle_between, double n_1 = 5.0)
{
    if (isnan(angle_between)) return 0.0;
    const double PI = 3.141592653589793;
    double theta = __dsub_rn(__ddiv_rn(PI, 2.0), __ddiv_rn(angle_between, 2.0));
    const double n_2 = 1.0;
    double theta_i = asin(__ddiv_rn(__dmul_rn(n_2, sin(theta)), n_1));
    double a = __dmul_rn(n_2, cos(theta_i)), b = __dmul_rn(n_1, cos(theta));
    double num = __dsub_rn(a, b);
    double denom = __dadd_rn(a, b);
    double q = __ddiv_rn(num, denom);
    double amp = -__dmul_rn(q, q);
    if (amp < -1) amp = -1;
    if (isnan(amp)) return 0.0;
    return -amp;
}

struct ReceiveParams {
    LiteralEnv env;
    const float *materials; // [n_tris] refractive index per triangle (rfrt_mesh_set_materials) or NULL
    const float *rx_verts;
    const double *rx_centers;
    const BvhNode *unit_nodes;
    const int32_t *unit_order;
    const float *unit_recs; // [n_faces*16] face records of the lockstep receiver query
    float inv_r;
    int32_t n_unit;
    int32_t n_faces;
    float3 tx;
    int32_t max_bounces;
    const uint4 *candidates;
    int64_t cand_capacity;
    unsigned long long *counters;
    double amp0, light_speed, sample_rate;
    uint32_t *rec_ray;
    int32_t *rec_rx;
    int32_t *rec_nverts;
    int64_t *rec_bin;
    double *rec_amp;
    double *rec_dist;
    float *rec_paths;
    int64_t rec_capacity;
    double *ir;      // direct mode: ir[rx * n_bins + bin] += amplitude instead of storing the record (or NULL)
    int64_t n_bins;
    int32_t stack_depth;
};

template <bool LSTACK, bool MT, bool NEAR>
__global__ void __launch_bounds__(TRACE_THREADS, NEAR ? RECV_MIN_CTAS : RECV_MIN_CTAS - 1) k_trace_receive(const ReceiveParams P)
{
    extern __shared__ __align__(16) int s_stack_raw[];
    int l_stack[LSTACK ? 64 : 1];
    float l_stack_t[LSTACK ? 64 : 1];
    int *stack = LSTACK ? l_stack : s_stack_raw + threadIdx.x;
    float *stack_t = LSTACK ? l_stack_t : reinterpret_cast<float *>(s_stack_raw + P.stack_depth * TRACE_THREADS) + threadIdx.x;
    constexpr int STRIDE = LSTACK ? 1 : TRACE_THREADS;
    // the receiver shape's face records live behind the stacks
    float4 *s_recs = reinterpret_cast<float4 *>(s_stack_raw + 2 * P.stack_depth * TRACE_THREADS);
    for (int i = threadIdx.x; i < 4 * P.n_faces; i += TRACE_THREADS) s_recs[i] = __ldg(reinterpret_cast<const float4 *>(P.unit_recs) + i);
    __syncthreads();
    int64_t n_cand = (int64_t)P.counters[RFRT_CTR_CANDIDATES];
    if (n_cand > P.cand_capacity) n_cand = P.cand_capacity;
    const int row = 3 * (P.max_bounces + 1);
    unsigned int n_direct = 0; // records binned directly by this thread
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_cand; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 cand = P.candidates[i];
        RecordSink sink;
        sink.last_rx_bounce = -1;
        sink.first_rx_bounce = -1;
        sink.want_faces = P.materials != nullptr;
        if (sink.want_faces)
            for (int v = 0; v <= P.max_bounces; ++v) sink.face[v] = -1;
        RxView rx;
        rx.verts = P.rx_verts + (int64_t)cand.y * P.n_unit * 3;
        rx.unit_nodes = P.unit_nodes; rx.unit_order = P.unit_order;
        rx.cx = (float)__ldg(P.rx_centers + 3 * (int64_t)cand.y); rx.cy = (float)__ldg(P.rx_centers + 3 * (int64_t)cand.y + 1);
        rx.cz = (float)__ldg(P.rx_centers + 3 * (int64_t)cand.y + 2); rx.inv_r = P.inv_r;
        literal_trace<MT, NEAR>(P.env, &rx, P.n_faces, P.tx, P.max_bounces, cand.x, stack, stack_t, STRIDE, sink, MT ? nullptr : s_recs);
        // a candidate raised at a later bounce than the replay's first receiver hit is a duplicate
        if (sink.last_rx_bounce < 0 || sink.first_rx_bounce != (int)cand.z) continue;
        int nverts = sink.last_rx_bounce + 2;
        // tracer.py:90-97: strip at the first vertex containing a NaN
        for (int v = 0; v < nverts; ++v)
            if (isnan(sink.path[3 * v]) || isnan(sink.path[3 * v + 1]) || isnan(sink.path[3 * v + 2])) { nverts = v; break; }
        // tracer.py:102-115
        double amplitude = P.amp0;
        double distance = 0.0;
        const float *p = sink.path;
        if (nverts >= 2) {
            for (int v = 0; v + 2 < nverts; ++v) {
                float s1x = __fsub_rn(p[3 * v + 3], p[3 * v]), s1y = __fsub_rn(p[3 * v + 4], p[3 * v + 1]),
                      s1z = __fsub_rn(p[3 * v + 5], p[3 * v + 2]);
                float s2x = __fsub_rn(p[3 * v + 6], p[3 * v + 3]), s2y = __fsub_rn(p[3 * v + 7], p[3 * v + 4]),
                      s2z = __fsub_rn(p[3 * v + 8], p[3 * v + 5]);
                float l1 = norm3_f32(s1x, s1y, s1z);
                float l2 = norm3_f32(s2x, s2y, s2z);
                float dot = __fadd_rn(__fadd_rn(__fmul_rn(s1x, s2x), __fmul_rn(s1y, s2y)), __fmul_rn(s1z, s2z));
                float q = __fdiv_rn(dot, __fmul_rn(l1, l2));
                float angle = (q > 1.0f || q < -1.0f || isnan(q)) ? __int_as_float(0x7fc00000)
                                                                   : __double2float_rn(acos((double)q));
                // tracer.py:43 hard-codes n_1 = 5; with a material table the vertex's triangle decides
                double n_1 = 5.0;
                if (P.materials && sink.face[v + 1] >= 0) n_1 = (double)__ldg(P.materials + sink.face[v + 1]);
                amplitude = __dmul_rn(amplitude, bounce_amplitude((double)angle, n_1));
                distance = __dadd_rn(distance, (double)l1);
            }
            const float *u = p + 3 * (nverts - 2), *w = p + 3 * (nverts - 1);
            distance = __dadd_rn(distance, (double)norm3_f32(__fsub_rn(u[0], w[0]), __fsub_rn(u[1], w[1]),
                                                             __fsub_rn(u[2], w[2])));
        }
        double samples = __dmul_rn(__ddiv_rn(distance, P.light_speed), P.sample_rate);
        long long bin = (long long)samples; // int() truncation, tracer.py:115
        if (P.ir) { // tracer.py:116-117 right here: no record list, no second pass over it
            if (bin >= 0 && bin < P.n_bins) atomicAdd(P.ir + (int64_t)cand.y * P.n_bins + bin, amplitude);
            ++n_direct;
            continue;
        }
        unsigned long long slot = atomicAdd(&P.counters[RFRT_CTR_RECORDS], 1ull);
        if ((int64_t)slot < P.rec_capacity) {
            P.rec_ray[slot] = cand.x;
            P.rec_rx[slot] = (int32_t)cand.y;
            P.rec_nverts[slot] = nverts;
            P.rec_bin[slot] = bin;
            P.rec_amp[slot] = amplitude;
            P.rec_dist[slot] = distance;
            if (P.rec_paths) {
                float *dst = P.rec_paths + (int64_t)slot * row;
                for (int k = 0; k < row; ++k) dst[k] = k < 3 * nverts ? p[k] : __int_as_float(0x7fc00000);
            }
        }
    }
    if (P.ir) {
        for (int o = 16; o > 0; o >>= 1) n_direct += __shfl_xor_sync(0xffffffffu, n_direct, o);
        if ((threadIdx.x & 31) == 0 && n_direct) atomicAdd(&P.counters[RFRT_CTR_RECORDS], (unsigned long long)n_direct);
    }
}

template <bool MT>
__global__ void __launch_bounds__(TRACE_THREADS)
k_query(LiteralEnv E, const float *__restrict__ origins, const float *__restrict__ dirs, int64_t n, float max_t,
        float *t_out, int32_t *face_out, int stack_depth)
{
    extern __shared__ int s_stack_raw[];
    int *stack = s_stack_raw + threadIdx.x;
    float *stack_t = reinterpret_cast<float *>(s_stack_raw + stack_depth * TRACE_THREADS) + threadIdx.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float3 p = make_float3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        float3 d = make_float3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        using Ray = typename std::conditional<MT, MtRay, WoopRay>::type;
        const Ray wr = tri_ray_setup<Ray>(p, d);
        SlabRay sr = slab_setup(p, d);
        Hit h;
        h.t = max_t; h.face = -1; h.slot = -1;
        closest_hit(E.nodes, E.tris, E.n_tris, wr, sr, stack, stack_t, TRACE_THREADS, h);
        t_out[i] = h.t;
        face_out[i] = h.face;
    }
}

// Every origin of a segment (the transmitter, then hit points) within 8 x the mesh's largest coordinate m of the
// coordinate origin: the origin-dependent slack of the slab test, 2^-21 * 8 m, then fits inside the boxes' padding of
// 1e-5 m and the walks use the plain test.  (The replay also starts walks at receiver hit points, which may lie
// anywhere: one outside that cube of half-size 8 m is on a straight line that has left the cube — it contains the
// transmitter and the whole mesh — for good, so no hit exists that a cull could lose; the ray only turns at
// environment hits.)  RFRT_SLAB_FAR=1 forces the general test (A/B runs, tests).
bool tx_is_near(const Mesh *m, const float *h_tx_pos)
{
    float m_coord = 0.0f, tx_coord = 0.0f;
    for (int k = 0; k < 6; ++k) m_coord = fmaxf(m_coord, fabsf(m->bvh.bounds[k]));
    for (int k = 0; k < 3; ++k) tx_coord = fmaxf(tx_coord, fabsf(h_tx_pos[k]));
    return tx_coord <= 8.0f * m_coord && m->bvh.pad >= 1.0e-5f * m_coord && !getenv("RFRT_SLAB_FAR");
}

int stack_depth_for(const Mesh *m, const RxSet *r)
{
    int d = m ? m->bvh.max_depth : 0;
    if (r && r->bvh.max_depth > d) d = r->bvh.max_depth;
    if (r && r->unit_bvh.max_depth > d) d = r->unit_bvh.max_depth;
    d += 2;
    if (d < 8) d = 8;
    return d;
}

size_t stack_bytes(int depth) { return (size_t)depth * TRACE_THREADS * 2 * sizeof(int); }

// all_smem: the kernel keeps its whole working set in shared memory (small scenes) — ask for the largest carve-out, so
// that the resident CTAs per SM are what the occupancy query below says and not what the driver's L1/shared split leaves
int grid_for(const void *kernel, size_t smem, int *out_grid, bool all_smem = false)
{
    int dev = 0, sms = 0, per_sm = 0;
    RFRT_CUDA(cudaGetDevice(&dev));
    RFRT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (smem > 48 * 1024) RFRT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (all_smem) RFRT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    RFRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRACE_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    *out_grid = sms * per_sm;
    return RFRT_OK;
}

// waves of at least this many rays are ordered by the counting sort over the 2^24 direction cells (its cost has a fixed
// part: ~0.3 GB of counter traffic), smaller ones by the radix sort; RFRT_RAY_ORDER=radix|cells forces one (A/B, tests)
constexpr int64_t RAY_CELLS_MIN_RAYS = 1 << 21;

// workspace of the direction-coherent ray order (BVH scenes): two key buffers + the sort's histograms + the cell counters
int reserve_ray_sort(Mesh *m, int64_t cap)
{
    if (m->ray_cap >= cap) return RFRT_OK;
    if (m->ray_keys[0]) cudaFree(m->ray_keys[0]);
    if (m->ray_keys[1]) cudaFree(m->ray_keys[1]);
    if (m->ray_hist) cudaFree(m->ray_hist);
    if (m->ray_cells) cudaFree(m->ray_cells);
    m->ray_keys[0] = m->ray_keys[1] = nullptr; m->ray_hist = nullptr; m->ray_cells = nullptr; m->ray_cap = 0;
    RFRT_CUDA(cudaMalloc(&m->ray_keys[0], sizeof(uint64_t) * cap));
    RFRT_CUDA(cudaMalloc(&m->ray_keys[1], sizeof(uint64_t) * cap));
    RFRT_CUDA(cudaMalloc(&m->ray_hist, sizeof(uint32_t) * 256 * ((size_t)sort_hist_blocks(cap) + 1)));
    const char *order_env = getenv("RFRT_RAY_ORDER");
    if (cap >= RAY_CELLS_MIN_RAYS || (order_env && strcmp(order_env, "cells") == 0)) RFRT_CUDA(cudaMalloc(&m->ray_cells, sizeof(uint32_t) * ((size_t)RAY_CELLS + RAY_CELLS / CELL_BLOCK)));
    m->ray_cap = cap;
    return RFRT_OK;
}

int upload_faces(const RxSet *r, cudaStream_t stream)
{
    RFRT_CUDA(cudaMemcpyToSymbolAsync(c_rx_faces, r->faces, sizeof(uint8_t) * 3 * (size_t)r->n_faces, 0,
                                      cudaMemcpyHostToDevice, stream));
    return RFRT_OK;
}

} // namespace
} // namespace rfrt

using namespace rfrt;

extern "C" int rfrt_ray_directions(int64_t ray_begin, int64_t ray_end, float *d_dirs, void *stream)
{
    int64_t n = ray_end - ray_begin;
    if (n < 0 || (n > 0 && !d_dirs)) { set_error("rfrt_ray_directions: bad arguments"); return RFRT_ERR_INVALID; }
    if (n == 0) return RFRT_OK;
    k_gen_dirs<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ray_begin, n, (float4 *)d_dirs);
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}

extern "C" int rfrt_mesh_reserve_rays(rfrt_handle mesh, int64_t max_chunk_rays)
{
    Mesh *m = get_mesh(mesh);
    if (!m) { set_error("rfrt_mesh_reserve_rays: unknown mesh handle"); return RFRT_ERR_HANDLE; }
    if (max_chunk_rays < 0) { set_error("rfrt_mesh_reserve_rays: bad arguments"); return RFRT_ERR_INVALID; }
    if (m->small || max_chunk_rays == 0) return RFRT_OK; // small scenes are swept, not walked: no ray order needed
    return reserve_ray_sort(m, max_chunk_rays);
}

extern "C" int rfrt_trace(rfrt_handle env_mesh, rfrt_handle rxset, const float *h_tx_pos, int32_t max_bounces,
                          int64_t ray_begin, int64_t ray_end, uint32_t flags, float *d_dir_scratch,
                          int64_t chunk_rays, uint64_t *d_counters, uint32_t *d_candidates, int64_t cand_capacity,
                          int32_t *d_hit_tri, float *d_hit_t, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool dirs_ready = (flags & RFRT_FLAG_DIRS_READY) != 0;
    Mesh *m = get_mesh(env_mesh);
    if (!m) { set_error("rfrt_trace: unknown environment mesh handle"); return RFRT_ERR_HANDLE; }
    RxSet *r = nullptr;
    if (rxset) {
        r = get_rxset(rxset);
        if (!r) { set_error("rfrt_trace: unknown receiver set handle"); return RFRT_ERR_HANDLE; }
        if (!d_candidates || cand_capacity <= 0) { set_error("rfrt_trace: candidate buffer required with receivers"); return RFRT_ERR_INVALID; }
    }
    int64_t n = ray_end - ray_begin;
    if (!h_tx_pos || !d_counters || max_bounces < 0 || max_bounces >= (1 << 23) || n < 0 || ray_end > (1ll << 32) || ray_begin < 0) {
        set_error("rfrt_trace: bad arguments (need tx_pos, counters, 0 <= ray ids <= 2^32, max_bounces < 2^23)");
        return RFRT_ERR_INVALID;
    }
    if (n == 0 || max_bounces == 0) return RFRT_OK;
    if (!d_dir_scratch) { set_error("rfrt_trace: d_dir_scratch required"); return RFRT_ERR_INVALID; }
    if (chunk_rays <= 0) chunk_rays = 1ll << 24;
    if (dirs_ready) chunk_rays = n; // the caller generated all directions with rfrt_ray_directions

    TraceParams P;
    P.nodes = m->bvh.nodes; P.tris = m->tris; P.n_tris = m->bvh.n_prims;
    P.normals = m->normals;
    P.small = m->small; P.small_pairs = m->small_pairs; P.small_extent = m->small_extent;
    for (int c = 0; c < 5; ++c) P.small_class[c] = m->small_class[c];
    P.rx_nodes = r ? r->bvh.nodes : nullptr; P.rx_order = r ? r->bvh.prim_order : nullptr;
    P.rx_verts = r ? r->verts : nullptr; P.rx_centers = r ? r->centers : nullptr;
    P.n_rx = r ? r->n_receivers : 0; P.n_unit = r ? r->n_unit : 0; P.n_faces = r ? r->n_faces : 0;
    P.rx_radius = r ? (float)r->radius : 0.0f;
    for (int k = 0; k < 3; ++k) {
        P.rx_lo[k] = r ? r->bvh.bounds[k] - r->bvh.pad : 0.0f;
        P.rx_hi[k] = r ? r->bvh.bounds[3 + k] + r->bvh.pad : 0.0f;
    }
    for (int k = 0; k < 3; ++k) {
        P.env_lo[k] = m->bvh.bounds[k] - m->bvh.pad;
        P.env_hi[k] = m->bvh.bounds[3 + k] + m->bvh.pad;
    }
    P.tx = make_float3(h_tx_pos[0], h_tx_pos[1], h_tx_pos[2]);
    P.max_bounces = max_bounces;
    P.dirs = (const float4 *)d_dir_scratch;
    P.counters = (unsigned long long *)d_counters;
    P.candidates = (uint4 *)d_candidates; P.cand_capacity = r ? cand_capacity : 0;
    P.hit_tri = d_hit_tri; P.hit_t = d_hit_t; P.dump_begin = ray_begin;
    // Receiver enumeration: per lane (stack) for sparse sets, warp-cooperative (queue) when a segment that crosses the
    // set overlaps many receivers at once — estimated as the receivers' cross-sections over the largest face of their
    // bounding box (C3's lattice: ~46; C2's: 3; C4's 16 receivers: 0.1).
    P.rx_coop = 0;
    if (r && r->n_receivers > 1) {
        const double ex = r->bvh.bounds[3] - r->bvh.bounds[0], ey = r->bvh.bounds[4] - r->bvh.bounds[1], ez = r->bvh.bounds[5] - r->bvh.bounds[2];
        double face = ex * ey > ey * ez ? ex * ey : ey * ez;
        if (ex * ez > face) face = ex * ez;
        const double d2 = 4.0 * r->radius * r->radius;
        P.rx_coop = (double)r->n_receivers * d2 >= 8.0 * (face > d2 ? face : d2) ? 1 : 0;
        if (getenv("RFRT_RX_COOP")) P.rx_coop = atoi(getenv("RFRT_RX_COOP")) ? 1 : 0; // tuning aid
    }
    P.stack_depth = stack_depth_for(m, P.rx_coop ? nullptr : r);
    // DUMP instantiations also accumulate the checksum
    const bool dump = d_hit_tri || d_hit_t || (flags & RFRT_FLAG_CHECKSUM);
    // small scenes: lockstep sweep over the scene staged in shared memory (see closest_hit_small)
    // (the small-scene sweep's candidate filter is derived for the watertight test: Moeller-Trumbore meshes walk the BVH)
    const bool mt = m->tri_test == RFRT_TRI_TEST_MT;
    const bool small = m->small && P.n_tris > 0 && !(flags & RFRT_FLAG_FORCE_BVH) && !mt;
    if (small) P.stack_depth = (r && !P.rx_coop) ? stack_depth_for(nullptr, r) : 1;
    const bool lstack = !small && P.stack_depth > 16; // deep tree: local-memory stack (<= 64 entries by construction)
    if (lstack && P.stack_depth > 64) { set_error("rfrt_trace: BVH deeper than 64 levels"); return RFRT_ERR_INVALID; }
    if (lstack) P.stack_depth = 0;
    size_t smem = stack_bytes(P.stack_depth) + sizeof(float4) * TRACE_THREADS + (P.rx_coop ? sizeof(int) * RX_COOP_INTS * (TRACE_THREADS / 32) : 0);
    typedef void (*kern_t)(const TraceParams);
    kern_t kern;
    if (small) {
        // several rays per lane: ray slots | scene image | one int column per thread for the receiver walk | queues
        smem = sizeof(float4) * 2 * SMALL_SLOTS * TRACE_THREADS + sizeof(float) * small_image_floats(m->small_pairs, (int)P.n_tris) +
               sizeof(int) * P.stack_depth * TRACE_THREADS + (P.rx_coop ? sizeof(int) * RX_COOP_INTS * (TRACE_THREADS / 32) : 0);
        // [cooperative enumeration | per-lane with candidate runs | per-lane, one candidate per step][pairs > 16][dump]
        static const kern_t small_kerns[3][2][2] = {
            {{k_trace_small<false, 1, false, false>, k_trace_small<true, 1, false, false>},
             {k_trace_small<false, 2, false, false>, k_trace_small<true, 2, false, false>}},
            {{k_trace_small<false, 1, false, true>, k_trace_small<true, 1, false, true>},
             {k_trace_small<false, 2, false, true>, k_trace_small<true, 2, false, true>}},
            {{k_trace_small<false, 1, true, false>, k_trace_small<true, 1, true, false>},
             {k_trace_small<false, 2, true, false>, k_trace_small<true, 2, true, false>}},
        };
        const int rx_mode = P.rx_coop ? 2 : (P.n_rx >= 64 ? 1 : 0);
        kern = small_kerns[rx_mode][m->small_pairs > 16 ? 1 : 0][dump ? 1 : 0];
    } else {
        static const kern_t walk_kerns[2][2][2][2] = {
            {{{k_trace_walk<false, false, false, false>, k_trace_walk<true, false, false, false>},
              {k_trace_walk<false, true, false, false>, k_trace_walk<true, true, false, false>}},
             {{k_trace_walk<false, false, true, false>, k_trace_walk<true, false, true, false>},
              {k_trace_walk<false, true, true, false>, k_trace_walk<true, true, true, false>}}},
            {{{k_trace_walk<false, false, false, true>, k_trace_walk<true, false, false, true>},
              {k_trace_walk<false, true, false, true>, k_trace_walk<true, true, false, true>}},
             {{k_trace_walk<false, false, true, true>, k_trace_walk<true, false, true, true>},
              {k_trace_walk<false, true, true, true>, k_trace_walk<true, true, true, true>}}},
        };
        // thresholds of the walk schedule (20 M-triangle terrain, segments/s: refill 8 / 12 / 16 / 24 lanes = 3.53 / 3.69 /
        // 3.89 / 3.91e9 at node_min 8; node_min 4 / 8 / 12 / 16 = 3.64 / 3.75 / 3.80 / 3.68e9 at refill 16); the
        // environment variables are tuning aids of scripts/walk_sweep.py
        P.walk_refill = getenv("RFRT_WALK_REFILL") ? atoi(getenv("RFRT_WALK_REFILL")) : 24;
        P.walk_node_min = getenv("RFRT_WALK_NODE_MIN") ? atoi(getenv("RFRT_WALK_NODE_MIN")) : 8;
        kern = walk_kerns[mt ? 1 : 0][P.rx_coop ? 1 : 0][lstack ? 1 : 0][dump ? 1 : 0];
        static const kern_t near_kerns[2][2][2] = {
            {{k_trace_walk<false, false, false, false, true>, k_trace_walk<true, false, false, false, true>},
             {k_trace_walk<false, true, false, false, true>, k_trace_walk<true, true, false, false, true>}},
            {{k_trace_walk<false, false, true, false, true>, k_trace_walk<true, false, true, false, true>},
             {k_trace_walk<false, true, true, false, true>, k_trace_walk<true, true, true, false, true>}},
        };
        if (!mt && tx_is_near(m, h_tx_pos)) kern = near_kerns[P.rx_coop ? 1 : 0][lstack ? 1 : 0][dump ? 1 : 0];
    }
    int grid = 0;
    int rc = grid_for((const void *)kern, smem, &grid, small);
    if (rc) return rc;
    // (the face table in __constant__ memory is only read by the replay / compat kernels)

    // BVH scenes: trace the rays of a chunk in direction-coherent order (workspace cached with the mesh)
    const bool sorted = !small && !(flags & RFRT_FLAG_NO_RAY_SORT) && n >= 4096;
    if (sorted) {
        // (sized once by rfrt_mesh_reserve_rays, which the Python layer calls when it builds a BVH scene; a caller that
        // did not reserve pays for the allocation here, on the first wave only)
        const int rc2 = reserve_ray_sort(m, n < chunk_rays ? n : chunk_rays);
        if (rc2) return rc2;
    }

    for (int64_t c0 = ray_begin; c0 < ray_end; c0 += chunk_rays) {
        int64_t cn = ray_end - c0 < chunk_rays ? ray_end - c0 : chunk_rays;
        const char *order_env = getenv("RFRT_RAY_ORDER");
        const bool by_cells = sorted && m->ray_cells && (order_env ? strcmp(order_env, "cells") == 0 : cn >= RAY_CELLS_MIN_RAYS);
        const bool fuse_gen = by_cells && !dirs_ready && !getenv("RFRT_NO_FUSED_GEN");
        if (!dirs_ready && !fuse_gen) k_gen_dirs<<<(unsigned)((cn + 255) / 256), 256, 0, stream>>>(c0, cn, (float4 *)d_dir_scratch);
        P.order = nullptr; P.order32 = nullptr;
        if (sorted) {
            const float3 blo = make_float3(P.env_lo[0], P.env_lo[1], P.env_lo[2]), bhi = make_float3(P.env_hi[0], P.env_hi[1], P.env_hi[2]);
            const unsigned key_blocks = (unsigned)((cn + 255) / 256);
            if (by_cells) {
                // 2^22 cells (the top 22 bits of the code): counters (16 MB) + order (4 B per ray) stay in L2 for a
                // wave of 2^24 rays; RFRT_RAY_CELL_BITS is a tuning aid
                int cell_bits = getenv("RFRT_RAY_CELL_BITS") ? atoi(getenv("RFRT_RAY_CELL_BITS")) : 22;
                cell_bits = cell_bits < 20 ? 20 : (cell_bits > 24 ? 24 : cell_bits);
                const int n_cells = 1 << cell_bits, cell_blocks = n_cells / CELL_BLOCK, cell_shift = 24 - cell_bits;
                uint32_t *block_sums = m->ray_cells + RAY_CELLS;
                RFRT_CUDA(cudaMemsetAsync(m->ray_cells, 0, sizeof(uint32_t) * n_cells, stream));
                if (fuse_gen) k_dir_keys<true, true><<<key_blocks, 256, 0, stream>>>((float4 *)d_dir_scratch, cn, m->ray_keys[0], P.tx, blo, bhi, m->ray_cells, cell_shift, c0);
                else k_dir_keys<true><<<key_blocks, 256, 0, stream>>>((float4 *)d_dir_scratch, cn, m->ray_keys[0], P.tx, blo, bhi, m->ray_cells, cell_shift);
                k_cells_reduce<<<cell_blocks, 256, 0, stream>>>(m->ray_cells, block_sums);
                k_cells_scan_sums<<<1, cell_blocks / 4, 0, stream>>>(block_sums);
                k_cells_apply<<<cell_blocks, 256, 0, stream>>>(m->ray_cells, block_sums);
                k_cells_place<<<key_blocks, 256, 0, stream>>>(m->ray_keys[0], cn, m->ray_cells, cell_shift, reinterpret_cast<uint32_t *>(m->ray_keys[1]));
                P.order32 = reinterpret_cast<const uint32_t *>(m->ray_keys[1]);
            } else {
            k_dir_keys<false><<<key_blocks, 256, 0, stream>>>((float4 *)d_dir_scratch, cn, m->ray_keys[0], P.tx, blo, bhi, nullptr, 0);
            // (radix passes over the 24-bit direction code, from its top: 3 / 2 / 1 passes = 3.90 / 3.95 / 3.65e9 segments/s
            // on the 20 M-triangle terrain, 5.88 / 5.84 / 5.18e9 on 2 M triangles; the variable is a tuning aid)
            const int sort_passes = getenv("RFRT_RAY_SORT_PASSES") ? atoi(getenv("RFRT_RAY_SORT_PASSES")) : 3;
            P.order = radix_sort_u64(m->ray_keys[0], m->ray_keys[1], m->ray_hist, cn, 32 + 8 * (3 - sort_passes), sort_passes, stream);
            }
        }
        RFRT_CUDA(cudaMemsetAsync(d_counters + RFRT_CTR_NEXT_RAY, 0, sizeof(uint64_t), stream));
        P.chunk_begin = c0; P.chunk_n = cn;
        int g = grid;
        int64_t need = (cn + TRACE_THREADS - 1) / TRACE_THREADS;
        if (need < g) g = (int)need;
        int64_t fb = cn / ((int64_t)g * (TRACE_THREADS / 32) * 64);
        P.fetch_block = fb < 32 ? 32 : (fb > 256 ? 256 : (int)fb);
        kern<<<g, TRACE_THREADS, smem, stream>>>(P);
    }
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}

extern "C" int rfrt_trace_receive(rfrt_handle env_mesh, rfrt_handle rxset, const float *h_tx_pos,
                                  int32_t max_bounces, const uint32_t *d_candidates, int64_t cand_capacity,
                                  uint64_t *d_counters, double amp0, double light_speed_mps, double sample_rate_hz,
                                  uint32_t *d_rec_ray, int32_t *d_rec_rx, int32_t *d_rec_nverts, int64_t *d_rec_bin,
                                  double *d_rec_amp, double *d_rec_dist, float *d_rec_paths, int64_t rec_capacity,
                                  double *d_ir, int64_t n_bins, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Mesh *m = get_mesh(env_mesh);
    RxSet *r = get_rxset(rxset);
    if (!m || !r) { set_error("rfrt_trace_receive: unknown handle"); return RFRT_ERR_HANDLE; }
    const bool direct = d_ir != nullptr;
    if (!h_tx_pos || !d_candidates || !d_counters || cand_capacity <= 0 || (direct && n_bins <= 0) ||
        (!direct && (!d_rec_ray || !d_rec_rx || !d_rec_nverts || !d_rec_bin || !d_rec_amp || !d_rec_dist || rec_capacity <= 0))) {
        set_error("rfrt_trace_receive: null buffer or empty capacity");
        return RFRT_ERR_INVALID;
    }
    if (max_bounces < 0 || max_bounces > MAX_RECV_BOUNCES) {
        set_error("rfrt_trace_receive: max_bounces must be in [0, 32]");
        return RFRT_ERR_INVALID;
    }
    ReceiveParams P;
    P.env.nodes = m->bvh.nodes; P.env.tris = m->tris; P.env.n_tris = m->bvh.n_prims;
    P.materials = m->materials;
    P.rx_verts = r->verts; P.n_unit = r->n_unit; P.n_faces = r->n_faces;
    P.rx_centers = r->centers; P.unit_nodes = r->unit_bvh.nodes; P.unit_order = r->unit_bvh.prim_order;
    P.unit_recs = r->unit_recs;
    P.inv_r = (float)(1.0 / r->radius);
    P.tx = make_float3(h_tx_pos[0], h_tx_pos[1], h_tx_pos[2]);
    P.max_bounces = max_bounces;
    P.candidates = (const uint4 *)d_candidates; P.cand_capacity = cand_capacity;
    P.counters = (unsigned long long *)d_counters;
    P.amp0 = amp0; P.light_speed = light_speed_mps; P.sample_rate = sample_rate_hz;
    P.rec_ray = d_rec_ray; P.rec_rx = d_rec_rx; P.rec_nverts = d_rec_nverts; P.rec_bin = d_rec_bin;
    P.rec_amp = d_rec_amp; P.rec_dist = d_rec_dist; P.rec_paths = d_rec_paths; P.rec_capacity = rec_capacity;
    P.ir = d_ir; P.n_bins = n_bins;
    // only the environment BVH is walked here (receiver query: lockstep sweep), except with the Moeller-Trumbore functor
    P.stack_depth = stack_depth_for(m, m->tri_test == RFRT_TRI_TEST_MT ? r : nullptr);
    const bool lstack = P.stack_depth > 16;
    if (lstack && P.stack_depth > 64) { set_error("rfrt_trace_receive: BVH deeper than 64 levels"); return RFRT_ERR_INVALID; }
    if (lstack) P.stack_depth = 0;
    const size_t smem = stack_bytes(P.stack_depth) + sizeof(float4) * 4 * (size_t)r->n_faces;
    int grid = 0;
    typedef void (*recv_kern_t)(const ReceiveParams);
    static const recv_kern_t recv_kerns[2][2][2] = {
        {{k_trace_receive<false, false, false>, k_trace_receive<false, false, true>},
         {k_trace_receive<true, false, false>, k_trace_receive<true, false, true>}},
        {{k_trace_receive<false, true, false>, k_trace_receive<false, true, true>},
         {k_trace_receive<true, true, false>, k_trace_receive<true, true, true>}}};
    // the replay's origins are the transmitter and hit points: with the transmitter within 8 x the mesh's largest
    // coordinate the boxes' padding covers the origin-dependent error and the plain slab test is enough
    const bool near_tx = tx_is_near(m, h_tx_pos);
    const recv_kern_t recv_kern = recv_kerns[m->tri_test == RFRT_TRI_TEST_MT ? 1 : 0][lstack ? 1 : 0][near_tx ? 1 : 0];
    int rc = grid_for((const void *)recv_kern, smem, &grid);
    if (rc) return rc;
    rc = upload_faces(r, stream);
    if (rc) return rc;
    recv_kern<<<grid, TRACE_THREADS, smem, stream>>>(P);
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}

extern "C" int rfrt_trace_paths_compat(rfrt_handle env_mesh, const float *h_tx_pos, rfrt_handle rxset,
                                       int64_t rx_index, int32_t max_bounces, int64_t ray_begin, int64_t n_rays,
                                       float *d_traced_paths, float *d_received_paths, uint32_t *d_row_mask,
                                       void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Mesh *m = get_mesh(env_mesh);
    if (!m) { set_error("rfrt_trace_paths_compat: unknown environment mesh handle"); return RFRT_ERR_HANDLE; }
    RxSet *r = nullptr;
    if (rxset) {
        r = get_rxset(rxset);
        if (!r) { set_error("rfrt_trace_paths_compat: unknown receiver set handle"); return RFRT_ERR_HANDLE; }
        if (rx_index < 0 || rx_index >= r->n_receivers) { set_error("rfrt_trace_paths_compat: rx_index out of range"); return RFRT_ERR_INVALID; }
    }
    if (!h_tx_pos || !d_traced_paths || !d_received_paths || !d_row_mask || max_bounces < 0 || n_rays < 0 ||
        ray_begin < 0 || ray_begin + n_rays > (1ll << 32)) {
        set_error("rfrt_trace_paths_compat: bad arguments");
        return RFRT_ERR_INVALID;
    }
    if (n_rays == 0) return RFRT_OK;
    LiteralEnv E{m->bvh.nodes, m->tris, m->bvh.n_prims};
    int depth = stack_depth_for(m, r);
    const size_t smem = stack_bytes(depth);
    int grid = 0;
    const bool mt = m->tri_test == RFRT_TRI_TEST_MT;
    int rc = grid_for(mt ? (const void *)k_trace_compat<true> : (const void *)k_trace_compat<false>, smem, &grid);
    RxView rxv{};
    if (r) {
        rxv.verts = r->verts + rx_index * r->n_unit * 3;
        rxv.unit_nodes = r->unit_bvh.nodes; rxv.unit_order = r->unit_bvh.prim_order;
        rxv.inv_r = (float)(1.0 / r->radius); // (centre: read by the kernel from r->centers — no host round trip)
    }
    if (rc) return rc;
    int64_t need = (n_rays + TRACE_THREADS - 1) / TRACE_THREADS;
    if (need < grid) grid = (int)need;
    if (r) { rc = upload_faces(r, stream); if (rc) return rc; }
    (mt ? k_trace_compat<true> : k_trace_compat<false>)<<<grid, TRACE_THREADS, smem, stream>>>(
        E, rxv, r ? r->centers + 3 * rx_index : nullptr, r ? r->n_faces : 0, make_float3(h_tx_pos[0], h_tx_pos[1], h_tx_pos[2]), max_bounces, ray_begin, n_rays,
        d_traced_paths, d_received_paths, d_row_mask, depth);
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}

extern "C" int rfrt_query_closest(rfrt_handle mesh, const float *d_origins, const float *d_dirs, int64_t n,
                                  float max_t, float *d_t, int32_t *d_face, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Mesh *m = get_mesh(mesh);
    if (!m) { set_error("rfrt_query_closest: unknown mesh handle"); return RFRT_ERR_HANDLE; }
    if (n < 0 || (n > 0 && (!d_origins || !d_dirs || !d_t || !d_face))) { set_error("rfrt_query_closest: bad arguments"); return RFRT_ERR_INVALID; }
    if (n == 0) return RFRT_OK;
    LiteralEnv E{m->bvh.nodes, m->tris, m->bvh.n_prims};
    int depth = stack_depth_for(m, nullptr);
    const size_t smem = stack_bytes(depth);
    int grid = 0;
    const bool mt = m->tri_test == RFRT_TRI_TEST_MT;
    int rc = grid_for(mt ? (const void *)k_query<true> : (const void *)k_query<false>, smem, &grid);
    if (rc) return rc;
    int64_t need = (n + TRACE_THREADS - 1) / TRACE_THREADS;
    if (need < grid) grid = (int)need;
    (mt ? k_query<true> : k_query<false>)<<<grid, TRACE_THREADS, smem, stream>>>(E, d_origins, d_dirs, n, max_t, d_t, d_face, depth);
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}
