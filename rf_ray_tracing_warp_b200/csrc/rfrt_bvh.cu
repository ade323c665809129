// rfrt_bvh.cu — GPU LBVH builder (replaces wp.Mesh's BVH build, /root/reference tracer.py:24,30).
//
//   primitive boxes -> scene bounds -> 63-bit Morton code of the box centre (21 bits per axis, all three axes
//   quantised with the SAME step = largest extent / 2^21, so the code's cells are cubes: a flat scene such as a
//   heightfield is then split in x / y only until the cells are as small as its relief, instead of being cut into
//   height bands at every third level) with the primitive index as the sort's value -> stable 8-bit LSD radix sort
//   (8 passes; hand-written: per-tile histogram, exclusive scan, stable ranked scatter) -> Karras 2012 hierarchy
//   emission (equal codes are split by their position in the sorted order) -> bottom-up refit with per-node
//   arrival counters.
//
// The hierarchy only accelerates the query: the hit rule (closest t, ties to the lowest triangle
// index) is order independent, so any valid BVH returns the oracle's answer.
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rfrt_internal.h"

namespace rfrt {

namespace {

constexpr int BVH_MAX_DEPTH = 60; // deeper hierarchies are rebuilt over coarser Morton cells (the walk's stack: 64 entries)
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

__device__ __forceinline__ int float_to_ordered(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// bounds_enc: 6 ordered ints {lo.xyz (min), hi.xyz (max)}
__global__ void k_bounds_init(int *bounds_enc)
{
    if (threadIdx.x < 3) bounds_enc[threadIdx.x] = float_to_ordered(FLT_MAX);
    else if (threadIdx.x < 6) bounds_enc[threadIdx.x] = float_to_ordered(-FLT_MAX);
}

__global__ void k_scene_bounds(const float4 *__restrict__ lo, const float4 *__restrict__ hi, int64_t n,
                               int *bounds_enc)
{
    float l[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, h[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = lo[i], b = hi[i];
        l[0] = fminf(l[0], a.x); l[1] = fminf(l[1], a.y); l[2] = fminf(l[2], a.z);
        h[0] = fmaxf(h[0], b.x); h[1] = fmaxf(h[1], b.y); h[2] = fmaxf(h[2], b.z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            l[k] = fminf(l[k], __shfl_xor_sync(0xffffffffu, l[k], o));
            h[k] = fmaxf(h[k], __shfl_xor_sync(0xffffffffu, h[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&bounds_enc[k], float_to_ordered(l[k]));
            atomicMax(&bounds_enc[3 + k], float_to_ordered(h[k]));
        }
    }
}

__global__ void k_bounds_decode(const int *bounds_enc, float *bounds)
{
    if (threadIdx.x < 6) bounds[threadIdx.x] = ordered_to_float(bounds_enc[threadIdx.x]);
}

__device__ __forceinline__ uint64_t expand_bits21(uint32_t v)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x001f00000000ffffull;
    x = (x | (x << 16)) & 0x001f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4 *__restrict__ lo, const float4 *__restrict__ hi, int64_t n,
                         const float *__restrict__ bounds, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, int legacy,
                         int drop_bits)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = lo[i], b = hi[i];
    float c[3] = {0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z)};
    const float ext = fmaxf(fmaxf(bounds[3] - bounds[0], bounds[4] - bounds[1]), bounds[5] - bounds[2]);
    const float scale = ext > 0.0f ? 2097152.0f / ext : 0.0f;
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float u = (c[k] - bounds[k]) * scale;
        u = fminf(fmaxf(u, 0.0f), 2097151.0f); // (NaN -> 0)
        q[k] = ((uint32_t)u >> drop_bits) << drop_bits; // coarser cells: fewer levels (see build_lbvh)
    }
    if (legacy) {
        // round 1's code, kept for A/B measurements (RFRT_BVH_LEGACY_MORTON=1): 10 bits per axis, every axis
        // normalised by its OWN extent (a heightfield is then cut into height bands at every third level)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float e = bounds[3 + k] - bounds[k];
            const float u = e > 0.0f ? (c[k] - bounds[k]) / e : 0.0f;
            q[k] = (uint32_t)fminf(fmaxf(u * 1024.0f, 0.0f), 1023.0f) << 11;
        }
    }
    keys[i] = (expand_bits21(q[0]) << 2) | (expand_bits21(q[1]) << 1) | expand_bits21(q[2]);
    vals[i] = (uint32_t)i;
}

// ---- radix sort ------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ ghist, int nblocks)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll 4
    for (int it = 0; it < SORT_ITEMS; ++it) {
        int64_t i = base + it * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    ghist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Exclusive scan of the digit-major histogram [256][nblocks] in two levels: every digit row by its own CTA (in place,
// row total to row_tot[digit]), then the 256 row totals by one CTA; the scatter adds the two.
__global__ void __launch_bounds__(256) k_scan_rows(uint32_t *hist, int nblocks, uint32_t *row_tot)
{
    __shared__ uint32_t warp_sums[8];
    uint32_t *row = hist + (int64_t)blockIdx.x * nblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (int base = 0; base < nblocks; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? row[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        uint32_t prefix = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) prefix += warp_sums[w];
            total += warp_sums[w];
        }
        if (i < nblocks) row[i] = carry + prefix + x - v;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_tot[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(256) k_scan_totals(uint32_t *row_tot)
{
    __shared__ uint32_t s[256];
    s[threadIdx.x] = row_tot[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int d = 0; d < 256; ++d) { uint32_t v = s[d]; s[d] = acc; acc += v; }
    }
    __syncthreads();
    row_tot[threadIdx.x] = s[threadIdx.x];
}

// VALS: a 32-bit value travels with every key
template <bool VALS>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, int64_t n, int shift,
               const uint32_t *__restrict__ gscan, const uint32_t *__restrict__ row_base, int nblocks,
               const uint32_t *__restrict__ vin, uint32_t *__restrict__ vout)
{
    __shared__ uint32_t base[256];
    __shared__ uint32_t cnt[SORT_THREADS / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    base[tid] = gscan[(int64_t)tid * nblocks + blockIdx.x] + row_base[tid];
    const int64_t tile = (int64_t)blockIdx.x * SORT_TILE;
    for (int it = 0; it < SORT_ITEMS; ++it) {
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) cnt[w][tid] = 0u;
        __syncthreads();
        int64_t i = tile + it * SORT_THREADS + tid;
        bool valid = i < n;
        uint64_t key = valid ? in[i] : 0ull;
        uint32_t d = valid ? ((uint32_t)(key >> shift) & 255u) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) cnt[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t off = base[d];
            for (int w = 0; w < warp; ++w) off += cnt[w][d];
            out[off + rank] = key;
            if (VALS) vout[off + rank] = vin[i];
        }
        __syncthreads();
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) s += cnt[w][tid];
        base[tid] += s;
    }
}

// ---- Karras hierarchy --------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t *__restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t x = keys[i] ^ keys[j];
    return x ? __clzll((long long)x) : 64 + __clz(i ^ j); // equal codes: split by position in the sorted order
}

// children[i] = (c0, c1) for internal node i; parent arrays for the refit
__global__ void k_karras(const uint64_t *__restrict__ keys, int n, int2 *__restrict__ children,
                         int *__restrict__ node_parent, int *__restrict__ leaf_parent)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int c0, c1;
    if (lo == gamma) { c0 = ~gamma; leaf_parent[gamma] = i; }
    else { c0 = gamma; node_parent[gamma] = i; }
    if (hi == gamma + 1) { c1 = ~(gamma + 1); leaf_parent[gamma + 1] = i; }
    else { c1 = gamma + 1; node_parent[gamma + 1] = i; }
    children[i] = make_int2(c0, c1);
    if (i == 0) node_parent[0] = -1;
}

__device__ __forceinline__ void store_child_box(BvhNode *node, int slot, const float lo[3], const float hi[3])
{
    float *f = reinterpret_cast<float *>(node);
    if (slot == 0) {
        f[0] = lo[0]; f[1] = lo[1]; f[2] = lo[2]; f[3] = hi[0]; f[4] = hi[1]; f[5] = hi[2];
    } else {
        f[6] = lo[0]; f[7] = lo[1]; f[8] = lo[2]; f[9] = hi[0]; f[10] = hi[1]; f[11] = hi[2];
    }
}
__device__ __forceinline__ void load_child_box(const BvhNode *node, int slot, float lo[3], float hi[3])
{
    const volatile float *f = reinterpret_cast<const volatile float *>(node);
    int o = slot * 6;
    lo[0] = f[o + 0]; lo[1] = f[o + 1]; lo[2] = f[o + 2]; hi[0] = f[o + 3]; hi[1] = f[o + 4]; hi[2] = f[o + 5];
}

__global__ void k_refit(const uint32_t *__restrict__ order, int n, const float4 *__restrict__ prim_lo,
                        const float4 *__restrict__ prim_hi, float pad, const int2 *__restrict__ children,
                        const int *__restrict__ node_parent, const int *__restrict__ leaf_parent,
                        int *__restrict__ arrive, BvhNode *nodes, int *max_depth)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t prim = order[i];
    float4 a = prim_lo[prim], b = prim_hi[prim];
    float lo[3] = {a.x - pad, a.y - pad, a.z - pad};
    float hi[3] = {b.x + pad, b.y + pad, b.z + pad};
    int child = ~i;
    int p = leaf_parent[i];
    int depth = 1;
    while (p >= 0) {
        int2 ch = children[p];
        int slot = (ch.x == child) ? 0 : 1;
        store_child_box(&nodes[p], slot, lo, hi);
        __threadfence();
        int old = atomicAdd(&arrive[p], 1);
        if (old == 0) return; // the sibling subtree finishes this node
        __threadfence();
        float lo2[3], hi2[3];
        load_child_box(&nodes[p], slot ^ 1, lo2, hi2);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], lo2[k]);
            hi[k] = fmaxf(hi[k], hi2[k]);
        }
        reinterpret_cast<int *>(&nodes[p])[12] = ch.x;
        reinterpret_cast<int *>(&nodes[p])[13] = ch.y;
        reinterpret_cast<int *>(&nodes[p])[14] = 0;
        reinterpret_cast<int *>(&nodes[p])[15] = 0;
        child = p;
        p = node_parent[p];
        ++depth;
    }
    (void)depth;
    (void)max_depth;
}

__global__ void k_depth(int n, const int *__restrict__ node_parent, const int *__restrict__ leaf_parent,
                        int *max_depth)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int depth = 0;
    if (i < n) {
        int p = leaf_parent[i];
        while (p >= 0) {
            ++depth;
            p = node_parent[p];
        }
    }
    for (int o = 16; o > 0; o >>= 1) depth = max(depth, __shfl_xor_sync(0xffffffffu, depth, o));
    if ((threadIdx.x & 31) == 0 && depth > 0) atomicMax(max_depth, depth);
}

__global__ void k_single_prim_node(const float4 *prim_lo, const float4 *prim_hi, float pad, BvhNode *nodes)
{
    float4 a = prim_lo[0], b = prim_hi[0];
    BvhNode nd;
    nd.q0 = make_float4(a.x - pad, a.y - pad, a.z - pad, b.x + pad);
    nd.q1 = make_float4(b.y + pad, b.z + pad, FLT_MAX, FLT_MAX);
    nd.q2 = make_float4(FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX); // empty second box: never entered
    nd.q3 = make_int4(~0, ~0, 0, 0);
    nodes[0] = nd;
}

} // namespace

// (stream-ordered frees on the stream the hierarchy was built on — no device-wide synchronisation as with cudaFree, and
// ordered after the work the caller enqueued on that stream; work on OTHER streams must be finished by the caller)
void free_bvh(Bvh *b)
{
    if (b->nodes) cudaFreeAsync(b->nodes, b->stream);
    if (b->prim_order) cudaFreeAsync(b->prim_order, b->stream);
    b->nodes = nullptr;
    b->prim_order = nullptr;
}

// Stream-ordered allocations come from the device's default memory pool; without a release threshold the pool gives
// its memory back at every synchronisation and the next build pays for cudaMalloc again.
void keep_pool_memory()
{
    static bool done = false;
    if (done) return;
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t threshold = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    done = true;
}

int build_lbvh(const float4 *d_lo, const float4 *d_hi, int64_t n, cudaStream_t stream, Bvh *out, float min_pad)
{
    keep_pool_memory();
    *out = Bvh();
    out->n_prims = n;
    out->stream = stream;
    if (n <= 0) return RFRT_OK;
    if (n >= (1ll << 31)) {
        set_error("build_lbvh: more than 2^31-1 primitives");
        return RFRT_ERR_INVALID;
    }
    const int nblocks = (int)((n + SORT_TILE - 1) / SORT_TILE);
    uint64_t *keys_a = nullptr, *keys_b = nullptr;
    uint32_t *vals_b = nullptr;
    uint32_t *ghist = nullptr;
    int *bounds_enc = nullptr, *node_parent = nullptr, *leaf_parent = nullptr, *arrive = nullptr, *d_depth = nullptr;
    float *d_bounds = nullptr;
    int2 *children = nullptr;
    const int64_t n_nodes = n > 1 ? n - 1 : 1;

    Temporaries tmp; // (out->nodes / out->prim_order belong to the caller's Bvh: free_bvh)
    RFRT_CUDA(tmp.alloc_async(&keys_a, sizeof(uint64_t) * n, stream));
    RFRT_CUDA(tmp.alloc_async(&keys_b, sizeof(uint64_t) * n, stream));
    RFRT_CUDA(tmp.alloc_async(&vals_b, sizeof(uint32_t) * n, stream));
    RFRT_CUDA(tmp.alloc_async(&ghist, sizeof(uint32_t) * 256 * ((size_t)nblocks + 1), stream));
    RFRT_CUDA(tmp.alloc_async(&bounds_enc, sizeof(int) * 8, stream));
    RFRT_CUDA(tmp.alloc_async(&d_bounds, sizeof(float) * 8, stream));
    RFRT_CUDA(tmp.alloc_async(&node_parent, sizeof(int) * n_nodes, stream));
    RFRT_CUDA(tmp.alloc_async(&leaf_parent, sizeof(int) * n, stream));
    RFRT_CUDA(tmp.alloc_async(&arrive, sizeof(int) * n_nodes, stream));
    RFRT_CUDA(tmp.alloc_async(&children, sizeof(int2) * n_nodes, stream));
    RFRT_CUDA(tmp.alloc_async(&d_depth, sizeof(int), stream));
    RFRT_CUDA(cudaMallocAsync(&out->nodes, sizeof(BvhNode) * n_nodes, stream));
    RFRT_CUDA(cudaMallocAsync(&out->prim_order, sizeof(int32_t) * n, stream));

    const int T = 256;
    const int nb = (int)((n + T - 1) / T);
    k_bounds_init<<<1, 32, 0, stream>>>(bounds_enc);
    k_scene_bounds<<<nb < 1184 ? nb : 1184, T, 0, stream>>>(d_lo, d_hi, n, bounds_enc);
    k_bounds_decode<<<1, 32, 0, stream>>>(bounds_enc, d_bounds);
    RFRT_CUDA(cudaMemcpyAsync(out->bounds, d_bounds, sizeof(float) * 6, cudaMemcpyDeviceToHost, stream));
    RFRT_CUDA(cudaStreamSynchronize(stream));
    float m = 0.0f;
    for (int k = 0; k < 6; ++k) m = fmaxf(m, fabsf(out->bounds[k]));
    // padding of the boxes: what the slab test and the exact test can disagree by because of the MESH's coordinates
    // (ulps of max |coordinate| = m: 1e-5 m is ~80 of them); the part that grows with the ray's origin is the slab
    // test's own per-ray offsets (rfrt_trace.cuh: SlabRay).  RFRT_BVH_PAD overrides the floor of environment meshes for
    // A/B runs (round 1: 1e-3).
    if (min_pad == BVH_PAD_MESH && getenv("RFRT_BVH_PAD")) min_pad = (float)atof(getenv("RFRT_BVH_PAD"));
    out->pad = fmaxf(min_pad, 1.0e-5f * m);

    // values: the primitive index; ping-pongs between out->prim_order and vals_b and ends in out->prim_order
    uint32_t *order = reinterpret_cast<uint32_t *>(out->prim_order);
    static const int legacy_morton = getenv("RFRT_BVH_LEGACY_MORTON") ? atoi(getenv("RFRT_BVH_LEGACY_MORTON")) : 0;
    // The walk's stack holds 64 entries, and a hierarchy over 63-bit codes can be deeper than that when the primitives
    // cluster at many scales (one level per code bit that splits something off).  Such a scene is rebuilt over coarser
    // cells — 14, then 10 bits per axis: at most 30 levels of code plus log2(n) levels of equal codes split by position.
    const int bit_choices[3] = {21, 14, 10};
    for (int attempt = 0; attempt < 3; ++attempt) {
        k_morton<<<nb, T, 0, stream>>>(d_lo, d_hi, n, d_bounds, keys_a, order, legacy_morton, 21 - bit_choices[attempt]);
        uint64_t *src = keys_a, *dst = keys_b;
        uint32_t *vsrc = order, *vdst = vals_b;
        for (int pass = 0; pass < 8; ++pass) {
            int shift = 8 * pass;
            k_sort_hist<<<nblocks, SORT_THREADS, 0, stream>>>(src, n, shift, ghist, nblocks);
            k_scan_rows<<<256, 256, 0, stream>>>(ghist, nblocks, ghist + 256ll * nblocks);
            k_scan_totals<<<1, 256, 0, stream>>>(ghist + 256ll * nblocks);
            k_sort_scatter<true><<<nblocks, SORT_THREADS, 0, stream>>>(src, dst, n, shift, ghist, ghist + 256ll * nblocks, nblocks, vsrc, vdst);
            uint64_t *tmpk = src; src = dst; dst = tmpk;
            uint32_t *tmpv = vsrc; vsrc = vdst; vdst = tmpv;
        }
        // after 8 passes the sorted codes are back in keys_a (src) and the sorted primitive indices in out->prim_order
        if (n == 1) {
            k_single_prim_node<<<1, 1, 0, stream>>>(d_lo, d_hi, out->pad, out->nodes);
            out->max_depth = 1;
            break;
        }
        RFRT_CUDA(cudaMemsetAsync(d_depth, 0, sizeof(int), stream));
        k_karras<<<(int)((n - 1 + T - 1) / T), T, 0, stream>>>(src, (int)n, children, node_parent, leaf_parent);
        k_depth<<<nb, T, 0, stream>>>((int)n, node_parent, leaf_parent, d_depth);
        RFRT_CUDA(cudaMemcpyAsync(&out->max_depth, d_depth, sizeof(int), cudaMemcpyDeviceToHost, stream));
        RFRT_CUDA(cudaStreamSynchronize(stream));
        const int depth_limit = getenv("RFRT_BVH_MAX_DEPTH") ? atoi(getenv("RFRT_BVH_MAX_DEPTH")) : BVH_MAX_DEPTH; // (test aid)
        if (out->max_depth > depth_limit && attempt < 2 && !legacy_morton) continue;
        RFRT_CUDA(cudaMemsetAsync(arrive, 0, sizeof(int) * n_nodes, stream));
        k_refit<<<nb, T, 0, stream>>>(order, (int)n, d_lo, d_hi, out->pad, children, node_parent, leaf_parent, arrive,
                                      out->nodes, d_depth);
        break;
    }
    out->n_nodes = n_nodes;
    RFRT_CUDA(cudaStreamSynchronize(stream));
    RFRT_CUDA(cudaGetLastError());

    return RFRT_OK;
}

// LSD radix sort of 64-bit keys on bits [shift, shift + 8*passes): returns the buffer holding the result (a or b).
// hist must hold 256 * (sort_hist_blocks(n) + 1) counters.
int64_t sort_hist_blocks(int64_t n) { return (n + SORT_TILE - 1) / SORT_TILE; }

uint64_t *radix_sort_u64(uint64_t *a, uint64_t *b, uint32_t *hist, int64_t n, int shift, int passes, cudaStream_t stream)
{
    const int nblocks = (int)sort_hist_blocks(n);
    uint64_t *src = a, *dst = b;
    for (int pass = 0; pass < passes; ++pass) {
        const int sh = shift + 8 * pass;
        k_sort_hist<<<nblocks, SORT_THREADS, 0, stream>>>(src, n, sh, hist, nblocks);
        k_scan_rows<<<256, 256, 0, stream>>>(hist, nblocks, hist + 256ll * nblocks);
        k_scan_totals<<<1, 256, 0, stream>>>(hist + 256ll * nblocks);
        k_sort_scatter<false><<<nblocks, SORT_THREADS, 0, stream>>>(src, dst, n, sh, hist, hist + 256ll * nblocks, nblocks, nullptr, nullptr);
        uint64_t *tmp = src; src = dst; dst = tmp;
    }
    return src;
}

} // namespace rfrt
