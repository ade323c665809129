// rfrt_internal.h — host/device structures shared by the translation units of librfrt.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/rfrt.h"

namespace rfrt {

// 64-byte BVH node: the two CHILD boxes live in the parent, so one node fetch (4 x 128-bit loads)
// decides both descents.  Boxes are stored already padded (see bvh pad rule in DESIGN.md).
//   q0 = (lo0.x, lo0.y, lo0.z, hi0.x)   q1 = (hi0.y, hi0.z, lo1.x, lo1.y)
//   q2 = (lo1.z, hi1.x, hi1.y, hi1.z)   q3 = (child0, child1, -, -) as int32
// child >= 0: internal node index; child < 0: leaf, ~child = sorted primitive slot.
struct __align__(16) BvhNode {
    float4 q0, q1, q2;
    int4 q3;
};
static_assert(sizeof(BvhNode) == 64, "node must be 64 bytes");

// 48-byte triangle in sorted (Morton) order: three 128-bit loads.
//   v0 = (a.x, a.y, a.z, b.x)  v1 = (b.y, b.z, c.x, c.y)  v2 = (c.z, as_float(triangle index), 0, 0)
struct __align__(16) BvhTri {
    float4 v0, v1, v2;
};
static_assert(sizeof(BvhTri) == 48, "triangle must be 48 bytes");

struct Bvh {
    BvhNode *nodes = nullptr; // n_nodes = max(n_prims - 1, 1) (0 when n_prims == 0)
    int32_t *prim_order = nullptr; // sorted slot -> primitive index
    int64_t n_prims = 0;
    int64_t n_nodes = 0;
    int32_t max_depth = 0;
    float bounds[6] = {0, 0, 0, 0, 0, 0};
    float pad = 0.0f;
    cudaStream_t stream = nullptr; // the stream the hierarchy was allocated on: free_bvh frees in its order
};

struct Mesh {
    Bvh bvh;
    BvhTri *tris = nullptr;      // sorted order, for traversal
    float *soup = nullptr;       // [n*9] original order (a,b,c)
    float4 *normals = nullptr;   // [n] sorted order: normalize(cross(b-a, c-a)) precomputed with the trace's own ops
    float *face_normals = nullptr; // [n*3] original order
    // small scenes: shared-memory image of the lockstep sweep (rfrt_small.cu), NULL when the scene does not fit:
    // recs[28*pairs] | nbr[4*n] | slot_tri[2*pairs] | soup[9*n] | face_normals[3*n] | tri_slot[n]   (first two: 16-byte rows)
    float *small = nullptr;
    int32_t small_pairs = 0;
    int32_t small_class[5] = {0, 0, 0, 0, 0}; // pair ranges of the plane classes (general, x-, y-, z-aligned)
    float small_extent = 0.0f;
    float build_ms = 0.0f;
    int32_t tri_test = 0;        // RFRT_TRI_TEST_WOOP (reference-faithful) / RFRT_TRI_TEST_MT (rfrt_mesh_set_triangle_test)
    float *materials = nullptr;  // [n] refractive index per triangle (rfrt_mesh_set_materials) or NULL = 5.0 everywhere
    // BVH scenes: workspace of the direction-coherent ray order (grown on demand by rfrt_trace, freed with the mesh)
    uint64_t *ray_keys[2] = {nullptr, nullptr};
    uint32_t *ray_hist = nullptr;
    uint32_t *ray_cells = nullptr; // counters of the 2^24 direction cells + their block sums (counting sort of a wave)
    int64_t ray_cap = 0;
};

struct RxSet {
    Bvh bvh;                  // over receiver bounding boxes
    Bvh unit_bvh;             // over the UNIT icosphere's triangles (shared by all receivers: translate + scale)
    float *verts = nullptr;   // [R * n_unit * 3] fp32
    int64_t n_receivers = 0;
    int32_t n_unit = 0;
    int32_t n_faces = 0;
    uint8_t faces[3 * 128];   // host copy of the face table (<= 128 faces), uploaded to __constant__
    double radius = 0.0;
    double *centers = nullptr; // [R*3] device copy
    float *unit_recs = nullptr; // [n_faces*16] unit-space face records of the receiver-query filter (rfrt_small.cu)
    cudaStream_t stream = nullptr; // creating stream (stream-ordered frees in rfrt_rxset_destroy)
};

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

#define RFRT_CUDA(call)                                                  \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) return ::rfrt::cuda_fail(_e, #call);      \
    } while (0)

// Device temporaries of a constructor: released when the scope ends, on success and on every early error return alike.
struct Temporaries {
    std::vector<void *> sync_ptrs;                           // from cudaMalloc
    std::vector<std::pair<void *, cudaStream_t>> async_ptrs; // from cudaMallocAsync (freed in stream order)
    std::vector<cudaEvent_t> events;
    Temporaries() = default;
    Temporaries(const Temporaries &) = delete;
    Temporaries &operator=(const Temporaries &) = delete;
    ~Temporaries()
    {
        for (void *q : sync_ptrs) cudaFree(q);
        for (auto &q : async_ptrs) cudaFreeAsync(q.first, q.second);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
    template <class T>
    cudaError_t alloc_async(T **p, size_t bytes, cudaStream_t stream)
    {
        cudaError_t e = cudaMallocAsync(p, bytes, stream);
        if (e == cudaSuccess) async_ptrs.push_back({*p, stream});
        return e;
    }
};

// Builds a BVH over n axis-aligned boxes (device arrays lo/hi as float4 per primitive, w ignored).
// Allocates bvh.nodes / bvh.prim_order.  Synchronises the stream.
// min_pad: floor of the boxes' padding (environment meshes: BVH_PAD_MESH, a formality — the padding is 1e-5 * max
// |coordinate|; receiver sets and their unit shape: BVH_PAD_RX = Warp's 1e-3)
constexpr float BVH_PAD_MESH = 1.0e-7f;
constexpr float BVH_PAD_RX = 1.0e-3f;
int build_lbvh(const float4 *d_lo, const float4 *d_hi, int64_t n, cudaStream_t stream, Bvh *out, float min_pad);
void free_bvh(Bvh *b);
void keep_pool_memory();
int64_t sort_hist_blocks(int64_t n);
uint64_t *radix_sort_u64(uint64_t *a, uint64_t *b, uint32_t *hist, int64_t n, int shift, int passes, cudaStream_t stream);

int small_scene_tables(const float *soup, int n_tris, float *recs, int32_t *slot_tri, int32_t *n_pairs, float *extent,
                       int32_t *class_begin);
void small_scene_neighbours(const float *soup, int n_tris, const int32_t *slot_tri, int n_pairs, double reach, uint32_t *nbr,
                            int32_t *tri_slot);
inline size_t small_image_floats(int n_pairs, int n_tris) { return 30 * (size_t)n_pairs + 17 * (size_t)n_tris; }
// self-re-hit shortcut (rfrt_trace.cuh): applies to hits of the ray's own triangle within SMALL_TAU_REL * extent
// (path length); neighbours are collected within SMALL_REACH_REL * extent.  A triangle that beats or ties the re-hit is
// hit within tau of the origin, and the origin lies on the ray's own triangle to fp32 rounding (~1e-6 * extent), so
// reach only has to exceed tau: 2.5 tau.  (Round 1: 1e-4 and 1e-3 — the boundary strip of width 2 reach + tau + tolerance
// then made one lane of every third re-hit trip walk its neighbour pairs alone.)
constexpr double SMALL_TAU_REL = 2.5e-5;
constexpr double SMALL_REACH_REL = 6.25e-5;

void unit_face_records(const double *unit_v, const int32_t *faces, int n_faces, float *recs);

Mesh *get_mesh(rfrt_handle h);
RxSet *get_rxset(rfrt_handle h);

} // namespace rfrt
