// rfrt_api.cu — handle registry, error reporting, mesh / receiver-set construction (C ABI).
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "rfrt_internal.h"
#include "rfrt_math.cuh"

namespace rfrt {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what)
{
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return RFRT_ERR_CUDA;
}

namespace {

std::mutex g_mutex;
std::unordered_map<rfrt_handle, Mesh *> g_meshes;
std::unordered_map<rfrt_handle, RxSet *> g_rxsets;
rfrt_handle g_next_handle = 0x1000;

__global__ void k_tri_boxes(const float *__restrict__ verts, int64_t n_verts, const int32_t *__restrict__ idx,
                            int64_t n_tris, float *__restrict__ soup, float4 *__restrict__ lo, float4 *__restrict__ hi,
                            int *bad)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_tris) return;
    float v[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int64_t vi = idx[3 * i + c];
        if (vi < 0 || vi >= n_verts) { *bad = 1; vi = 0; }
        v[3 * c] = verts[3 * vi]; v[3 * c + 1] = verts[3 * vi + 1]; v[3 * c + 2] = verts[3 * vi + 2];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) soup[9 * i + k] = v[k];
    lo[i] = make_float4(fminf(fminf(v[0], v[3]), v[6]), fminf(fminf(v[1], v[4]), v[7]), fminf(fminf(v[2], v[5]), v[8]), 0.f);
    hi[i] = make_float4(fmaxf(fmaxf(v[0], v[3]), v[6]), fmaxf(fmaxf(v[1], v[4]), v[7]), fmaxf(fmaxf(v[2], v[5]), v[8]), 0.f);
}

// Also precomputes each triangle's reflection normal (kernel.py:82,96: the normal mesh_query_ray returns) with the
// very same operation sequence the trace would use — it does not depend on the ray, so it is hoisted to build time.
__global__ void k_pack_tris(const float *__restrict__ soup, const int32_t *__restrict__ order, int64_t n,
                            BvhTri *__restrict__ tris, float4 *__restrict__ normals, float *__restrict__ face_normals)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    int32_t f = order[s];
    const float *v = soup + 9 * (int64_t)f;
    float3 nrm = tri_normal(make_float3(v[0], v[1], v[2]), make_float3(v[3], v[4], v[5]), make_float3(v[6], v[7], v[8]));
    normals[s] = make_float4(nrm.x, nrm.y, nrm.z, 0.f);
    face_normals[3 * (int64_t)f] = nrm.x; face_normals[3 * (int64_t)f + 1] = nrm.y; face_normals[3 * (int64_t)f + 2] = nrm.z;
    BvhTri t;
    t.v0 = make_float4(v[0], v[1], v[2], v[3]);
    t.v1 = make_float4(v[4], v[5], v[6], v[7]);
    t.v2 = make_float4(v[8], __int_as_float(f), 0.f, 0.f);
    tris[s] = t;
}

// receiver k, unit vertex j:  float32(center + radius*unit)  in fp64, one rounding (tracer.py:27-28)
__global__ void k_rx_vertices(const double *__restrict__ centers, int64_t n_rx, double radius,
                              const double *__restrict__ unit, int n_unit, float *__restrict__ verts,
                              float4 *__restrict__ lo, float4 *__restrict__ hi)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_rx) return;
    double c[3] = {centers[3 * k], centers[3 * k + 1], centers[3 * k + 2]};
    float l[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, h[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int j = 0; j < n_unit; ++j) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float x = __double2float_rn(__dadd_rn(c[a], __dmul_rn(radius, unit[3 * j + a])));
            verts[(k * n_unit + j) * 3 + a] = x;
            l[a] = fminf(l[a], x);
            h[a] = fmaxf(h[a], x);
        }
    }
    lo[k] = make_float4(l[0], l[1], l[2], 0.f);
    hi[k] = make_float4(h[0], h[1], h[2], 0.f);
}

// Device memory of a mesh / receiver set (also the clean-up of a half-built one: every pointer starts as NULL)
void release_mesh(Mesh *m)
{
    free_bvh(&m->bvh);
    if (m->tris) cudaFree(m->tris);
    if (m->soup) cudaFree(m->soup);
    if (m->normals) cudaFree(m->normals);
    if (m->face_normals) cudaFree(m->face_normals);
    if (m->small) cudaFree(m->small);
    if (m->materials) cudaFree(m->materials);
    if (m->ray_keys[0]) cudaFree(m->ray_keys[0]);
    if (m->ray_keys[1]) cudaFree(m->ray_keys[1]);
    if (m->ray_hist) cudaFree(m->ray_hist);
    if (m->ray_cells) cudaFree(m->ray_cells);
    *m = Mesh();
}

void release_rxset(RxSet *r)
{
    const cudaStream_t stream = r->stream; // the creating stream: frees are ordered after the work enqueued on it
    free_bvh(&r->bvh);
    free_bvh(&r->unit_bvh);
    if (r->verts) cudaFreeAsync(r->verts, stream);
    if (r->centers) cudaFreeAsync(r->centers, stream);
    if (r->unit_recs) cudaFreeAsync(r->unit_recs, stream);
    r->verts = nullptr; r->centers = nullptr; r->unit_recs = nullptr;
}

struct MeshDeleter {
    void operator()(Mesh *m) const { release_mesh(m); delete m; }
};
struct RxSetDeleter {
    void operator()(RxSet *r) const { release_rxset(r); delete r; }
};

} // namespace

Mesh *get_mesh(rfrt_handle h)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_meshes.find(h);
    return it == g_meshes.end() ? nullptr : it->second;
}

RxSet *get_rxset(rfrt_handle h)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_rxsets.find(h);
    return it == g_rxsets.end() ? nullptr : it->second;
}

} // namespace rfrt

using namespace rfrt;

extern "C" int rfrt_version(void) { return RFRT_VERSION; }

extern "C" const char *rfrt_last_error(void) { return g_last_error.c_str(); }

extern "C" int rfrt_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor)
{
    int dev = 0, n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        set_error("rfrt_device_info: no CUDA device visible");
        return RFRT_ERR_NO_DEVICE;
    }
    RFRT_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) { RFRT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (cc_major) { RFRT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { RFRT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return RFRT_OK;
}

extern "C" int rfrt_mesh_create(const float *d_vertices_xyz, int64_t n_vertices, const int32_t *d_indices,
                                int64_t n_triangles, void *stream_, rfrt_handle *out_mesh)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!out_mesh || n_vertices < 0 || n_triangles < 0 || (n_triangles > 0 && (!d_vertices_xyz || !d_indices))) {
        set_error("rfrt_mesh_create: bad arguments");
        return RFRT_ERR_INVALID;
    }
    std::unique_ptr<Mesh, MeshDeleter> m(new Mesh());
    Temporaries tmp;
    cudaEvent_t e0, e1;
    RFRT_CUDA(cudaEventCreate(&e0));
    tmp.events.push_back(e0);
    RFRT_CUDA(cudaEventCreate(&e1));
    tmp.events.push_back(e1);
    RFRT_CUDA(cudaEventRecord(e0, stream));
    if (n_triangles > 0) {
        float4 *lo = nullptr, *hi = nullptr;
        int *bad = nullptr;
        RFRT_CUDA(cudaMalloc(&m->soup, sizeof(float) * 9 * n_triangles));
        RFRT_CUDA(cudaMalloc(&lo, sizeof(float4) * n_triangles));
        tmp.sync_ptrs.push_back(lo);
        RFRT_CUDA(cudaMalloc(&hi, sizeof(float4) * n_triangles));
        tmp.sync_ptrs.push_back(hi);
        RFRT_CUDA(cudaMalloc(&bad, sizeof(int)));
        tmp.sync_ptrs.push_back(bad);
        RFRT_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
        const int T = 256;
        const unsigned nb = (unsigned)((n_triangles + T - 1) / T);
        k_tri_boxes<<<nb, T, 0, stream>>>(d_vertices_xyz, n_vertices, d_indices, n_triangles, m->soup, lo, hi, bad);
        int h_bad = 0;
        RFRT_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
        RFRT_CUDA(cudaStreamSynchronize(stream));
        if (h_bad) {
            set_error("rfrt_mesh_create: face index out of range");
            return RFRT_ERR_INVALID;
        }
        int rc = build_lbvh(lo, hi, n_triangles, stream, &m->bvh, BVH_PAD_MESH);
        if (rc) return rc;
        RFRT_CUDA(cudaMalloc(&m->tris, sizeof(BvhTri) * n_triangles));
        RFRT_CUDA(cudaMalloc(&m->normals, sizeof(float4) * n_triangles));
        RFRT_CUDA(cudaMalloc(&m->face_normals, sizeof(float) * 3 * n_triangles));
        k_pack_tris<<<nb, T, 0, stream>>>(m->soup, m->bvh.prim_order, n_triangles, m->tris, m->normals, m->face_normals);
        RFRT_CUDA(cudaGetLastError());
        RFRT_CUDA(cudaEventRecord(e1, stream));
        RFRT_CUDA(cudaStreamSynchronize(stream));
        RFRT_CUDA(cudaEventElapsedTime(&m->build_ms, e0, e1));
        if (n_triangles <= RFRT_SMALL_MAX_TRIS) {
            // small scene: the shared-memory image of the lockstep sweep (candidate-filter tables + exact-test data)
            const int n = (int)n_triangles;
            std::vector<float> soup(9 * n), fnorm(3 * n), recs(28 * (RFRT_SMALL_MAX_TRIS / 2));
            std::vector<int32_t> slot_tri(RFRT_SMALL_MAX_TRIS);
            RFRT_CUDA(cudaMemcpy(soup.data(), m->soup, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost));
            RFRT_CUDA(cudaMemcpy(fnorm.data(), m->face_normals, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
            int32_t n_pairs = 0;
            if (small_scene_tables(soup.data(), n, recs.data(), slot_tri.data(), &n_pairs, &m->small_extent, m->small_class) == RFRT_OK) {
                m->small_pairs = n_pairs;
                std::vector<float> image(small_image_floats(n_pairs, n));
                float *w = image.data();
                std::vector<uint32_t> nbr(4 * n);
                std::vector<int32_t> tri_slot(n);
                small_scene_neighbours(soup.data(), n, slot_tri.data(), n_pairs, SMALL_REACH_REL * (double)m->small_extent, nbr.data(),
                                       tri_slot.data());
                memcpy(w, recs.data(), sizeof(float) * 28 * n_pairs); w += 28 * n_pairs;
                memcpy(w, nbr.data(), sizeof(uint32_t) * 4 * n); w += 4 * n;
                memcpy(w, slot_tri.data(), sizeof(int32_t) * 2 * n_pairs); w += 2 * n_pairs;
                memcpy(w, soup.data(), sizeof(float) * 9 * n); w += 9 * n;
                memcpy(w, fnorm.data(), sizeof(float) * 3 * n); w += 3 * n;
                memcpy(w, tri_slot.data(), sizeof(int32_t) * n);
                RFRT_CUDA(cudaMalloc(&m->small, sizeof(float) * image.size()));
                RFRT_CUDA(cudaMemcpy(m->small, image.data(), sizeof(float) * image.size(), cudaMemcpyHostToDevice));
            } // else: too many distinct planes for 64 slots -> BVH path
        }
    }
    std::lock_guard<std::mutex> lock(g_mutex);
    rfrt_handle h = g_next_handle++;
    g_meshes[h] = m.release();
    *out_mesh = h;
    return RFRT_OK;
}

extern "C" int rfrt_mesh_destroy(rfrt_handle mesh)
{
    Mesh *m = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_meshes.find(mesh);
        if (it == g_meshes.end()) { set_error("rfrt_mesh_destroy: unknown handle"); return RFRT_ERR_HANDLE; }
        m = it->second;
        g_meshes.erase(it);
    }
    release_mesh(m);
    delete m;
    return RFRT_OK;
}

extern "C" int rfrt_mesh_set_materials(rfrt_handle mesh, const float *d_refractive_index, void *stream_)
{
    Mesh *m = get_mesh(mesh);
    if (!m) { set_error("rfrt_mesh_set_materials: unknown handle"); return RFRT_ERR_HANDLE; }
    if (!d_refractive_index) {
        if (m->materials) cudaFree(m->materials);
        m->materials = nullptr;
        return RFRT_OK;
    }
    if (!m->materials) RFRT_CUDA(cudaMalloc(&m->materials, sizeof(float) * (size_t)(m->bvh.n_prims > 0 ? m->bvh.n_prims : 1)));
    RFRT_CUDA(cudaMemcpyAsync(m->materials, d_refractive_index, sizeof(float) * (size_t)m->bvh.n_prims, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream_));
    return RFRT_OK;
}

extern "C" int rfrt_mesh_set_triangle_test(rfrt_handle mesh, int32_t kind)
{
    Mesh *m = get_mesh(mesh);
    if (!m) { set_error("rfrt_mesh_set_triangle_test: unknown handle"); return RFRT_ERR_HANDLE; }
    if (kind != RFRT_TRI_TEST_WOOP && kind != RFRT_TRI_TEST_MT) {
        set_error("rfrt_mesh_set_triangle_test: kind must be RFRT_TRI_TEST_WOOP or RFRT_TRI_TEST_MT");
        return RFRT_ERR_INVALID;
    }
    m->tri_test = kind;
    return RFRT_OK;
}

extern "C" int rfrt_mesh_info(rfrt_handle mesh, int64_t *n_triangles, int64_t *n_nodes, float *h_bounds6,
                              int32_t *max_depth, float *build_ms)
{
    Mesh *m = get_mesh(mesh);
    if (!m) { set_error("rfrt_mesh_info: unknown handle"); return RFRT_ERR_HANDLE; }
    if (n_triangles) *n_triangles = m->bvh.n_prims;
    if (n_nodes) *n_nodes = m->bvh.n_nodes;
    if (h_bounds6) for (int k = 0; k < 6; ++k) h_bounds6[k] = m->bvh.bounds[k];
    if (max_depth) *max_depth = m->bvh.max_depth;
    if (build_ms) *build_ms = m->build_ms;
    return RFRT_OK;
}

extern "C" int rfrt_mesh_export(rfrt_handle mesh, float *d_nodes, int32_t *d_tri_order, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Mesh *m = get_mesh(mesh);
    if (!m) { set_error("rfrt_mesh_export: unknown handle"); return RFRT_ERR_HANDLE; }
    if (d_nodes && m->bvh.n_nodes > 0)
        RFRT_CUDA(cudaMemcpyAsync(d_nodes, m->bvh.nodes, sizeof(BvhNode) * m->bvh.n_nodes, cudaMemcpyDeviceToDevice, stream));
    if (d_tri_order && m->bvh.n_prims > 0)
        RFRT_CUDA(cudaMemcpyAsync(d_tri_order, m->bvh.prim_order, sizeof(int32_t) * m->bvh.n_prims, cudaMemcpyDeviceToDevice, stream));
    return RFRT_OK;
}

extern "C" int rfrt_rxset_create(const double *d_centers_xyz, int64_t n_receivers, double radius,
                                 const double *h_unit_vertices, int32_t n_unit_vertices, const int32_t *h_faces,
                                 int32_t n_faces, void *stream_, rfrt_handle *out_rxset)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!out_rxset || !d_centers_xyz || n_receivers <= 0 || !h_unit_vertices || !h_faces || n_unit_vertices <= 0 ||
        n_unit_vertices > 255 || n_faces <= 0 || n_faces > 128 || !(radius > 0.0)) {
        set_error("rfrt_rxset_create: bad arguments (need 1..255 unit vertices, 1..128 faces, radius > 0)");
        return RFRT_ERR_INVALID;
    }
    keep_pool_memory();
    std::unique_ptr<RxSet, RxSetDeleter> r(new RxSet());
    Temporaries tmp;
    r->n_receivers = n_receivers; r->n_unit = n_unit_vertices; r->n_faces = n_faces; r->radius = radius;
    r->stream = stream;
    for (int i = 0; i < 3 * n_faces; ++i) {
        if (h_faces[i] < 0 || h_faces[i] >= n_unit_vertices) { set_error("rfrt_rxset_create: face index out of range"); return RFRT_ERR_INVALID; }
        r->faces[i] = (uint8_t)h_faces[i];
    }
    double *d_unit = nullptr;
    float4 *lo = nullptr, *hi = nullptr;
    RFRT_CUDA(cudaMallocAsync(&d_unit, sizeof(double) * 3 * n_unit_vertices, stream));
    tmp.async_ptrs.push_back({d_unit, stream});
    RFRT_CUDA(cudaMallocAsync(&r->centers, sizeof(double) * 3 * n_receivers, stream));
    RFRT_CUDA(cudaMallocAsync(&r->verts, sizeof(float) * 3 * n_unit_vertices * n_receivers, stream));
    RFRT_CUDA(cudaMallocAsync(&lo, sizeof(float4) * n_receivers, stream));
    tmp.async_ptrs.push_back({lo, stream});
    RFRT_CUDA(cudaMallocAsync(&hi, sizeof(float4) * n_receivers, stream));
    tmp.async_ptrs.push_back({hi, stream});
    RFRT_CUDA(cudaMemcpyAsync(d_unit, h_unit_vertices, sizeof(double) * 3 * n_unit_vertices, cudaMemcpyHostToDevice, stream));
    RFRT_CUDA(cudaMemcpyAsync(r->centers, d_centers_xyz, sizeof(double) * 3 * n_receivers, cudaMemcpyDeviceToDevice, stream));
    const int T = 128;
    k_rx_vertices<<<(unsigned)((n_receivers + T - 1) / T), T, 0, stream>>>(r->centers, n_receivers, radius, d_unit,
                                                                           n_unit_vertices, r->verts, lo, hi);
    RFRT_CUDA(cudaGetLastError());
    int rc = build_lbvh(lo, hi, n_receivers, stream, &r->bvh, BVH_PAD_RX);
    if (rc) return rc;
    {
        // BVH over the unit icosphere's faces, used (after mapping the ray into unit space) to prune the exact
        // per-receiver triangle tests.  Boxes are inflated: the mapping (o - c) / r is only approximate in fp32.
        std::vector<float4> ulo(n_faces), uhi(n_faces);
        for (int f = 0; f < n_faces; ++f) {
            float l[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, h[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int c = 0; c < 3; ++c)
                for (int a = 0; a < 3; ++a) {
                    float x = (float)h_unit_vertices[3 * h_faces[3 * f + c] + a];
                    l[a] = fminf(l[a], x); h[a] = fmaxf(h[a], x);
                }
            const float infl = 4.0e-3f;
            ulo[f] = make_float4(l[0] - infl, l[1] - infl, l[2] - infl, 0.f);
            uhi[f] = make_float4(h[0] + infl, h[1] + infl, h[2] + infl, 0.f);
        }
        float4 *dlo = nullptr, *dhi = nullptr;
        RFRT_CUDA(cudaMallocAsync(&dlo, sizeof(float4) * n_faces, stream));
        tmp.async_ptrs.push_back({dlo, stream});
        RFRT_CUDA(cudaMallocAsync(&dhi, sizeof(float4) * n_faces, stream));
        tmp.async_ptrs.push_back({dhi, stream});
        RFRT_CUDA(cudaMemcpyAsync(dlo, ulo.data(), sizeof(float4) * n_faces, cudaMemcpyHostToDevice, stream));
        RFRT_CUDA(cudaMemcpyAsync(dhi, uhi.data(), sizeof(float4) * n_faces, cudaMemcpyHostToDevice, stream));
        rc = build_lbvh(dlo, dhi, n_faces, stream, &r->unit_bvh, BVH_PAD_RX); // synchronises the stream
        if (rc) return rc;
    }
    {
        std::vector<float> recs(16 * (size_t)n_faces);
        unit_face_records(h_unit_vertices, h_faces, n_faces, recs.data());
        RFRT_CUDA(cudaMallocAsync(&r->unit_recs, sizeof(float) * recs.size(), stream));
        RFRT_CUDA(cudaMemcpyAsync(r->unit_recs, recs.data(), sizeof(float) * recs.size(), cudaMemcpyHostToDevice, stream));
        RFRT_CUDA(cudaStreamSynchronize(stream)); // (recs is a host temporary)
    }
    std::lock_guard<std::mutex> lock(g_mutex);
    rfrt_handle h = g_next_handle++;
    g_rxsets[h] = r.release();
    *out_rxset = h;
    return RFRT_OK;
}

extern "C" int rfrt_rxset_destroy(rfrt_handle rxset)
{
    RxSet *r = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_rxsets.find(rxset);
        if (it == g_rxsets.end()) { set_error("rfrt_rxset_destroy: unknown handle"); return RFRT_ERR_HANDLE; }
        r = it->second;
        g_rxsets.erase(it);
    }
    release_rxset(r);
    delete r;
    return RFRT_OK;
}

extern "C" int rfrt_rxset_export(rfrt_handle rxset, float *d_vertices, void *stream_)
{
    RxSet *r = get_rxset(rxset);
    if (!r) { set_error("rfrt_rxset_export: unknown handle"); return RFRT_ERR_HANDLE; }
    if (!d_vertices) { set_error("rfrt_rxset_export: null buffer"); return RFRT_ERR_INVALID; }
    RFRT_CUDA(cudaMemcpyAsync(d_vertices, r->verts, sizeof(float) * 3 * r->n_unit * r->n_receivers,
                              cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
    return RFRT_OK;
}
