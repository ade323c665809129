/*
 * rfrt.h — C ABI of librfrt.so, the B200 (sm_100a) RF ray-tracing hot path.
 *
 * This is the drop-in boundary for the ONE path the reference accelerates on a device:
 *   tracer.py:22-24   wp.array(vertices) / wp.array(faces) / wp.Mesh(...)      -> rfrt_mesh_create
 *   tracer.py:26-30   Tracer._generate_rx_mesh (icosphere -> wp.Mesh)          -> rfrt_rxset_create
 *   tracer.py:75-80   wp.launch(kernel.trace_paths_kernel, dim=(N,1,1), inputs=[env.id, tx_pos,
 *                     rx.id, max_bounces, traced_paths, received_paths, row_mask]);
 *                     wp.synchronize_device()                                   -> rfrt_trace_paths_compat
 *                                                                                  (same 7 inputs, dense)
 *                                                                               -> rfrt_trace + rfrt_trace_receive
 *                                                                                  (streaming: no N x (B+1) arrays)
 *   tracer.py:101-117 impulse-response binning (+ _bounce_amplitude :34-61)     -> fused in rfrt_trace_receive,
 *                                                                                  rfrt_bin_ir
 *   main.py:39,46-55 / coverage.py:45-55  RX power from the impulse response    -> rfrt_rx_power
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  Pointers named d_* are DEVICE pointers into buffers
 *     the CALLER owns (e.g. torch CUDA tensors: tensor.data_ptr()); h_* are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); NULL = legacy
 *     default stream.  Calls only enqueue work; they do not synchronise unless stated.
 *   - every function returns 0 on success or a negative rfrt_status; rfrt_last_error() gives the
 *     message of the last failure on the calling thread.  No C++ exception crosses this boundary.
 *   - the library owns only what lives behind a handle (BVH nodes, re-ordered triangles, receiver
 *     vertices, the ray-sort workspace of big scenes) and frees it in *_destroy.  *_destroy frees in the order of the
 *     stream the object was created on: work that uses the handle on OTHER streams must have finished.
 *   - concurrency: calls on DIFFERENT environment handles may run on different streams at the same time; calls that
 *     share an environment handle, and replay / compat calls whose receiver sets have different face tables (the
 *     table lives in __constant__ memory), must be stream-ordered.
 */
#ifndef RFRT_H
#define RFRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFRT_VERSION 100

#if defined(__GNUC__)
#define RFRT_API __attribute__((visibility("default")))
#else
#define RFRT_API
#endif

typedef uint64_t rfrt_handle;

enum rfrt_status {
    RFRT_OK = 0,
    RFRT_ERR_INVALID = -1,  /* bad argument */
    RFRT_ERR_CUDA = -2,     /* a CUDA runtime call or launch failed */
    RFRT_ERR_HANDLE = -3,   /* unknown / stale handle */
    RFRT_ERR_NO_DEVICE = -4 /* no sm_100 device visible */
};

/* rfrt_trace flags */
#define RFRT_FLAG_NONE 0u
#define RFRT_FLAG_FORCE_BVH 2u  /* always walk the BVH (default: scenes of <= 64 triangles use the lockstep sweep) */
#define RFRT_FLAG_DIRS_READY 1u /* d_dir_scratch already holds rfrt_ray_directions(ray_begin, ray_end): one wave */
#define RFRT_FLAG_NO_RAY_SORT 16u /* BVH scenes: keep ray-id order (default: each wave is traced in direction-coherent order) */
#define RFRT_FLAG_CHECKSUM 8u   /* also accumulate RFRT_CTR_CHECKSUM (costs a few instructions per segment) */

#define RFRT_SMALL_MAX_TRIS 64  /* scenes that fit this many filter slots take the lockstep sweep instead of the BVH walk */

/* layout of the u64 counter block written by rfrt_trace / rfrt_trace_receive */
#define RFRT_CTR_SEGMENTS 0    /* traced ray segments (alive bounce iterations), SURVEY.md 8d */
#define RFRT_CTR_CANDIDATES 1  /* (ray, receiver) candidates appended (may exceed capacity) */
#define RFRT_CTR_RECORDS 2     /* received records appended (may exceed capacity) */
#define RFRT_CTR_ENV_HITS 3    /* environment hits among the segments */
#define RFRT_CTR_NEXT_RAY 4    /* internal: persistent-kernel ray fetch cursor */
#define RFRT_CTR_NEXT_CAND 5   /* internal: candidate fetch cursor */
#define RFRT_CTR_CHECKSUM 6    /* with RFRT_FLAG_CHECKSUM: sum over segments of hash(ray id, bounce, hit triangle, bits of t)
                                  (mod 2^64; order-independent, so it compares whole runs of different kernels / GPU counts) */
#define RFRT_CTR_QUEUE_OVERFLOW 7 /* receiver-enumeration queue overflows (must stay 0: results would be incomplete) */
#define RFRT_CTR_NODE_VISITS 8 /* with RFRT_FLAG_CHECKSUM, BVH scenes: internal nodes fetched by the environment walks */
#define RFRT_CTR_TRI_TESTS 9   /* with RFRT_FLAG_CHECKSUM, BVH scenes: exact triangle tests run by the environment walks */
#define RFRT_CTR_COUNT 10

RFRT_API int rfrt_version(void);
RFRT_API const char *rfrt_last_error(void);

/* SM count / compute capability of the current device. */
RFRT_API int rfrt_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Environment mesh + LBVH.   Replaces wp.Mesh(points, velocities=None, indices) at tracer.py:22-24.
 *   d_vertices_xyz : [n_vertices*3] float32      d_indices : [n_triangles*3] int32
 * Builds (on `stream`): per-triangle bounds -> 63-bit Morton codes on a cubic grid -> 8-bit LSD radix sort ->
 * Karras hierarchy -> bottom-up refit.  Triangle index == row of d_indices (== STL facet order).
 * Synchronises the stream once before returning (the handle is ready to use on any stream).
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_mesh_create(const float *d_vertices_xyz, int64_t n_vertices, const int32_t *d_indices,
                     int64_t n_triangles, void *stream, rfrt_handle *out_mesh);
RFRT_API int rfrt_mesh_destroy(rfrt_handle mesh);
/* Sizes the mesh's ray-order workspace (BVH scenes trace each wave in direction-coherent order: 16 B per ray of the
 * largest wave) so that rfrt_trace never allocates; without it the first rfrt_trace call of a larger wave allocates. */
RFRT_API int rfrt_mesh_reserve_rays(rfrt_handle mesh, int64_t max_chunk_rays);
/* Material table for the reference-mode amplitude: tracer.py:43 hard-codes the refractive index n_1 = 5.0 in
 * _bounce_amplitude; with a table the triangle of each path vertex supplies n_1 (receiver vertices keep 5.0).
 *   d_refractive_index : [n_triangles] float32 (device; copied), or NULL to restore the reference's constant. */
RFRT_API int rfrt_mesh_set_materials(rfrt_handle mesh, const float *d_refractive_index, void *stream);
/* The ray/triangle functor behind every closest-hit query on this mesh and on the receivers traced against it
 * (the `mesh_query_ray` of kernel.py:71,82):
 *   RFRT_TRI_TEST_WOOP (default) the watertight Woop/Benthin/Wald test in the operation order of Warp's
 *                      intersect_ray_tri_woop — the reference's arithmetic, bit-exact against the oracle;
 *   RFRT_TRI_TEST_MT   Moeller-Trumbore (two-sided, no epsilon, one fp32 rounding per operation), bit-exact against
 *                      the oracle's mt_tri.  NOT reference behaviour: the two functors round differently, so rays
 *                      that graze an edge or re-hit their own triangle at t ~ 0 diverge (DESIGN.md reports the rate).
 *                      Scenes of <= 64 triangles walk the BVH in this mode (the lockstep sweep's filter is derived
 *                      for the watertight test); rfrt_trace_physical rejects such a mesh. */
#define RFRT_TRI_TEST_WOOP 0
#define RFRT_TRI_TEST_MT 1
RFRT_API int rfrt_mesh_set_triangle_test(rfrt_handle mesh, int32_t kind);
/* bounds6 = {lo.xyz, hi.xyz} (unpadded); any output pointer may be NULL. */
RFRT_API int rfrt_mesh_info(rfrt_handle mesh, int64_t *n_triangles, int64_t *n_nodes, float *h_bounds6,
                   int32_t *max_depth, float *build_ms);
/* Debug/test export of the built hierarchy into caller-owned device buffers:
 *   d_nodes      : [n_nodes*16] float32 (64-byte nodes)  or NULL
 *   d_tri_order  : [n_triangles] int32 — triangle index stored at each sorted slot  or NULL */
RFRT_API int rfrt_mesh_export(rfrt_handle mesh, float *d_nodes, int32_t *d_tri_order, void *stream);

/* Host-only helper (no device work): the candidate-filter tables rfrt_mesh_create builds for a small scene,
 * exposed so that the filter's "superset of the exact test" property can be checked on the CPU.
 *   h_soup [n*9] (a, b, c per triangle); outputs h_recs [64*14] (28 floats per pair of coplanar triangles:
 *   plane (n.xyz, d) + 2 x 3 edge functions (m.xyz, c)), h_slot_tri [64] (slot -> triangle), *n_pairs,
 *   *extent (max |coordinate|), h_nbr [n*4] or NULL (neighbour pair masks of the self-re-hit shortcut: interior,
 *   boundary, and both restricted to lower triangle indices; bit k = pair k holds another triangle within
 *   1e-3 * extent of this one), h_class_begin [5] or NULL (pairs are ordered general / x- / y- / z-aligned planes;
 *   class c = pairs [h_class_begin[c], h_class_begin[c+1])).  RFRT_ERR_INVALID when the scene needs more than 64 slots
 *   (it then takes the BVH path).  Layout and tolerance: csrc/rfrt_small.cu. */
RFRT_API int rfrt_small_scene_tables(const float *h_soup, int32_t n_triangles, float *h_recs, int32_t *h_slot_tri,
                            int32_t *n_pairs, float *extent, uint32_t *h_nbr, int32_t *h_class_begin);

/* ---------------------------------------------------------------------------------------------
 * Receiver set.  Replaces Tracer._generate_rx_mesh (tracer.py:26-30), batched over R receivers:
 * receiver k is the mesh  float32(center_k + radius * unit_vertex_j)  (fp64 arithmetic, one rounding)
 * with the face table given.  A BVH over the receivers' bounding boxes is built for the trace.
 *   d_centers_xyz   : [R*3] float64 (device)
 *   h_unit_vertices : [n_unit_vertices*3] float64 (host) — icosphere(subdivisions=1): 42
 *   h_faces         : [n_faces*3] int32 (host) — 80 faces
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_rxset_create(const double *d_centers_xyz, int64_t n_receivers, double radius,
                      const double *h_unit_vertices, int32_t n_unit_vertices, const int32_t *h_faces,
                      int32_t n_faces, void *stream, rfrt_handle *out_rxset);
RFRT_API int rfrt_rxset_destroy(rfrt_handle rxset);
/* copies the generated receiver vertices ([R*n_unit_vertices*3] float32) to a caller device buffer */
RFRT_API int rfrt_rxset_export(rfrt_handle rxset, float *d_vertices, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Deterministic ray directions (kernel.py:51-52: rand_init(tid) + sample_unit_sphere_surface).
 *   d_dirs : [(ray_end-ray_begin)*4] float32 (x, y, z, 0)
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_ray_directions(int64_t ray_begin, int64_t ray_end, float *d_dirs, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Streaming trace of the environment trajectory of rays [ray_begin, ray_end)  (kernel.py:48-98 with
 * the receiver hits factored out): per segment one closest-hit query against the environment BVH
 * (kernel.py:82) and a receiver test (kernel.py:71,85) against every receiver of `rxset`.  A ray
 * that hits receiver k before the environment appends the candidate (ray id, k, bounce); candidates
 * are later replayed literally by rfrt_trace_receive (which drops repeats of a (ray, k) pair).  No N x (B+1) array is ever materialised.
 *   h_tx_pos        : 3 floats (host)
 *   max_bounces     : bounce iterations per ray (kernel.py:57), 0 <= max_bounces < 2^23
 *   d_dir_scratch   : [min(chunk, n)*4] float32 workspace (directions of the chunk in flight)
 *   chunk_rays      : rays generated per wave (0 = default 2^24)
 *   d_counters      : [RFRT_CTR_COUNT] uint64, zeroed by the CALLER before the first call of a job;
 *                     accumulates across calls
 *   d_candidates    : [cand_capacity*4] uint32 quads (ray id, receiver, bounce, 0) or NULL when rxset == 0
 *   d_hit_tri/d_hit_t : optional dense [n*max_bounces] parity dumps (int32 triangle index, -1 = miss /
 *                     dead; float32 hit distance, 0 = miss / dead); NULL to skip
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_trace(rfrt_handle env_mesh, rfrt_handle rxset, const float *h_tx_pos, int32_t max_bounces,
               int64_t ray_begin, int64_t ray_end, uint32_t flags, float *d_dir_scratch,
               int64_t chunk_rays, uint64_t *d_counters, uint32_t *d_candidates, int64_t cand_capacity,
               int32_t *d_hit_tri, float *d_hit_t, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Literal replay of kernel.py:38-98 for each candidate (ray id, receiver k) against (env, receiver k),
 * followed by the per-path post-processing of tracer.py:90-117 (NaN-prefix strip is implicit: the
 * record carries n_vertices) — Fresnel amplitude product, fp64 distance, delay bin.
 *   d_candidates / d_cand_count : pairs from rfrt_trace and the device counter holding their number
 *                                 (= d_counters + RFRT_CTR_CANDIDATES); at most cand_capacity are read
 *   amp0            : tx_power / tx_num_rays          (tracer.py:103)
 *   samples_per_m   : used as bin = int((distance / light_speed) * sample_rate)  (tracer.py:115)
 * Outputs, one record per RECEIVED (ray, receiver) pair, appended in arbitrary order
 * (count in d_counters[RFRT_CTR_RECORDS]; records beyond rec_capacity are counted but not stored):
 *   d_rec_ray [cap] uint32, d_rec_rx [cap] int32, d_rec_nverts [cap] int32, d_rec_bin [cap] int64,
 *   d_rec_amp [cap] float64, d_rec_dist [cap] float64,
 *   d_rec_paths [cap*(max_bounces+1)*3] float32 (NaN padded like tracer.py:67-71) or NULL
 * Direct mode (d_ir != NULL: [n_receivers*n_bins] float64, accumulated into): every received pair is binned right
 * away — ir[k][bin] += amplitude if 0 <= bin < n_bins (tracer.py:116-117, fp64 atomics, order-free) — and no record is
 * stored (the d_rec_* arrays may be NULL; RFRT_CTR_RECORDS still counts): the dense coverage maps' path, which needs
 * neither the paths nor a second pass over a record list.
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_trace_receive(rfrt_handle env_mesh, rfrt_handle rxset, const float *h_tx_pos, int32_t max_bounces,
                       const uint32_t *d_candidates, int64_t cand_capacity, uint64_t *d_counters,
                       double amp0, double light_speed_mps, double sample_rate_hz, uint32_t *d_rec_ray,
                       int32_t *d_rec_rx, int32_t *d_rec_nverts, int64_t *d_rec_bin, double *d_rec_amp,
                       double *d_rec_dist, float *d_rec_paths, int64_t rec_capacity, double *d_ir, int64_t n_bins,
                       void *stream);

/* ---------------------------------------------------------------------------------------------
 * Impulse-response binning (tracer.py:101,116-117):  ir[rx][bin] += amp  if bin < n_bins.
 *   d_ir : [n_receivers*n_bins] float64, zeroed by the caller.
 *   d_n_records : optional device counter (e.g. d_counters + RFRT_CTR_RECORDS); when non-NULL only
 *                 min(*d_n_records, n_records) records are read, so no host round trip is needed.
 *   deterministic != 0 : records must be sorted by (rx, ray id); each receiver's records are then
 *                        summed by one thread in ray-id order — the reference's own order — so the
 *                        result is bit-reproducible and independent of the GPU count.
 *   deterministic == 0 : shared-memory privatised histogram per receiver tile, flushed with atomics.
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_bin_ir(const int32_t *d_rec_rx, const int64_t *d_rec_bin, const double *d_rec_amp, int64_t n_records,
                const uint64_t *d_n_records, int64_t n_receivers, int64_t n_bins, int32_t deterministic,
                double *d_ir, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Received-record pipeline (tracer.py:84-87 compaction order, tracer.py:101-117 binning) — the library's own
 * radix sort and run sums; the Python layer only allocates.
 *
 * SEGMENT = fixed-capacity, self-describing block of received records, the unit of the multi-GPU exchange: every
 * rank packs its records into one segment of the same capacity and the ranks all-gather them (one collective; the
 * counts ride in the headers, so no host round trip is needed).  Layout (sections 16-byte aligned, cap = capacity):
 *   u64 header[16] : [0] records produced (> [1] means overflow), [1] records that fit, [2..11] the job's counter block
 *   u32 ray[cap] | i32 rx[cap] | i32 nverts[cap] | i64 bin[cap] | f64 amp[cap] | f64 dist[cap] | f32 paths[cap*path_floats]
 * rfrt_record_segment_bytes / rfrt_records_workspace_bytes are host-only size queries.
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_record_segment_bytes(int64_t capacity, int32_t path_floats, int64_t *out_bytes);
RFRT_API int rfrt_records_workspace_bytes(int64_t n_slots, int64_t *out_bytes);
/* Packs the records rfrt_trace_receive appended (count = d_counters[RFRT_CTR_RECORDS]) into d_segment.
 *   d_rec_paths may be NULL (then the segment's path section stays unwritten); path_floats = (max_bounces+1)*3 */
RFRT_API int rfrt_records_pack(const uint64_t *d_counters, const uint32_t *d_rec_ray, const int32_t *d_rec_rx,
                      const int32_t *d_rec_nverts, const int64_t *d_rec_bin, const double *d_rec_amp,
                      const double *d_rec_dist, const float *d_rec_paths, int64_t rec_capacity, int32_t path_floats,
                      void *d_segment, int64_t seg_capacity, void *stream);
/* Merges n_segments consecutive segments (stride = rfrt_record_segment_bytes) into records sorted by
 * (receiver, ray id) — per receiver the reference's own order (tracer.py:87,102).
 *   outputs: arrays of n_segments*seg_capacity entries (d_paths may be NULL); the first d_summary[0] are valid
 *   d_summary [16] u64: [0] records stored, [1] segments that overflowed, [2..11] counter block summed over segments,
 *                       [12] max records produced by one segment, [13] max candidates of one segment
 *   d_workspace: rfrt_records_workspace_bytes(n_segments*seg_capacity) bytes */
RFRT_API int rfrt_records_sort(const void *d_segments, int64_t n_segments, int64_t seg_capacity, int32_t path_floats,
                      int64_t n_receivers, uint32_t *d_ray, int32_t *d_rx, int32_t *d_nverts, int64_t *d_bin,
                      double *d_amp, double *d_dist, float *d_paths, uint64_t *d_summary, void *d_workspace,
                      int64_t workspace_bytes, void *stream);
/* Arrival sums of (receiver, ray id)-ordered records: every (receiver, bin) run is added sequentially in ray-id order,
 * exactly like `impulse_response[delay_samples] += amplitude` (tracer.py:116-117) — bit-reproducible, independent of
 * the GPU count, parallel over records.  Records with bin outside [0, n_bins) are skipped (tracer.py:116); runs that
 * sum to exactly 0 are dropped from the arrival list.
 *   n_slots / d_n_records : as rfrt_bin_ir (d_n_records may point at d_summary[0])
 *   d_arr_offsets [n_receivers+1] int64, d_arr_bin [n_slots] int32, d_arr_amp [n_slots] float64: CSR for rfrt_rx_power
 *                 (all three NULL to skip)
 *   d_ir : NULL or [n_receivers*n_bins] float64 zeroed by the caller: ir[rx][bin] = the run's sum
 *   d_workspace: rfrt_records_workspace_bytes(n_slots) bytes */
RFRT_API int rfrt_arrivals_build(const int32_t *d_rec_rx, const int64_t *d_rec_bin, const double *d_rec_amp, int64_t n_slots,
                        const uint64_t *d_n_records, int64_t n_receivers, int64_t n_bins, int64_t *d_arr_offsets,
                        int32_t *d_arr_bin, double *d_arr_amp, double *d_ir, void *d_workspace, int64_t workspace_bytes,
                        void *stream);

/* ---------------------------------------------------------------------------------------------
 * RX power of main.py:39,46-55 / coverage.py:45-55 for each receiver, from its SPARSE arrivals
 * (the non-zero impulse-response bins):  t = linspace(0, window, n_bins); s_tx = sin(2 pi f t);
 * s_rx = convolve(ir, s_tx, "same"); power = mean(s_rx[s_rx != 0]^2)   (NaN if nothing is non-zero).
 *   arrivals are CSR: d_arr_offsets [n_receivers+1] int64, d_arr_bin [nnz] int32 (ascending per
 *   receiver), d_arr_amp [nnz] float64.
 *   d_stx_table : [n_bins] float64 scratch (filled with s_tx by this call)
 *   d_power : [n_receivers] float64 (linear mean-square power; dBm = 10 log10(power / 1e-3))
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_rx_power(const int64_t *d_arr_offsets, const int32_t *d_arr_bin, const double *d_arr_amp,
                  int64_t n_receivers, int64_t n_bins, double sample_window_s, double carrier_hz,
                  double *d_stx_table, double *d_power, void *stream);

/* Same quantity straight from DENSE impulse-response rows d_ir [n_receivers*n_bins] (one CTA per receiver stages
 * the row's non-zero bins in shared memory and evaluates the convolution through prefix sums of phasors:
 * O(L + nnz) per receiver instead of O(L * nnz); agrees with rfrt_rx_power to ~1e-12 relative). */
RFRT_API int rfrt_rx_power_dense(const double *d_ir, int64_t n_receivers, int64_t n_bins, double sample_window_s,
                        double carrier_hz, double *d_power, void *stream);

/* ---------------------------------------------------------------------------------------------
 * "Physical" mode (what the reference's quirks Q1-Q6 approximate; NOT reference behaviour).
 * Same deterministic rays (kernel.py:51-52) and the same exact closest-hit test (kernel.py:82), but the triangle a ray
 * has just left is excluded from its next query, receivers are analytic spheres crossed once per segment, and every
 * arrival adds the field  E = (L lambda / (pi N r^2)) * prod Gamma_i * exp(-j 2 pi L / lambda)  (free-space loss x
 * reception-sphere weight x Fresnel amplitude coefficients x carrier phase; L = unfolded path length to the point of
 * closest approach) to its receiver: received power = P_tx * |d_field[k]|^2.
 *   rxset          : receiver set whose unit shape is the cube (+-1)^3, so that its boxes are centre +- radius
 *                    (the Python layer builds it); 0 = trace without receivers
 *   n_rays_total   : N of the formula (rays of the whole job, all GPUs)
 *   d_materials    : [n_triangles] float32 refractive index per triangle, or NULL (5.0 everywhere = tracer.py:43)
 *   d_field        : [n_receivers*2] float64 (re, im), ACCUMULATED (zeroed by the caller)
 *   d_ir           : optional [n_receivers*n_bins*2] float64 complex impulse response, accumulated; bin =
 *                    int(L / light_speed * sample_rate) (tracer.py:115)
 *   d_counters     : as rfrt_trace (RFRT_CTR_RECORDS counts the arrivals)
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_trace_physical(rfrt_handle env_mesh, rfrt_handle rxset, const float *h_tx_pos, int32_t max_bounces,
                        int64_t ray_begin, int64_t ray_end, int64_t n_rays_total, double carrier_hz,
                        double light_speed_mps, double sample_rate_hz, int64_t n_bins, const float *d_materials,
                        float *d_dir_scratch, int64_t chunk_rays, uint64_t *d_counters, double *d_field, double *d_ir,
                        void *stream);

/* ---------------------------------------------------------------------------------------------
 * Compatibility launch with the reference kernel's exact contract (kernel.py:38-47 as launched at
 * tracer.py:75-79): dense outputs for ray ids [ray_begin, ray_begin+n_rays) against receiver
 * `rx_index` of `rxset`.  The caller pre-fills d_traced_paths / d_received_paths with NaN and zeroes
 * d_row_mask, exactly as tracer.py:67-72 does.
 *   d_traced_paths, d_received_paths : [n_rays*(max_bounces+1)*3] float32   d_row_mask : [n_rays] uint32
 * ------------------------------------------------------------------------------------------- */
RFRT_API int rfrt_trace_paths_compat(rfrt_handle env_mesh, const float *h_tx_pos, rfrt_handle rxset, int64_t rx_index,
                            int32_t max_bounces, int64_t ray_begin, int64_t n_rays, float *d_traced_paths,
                            float *d_received_paths, uint32_t *d_row_mask, void *stream);

/* Single-query probe for tests: closest hit of n rays (d_origins/d_dirs [n*3]) against the mesh BVH.
 *   d_t [n] float32 (max_t where missed), d_face [n] int32 (-1 where missed) */
RFRT_API int rfrt_query_closest(rfrt_handle mesh, const float *d_origins, const float *d_dirs, int64_t n, float max_t,
                       float *d_t, int32_t *d_face, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RFRT_H */
