// rfrt_records.cu — the received-record pipeline behind the C ABI (no torch compute on the product path):
//
//   rfrt_records_pack     job records -> one fixed-capacity, self-describing SEGMENT (the unit of the multi-GPU
//                         exchange: every rank all-gathers one equal-sized segment, counts ride in the header, so
//                         the exchange needs no host round trip)
//   rfrt_records_sort     S segments -> records in (receiver, ray id) order = the reference's own order
//                         (tracer.py:87 keeps ray-id order, tracer.py:102 iterates in it); own LSD radix sort of the
//                         unique keys + a binary-search scatter
//   rfrt_arrivals_build   (receiver, ray id)-ordered records -> per (receiver, bin) arrival sums, each run added
//                         sequentially in ray-id order exactly like `impulse_response[bin] += amp` (tracer.py:116-117):
//                         CSR for rfrt_rx_power and/or the dense impulse responses.  Bit-reproducible and independent
//                         of the GPU count, and parallel over records (the per-receiver serial walk it replaces was
//                         O(records per receiver) deep).
#include <cstdio>

#include "rfrt_internal.h"

namespace rfrt {
namespace {

constexpr int REC_THREADS = 256;
constexpr int SEG_HEADER_U64 = 16;

// byte offsets of the sections of a segment (all 16-byte aligned)
struct SegLayout {
    int64_t ray, rx, nverts, bin, amp, dist, paths, total;
};
__host__ __device__ inline int64_t align16(int64_t x) { return (x + 15) & ~(int64_t)15; }
__host__ __device__ inline SegLayout seg_layout(int64_t cap, int row)
{
    SegLayout L;
    L.ray = 8 * SEG_HEADER_U64;
    L.rx = L.ray + align16(4 * cap);
    L.nverts = L.rx + align16(4 * cap);
    L.bin = L.nverts + align16(4 * cap);
    L.amp = L.bin + align16(8 * cap);
    L.dist = L.amp + align16(8 * cap);
    L.paths = L.dist + align16(8 * cap);
    L.total = L.paths + align16(4 * cap * (int64_t)row);
    return L;
}

struct RecArrays {
    uint32_t *ray;
    int32_t *rx;
    int32_t *nverts;
    int64_t *bin;
    double *amp;
    double *dist;
    float *paths; // may be NULL
};

__global__ void __launch_bounds__(REC_THREADS)
k_rec_pack(const unsigned long long *__restrict__ counters, RecArrays in, int64_t in_cap, int row, unsigned char *seg, int64_t cap)
{
    const SegLayout L = seg_layout(cap, row);
    const unsigned long long produced = counters[RFRT_CTR_RECORDS];
    int64_t n = (int64_t)produced < in_cap ? (int64_t)produced : in_cap;
    if (n > cap) n = cap;
    if (blockIdx.x == 0 && threadIdx.x < SEG_HEADER_U64) {
        unsigned long long *h = reinterpret_cast<unsigned long long *>(seg);
        unsigned long long v = 0ull;
        if (threadIdx.x == 0) v = produced;                 // true count: > cap (or > in_cap) means overflow
        else if (threadIdx.x == 1) v = (unsigned long long)(in_cap < cap ? in_cap : cap);
        else if (threadIdx.x < 2 + RFRT_CTR_COUNT) v = counters[threadIdx.x - 2];
        h[threadIdx.x] = v;
    }
    uint32_t *o_ray = reinterpret_cast<uint32_t *>(seg + L.ray);
    int32_t *o_rx = reinterpret_cast<int32_t *>(seg + L.rx), *o_nv = reinterpret_cast<int32_t *>(seg + L.nverts);
    int64_t *o_bin = reinterpret_cast<int64_t *>(seg + L.bin);
    double *o_amp = reinterpret_cast<double *>(seg + L.amp), *o_dist = reinterpret_cast<double *>(seg + L.dist);
    float *o_paths = reinterpret_cast<float *>(seg + L.paths);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        o_ray[i] = in.ray[i]; o_rx[i] = in.rx[i]; o_nv[i] = in.nverts[i];
        o_bin[i] = in.bin[i]; o_amp[i] = in.amp[i]; o_dist[i] = in.dist[i];
    }
    if (row > 0 && in.paths) {
        const int64_t m = n * row;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
            o_paths[i] = in.paths[i];
    }
}

__device__ __forceinline__ int64_t seg_count(const unsigned char *seg)
{
    const unsigned long long *h = reinterpret_cast<const unsigned long long *>(seg);
    return (int64_t)(h[0] < h[1] ? h[0] : h[1]);
}

// key of slot i = s * cap + j: (receiver << 32 | ray id) for stored records, `invalid` (larger than every valid key)
// for the empty tail of a segment.  Block 0 also reduces the headers into the summary.
__global__ void __launch_bounds__(REC_THREADS)
k_rec_keys(const unsigned char *__restrict__ segs, int64_t n_seg, int64_t seg_bytes, int64_t cap, int row, uint64_t invalid,
           uint64_t *__restrict__ keys, unsigned long long *__restrict__ summary)
{
    const SegLayout L = seg_layout(cap, row);
    const int64_t n = n_seg * cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / cap, j = i - s * cap;
        const unsigned char *seg = segs + s * seg_bytes;
        uint64_t key = invalid;
        if (j < seg_count(seg)) {
            const uint32_t ray = reinterpret_cast<const uint32_t *>(seg + L.ray)[j];
            const int32_t rx = reinterpret_cast<const int32_t *>(seg + L.rx)[j];
            key = ((uint64_t)(uint32_t)rx << 32) | (uint64_t)ray;
            if (key > invalid) key = invalid; // (receiver index out of range: cannot happen for library-produced records)
        }
        keys[i] = key;
    }
    if (blockIdx.x == 0 && threadIdx.x < SEG_HEADER_U64) {
        // [0] records stored, [1] segments that overflowed, [2 .. 2+CTR_COUNT) counters summed over segments,
        // [12] / [13] largest record / candidate count of one segment (the capacities every rank needs)
        unsigned long long v = 0ull;
        for (int64_t s = 0; s < n_seg; ++s) {
            const unsigned long long *h = reinterpret_cast<const unsigned long long *>(segs + s * seg_bytes);
            if (threadIdx.x == 0) v += h[0] < h[1] ? h[0] : h[1];
            else if (threadIdx.x == 1) v += h[0] > h[1] ? 1ull : 0ull;
            else if (threadIdx.x < 2 + RFRT_CTR_COUNT) v += h[threadIdx.x];
            else if (threadIdx.x == 2 + RFRT_CTR_COUNT) v = h[0] > v ? h[0] : v;
            else if (threadIdx.x == 3 + RFRT_CTR_COUNT) v = h[2 + RFRT_CTR_CANDIDATES] > v ? h[2 + RFRT_CTR_CANDIDATES] : v;
        }
        summary[threadIdx.x] = v;
    }
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t *__restrict__ a, int64_t n, uint64_t key)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// every stored record finds its rank among the sorted keys (keys are unique: one record per (ray, receiver)) and is
// written there
__global__ void __launch_bounds__(REC_THREADS)
k_rec_scatter(const unsigned char *__restrict__ segs, int64_t n_seg, int64_t seg_bytes, int64_t cap, int row,
              const uint64_t *__restrict__ sorted, RecArrays out)
{
    const SegLayout L = seg_layout(cap, row);
    const int64_t n = n_seg * cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / cap, j = i - s * cap;
        const unsigned char *seg = segs + s * seg_bytes;
        if (j >= seg_count(seg)) continue;
        const uint32_t ray = reinterpret_cast<const uint32_t *>(seg + L.ray)[j];
        const int32_t rx = reinterpret_cast<const int32_t *>(seg + L.rx)[j];
        const int64_t pos = lower_bound_u64(sorted, n, ((uint64_t)(uint32_t)rx << 32) | (uint64_t)ray);
        out.ray[pos] = ray; out.rx[pos] = rx;
        out.nverts[pos] = reinterpret_cast<const int32_t *>(seg + L.nverts)[j];
        out.bin[pos] = reinterpret_cast<const int64_t *>(seg + L.bin)[j];
        out.amp[pos] = reinterpret_cast<const double *>(seg + L.amp)[j];
        out.dist[pos] = reinterpret_cast<const double *>(seg + L.dist)[j];
        if (row > 0 && out.paths) {
            const float *src = reinterpret_cast<const float *>(seg + L.paths) + j * row;
            float *dst = out.paths + pos * row;
            for (int k = 0; k < row; ++k) dst[k] = src[k];
        }
    }
}

// ---- arrivals ------------------------------------------------------------------------------------------------------
// key = ((receiver * n_bins + bin) << idx_bits) | record index for records with 0 <= bin < n_bins, all ones otherwise
__global__ void __launch_bounds__(REC_THREADS)
k_arr_keys(const int32_t *__restrict__ rx, const int64_t *__restrict__ bin, int64_t n_slots, const unsigned long long *d_n,
           int64_t n_rx, int64_t n_bins, int idx_bits, uint64_t *__restrict__ keys)
{
    int64_t n = n_slots;
    if (d_n && (int64_t)*d_n < n) n = (int64_t)*d_n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_slots; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t key = ~0ull;
        if (i < n) {
            const int64_t b = bin[i], k = rx[i];
            if (b >= 0 && b < n_bins && k >= 0 && k < n_rx) key = ((uint64_t)(k * n_bins + b) << idx_bits) | (uint64_t)i;
        }
        keys[i] = key;
    }
}

// amplitudes in sorted order (contiguous reads for the run sums below)
__global__ void __launch_bounds__(REC_THREADS)
k_arr_gather_amp(const uint64_t *__restrict__ keys, int64_t n, int idx_bits, const double *__restrict__ amp, double *__restrict__ samp)
{
    const uint64_t mask = (1ull << idx_bits) - 1ull;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[p];
        samp[p] = key == ~0ull ? 0.0 : amp[key & mask];
    }
}

// the head of every run of equal (receiver, bin) adds the run's amplitudes in order — ((0 + a1) + a2) + ... — which is
// `impulse_response[bin] += amp` over ascending ray ids (tracer.py:102,116-117); flag = the sum is non-zero (np.convolve
// and np.nonzero see the summed impulse response: exact zeros contribute nothing)
__global__ void __launch_bounds__(REC_THREADS)
k_arr_heads(const uint64_t *__restrict__ keys, int64_t n, int idx_bits, double *samp, uint32_t *__restrict__ flags,
            uint32_t *__restrict__ block_counts)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool flag = false;
    if (p < n) {
        const uint64_t key = keys[p];
        if (key != ~0ull) {
            const uint64_t g = key >> idx_bits;
            const bool head = p == 0 || (keys[p - 1] >> idx_bits) != g;
            if (head) {
                double sum = __dadd_rn(0.0, samp[p]);
                for (int64_t q = p + 1; q < n && (keys[q] >> idx_bits) == g; ++q) sum = __dadd_rn(sum, samp[q]);
                samp[p] = sum;
                flag = sum != 0.0;
            }
        }
        flags[p] = flag ? 1u : 0u;
    }
    const int cnt = __syncthreads_count(flag);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)cnt;
}

// exclusive scan of the block counts by one CTA; total -> block_counts[n_blocks]
__global__ void __launch_bounds__(1024) k_arr_scan_blocks(uint32_t *block_counts, int64_t n_blocks)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_counts[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        uint32_t prefix = 0u, total = 0u;
        for (int w = 0; w < 32; ++w) {
            const uint32_t s = warp_sums[w];
            if (w < warp) prefix += s;
            total += s;
        }
        const uint32_t carry = carry_s;
        if (i < n_blocks) block_counts[i] = carry + prefix + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_counts[n_blocks] = carry_s;
}

// compaction of the flagged run heads, in order: arrival list (group key, bin, amplitude) and the dense rows
__global__ void __launch_bounds__(REC_THREADS)
k_arr_compact(const uint64_t *__restrict__ keys, int64_t n, int idx_bits, const double *__restrict__ samp,
              const uint32_t *__restrict__ flags, const uint32_t *__restrict__ block_offsets, int64_t n_bins,
              uint64_t *__restrict__ arr_group, int32_t *__restrict__ arr_bin, double *__restrict__ arr_amp, double *__restrict__ ir)
{
    __shared__ uint32_t warp_cnt[REC_THREADS / 32];
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool flag = p < n && flags[p] != 0u;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_cnt[warp] = (uint32_t)__popc(m);
    __syncthreads();
    if (!flag) return;
    uint32_t off = block_offsets[blockIdx.x];
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    off += (uint32_t)__popc(m & ((1u << lane) - 1u));
    const uint64_t g = keys[p] >> idx_bits;
    const double a = samp[p];
    arr_group[off] = g;
    if (arr_bin) arr_bin[off] = (int32_t)(g % (uint64_t)n_bins);
    if (arr_amp) arr_amp[off] = a;
    if (ir) ir[g] = a;
}

// CSR offsets: offsets[k] = first arrival of receiver k (k = n_rx: the arrival count)
__global__ void __launch_bounds__(REC_THREADS)
k_arr_offsets(const uint64_t *__restrict__ arr_group, const uint32_t *__restrict__ total, int64_t n_rx, int64_t n_bins,
              int64_t *__restrict__ offsets)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > n_rx) return;
    const int64_t n = (int64_t)*total;
    offsets[k] = k == n_rx ? n : lower_bound_u64(arr_group, n, (uint64_t)(k * n_bins));
}

inline int bits_for(uint64_t v) // number of bits needed to represent v (0 -> 0)
{
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

inline unsigned grid_1d(int64_t n)
{
    int64_t nb = (n + REC_THREADS - 1) / REC_THREADS;
    if (nb < 1) nb = 1;
    if (nb > 148 * 16) nb = 148 * 16;
    return (unsigned)nb;
}

// workspace carving (16-byte aligned sections)
struct Carver {
    unsigned char *p;
    int64_t left;
    template <class T> T *take(int64_t count)
    {
        const int64_t bytes = align16((int64_t)sizeof(T) * count);
        if (bytes > left) return nullptr;
        T *r = reinterpret_cast<T *>(p);
        p += bytes; left -= bytes;
        return r;
    }
};

int64_t workspace_need(int64_t n)
{
    const int64_t blocks = sort_hist_blocks(n) + 1;
    return 4 * align16(8 * n) + align16(4 * n) + align16(4 * 256 * (blocks + 1)) + align16(4 * ((n + REC_THREADS - 1) / REC_THREADS + 2)) + 256;
}

} // namespace
} // namespace rfrt

using namespace rfrt;

extern "C" int rfrt_record_segment_bytes(int64_t capacity, int32_t path_floats, int64_t *out_bytes)
{
    if (capacity < 0 || path_floats < 0 || !out_bytes) { set_error("rfrt_record_segment_bytes: bad arguments"); return RFRT_ERR_INVALID; }
    *out_bytes = seg_layout(capacity, path_floats).total;
    return RFRT_OK;
}

extern "C" int rfrt_records_workspace_bytes(int64_t n_slots, int64_t *out_bytes)
{
    if (n_slots < 0 || !out_bytes) { set_error("rfrt_records_workspace_bytes: bad arguments"); return RFRT_ERR_INVALID; }
    *out_bytes = workspace_need(n_slots > 0 ? n_slots : 1);
    return RFRT_OK;
}

extern "C" int rfrt_records_pack(const uint64_t *d_counters, const uint32_t *d_rec_ray, const int32_t *d_rec_rx,
                                 const int32_t *d_rec_nverts, const int64_t *d_rec_bin, const double *d_rec_amp,
                                 const double *d_rec_dist, const float *d_rec_paths, int64_t rec_capacity,
                                 int32_t path_floats, void *d_segment, int64_t seg_capacity, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_counters || !d_rec_ray || !d_rec_rx || !d_rec_nverts || !d_rec_bin || !d_rec_amp || !d_rec_dist || !d_segment ||
        rec_capacity <= 0 || seg_capacity <= 0 || path_floats < 0) {
        set_error("rfrt_records_pack: null buffer or empty capacity");
        return RFRT_ERR_INVALID;
    }
    RecArrays in{const_cast<uint32_t *>(d_rec_ray), const_cast<int32_t *>(d_rec_rx), const_cast<int32_t *>(d_rec_nverts),
                 const_cast<int64_t *>(d_rec_bin), const_cast<double *>(d_rec_amp), const_cast<double *>(d_rec_dist),
                 const_cast<float *>(d_rec_paths)};
    const int64_t m = rec_capacity < seg_capacity ? rec_capacity : seg_capacity;
    k_rec_pack<<<grid_1d(m), REC_THREADS, 0, stream>>>((const unsigned long long *)d_counters, in, rec_capacity,
                                                       d_rec_paths ? path_floats : 0, (unsigned char *)d_segment, seg_capacity);
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}

extern "C" int rfrt_records_sort(const void *d_segments, int64_t n_segments, int64_t seg_capacity, int32_t path_floats,
                                 int64_t n_receivers, uint32_t *d_ray, int32_t *d_rx, int32_t *d_nverts, int64_t *d_bin,
                                 double *d_amp, double *d_dist, float *d_paths, uint64_t *d_summary, void *d_workspace,
                                 int64_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_segments || n_segments <= 0 || seg_capacity <= 0 || path_floats < 0 || n_receivers <= 0 || !d_ray || !d_rx ||
        !d_nverts || !d_bin || !d_amp || !d_dist || !d_summary || !d_workspace) {
        set_error("rfrt_records_sort: bad arguments");
        return RFRT_ERR_INVALID;
    }
    const int64_t n = n_segments * seg_capacity;
    if (n >= (1ll << 31)) { set_error("rfrt_records_sort: more than 2^31-1 record slots"); return RFRT_ERR_INVALID; }
    if (workspace_bytes < workspace_need(n)) { set_error("rfrt_records_sort: workspace too small (rfrt_records_workspace_bytes)"); return RFRT_ERR_INVALID; }
    Carver ws{(unsigned char *)d_workspace, workspace_bytes};
    uint64_t *ka = ws.take<uint64_t>(n), *kb = ws.take<uint64_t>(n);
    uint32_t *hist = ws.take<uint32_t>(256 * (sort_hist_blocks(n) + 2));
    if (!ka || !kb || !hist) { set_error("rfrt_records_sort: workspace too small"); return RFRT_ERR_INVALID; }
    const int row = d_paths ? path_floats : 0;
    const int64_t seg_bytes = seg_layout(seg_capacity, path_floats).total;
    // sorted bits: 32 (ray id) + enough to hold the value n_receivers itself, which marks the empty slots
    const uint64_t invalid = ((uint64_t)n_receivers << 32) | 0xffffffffull;
    const int passes = (32 + bits_for((uint64_t)n_receivers) + 7) / 8;
    k_rec_keys<<<grid_1d(n), REC_THREADS, 0, stream>>>((const unsigned char *)d_segments, n_segments, seg_bytes, seg_capacity,
                                                       path_floats, invalid, ka, (unsigned long long *)d_summary);
    const uint64_t *sorted = radix_sort_u64(ka, kb, hist, n, 0, passes, stream);
    RecArrays out{d_ray, d_rx, d_nverts, d_bin, d_amp, d_dist, row ? d_paths : nullptr};
    k_rec_scatter<<<grid_1d(n), REC_THREADS, 0, stream>>>((const unsigned char *)d_segments, n_segments, seg_bytes, seg_capacity,
                                                          path_floats, sorted, out);
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}

extern "C" int rfrt_arrivals_build(const int32_t *d_rec_rx, const int64_t *d_rec_bin, const double *d_rec_amp, int64_t n_slots,
                                   const uint64_t *d_n_records, int64_t n_receivers, int64_t n_bins, int64_t *d_arr_offsets,
                                   int32_t *d_arr_bin, double *d_arr_amp, double *d_ir, void *d_workspace,
                                   int64_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_slots < 0 || n_receivers <= 0 || n_bins <= 0 || !d_workspace || (n_slots > 0 && (!d_rec_rx || !d_rec_bin || !d_rec_amp)) ||
        (!d_arr_offsets && !d_ir) || (d_arr_offsets && (!d_arr_bin || !d_arr_amp))) {
        set_error("rfrt_arrivals_build: bad arguments");
        return RFRT_ERR_INVALID;
    }
    if (n_slots >= (1ll << 31) || n_bins >= (1ll << 31)) { set_error("rfrt_arrivals_build: more than 2^31-1 slots or bins"); return RFRT_ERR_INVALID; }
    const int64_t n = n_slots > 0 ? n_slots : 1;
    const int idx_bits = bits_for((uint64_t)(n - 1)) > 0 ? bits_for((uint64_t)(n - 1)) : 1;
    const int group_bits = bits_for((uint64_t)n_receivers * (uint64_t)n_bins - 1ull);
    if (group_bits + idx_bits > 63) {
        set_error("rfrt_arrivals_build: receivers x bins x records does not fit a 63-bit sort key (use the order-free rfrt_bin_ir)");
        return RFRT_ERR_INVALID;
    }
    if (workspace_bytes < workspace_need(n)) { set_error("rfrt_arrivals_build: workspace too small (rfrt_records_workspace_bytes)"); return RFRT_ERR_INVALID; }
    Carver ws{(unsigned char *)d_workspace, workspace_bytes};
    const int64_t n_blocks = (n + REC_THREADS - 1) / REC_THREADS;
    uint64_t *ka = ws.take<uint64_t>(n), *kb = ws.take<uint64_t>(n);
    double *samp = ws.take<double>(n);
    uint64_t *arr_group = ws.take<uint64_t>(n);
    uint32_t *flags = ws.take<uint32_t>(n);
    uint32_t *hist = ws.take<uint32_t>(256 * (sort_hist_blocks(n) + 2));
    uint32_t *block_counts = ws.take<uint32_t>(n_blocks + 2);
    if (!ka || !kb || !samp || !arr_group || !flags || !hist || !block_counts) { set_error("rfrt_arrivals_build: workspace too small"); return RFRT_ERR_INVALID; }
    if (n_slots == 0) {
        if (d_arr_offsets) RFRT_CUDA(cudaMemsetAsync(d_arr_offsets, 0, sizeof(int64_t) * (n_receivers + 1), stream));
        return RFRT_OK;
    }
    k_arr_keys<<<grid_1d(n), REC_THREADS, 0, stream>>>(d_rec_rx, d_rec_bin, n, (const unsigned long long *)d_n_records, n_receivers,
                                                       n_bins, idx_bits, ka);
    // stable LSD passes over the group bits only: inside a group the record (= ray id) order survives
    const uint64_t *sorted = radix_sort_u64(ka, kb, hist, n, idx_bits, (group_bits + 7) / 8 > 0 ? (group_bits + 7) / 8 : 1, stream);
    k_arr_gather_amp<<<grid_1d(n), REC_THREADS, 0, stream>>>(sorted, n, idx_bits, d_rec_amp, samp);
    k_arr_heads<<<(unsigned)n_blocks, REC_THREADS, 0, stream>>>(sorted, n, idx_bits, samp, flags, block_counts);
    k_arr_scan_blocks<<<1, 1024, 0, stream>>>(block_counts, n_blocks);
    k_arr_compact<<<(unsigned)n_blocks, REC_THREADS, 0, stream>>>(sorted, n, idx_bits, samp, flags, block_counts, n_bins, arr_group,
                                                                 d_arr_offsets ? d_arr_bin : nullptr, d_arr_offsets ? d_arr_amp : nullptr, d_ir);
    if (d_arr_offsets)
        k_arr_offsets<<<(unsigned)((n_receivers + 1 + REC_THREADS - 1) / REC_THREADS), REC_THREADS, 0, stream>>>(
            arr_group, block_counts + n_blocks, n_receivers, n_bins, d_arr_offsets);
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}
