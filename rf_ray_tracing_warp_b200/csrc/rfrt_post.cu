// rfrt_post.cu — impulse-response binning and receiver power (C ABI).
//
//   rfrt_bin_ir    tracer.py:101,116-117   ir[rx][bin] += amplitude  if bin < n_bins
//   rfrt_rx_power  main.py:39,46-55 / coverage.py:45-55, evaluated from the sparse arrivals instead of a
//                  dense O(L^2) np.convolve
#include <cmath>

#include "rfrt_internal.h"

namespace rfrt {
namespace {

// order-free: one fp64 atomic per record
__global__ void k_bin_atomic(const int32_t *__restrict__ rx, const int64_t *__restrict__ bin,
                             const double *__restrict__ amp, int64_t n, const unsigned long long *d_n, int64_t n_rx,
                             int64_t n_bins, double *ir)
{
    if (d_n && (int64_t)*d_n < n) n = (int64_t)*d_n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = bin[i];
        int64_t k = rx[i];
        if (b >= 0 && b < n_bins && k >= 0 && k < n_rx) atomicAdd(&ir[k * n_bins + b], amp[i]);
    }
}

// Privatised variant for few receivers / many records (one receiver's histogram fits in shared memory): each CTA owns
// a contiguous slice of the records and, receiver by receiver, accumulates them into its shared-memory window, then
// flushes the window's non-zero bins with one global atomic each.  The shared-memory adds are WARP-AGGREGATED: lanes
// whose records fall into the same bin (__match_any_sync) are summed by their lowest lane in lane order and that lane
// issues the one atomic — fp64 shared atomics are compare-and-swap loops, and the arrivals of a receiver cluster in a
// few bins (every line-of-sight ray of a receiver has the same delay), so same-address conflicts are the common case.
__global__ void __launch_bounds__(512)
k_bin_privatised(const int32_t *__restrict__ rx, const int64_t *__restrict__ bin, const double *__restrict__ amp, int64_t n,
                 const unsigned long long *d_n, int64_t n_rx, int64_t n_bins, double *ir, int64_t per_block)
{
    extern __shared__ double s_hist[];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if (d_n && (int64_t)*d_n < n) n = (int64_t)*d_n;
    const int64_t begin = blockIdx.x * per_block;
    const int64_t end = begin + per_block < n ? begin + per_block : n;
    if (begin >= end) return;
    for (int64_t k = 0; k < n_rx; ++k) {
        for (int64_t b = threadIdx.x; b < n_bins; b += blockDim.x) s_hist[b] = 0.0;
        __syncthreads();
        // warp-uniform trip count: every lane takes part in the votes of every trip
        for (int64_t base = begin + (threadIdx.x & ~31); base < end; base += blockDim.x) {
            const int64_t i = base + lane;
            int64_t b = -1;
            double a = 0.0;
            if (i < end && rx[i] == k) { b = bin[i]; a = amp[i]; }
            const bool valid = b >= 0 && b < n_bins;
            const int tag = valid ? (int)b : -1 - lane; // lanes without a record match nobody
            const unsigned peers = __match_any_sync(FULL, tag);
            const int iters = __reduce_max_sync(FULL, (unsigned)__popc(peers));
            unsigned rest = peers;
            double sum = 0.0;
            for (int it = 0; it < iters; ++it) { // lane order: the group's sum does not depend on the schedule
                const int src = rest ? __ffs((int)rest) - 1 : lane;
                const double v = __shfl_sync(FULL, a, src);
                if (rest) { sum += v; rest &= rest - 1u; }
            }
            if (valid && lane == __ffs((int)peers) - 1) atomicAdd(&s_hist[b], sum);
        }
        __syncthreads();
        for (int64_t b = threadIdx.x; b < n_bins; b += blockDim.x) {
            const double v = s_hist[b];
            if (v != 0.0) atomicAdd(&ir[k * n_bins + b], v);
        }
        __syncthreads();
    }
}

// deterministic: records sorted by (rx, ray id); receiver k's records are added by one thread in order
__global__ void k_bin_ordered(const int32_t *__restrict__ rx, const int64_t *__restrict__ bin,
                              const double *__restrict__ amp, int64_t n, const unsigned long long *d_n, int64_t n_rx,
                              int64_t n_bins, double *ir)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_rx) return;
    if (d_n && (int64_t)*d_n < n) n = (int64_t)*d_n;
    int64_t lo = 0, hi = n; // lower_bound of k
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rx[mid] < k) lo = mid + 1; else hi = mid;
    }
    for (int64_t i = lo; i < n && rx[i] == k; ++i) {
        int64_t b = bin[i];
        if (b >= 0 && b < n_bins) ir[k * n_bins + b] = __dadd_rn(ir[k * n_bins + b], amp[i]);
    }
}

// same result with gridDim.y CTAs per receiver: CTA s owns the 256-bin blocks j with j % gridDim.y == s and its thread t
// the bins b of those blocks with b % 256 == t; the receiver's records are staged tile by tile in shared memory
// (coalesced) and every thread walks the tile in order, so every bin still sees its additions in ray-id order (few
// receivers, many records each).  (rfrt_arrivals_build is the scalable ordered path; this one needs no workspace.)
__global__ void __launch_bounds__(256) k_bin_ordered_cta(const int32_t *__restrict__ rx, const int64_t *__restrict__ bin,
                                                         const double *__restrict__ amp, int64_t n,
                                                         const unsigned long long *d_n, int64_t n_bins, double *ir)
{
    __shared__ int64_t s_bin[256];
    __shared__ double s_amp[256];
    const int64_t k = blockIdx.x;
    if (d_n && (int64_t)*d_n < n) n = (int64_t)*d_n;
    int64_t lo = 0, hi = n; // [lo, hi) = the records of receiver k (records are sorted by receiver)
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rx[mid] < k) lo = mid + 1; else hi = mid;
    }
    int64_t end = lo, top = n;
    while (end < top) {
        int64_t mid = (end + top) >> 1;
        if (rx[mid] <= k) end = mid + 1; else top = mid;
    }
    double *row = ir + k * n_bins;
    for (int64_t base = lo; base < end; base += 256) {
        const int64_t i = base + threadIdx.x;
        s_bin[threadIdx.x] = i < end ? bin[i] : -1;
        s_amp[threadIdx.x] = i < end ? amp[i] : 0.0;
        __syncthreads();
        const int cnt = (int)(end - base < 256 ? end - base : 256);
        for (int j = 0; j < cnt; ++j) {
            const int64_t b = s_bin[j];
            if (b >= 0 && b < n_bins && (int)(b & 255) == (int)threadIdx.x && (int)((b >> 8) % gridDim.y) == (int)blockIdx.y)
                row[b] = __dadd_rn(row[b], s_amp[j]);
        }
        __syncthreads();
    }
}

// s_tx[m] = sin((2*pi*f) * t_m),  t = np.linspace(0, window, n_bins)
__global__ void k_stx_table(int64_t n_bins, double window, double carrier, double *table)
{
    int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= n_bins) return;
    const double K = __dmul_rn(__dmul_rn(2.0, 3.141592653589793), carrier);
    double t;
    if (n_bins == 1) t = 0.0;
    else if (m == n_bins - 1) t = window;
    else t = __dmul_rn((double)m, __ddiv_rn(window, (double)(n_bins - 1)));
    table[m] = sin(__dmul_rn(K, t));
}

// one CTA per receiver: s_rx[n] = sum_j a_j * s_tx[n + (L-1)/2 - b_j]  (np.convolve(..., "same"));
// power = mean of s_rx^2 over the samples that are != 0  (np.nonzero selection)
__global__ void __launch_bounds__(256)
k_rx_power(const int64_t *__restrict__ offsets, const int32_t *__restrict__ abin, const double *__restrict__ aamp,
           int64_t n_bins, const double *__restrict__ stx, double *__restrict__ power)
{
    const int64_t k = blockIdx.x;
    const int64_t j0 = offsets[k], j1 = offsets[k + 1];
    const int64_t half = (n_bins - 1) / 2;
    double sum = 0.0;
    unsigned long long cnt = 0;
    for (int64_t n = threadIdx.x; n < n_bins; n += blockDim.x) {
        double s = 0.0;
        for (int64_t j = j0; j < j1; ++j) {
            int64_t m = n + half - (int64_t)abin[j];
            if (m >= 0 && m < n_bins) s = __dadd_rn(s, __dmul_rn(aamp[j], __ldg(stx + m)));
        }
        if (s != 0.0) { sum += s * s; ++cnt; }
    }
    __shared__ double s_sum[256];
    __shared__ unsigned long long s_cnt[256];
    s_sum[threadIdx.x] = sum;
    s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
            s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) power[k] = s_cnt[0] ? s_sum[0] / (double)s_cnt[0] : nan("");
}


// ---- receiver power straight from dense impulse-response rows, O(L + nnz) per receiver -------------------------
// s_rx[n] = sum_j a_j sin(arg(q - b_j)),  q = n + (L-1)/2, over the arrivals with 0 <= q - b_j <= L-1.  With
// arg(m) = K*(m*step) the sum is  Im( e^{i arg(q)} * sum_j a_j e^{-i arg(b_j)} ), and because arrivals are sorted by
// bin the valid j form a contiguous range -> a prefix-sum difference.  Agrees with the direct sum to ~1e-13
// relative; a sample counts as non-zero exactly when the direct sum has a non-zero term (np.nonzero, main.py:48).
constexpr int RXP_THREADS = 256;
// arrivals staged in shared memory per receiver: a first launch with the small capacity (20 KB -> ~10 CTAs per SM; one
// 123 KB CTA per SM left the row reads latency-bound: 25 ms for 65 536 x 10 000 bins) marks the rare denser rows with
// power = -1, a second launch with the large capacity redoes just those
constexpr int RXP_CAP_SMALL = 1024;
constexpr int RXP_CAP_BIG = 6144;

__device__ __forceinline__ double stx_arg(int64_t m, int64_t n_bins, double window, double K)
{
    double t;
    if (n_bins == 1) t = 0.0;
    else if (m == n_bins - 1) t = window;
    else t = __dmul_rn((double)m, __ddiv_rn(window, (double)(n_bins - 1)));
    return __dmul_rn(K, t);
}

template <int RXP_CAP, bool SECOND>
__global__ void __launch_bounds__(RXP_THREADS)
k_rx_power_dense(const double *__restrict__ ir, int64_t n_bins, double window, double carrier, double *__restrict__ power)
{
    if (SECOND && power[blockIdx.x] != -1.0) return; // only the rows the first launch could not stage
    extern __shared__ double s_dyn[];
    double *s_re = s_dyn;                       // [RXP_CAP+1]  after the scan: exclusive prefix sums of the phasors
    double *s_im = s_dyn + (RXP_CAP + 1);       // [RXP_CAP+1]
    int *s_bin = reinterpret_cast<int *>(s_dyn + 2 * (RXP_CAP + 1)); // [RXP_CAP]
    __shared__ int s_warp_cnt[RXP_THREADS / 32];
    __shared__ int s_total;
    __shared__ double s_sum[RXP_THREADS];
    __shared__ unsigned long long s_cnt[RXP_THREADS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *row = ir + (int64_t)blockIdx.x * n_bins;
    const double K = __dmul_rn(__dmul_rn(2.0, 3.141592653589793), carrier);
    const int64_t half = (n_bins - 1) / 2;

    // 1. ordered compaction of the non-zero bins.  The row is taken in blocks of 8 x 2048 bins: every warp owns a
    //    contiguous 2048-bin slice of the block, reads it with 64 independent loads per lane and remembers which of its
    //    bins are non-zero in a 64-bit mask (pass a: no barrier, the loads of all warps are in flight together — one
    //    barrier-separated 256-bin step at a time left this kernel latency-bound at 13 % of the HBM peak); the slices'
    //    counts give every warp its offset (one barrier), and only the non-zero bins are read again, from cache, to
    //    be staged in bin order (pass b).
    if (tid == 0) s_total = 0;
    __syncthreads();
    constexpr int SLICE = 2048;
    for (int64_t blk = 0; blk < n_bins; blk += (int64_t)SLICE * (RXP_THREADS / 32)) {
        const int64_t slice = blk + (int64_t)warp * SLICE;
        unsigned long long nzmask = 0ull;
#pragma unroll 8
        for (int it = 0; it < SLICE / 32; ++it) {
            const int64_t b = slice + it * 32 + lane;
            const double a = b < n_bins ? row[b] : 0.0;
            if (a != 0.0) nzmask |= 1ull << it;
        }
        int cnt = __popcll(nzmask);
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) s_warp_cnt[warp] = cnt;
        __syncthreads();
        int off = s_total;
        for (int w = 0; w < warp; ++w) off += s_warp_cnt[w];
        if (cnt > 0) {
            for (int it = 0; it < SLICE / 32; ++it) {
                const bool nz = (nzmask >> it) & 1ull;
                const unsigned m = __ballot_sync(0xffffffffu, nz);
                if (m == 0u) continue;
                const int pos = off + __popc(m & ((1u << lane) - 1u));
                if (nz && pos < RXP_CAP) {
                    const int64_t b = slice + it * 32 + lane;
                    const double a = row[b];
                    double sn, cs;
                    sincos(stx_arg(b, n_bins, window, K), &sn, &cs);
                    s_bin[pos] = (int)b;
                    s_re[pos] = a * cs;   // a * e^{-i arg(b)}
                    s_im[pos] = -a * sn;
                }
                off += __popc(m);
            }
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < RXP_THREADS / 32; ++w) t += s_warp_cnt[w];
            s_total += t;
        }
        __syncthreads();
    }
    const int nnz = s_total;
    double sum = 0.0;
    unsigned long long cnt = 0;
    if (nnz <= RXP_CAP) {
        // 2. exclusive prefix sums (sequential: nnz is small and the order is fixed -> reproducible)
        if (tid == 0) {
            double ar = 0.0, ai = 0.0;
            for (int j = 0; j < nnz; ++j) {
                double r = s_re[j], i = s_im[j];
                s_re[j] = ar; s_im[j] = ai;
                ar += r; ai += i;
            }
            s_re[nnz] = ar; s_im[nnz] = ai;
        }
        __syncthreads();
        // 3. samples, in closed form per RUN.  Sample q = n + half sees the arrivals with q - (L-1) <= b_j <= q, a
        //    contiguous range [lo, hi) of the sorted bins; as q goes up an arrival enters at q = b_j (> half) and leaves
        //    at q = b_j + L, and every enter (q <= L-1) precedes every leave (q >= L): at most nnz + 1 runs of samples
        //    with a constant phasor sum Z = P[hi] - P[lo].  Inside a run s_rx[q] = Im(e^{i d q} Z) with d = K * step, so
        //        sum_q s_rx[q]^2 = ( m |Z|^2 - Re( Z^2 * sum_q e^{2 i d q} ) ) / 2,   a geometric series:
        //    O(nnz) work per receiver instead of O(L) (one thread per run), which leaves this kernel with reading the
        //    rows.  Samples that are exactly zero do not count (np.nonzero, main.py:48): empty windows, and the sample
        //    q = b_j of a window that holds only arrival j (its one term is a_j sin(0)); agrees with the per-sample
        //    evaluation to ~1e-12 relative.
        const double d = __dmul_rn(K, __ddiv_rn(window, (double)(n_bins > 1 ? n_bins - 1 : 1)));
        double s2d, c2d;
        sincos(2.0 * d, &s2d, &c2d);
        const double den_re = 1.0 - c2d, den_im = -s2d;           // 1 - e^{2 i d}
        const double den2 = den_re * den_re + den_im * den_im;
        int j_enter = 0;
        { int a0 = 0, b0 = nnz; while (a0 < b0) { int m = (a0 + b0) >> 1; if (s_bin[m] <= half) a0 = m + 1; else b0 = m; } j_enter = a0; }
        int n_leave = 0; // arrivals with b_j <= half - 1
        { int a0 = 0, b0 = nnz; while (a0 < b0) { int m = (a0 + b0) >> 1; if (s_bin[m] < half) a0 = m + 1; else b0 = m; } n_leave = a0; }
        const int n_enter = nnz - j_enter;
        const int n_runs = nnz > 0 ? 1 + n_enter + n_leave : 0;
        const int64_t q_end = half + n_bins;
        for (int r = tid; r < n_runs; r += RXP_THREADS) {
            int lo, hi;
            int64_t qa, qb;
            if (r <= n_enter) {
                lo = 0; hi = j_enter + r;
                qa = r == 0 ? half : (int64_t)s_bin[j_enter + r - 1];
                qb = r < n_enter ? (int64_t)s_bin[j_enter + r] : (n_leave > 0 ? (int64_t)s_bin[0] + n_bins : q_end);
            } else {
                const int k = r - n_enter; // 1 .. n_leave
                lo = k; hi = nnz;
                qa = (int64_t)s_bin[k - 1] + n_bins;
                qb = k < n_leave ? (int64_t)s_bin[k] + n_bins : q_end;
            }
            if (qb > q_end) qb = q_end;
            const int64_t m = qb - qa;
            if (m <= 0 || hi <= lo) continue;
            const double zr = s_re[hi] - s_re[lo], zi = s_im[hi] - s_im[lo];
            // geometric series G = e^{2 i d qa} (1 - e^{2 i d m}) / (1 - e^{2 i d})
            double sa, ca, sm, cm;
            sincos(2.0 * d * (double)qa, &sa, &ca);
            sincos(2.0 * d * (double)m, &sm, &cm);
            double g_re, g_im;
            if (den2 > 1.0e-24) {
                const double nr = 1.0 - cm, ni = -sm;                               // 1 - e^{2 i d m}
                const double qr = (nr * den_re + ni * den_im) / den2, qi = (ni * den_re - nr * den_im) / den2;
                g_re = ca * qr - sa * qi; g_im = ca * qi + sa * qr;
            } else { // 2 d is a multiple of 2 pi: every term of the series equals e^{2 i d qa}
                g_re = ca * (double)m; g_im = sa * (double)m;
            }
            // s = Im(e^{i phi} Z) = |Z| sin(phi + psi):  sum s^2 = (m |Z|^2 - Re(conj-free Z^2 G)) / 2 with Z^2 = (zi + i zr)^2
            // rotated so that the sine's phase is right: sin(phi) zr + cos(phi) zi = Im(e^{i phi} (zr + i zi))
            const double z2_re = zr * zr - zi * zi, z2_im = 2.0 * zr * zi;        // (zr + i zi)^2
            const double mag2 = zr * zr + zi * zi;
            // Im(w)^2 = (|w|^2 - Re(w^2)) / 2 with w = e^{i phi} (zr + i zi):  Re(w^2) summed = Re(Z^2 G)
            const double run = 0.5 * ((double)m * mag2 - (z2_re * g_re - z2_im * g_im));
            unsigned long long c = (unsigned long long)m;
            if (hi - lo == 1 && (int64_t)s_bin[lo] == qa) --c; // q = b_j with only arrival j in the window: sin(0)
            if (mag2 != 0.0) { sum += run; cnt += c; }
        }
    } else if (!SECOND) {
        if (tid == 0) power[blockIdx.x] = -1.0; // too dense for the small staging area: left to the second launch
        return;
    } else {
        // fallback for very dense rows: the direct O(L * nnz) sum straight from the row
        for (int64_t n = tid; n < n_bins; n += RXP_THREADS) {
            const int64_t q = n + half;
            double s = 0.0;
            int64_t b0 = q - (n_bins - 1); if (b0 < 0) b0 = 0;
            int64_t b1 = q < n_bins - 1 ? q : n_bins - 1;
            for (int64_t b = b0; b <= b1; ++b) {
                const double a = row[b];
                if (a != 0.0) s = __dadd_rn(s, __dmul_rn(a, sin(stx_arg(q - b, n_bins, window, K))));
            }
            if (s != 0.0) { sum += s * s; ++cnt; }
        }
    }
    s_sum[tid] = sum;
    s_cnt[tid] = cnt;
    __syncthreads();
    for (int o = RXP_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) { s_sum[tid] += s_sum[tid + o]; s_cnt[tid] += s_cnt[tid + o]; }
        __syncthreads();
    }
    if (tid == 0) power[blockIdx.x] = s_cnt[0] ? s_sum[0] / (double)s_cnt[0] : nan("");
}

} // namespace
} // namespace rfrt

using namespace rfrt;

extern "C" int rfrt_bin_ir(const int32_t *d_rec_rx, const int64_t *d_rec_bin, const double *d_rec_amp,
                           int64_t n_records, const uint64_t *d_n_records, int64_t n_receivers, int64_t n_bins,
                           int32_t deterministic, double *d_ir, void *stream_)
{
    const unsigned long long *d_n = (const unsigned long long *)d_n_records;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_records < 0 || n_receivers <= 0 || n_bins < 0 || !d_ir ||
        (n_records > 0 && (!d_rec_rx || !d_rec_bin || !d_rec_amp))) {
        set_error("rfrt_bin_ir: bad arguments");
        return RFRT_ERR_INVALID;
    }
    if (n_records == 0 || n_bins == 0) return RFRT_OK;
    if (deterministic) {
        if (n_receivers <= 2048 && n_records >= 8 * n_receivers) {
            int64_t split = 296 / n_receivers; // ~2 CTAs per SM in flight
            const int64_t blocks256 = (n_bins + 255) / 256;
            if (split > blocks256) split = blocks256;
            if (split > 16) split = 16;
            if (split < 1) split = 1;
            k_bin_ordered_cta<<<dim3((unsigned)n_receivers, (unsigned)split), 256, 0, stream>>>(d_rec_rx, d_rec_bin, d_rec_amp,
                                                                                              n_records, d_n, n_bins, d_ir);
        }
        else
            k_bin_ordered<<<(unsigned)((n_receivers + 127) / 128), 128, 0, stream>>>(d_rec_rx, d_rec_bin, d_rec_amp,
                                                                                     n_records, d_n, n_receivers, n_bins, d_ir);
    } else {
        const size_t smem = sizeof(double) * (size_t)n_bins;
        if (n_receivers <= 4 && smem <= 200 * 1024 && n_records >= (1 << 16)) {
            int dev = 0, sms = 0;
            RFRT_CUDA(cudaGetDevice(&dev));
            RFRT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            RFRT_CUDA(cudaFuncSetAttribute((const void *)k_bin_privatised, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int ctas = smem <= 100 * 1024 ? 2 * sms : sms; // two resident CTAs per SM when the window allows it
            int64_t per_block = (n_records + ctas - 1) / ctas;
            per_block = (per_block + 31) & ~(int64_t)31;
            k_bin_privatised<<<ctas, 512, smem, stream>>>(d_rec_rx, d_rec_bin, d_rec_amp, n_records, d_n, n_receivers,
                                                         n_bins, d_ir, per_block);
        } else {
            int64_t nb = (n_records + 255) / 256;
            if (nb > 4096) nb = 4096;
            k_bin_atomic<<<(unsigned)nb, 256, 0, stream>>>(d_rec_rx, d_rec_bin, d_rec_amp, n_records, d_n, n_receivers,
                                                           n_bins, d_ir);
        }
    }
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}

extern "C" int rfrt_rx_power(const int64_t *d_arr_offsets, const int32_t *d_arr_bin, const double *d_arr_amp,
                             int64_t n_receivers, int64_t n_bins, double sample_window_s, double carrier_hz,
                             double *d_stx_table, double *d_power, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_arr_offsets || n_receivers <= 0 || n_bins <= 0 || !d_power || !d_stx_table) {
        set_error("rfrt_rx_power: bad arguments");
        return RFRT_ERR_INVALID;
    }
    k_stx_table<<<(unsigned)((n_bins + 255) / 256), 256, 0, stream>>>(n_bins, sample_window_s, carrier_hz, d_stx_table);
    k_rx_power<<<(unsigned)n_receivers, 256, 0, stream>>>(d_arr_offsets, d_arr_bin, d_arr_amp, n_bins, d_stx_table,
                                                          d_power);
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}

extern "C" int rfrt_rx_power_dense(const double *d_ir, int64_t n_receivers, int64_t n_bins, double sample_window_s,
                                   double carrier_hz, double *d_power, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_ir || n_receivers <= 0 || n_bins <= 0 || n_bins >= (1ll << 31) || !d_power) {
        set_error("rfrt_rx_power_dense: bad arguments");
        return RFRT_ERR_INVALID;
    }
    const size_t smem_small = sizeof(double) * 2 * (RXP_CAP_SMALL + 1) + sizeof(int) * RXP_CAP_SMALL;
    const size_t smem_big = sizeof(double) * 2 * (RXP_CAP_BIG + 1) + sizeof(int) * RXP_CAP_BIG;
    RFRT_CUDA(cudaFuncSetAttribute((const void *)k_rx_power_dense<RXP_CAP_BIG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem_big));
    k_rx_power_dense<RXP_CAP_SMALL, false><<<(unsigned)n_receivers, RXP_THREADS, smem_small, stream>>>(d_ir, n_bins, sample_window_s,
                                                                                                     carrier_hz, d_power);
    k_rx_power_dense<RXP_CAP_BIG, true><<<(unsigned)n_receivers, RXP_THREADS, smem_big, stream>>>(d_ir, n_bins, sample_window_s,
                                                                                                carrier_hz, d_power);
    RFRT_CUDA(cudaGetLastError());
    return RFRT_OK;
}
