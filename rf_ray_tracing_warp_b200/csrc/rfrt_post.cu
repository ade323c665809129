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

// Privatised variant for few receivers / many records: each CTA owns a contiguous slice of the records,
// accumulates the bins that fall into its shared-memory window with shared atomics and flushes the
// non-zero bins with one global atomic each.  Used when one receiver's histogram fits in shared memory.
__global__ void k_bin_privatised(const int32_t *__restrict__ rx, const int64_t *__restrict__ bin,
                                 const double *__restrict__ amp, int64_t n, const unsigned long long *d_n, int64_t n_rx,
                                 int64_t n_bins, double *ir, int64_t per_block)
{
    extern __shared__ double s_hist[];
    if (d_n && (int64_t)*d_n < n) n = (int64_t)*d_n;
    const int64_t begin = blockIdx.x * per_block;
    const int64_t end = begin + per_block < n ? begin + per_block : n;
    for (int64_t k = 0; k < n_rx; ++k) {
        for (int64_t b = threadIdx.x; b < n_bins; b += blockDim.x) s_hist[b] = 0.0;
        __syncthreads();
        for (int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
            int64_t b = bin[i];
            if (rx[i] == k && b >= 0 && b < n_bins) atomicAdd(&s_hist[b], amp[i]);
        }
        __syncthreads();
        for (int64_t b = threadIdx.x; b < n_bins; b += blockDim.x) {
            double v = s_hist[b];
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
        k_bin_ordered<<<(unsigned)((n_receivers + 127) / 128), 128, 0, stream>>>(d_rec_rx, d_rec_bin, d_rec_amp,
                                                                                 n_records, d_n, n_receivers, n_bins, d_ir);
    } else {
        const size_t smem = sizeof(double) * (size_t)n_bins;
        if (n_receivers <= 4 && smem <= 200 * 1024 && n_records >= (1 << 16)) {
            int dev = 0, sms = 0;
            RFRT_CUDA(cudaGetDevice(&dev));
            RFRT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            RFRT_CUDA(cudaFuncSetAttribute((const void *)k_bin_privatised, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int64_t per_block = (n_records + sms - 1) / sms;
            k_bin_privatised<<<sms, 512, smem, stream>>>(d_rec_rx, d_rec_bin, d_rec_amp, n_records, d_n, n_receivers,
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
