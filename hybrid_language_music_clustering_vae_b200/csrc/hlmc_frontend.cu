// Kernels either side of the feature kernels:
//
//   frontend_kernel    librosa.load's arithmetic on device: int16 -> float32 (x / 32768), interleaved
//                      channels -> mono (np.mean over channels), polyphase resampling to the plan's rate
//                      (librosa.resample(res_type="polyphase") = scipy.signal.resample_poly) and the
//                      scripts' right zero pad ([R] src/1_preprocessing.py:137-153).
//   tab_* kernels      the scripts' tabular normalisation of the (N, 370) / (N, 290) float64 feature
//                      matrix: inf -> nan, SimpleImputer(mean), StandardScaler
//                      ([R] src/1_preprocessing.py:303-311, src/1_preprocessing_advanced.py:384-391).
//
// All HBM-bound streaming work: coalesced loads, grid sized from the data, no shared-memory staging needed.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "hlmc_internal.h"

namespace hlmc {

void count_launch(int n);

// ---------------------------------------------------------------------------
// front end
// ---------------------------------------------------------------------------
template <int FMT, int CH>
__device__ __forceinline__ float mono_sample(const void* __restrict__ raw, long long i, int channels) {
    // FMT 0: float32, 1: int16.  CH = 1, 2 or 0 (run-time count).  soundfile hands librosa float32 = int16 / 32768;
    // librosa.to_mono = np.mean(y, axis=0): float32 adds in channel order, then one division by the count.
    if (FMT == 1) {
        const int16_t* p = static_cast<const int16_t*>(raw);
        if (CH == 1) return float(__ldg(p + i)) * (1.0f / 32768.0f);
        if (CH == 2) {
            const int v = __ldg(reinterpret_cast<const int*>(p) + i);
            const float a = float(int16_t(v & 0xffff)) * (1.0f / 32768.0f);
            const float b = float(int16_t(v >> 16)) * (1.0f / 32768.0f);
            return (a + b) / 2.0f;
        }
        float s = float(__ldg(p + i * channels)) * (1.0f / 32768.0f);
        for (int c = 1; c < channels; ++c) s += float(__ldg(p + i * channels + c)) * (1.0f / 32768.0f);
        return s / float(channels);
    } else {
        const float* p = static_cast<const float*>(raw);
        if (CH == 1) return __ldg(p + i);
        if (CH == 2) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(p) + i);
            return (v.x + v.y) / 2.0f;
        }
        float s = __ldg(p + i * channels);
        for (int c = 1; c < channels; ++c) s += __ldg(p + i * channels + c);
        return s / float(channels);
    }
}

template <int FMT, int CH>
__global__ void __launch_bounds__(256) frontend_kernel(const FrontArgs a) {
    const long long per = a.n_total;
    const long long total = a.B * per;
    const size_t esz = (FMT == 1) ? 2 : 4;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long b = g / per, j = g - b * per;
        const char* raw = static_cast<const char*>(a.raw) + (size_t)b * a.raw_pitch * a.channels * esz;
        float v = 0.0f;
        long long n_in = a.n_in, n_out = a.n_out;
        if (a.valid != nullptr) {                  // this clip's own length (a file shorter than `duration`)
            n_in = min(max(a.valid[b], 0LL), a.n_in);
            n_out = min((n_in * a.up + a.down - 1) / a.down, a.n_total);
        }
        if (j < n_out) {
            if (a.up == 1 && a.down == 1) {
                v = mono_sample<FMT, CH>(raw, j, a.channels);
            } else {
                // y[m] = sum_i x[i] h[m*down - n_pre_pad - i*up]; taps regrouped by phase on the host:
                // hpoly[p][t] = h[p + t*up], so output m reads one contiguous row
                const long long c = (j + a.n_pre_remove) * (long long)a.down - a.n_pre_pad;
                const long long ihi = c / a.up;
                const int p = (int)(c - ihi * a.up);
                const float* hp = a.hpoly + (size_t)p * a.tpp;
                int t1 = a.tpp - 1;                              // oldest tap first, as scipy's upfirdn does
                if (ihi - t1 < 0) t1 = (int)ihi;
                int t0 = 0;
                if (ihi >= n_in) t0 = (int)(ihi - n_in + 1);
                for (int t = t1; t >= t0; --t)
                    v = fmaf(mono_sample<FMT, CH>(raw, ihi - t, a.channels), __ldg(hp + t), v);
            }
        }
        a.out[(size_t)b * a.pitch + j] = v;
    }
}

template <int FMT>
static cudaError_t launch_frontend_fmt(const FrontArgs& a, int grid, cudaStream_t st) {
    if (a.channels == 1) frontend_kernel<FMT, 1><<<grid, 256, 0, st>>>(a);
    else if (a.channels == 2) frontend_kernel<FMT, 2><<<grid, 256, 0, st>>>(a);
    else frontend_kernel<FMT, 0><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_frontend(const FrontArgs& a, int num_sms, cudaStream_t stream) {
    const long long total = a.B * a.n_total;
    if (total <= 0) return cudaSuccess;
    long long grid = (total + 255) / 256;
    const long long cap = (long long)num_sms * 16;
    if (grid > cap) grid = cap;
    count_launch(1);
    return (a.fmt == 1) ? launch_frontend_fmt<1>(a, (int)grid, stream) : launch_frontend_fmt<0>(a, (int)grid, stream);
}

// ---------------------------------------------------------------------------
// tabular normalisation, float64.  N rows x D columns, row-major; D is a few hundred, N the number of
// clips: one warp walks 32 consecutive columns of a block of rows (coalesced 256-byte reads) and the row
// blocks meet through double atomics (order-independent to ~1 ulp of the sum; the tests allow 1e-12).
// ---------------------------------------------------------------------------
constexpr int kTabRowsPerBlock = 256;

__global__ void tab_impute_stats_kernel(const double* __restrict__ x, long long N, long long D,
                                        double* __restrict__ sum, unsigned long long* __restrict__ count) {
    const long long c = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
    const int wy = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long r0 = (long long)blockIdx.y * kTabRowsPerBlock;
    const long long r1 = (r0 + kTabRowsPerBlock < N) ? r0 + kTabRowsPerBlock : N;
    if (c >= D) return;
    double s = 0.0;
    unsigned long long k = 0;
    for (long long r = r0 + wy; r < r1; r += nw) {
        const double v = x[r * D + c];
        if (isfinite(v)) { s += v; ++k; }              // inf -> nan, and nan is "missing" for the imputer
    }
    if (k) {
        atomicAdd(sum + c, s);
        atomicAdd(count + c, k);
    }
}

// two passes like sklearn's _incremental_mean_and_var: (1) column sums, (2) sums of d and d^2 about the mean
__global__ void tab_sum_kernel(const double* __restrict__ x, long long N, long long D,
                               const double* __restrict__ fill, double* __restrict__ sum) {
    const long long c = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
    const int wy = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long r0 = (long long)blockIdx.y * kTabRowsPerBlock;
    const long long r1 = (r0 + kTabRowsPerBlock < N) ? r0 + kTabRowsPerBlock : N;
    if (c >= D) return;
    const double f = fill ? fill[c] : 0.0;
    double s = 0.0;
    for (long long r = r0 + wy; r < r1; r += nw) {
        const double v = x[r * D + c];
        s += isfinite(v) ? v : f;
    }
    atomicAdd(sum + c, s);
}
__global__ void tab_dev_kernel(const double* __restrict__ x, long long N, long long D,
                               const double* __restrict__ fill, const double* __restrict__ sum,
                               double* __restrict__ d1, double* __restrict__ d2) {
    const long long c = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
    const int wy = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long r0 = (long long)blockIdx.y * kTabRowsPerBlock;
    const long long r1 = (r0 + kTabRowsPerBlock < N) ? r0 + kTabRowsPerBlock : N;
    if (c >= D) return;
    const double f = fill ? fill[c] : 0.0;
    const double mu = sum[c] / (double)N;
    double a1 = 0.0, a2 = 0.0;
    for (long long r = r0 + wy; r < r1; r += nw) {
        const double v = x[r * D + c];
        const double d = (isfinite(v) ? v : f) - mu;
        a1 += d;
        a2 = fma(d, d, a2);
    }
    atomicAdd(d1 + c, a1);
    atomicAdd(d2 + c, a2);
}
__global__ void tab_finish_kernel(long long N, long long D, const double* __restrict__ sum,
                                  const double* __restrict__ d1, const double* __restrict__ d2,
                                  double* __restrict__ mean, double* __restrict__ m2) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    mean[c] = (N > 0) ? sum[c] / (double)N : 0.0;
    m2[c] = (N > 0) ? d2[c] - d1[c] * d1[c] / (double)N : 0.0;
}

__global__ void tab_impute_scale_kernel(const double* __restrict__ x, long long N, long long D,
                                        const int* __restrict__ cols, long long Do,
                                        const double* __restrict__ fill, const double* __restrict__ mean,
                                        const double* __restrict__ scale, double* __restrict__ imputed,
                                        double* __restrict__ scaled) {
    const long long total = N * Do;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long r = g / Do, j = g - r * Do;
        const int c = cols ? cols[j] : (int)j;
        double v = x[r * D + c];
        if (!isfinite(v)) v = fill[c];
        if (imputed) imputed[g] = v;
        if (scaled) scaled[g] = (v - mean[j]) / scale[j];        // sklearn: X -= mean_; X /= scale_
    }
}

static dim3 tab_grid(long long N, long long D) {
    return dim3((unsigned)((D + 31) / 32), (unsigned)((N + kTabRowsPerBlock - 1) / kTabRowsPerBlock));
}

cudaError_t launch_impute_stats(const double* x, long long N, long long D, double* sum, long long* count,
                                cudaStream_t st) {
    if (D <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(sum, 0, (size_t)D * 8, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(count, 0, (size_t)D * 8, st);
    if (e != cudaSuccess || N <= 0) return e;
    tab_impute_stats_kernel<<<tab_grid(N, D), 256, 0, st>>>(x, N, D, sum, reinterpret_cast<unsigned long long*>(count));
    count_launch(1);
    return cudaGetLastError();
}

// scratch: 3*D doubles (sum, d1, d2), zeroed here
cudaError_t launch_scaler_stats_f64(const double* x, long long N, long long D, const double* fill, double* mean,
                                    double* m2, double* scratch, cudaStream_t st) {
    if (D <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)3 * D * 8, st);
    if (e != cudaSuccess) return e;
    double *sum = scratch, *d1 = scratch + D, *d2 = scratch + 2 * D;
    if (N > 0) {
        tab_sum_kernel<<<tab_grid(N, D), 256, 0, st>>>(x, N, D, fill, sum);
        tab_dev_kernel<<<tab_grid(N, D), 256, 0, st>>>(x, N, D, fill, sum, d1, d2);
        count_launch(2);
    }
    tab_finish_kernel<<<(unsigned)((D + 127) / 128), 128, 0, st>>>(N, D, sum, d1, d2, mean, m2);
    count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_impute_scale(const double* x, long long N, long long D, const int* cols, long long Do,
                                const double* fill, const double* mean, const double* scale, double* imputed,
                                double* scaled, cudaStream_t st) {
    const long long total = N * Do;
    if (total <= 0) return cudaSuccess;
    long long grid = (total + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    tab_impute_scale_kernel<<<(unsigned)grid, 256, 0, st>>>(x, N, D, cols, Do, fill, mean, scale, imputed, scaled);
    count_launch(1);
    return cudaGetLastError();
}

}  // namespace hlmc
