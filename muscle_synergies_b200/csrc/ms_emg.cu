// EMG envelope chain on the GPU ("next" row 1 of SURVEY.md section 8f): the step between the
// cut windows and the NMF stage.
//
//   ms_channel_means          zero_center:   column means                       analysis.py:230-249
//   ms_rms_envelope           rms:           sqrt(convolve(x^2, ones(w)/w, "same"))   analysis.py:435-507
//   ms_time_normalize_windows time_normalize (linear interp1d onto linspace(0,1,R))    analysis.py:551-594
//                             + normalize (divide by the column max |.|)               analysis.py:510-525
//
// float64 throughout; sums are ordered differently from numpy's, so parity with the reference
// functions is to a stated tolerance (tests/test_emg_gpu.py), not bit-exact.
#include <stdio.h>

#include "ms_common.cuh"

// ---- column means ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    ms_channel_sum_kernel(const double* __restrict__ src, int64_t stride, int64_t n, double* __restrict__ sums) {
    const int c = blockIdx.y;
    const double* __restrict__ x = src + (int64_t)c * stride;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += x[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += s[w];
        atomicAdd(&sums[c], t);
    }
}
__global__ void ms_scale_kernel(double* __restrict__ v, int n_ch, double factor) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_ch) v[c] *= factor;
}

extern "C" int ms_channel_means(const double* d_src, int64_t stride, int32_t n_channels, int64_t n, double* d_mean,
                                void* stream) {
    if (!d_src || !d_mean || n_channels < 1 || n < 1) return MS_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    MS_CUDA_CHECK(cudaMemsetAsync(d_mean, 0, sizeof(double) * n_channels, st));
    unsigned gx = (unsigned)((n + 256 * 16 - 1) / (256 * 16));
    if (gx < 1) gx = 1;
    if (gx > 512) gx = 512;
    ms_channel_sum_kernel<<<dim3(gx, n_channels), 256, 0, st>>>(d_src, stride, n, d_mean);
    MS_COUNT_LAUNCH();
    ms_scale_kernel<<<(n_channels + 63) / 64, 64, 0, st>>>(d_mean, n_channels, 1.0 / (double)n);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

// ---- moving RMS ---------------------------------------------------------------------------------------
// np.convolve(sq, ones(w)/w, "same")[i] = (1/w) * sum_{k = i - (w-1-h)}^{i + h} sq[k], h = (w-1)//2, zero outside
#define RMS_RUN 128
__global__ void __launch_bounds__(128)
    ms_rms_kernel(const double* __restrict__ src, int64_t stride, int64_t n, const double* __restrict__ mean, int window,
                  double* __restrict__ out, int64_t out_stride) {
    const int c = blockIdx.y;
    const double mu = mean ? mean[c] : 0.0;
    const double* __restrict__ x = src + (int64_t)c * stride;
    double* __restrict__ y = out + (int64_t)c * out_stride;
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * RMS_RUN;
    if (i0 >= n) return;
    const int h = (window - 1) / 2;
    const int64_t lo0 = i0 - (window - 1 - h), hi0 = i0 + h;  // inclusive range of the first output
    double acc = 0.0;
    for (int64_t k = max(lo0, (int64_t)0); k <= min(hi0, n - 1); k++) {
        const double v = x[k] - mu;
        acc += v * v;
    }
    const double inv = 1.0 / (double)window;
    const int64_t i1 = min(n, i0 + RMS_RUN);
    for (int64_t i = i0; i < i1; i++) {
        y[i] = sqrt(fmax(acc, 0.0) * inv);
        // slide: drop the oldest sample, take the next one
        const int64_t drop = i - (window - 1 - h), take = i + h + 1;
        if (drop >= 0 && drop < n) {
            const double v = x[drop] - mu;
            acc -= v * v;
        }
        if (take < n) {
            const double v = x[take] - mu;
            acc += v * v;
        }
    }
}

extern "C" int ms_rms_envelope(const double* d_src, int64_t stride, int32_t n_channels, int64_t n, const double* d_mean,
                               int32_t window, double* d_out, int64_t out_stride, void* stream) {
    if (!d_src || !d_out || n_channels < 1 || n < 1 || window < 1 || window > n) return MS_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t threads = (n + RMS_RUN - 1) / RMS_RUN;
    ms_rms_kernel<<<dim3((unsigned)((threads + 127) / 128), n_channels), 128, 0, st>>>(d_src, stride, n, d_mean, window,
                                                                                     d_out, out_stride);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

// ---- time normalisation + amplitude normalisation per window -------------------------------------------
// out[w][j][c], j < reduce_to: value of channel c at normalised time j / (reduce_to - 1) of window w,
// linear interpolation between the window's samples; then divided by max_j |out[w][j][c]|.
__global__ void __launch_bounds__(256)
    ms_time_normalize_kernel(const double* __restrict__ env, int64_t stride, int n_ch, const int64_t* __restrict__ starts,
                             const int64_t* __restrict__ stops, int reduce_to, int normalize, double* __restrict__ out) {
    extern __shared__ unsigned long long s_max[];  // per channel: bits of the max |value| (non-negative doubles order like integers)
    const int w = blockIdx.x;
    const int64_t a = starts[w], b = stops[w];
    const int64_t len = b - a;
    double* __restrict__ o = out + (int64_t)w * reduce_to * n_ch;
    for (int c = threadIdx.x; c < n_ch; c += blockDim.x) s_max[c] = 0ull;
    __syncthreads();
    const int total = reduce_to * n_ch;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int j = e / n_ch, c = e - j * n_ch;
        double v = 0.0;
        if (len >= 2) {
            const double t = reduce_to > 1 ? (double)j * (double)(len - 1) / (double)(reduce_to - 1) : 0.0;
            int64_t k = (int64_t)floor(t);
            if (k > len - 2) k = len - 2;
            const double f = t - (double)k;
            const double* __restrict__ x = env + (int64_t)c * stride + a + k;
            const double y0 = x[0], y1 = x[1];
            v = y0 + f * (y1 - y0);
        } else if (len == 1) {
            v = env[(int64_t)c * stride + a];
        }
        o[e] = v;
        if (normalize) atomicMax(&s_max[c], (unsigned long long)__double_as_longlong(fabs(v)));
    }
    if (!normalize) return;
    __syncthreads();
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int c = e % n_ch;
        o[e] = o[e] / __longlong_as_double((long long)s_max[c]);
    }
}

extern "C" int ms_time_normalize_windows(const double* d_env, int64_t stride, int32_t n_channels, const int64_t* d_starts,
                                         const int64_t* d_stops, int32_t n_windows, int32_t reduce_to, int32_t normalize,
                                         double* d_out, void* stream) {
    if (n_windows == 0) return MS_OK;
    if (!d_env || !d_starts || !d_stops || !d_out || n_channels < 1 || n_windows < 0 || reduce_to < 1)
        return MS_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    ms_time_normalize_kernel<<<n_windows, 256, sizeof(unsigned long long) * n_channels, st>>>(
        d_env, stride, n_channels, d_starts, d_stops, reduce_to, normalize, d_out);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}
