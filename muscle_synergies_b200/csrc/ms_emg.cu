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
#include <string.h>

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

// Block version (windows that fit shared memory): a CTA owns RMS_SPAN consecutive outputs of one channel, stages
// the squares of the RMS_SPAN + window - 1 samples they need with coalesced loads, turns them into prefix sums
// in place (each thread scans a contiguous chunk, chunk totals are scanned across the block), and every output
// is one difference of two prefix values - coalesced stores, a handful of shared-memory reads per output.
// The prefix sums are kept as unevaluated pairs hi + lo (error-free TwoSum), so that the difference is as
// accurate in a quiet stretch next to a burst as the direct window sum numpy computes.
#define RMS_SPAN 4096
#define RMS_THREADS 256
struct MsDD {
    double hi, lo;
};
__device__ __forceinline__ MsDD ms_dd_add(MsDD a, MsDD b) {
    const double s = a.hi + b.hi;
    const double bb = s - a.hi;
    const double err = (a.hi - (s - bb)) + (b.hi - bb);
    const double lo = (a.lo + b.lo) + err;
    MsDD r;
    r.hi = s + lo;
    r.lo = lo - (r.hi - s);
    return r;
}
__global__ void __launch_bounds__(RMS_THREADS)
    ms_rms_block_kernel(const double* __restrict__ src, int64_t stride, int64_t n, const double* __restrict__ mean,
                        int window, int chunk, double* __restrict__ out, int64_t out_stride) {
    extern __shared__ double s_dyn[];  // hi[RMS_THREADS * chunk], lo[RMS_THREADS * chunk]
    __shared__ MsDD s_warp[RMS_THREADS / 32];
    const int c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double mu = mean ? mean[c] : 0.0;
    const double* __restrict__ x = src + (int64_t)c * stride;
    const int h = (window - 1) / 2;
    const int64_t o0 = (int64_t)blockIdx.x * RMS_SPAN;       // first output of this CTA
    const int64_t in0 = o0 - (window - 1 - h);               // first sample it needs
    const int need = RMS_SPAN + window - 1, total = RMS_THREADS * chunk;  // total >= need
    double* s_hi = s_dyn;
    double* s_lo = s_dyn + total;
    for (int k = tid; k < total; k += RMS_THREADS) {
        const int64_t g = in0 + k;
        double v = 0.0;
        if (k < need && g >= 0 && g < n) {
            v = x[g] - mu;
            v *= v;
        }
        s_hi[k] = v;
    }
    __syncthreads();
    // inclusive scan of this thread's chunk, in place
    MsDD run = {0.0, 0.0};
    double* mine_hi = s_hi + tid * chunk;
    double* mine_lo = s_lo + tid * chunk;
    for (int k = 0; k < chunk; k++) {
        const MsDD v = {mine_hi[k], 0.0};
        run = ms_dd_add(run, v);
        mine_hi[k] = run.hi;
        mine_lo[k] = run.lo;
    }
    // exclusive scan of the chunk totals across the block
    MsDD inc = run;
    for (int o = 1; o < 32; o <<= 1) {
        MsDD t;
        t.hi = __shfl_up_sync(0xffffffffu, inc.hi, o);
        t.lo = __shfl_up_sync(0xffffffffu, inc.lo, o);
        if (lane >= o) inc = ms_dd_add(t, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    // base = everything before this thread's chunk = (warps before) + (lanes before in this warp)
    MsDD before = {__shfl_up_sync(0xffffffffu, inc.hi, 1), __shfl_up_sync(0xffffffffu, inc.lo, 1)};
    if (lane == 0) before.hi = before.lo = 0.0;
    MsDD base = {0.0, 0.0};
    for (int w = 0; w < warp; w++) base = ms_dd_add(base, s_warp[w]);
    base = ms_dd_add(base, before);
    for (int k = 0; k < chunk; k++) {
        const MsDD v = {mine_hi[k], mine_lo[k]};
        const MsDD r = ms_dd_add(base, v);
        mine_hi[k] = r.hi;
        mine_lo[k] = r.lo;
    }
    __syncthreads();
    // prefix k = sum of squares of samples in0 .. in0 + k; output o needs samples o - (window-1-h) .. o + h
    const double inv = 1.0 / (double)window;
    double* __restrict__ y = out + (int64_t)c * out_stride;
    for (int k = tid; k < RMS_SPAN; k += RMS_THREADS) {
        const int64_t o = o0 + k;
        if (o >= n) break;
        const int b = k + window - 1;
        double sum = s_hi[b], low = s_lo[b];
        if (k > 0) {
            sum -= s_hi[k - 1];
            low -= s_lo[k - 1];
        }
        y[o] = sqrt(fmax(sum + low, 0.0) * inv);
    }
}

extern "C" int ms_rms_envelope(const double* d_src, int64_t stride, int32_t n_channels, int64_t n, const double* d_mean,
                               int32_t window, double* d_out, int64_t out_stride, void* stream) {
    if (!d_src || !d_out || n_channels < 1 || n < 1 || window < 1 || window > n) return MS_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int chunk = (RMS_SPAN + window - 1 + RMS_THREADS - 1) / RMS_THREADS;
    chunk |= 1;  // odd chunk length: neighbouring threads' chunks start in different banks
    const size_t smem = sizeof(double) * 2 * (size_t)RMS_THREADS * chunk;
    if (smem <= 200 * 1024) {
        MS_CUDA_CHECK(cudaFuncSetAttribute(ms_rms_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const dim3 grid((unsigned)((n + RMS_SPAN - 1) / RMS_SPAN), n_channels);
        ms_rms_block_kernel<<<grid, RMS_THREADS, smem, st>>>(d_src, stride, n, d_mean, window, chunk, d_out, out_stride);
    } else {  // very long windows: the running-sum kernel
        const int64_t threads = (n + RMS_RUN - 1) / RMS_RUN;
        ms_rms_kernel<<<dim3((unsigned)((threads + 127) / 128), n_channels), 128, 0, st>>>(d_src, stride, n, d_mean,
                                                                                         window, d_out, out_stride);
    }
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

// ---- IIR filtering: cascaded second-order sections, forward or zero-lag -------------------------------
// digital_filter (analysis.py:314-432) applies scipy.signal.sosfilt / sosfiltfilt along time; linear_envelope
// (analysis.py:252-311) is zero-centre -> abs -> that low-pass.  A recursion along time is a serial
// chain per channel, so the signal is cut into chunks of SOS_CHUNK samples that run in parallel:
//   transition  M = state after SOS_CHUNK steps of zero input from each unit state (the filter is linear)
//   zero-state  every chunk from a zero state                     -> its final state f_c
//   carry       per channel, over the chunks: z_{c+1} = M z_c + f_c   (the only serial part)
//   final       every chunk again from its true start state z_c   -> the outputs
// Each section is scipy's direct-form-II-transposed step (scipy/signal/_sosfilt.pyx:_sosfilt_float).
// sosfiltfilt = odd extension by padlen samples on both ends, forward run started at zi * ext[0], backward run
// over the reversed result started at zi * (its first sample), reversed again, extension dropped
// (scipy/signal/_signaltools.py:sosfiltfilt); the extension and the rectification are generated on the fly.
#define SOS_CHUNK 1024
#define SOS_MAX_SECTIONS 8

struct MsSosCoef {
    double b0[SOS_MAX_SECTIONS], b1[SOS_MAX_SECTIONS], b2[SOS_MAX_SECTIONS], a1[SOS_MAX_SECTIONS], a2[SOS_MAX_SECTIONS];
    double zi[2 * SOS_MAX_SECTIONS];  // steady-state initial condition per unit input (sosfilt_zi), zero for a plain sosfilt
};
struct MsSosIo {
    const double* src;  // forward: the recording [channel][stride]; backward: the forward result [channel][N]
    int64_t src_stride;
    int64_t n;       // samples of the recording
    int64_t padlen;  // odd extension on both sides (0: none)
    const double* mean;  // forward only: subtracted first (NULL: nothing)
    int rectify;         // forward only: |x - mean|
    int backward;        // 0: forward over the extended recording, 1: backward over src
    double* dst;         // [channel][dst_stride]; forward: all N samples, backward: the extension dropped
    int64_t dst_stride;
};

__device__ __forceinline__ double ms_sos_value(const MsSosIo& io, const double* __restrict__ x, double mu, int64_t k) {
    const double v = x[k] - mu;
    return io.rectify ? fabs(v) : v;
}
// sample i of the sequence this pass filters (N = n + 2 padlen of them)
__device__ __forceinline__ double ms_sos_input(const MsSosIo& io, const double* __restrict__ x, double mu, int64_t i,
                                               int64_t N) {
    if (io.backward) return x[N - 1 - i];
    const int64_t j = i - io.padlen;
    if (j < 0) return 2.0 * ms_sos_value(io, x, mu, 0) - ms_sos_value(io, x, mu, io.padlen - i);
    if (j >= io.n) return 2.0 * ms_sos_value(io, x, mu, io.n - 1) - ms_sos_value(io, x, mu, io.n - 2 - (j - io.n));
    return ms_sos_value(io, x, mu, j);
}

template <int S>
__device__ __forceinline__ double ms_sos_step(const MsSosCoef& c, double (&z)[2 * S], double x) {
#pragma unroll
    for (int s = 0; s < S; s++) {
        const double y = c.b0[s] * x + z[2 * s];
        z[2 * s] = c.b1[s] * x - c.a1[s] * y + z[2 * s + 1];
        z[2 * s + 1] = c.b2[s] * x - c.a2[s] * y;
        x = y;
    }
    return x;
}

template <int S>
__global__ void ms_sos_transition_kernel(MsSosCoef c, double* __restrict__ M) {
    const int j = threadIdx.x;  // unit state j
    if (j >= 2 * S) return;
    double z[2 * S];
#pragma unroll
    for (int i = 0; i < 2 * S; i++) z[i] = i == j ? 1.0 : 0.0;
    for (int t = 0; t < SOS_CHUNK; t++) ms_sos_step<S>(c, z, 0.0);
#pragma unroll
    for (int i = 0; i < 2 * S; i++) M[i * 2 * S + j] = z[i];
}

// FINAL = 0: zero-state run, writes the chunk's final state; FINAL = 1: run from the carried state, writes the outputs
template <int S, int FINAL>
__global__ void __launch_bounds__(64)
    ms_sos_chunk_kernel(MsSosCoef c, MsSosIo io, int64_t N, int64_t n_chunks, double* __restrict__ state) {
    const int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    if (chunk >= n_chunks) return;
    const double* __restrict__ x = io.src + (int64_t)ch * io.src_stride;
    const double mu = (!io.backward && io.mean) ? io.mean[ch] : 0.0;
    double* __restrict__ st = state + ((int64_t)ch * n_chunks + chunk) * (2 * S);
    double z[2 * S];
#pragma unroll
    for (int i = 0; i < 2 * S; i++) z[i] = FINAL ? st[i] : 0.0;
    const int64_t i0 = chunk * SOS_CHUNK, i1 = min(N, i0 + SOS_CHUNK);
    if (!FINAL) {
        for (int64_t i = i0; i < i1; i++) ms_sos_step<S>(c, z, ms_sos_input(io, x, mu, i, N));
#pragma unroll
        for (int k = 0; k < 2 * S; k++) st[k] = z[k];
    } else if (!io.backward) {
        double* __restrict__ y = io.dst + (int64_t)ch * io.dst_stride;
        for (int64_t i = i0; i < i1; i++) y[i] = ms_sos_step<S>(c, z, ms_sos_input(io, x, mu, i, N));
    } else {
        double* __restrict__ y = io.dst + (int64_t)ch * io.dst_stride;
        for (int64_t i = i0; i < i1; i++) {
            const double v = ms_sos_step<S>(c, z, ms_sos_input(io, x, mu, i, N));
            const int64_t j = N - 1 - i - io.padlen;  // position in the recording
            if (j >= 0 && j < io.n) y[j] = v;
        }
    }
}

// one thread per channel: turns the zero-state final states into every chunk's start state, in place
template <int S>
__global__ void ms_sos_carry_kernel(MsSosCoef c, MsSosIo io, int64_t N, int64_t n_chunks, int n_channels,
                                    const double* __restrict__ M, double* __restrict__ state) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_channels) return;
    const double* __restrict__ x = io.src + (int64_t)ch * io.src_stride;
    const double mu = (!io.backward && io.mean) ? io.mean[ch] : 0.0;
    const double first = ms_sos_input(io, x, mu, 0, N);
    double m[2 * S][2 * S], z[2 * S];
#pragma unroll
    for (int i = 0; i < 2 * S; i++) {
        z[i] = c.zi[i] * first;
#pragma unroll
        for (int j = 0; j < 2 * S; j++) m[i][j] = M[i * 2 * S + j];
    }
    double* __restrict__ st = state + (int64_t)ch * n_chunks * (2 * S);
    double f[2 * S];
#pragma unroll
    for (int i = 0; i < 2 * S; i++) f[i] = st[i];
    for (int64_t k = 0; k < n_chunks; k++) {
        double fn[2 * S];
        const int64_t kn = k + 1 < n_chunks ? k + 1 : k;  // next chunk's loads overlap this chunk's arithmetic
#pragma unroll
        for (int i = 0; i < 2 * S; i++) fn[i] = st[kn * (2 * S) + i];
        double zn[2 * S];
#pragma unroll
        for (int i = 0; i < 2 * S; i++) {
            st[k * (2 * S) + i] = z[i];
            double acc = f[i];
#pragma unroll
            for (int j = 0; j < 2 * S; j++) acc += m[i][j] * z[j];
            zn[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 2 * S; i++) {
            z[i] = zn[i];
            f[i] = fn[i];
        }
    }
}

template <int S>
static int ms_sos_run(const MsSosCoef& c, MsSosIo io, int n_channels, double* M, double* state, cudaStream_t st) {
    const int64_t N = io.n + 2 * io.padlen;
    const int64_t n_chunks = (N + SOS_CHUNK - 1) / SOS_CHUNK;
    const dim3 grid((unsigned)((n_chunks + 63) / 64), n_channels);
    ms_sos_chunk_kernel<S, 0><<<grid, 64, 0, st>>>(c, io, N, n_chunks, state);
    MS_COUNT_LAUNCH();
    ms_sos_carry_kernel<S><<<(n_channels + 31) / 32, 32, 0, st>>>(c, io, N, n_chunks, n_channels, M, state);
    MS_COUNT_LAUNCH();
    ms_sos_chunk_kernel<S, 1><<<grid, 64, 0, st>>>(c, io, N, n_chunks, state);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

template <int S>
static int ms_sosfilt_impl(const double* d_src, int64_t stride, int32_t n_channels, int64_t n, const MsSosCoef& c,
                           int64_t padlen, int zero_lag, const double* d_mean, int rectify, double* d_out,
                           int64_t out_stride, double* work, cudaStream_t st) {
    const int64_t N = n + 2 * padlen;
    const int64_t n_chunks = (N + SOS_CHUNK - 1) / SOS_CHUNK;
    double* M = work;
    double* state = M + 4 * SOS_MAX_SECTIONS * SOS_MAX_SECTIONS;
    double* tmp = state + (int64_t)n_channels * n_chunks * 2 * SOS_MAX_SECTIONS;
    ms_sos_transition_kernel<S><<<1, 32, 0, st>>>(c, M);
    MS_COUNT_LAUNCH();
    MsSosIo io;
    io.src = d_src, io.src_stride = stride, io.n = n, io.padlen = padlen, io.mean = d_mean, io.rectify = rectify;
    io.backward = 0;
    if (!zero_lag) {  // sosfilt: forward, from rest, straight into the output
        io.dst = d_out, io.dst_stride = out_stride;
        return ms_sos_run<S>(c, io, n_channels, M, state, st);
    }
    io.dst = tmp, io.dst_stride = N;
    int rc = ms_sos_run<S>(c, io, n_channels, M, state, st);
    if (rc != MS_OK) return rc;
    io.src = tmp, io.src_stride = N, io.mean = nullptr, io.rectify = 0, io.backward = 1;
    io.dst = d_out, io.dst_stride = out_stride;
    return ms_sos_run<S>(c, io, n_channels, M, state, st);
}

extern "C" size_t ms_sosfilt_workspace_bytes(int64_t n, int32_t n_channels, int64_t padlen, int32_t zero_lag) {
    if (n < 0 || n_channels < 0 || padlen < 0) return 0;
    const int64_t N = n + 2 * padlen;
    const int64_t n_chunks = (N + SOS_CHUNK - 1) / SOS_CHUNK;
    size_t doubles = 4 * SOS_MAX_SECTIONS * SOS_MAX_SECTIONS + (size_t)n_channels * n_chunks * 2 * SOS_MAX_SECTIONS;
    if (zero_lag) doubles += (size_t)n_channels * N;
    return doubles * sizeof(double);
}

extern "C" int ms_sosfilt(const double* d_src, int64_t stride, int32_t n_channels, int64_t n, const double* h_sos,
                          int32_t n_sections, const double* h_zi, int64_t padlen, int32_t zero_lag, const double* d_mean,
                          int32_t rectify, double* d_out, int64_t out_stride, void* d_work, void* stream) {
    if (!d_src || !d_out || !h_sos || !d_work || n_channels < 1 || n < 1 || n_sections < 1 ||
        n_sections > SOS_MAX_SECTIONS || padlen < 0)
        return MS_E_INVALID;
    if (zero_lag && (!h_zi || n <= padlen)) return MS_E_INVALID;
    if (!zero_lag && padlen != 0) return MS_E_INVALID;
    MsSosCoef c;
    memset(&c, 0, sizeof(c));
    for (int s = 0; s < n_sections; s++) {
        const double* r = h_sos + 6 * s;  // b0 b1 b2 a0 a1 a2, a0 == 1 (scipy's sos layout)
        if (r[3] != 1.0) return MS_E_INVALID;
        c.b0[s] = r[0], c.b1[s] = r[1], c.b2[s] = r[2], c.a1[s] = r[4], c.a2[s] = r[5];
        if (zero_lag) c.zi[2 * s] = h_zi[2 * s], c.zi[2 * s + 1] = h_zi[2 * s + 1];
    }
    cudaStream_t st = (cudaStream_t)stream;
    double* work = (double*)d_work;
#define MS_SOS_CASE(S)                                                                                                  \
    case S:                                                                                                             \
        return ms_sosfilt_impl<S>(d_src, stride, n_channels, n, c, padlen, zero_lag, d_mean, rectify, d_out, out_stride, \
                                  work, st);
    switch (n_sections) {
        MS_SOS_CASE(1)
        MS_SOS_CASE(2)
        MS_SOS_CASE(3)
        MS_SOS_CASE(4)
        MS_SOS_CASE(5)
        MS_SOS_CASE(6)
        MS_SOS_CASE(7)
        MS_SOS_CASE(8)
    }
#undef MS_SOS_CASE
    return MS_E_INVALID;
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
