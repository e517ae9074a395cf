// Trial windowing kernels (project/segment.py of the reference).
//
//   ms_phase_valid_kernel   per sample: does a run of >= min_phase_size samples with exactly
//                           one / exactly two loaded force plates start here?  -> two bitmaps
//   ms_transition_chase_kernel  one CTA: the alternating 1-leg / 2-leg search of
//                           _transition_indices (segment.py:667-755) over the bitmaps
//   ms_cut_windows_kernel   DeviceData.__getitem__(slice) (user_data.py:727-731) for a batch
//                           of windows over a channel-major array
#include <stdio.h>

#include "ms_common.cuh"

extern "C" int64_t ms_transitions_workspace_bytes(int64_t n) { return 2 * ((n + 31) / 32 + 1) * 4; }

// loaded(i) = value != 0; NaN != 0 is true, as in numpy (segment.py:716-721)
__global__ void __launch_bounds__(256)
    ms_phase_valid_kernel(const double* __restrict__ left, const double* __restrict__ right, int64_t n, int min_phase,
                          uint32_t* __restrict__ valid1, uint32_t* __restrict__ valid2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool v1 = false, v2 = false;
    if (i < n) {
        // correct_activation[ind : ind + min_phase_size].all() - the slice is cut at the end
        // of the signal (segment.py:730)
        const int64_t end = min(n, i + (int64_t)min_phase);
        v1 = v2 = true;
        for (int64_t j = i; j < end; j++) {
            const bool a = left[j] != 0.0, b = right[j] != 0.0;
            v1 = v1 && (a != b);
            v2 = v2 && (a && b);
            if (!v1 && !v2) break;
        }
        if (min_phase <= 0) {  // an empty slice is all-true, but look_for only visits true samples
            const bool a = left[i] != 0.0, b = right[i] != 0.0;
            v1 = a != b;
            v2 = a && b;
        }
    }
    const uint32_t w1 = __ballot_sync(0xffffffffu, v1), w2 = __ballot_sync(0xffffffffu, v2);
    if ((threadIdx.x & 31) == 0 && i < n + 32) {
        const int64_t w = i >> 5;
        if (w < (n + 31) / 32) {
            valid1[w] = w1;
            valid2[w] = w2;
        }
    }
}

#define CHASE_THREADS 1024

__global__ void __launch_bounds__(CHASE_THREADS)
    ms_transition_chase_kernel(const double* __restrict__ left, const double* __restrict__ right, int64_t n,
                               const uint32_t* __restrict__ valid1, const uint32_t* __restrict__ valid2,
                               int num_segments, int64_t* __restrict__ transitions, int32_t* __restrict__ loaded,
                               int32_t* __restrict__ n_found) {
    __shared__ long long s_min;
    const int tid = threadIdx.x;
    const int64_t n_words = (n + 31) / 32;
    int64_t cursor = 0;
    int found = 0;
    for (int s = 0; s < num_segments; s++) {
        const uint32_t* bm = (s & 1) ? valid2 : valid1;  // 1 leg, 2 legs, 1 leg, ...
        // first set bit at index >= cursor
        int64_t base = cursor >> 5;
        long long hit = -1;
        while (base < n_words) {
            if (tid == 0) s_min = 0x7fffffffffffffffll;
            __syncthreads();
            // two words per thread and round (independent loads): 65536 samples per round, so that a search
            // across one gait phase usually ends in its first round - the rounds are what this kernel costs
            const int64_t wa = base + tid, wb = wa + CHASE_THREADS;
            uint32_t word_a = wa < n_words ? bm[wa] : 0u;
            const uint32_t word_b = wb < n_words ? bm[wb] : 0u;
            if (wa == (cursor >> 5)) word_a &= ~((1u << (cursor & 31)) - 1u);
            if (word_a)
                atomicMin(&s_min, (long long)(wa * 32 + __ffs(word_a) - 1));
            else if (word_b)
                atomicMin(&s_min, (long long)(wb * 32 + __ffs(word_b) - 1));
            __syncthreads();
            const long long m = s_min;
            __syncthreads();
            if (m != 0x7fffffffffffffffll) {
                hit = m;
                break;
            }
            base += 2 * CHASE_THREADS;
        }
        if (hit < 0) break;
        cursor = hit;
        if (tid == 0) {
            transitions[s] = hit;
            if (loaded) loaded[s] = (left[hit] != 0.0 ? 1 : 0) | (right[hit] != 0.0 ? 2 : 0);
        }
        found++;
    }
    if (tid == 0) *n_found = found;
}

extern "C" int ms_find_transitions(const double* d_left_fz, const double* d_right_fz, int64_t n, int32_t min_phase_size,
                                   int32_t num_segments, void* d_work, int64_t* d_transitions, int32_t* d_loaded,
                                   int32_t* d_n_found, void* stream) {
    if (!d_left_fz || !d_right_fz || !d_work || !d_transitions || !d_n_found || n < 0 || num_segments < 0)
        return MS_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* v1 = (uint32_t*)d_work;
    uint32_t* v2 = v1 + ((n + 31) / 32 + 1);
    if (n > 0) {
        const int64_t padded = (n + 31) / 32 * 32;
        const unsigned blocks = (unsigned)((padded + 255) / 256);
        ms_phase_valid_kernel<<<blocks, 256, 0, st>>>(d_left_fz, d_right_fz, n, min_phase_size, v1, v2);
        MS_COUNT_LAUNCH();
        MS_CUDA_CHECK(cudaGetLastError());
    }
    ms_transition_chase_kernel<<<1, CHASE_THREADS, 0, st>>>(d_left_fz, d_right_fz, n, v1, v2, num_segments,
                                                           d_transitions, d_loaded, d_n_found);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    ms_cut_windows_kernel(const double* __restrict__ src, int64_t src_stride, int n_channels,
                          const int64_t* __restrict__ starts, const int64_t* __restrict__ stops,
                          const int64_t* __restrict__ out_offsets, double* __restrict__ out) {
    const int w = blockIdx.x / n_channels, c = blockIdx.x % n_channels;
    const int64_t a = starts[w], b = stops[w];
    const int64_t len = b > a ? b - a : 0;
    const double* __restrict__ s = src + (int64_t)c * src_stride + a;
    double* __restrict__ d = out + out_offsets[w] + (int64_t)c * len;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.y * blockDim.x)
        d[i] = s[i];
}

extern "C" int ms_cut_windows(const double* d_src, int64_t src_stride, int32_t n_channels, const int64_t* d_starts,
                              const int64_t* d_stops, const int64_t* d_out_offsets, int32_t n_windows, double* d_out,
                              int64_t max_window_len, void* stream) {
    if (n_windows == 0 || n_channels == 0) return MS_OK;
    if (!d_src || !d_starts || !d_stops || !d_out_offsets || !d_out || n_windows < 0 || n_channels < 0)
        return MS_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned gy = (unsigned)((max_window_len + 256 * 8 - 1) / (256 * 8));
    if (gy < 1) gy = 1;
    if (gy > 64) gy = 64;
    dim3 grid((unsigned)n_windows * (unsigned)n_channels, gy);
    ms_cut_windows_kernel<<<grid, 256, 0, st>>>(d_src, src_stride, n_channels, d_starts, d_stops, d_out_offsets, d_out);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

// ---------------------------------------------------------------------------------------------------
// The 32 phase windows (or 8 cycle windows) of a trial as row ranges of one device, computed ON the GPU from
// the transitions still in device memory, so the gather can be queued behind the search without a host
// round trip.  Same arithmetic as the host path: phase i of trecho k runs from transition 10k+1+i to
// transition 10k+2+i minus one (_organize_transitions, segment.py:862-917), both converted to (frame,
// subframe) and back by the device's own tracker (user_data.py:513-661), then used as df.iloc[a:b]:
//   a = t_j / divisor, b = (t_{j+1} - 1) / divisor      divisor = 1 for the force plate / EMG section,
//                                                        num_subframes for trajectory markers.
__global__ void ms_plan_phase_windows_kernel(const int64_t* __restrict__ transitions, const int32_t* __restrict__ n_found,
                                             int num_segments, int cycles, int64_t divisor, int64_t n_rows, int n_channels,
                                             int64_t* __restrict__ starts, int64_t* __restrict__ stops,
                                             int64_t* __restrict__ offsets) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int per_trecho = cycles ? 2 : 8, span = cycles ? 4 : 1, n_windows = 4 * per_trecho;
    const bool ok = *n_found >= num_segments && num_segments >= 40;
    int64_t off = 0;
    for (int w = 0; w < n_windows; w++) {
        int64_t a = 0, b = 0;
        if (ok) {
            const int j = 10 * (w / per_trecho) + 1 + (w % per_trecho) * span;
            a = transitions[j] / divisor;
            b = (transitions[j + span] - 1) / divisor;
            a = a < 0 ? 0 : (a > n_rows ? n_rows : a);
            b = b < a ? a : (b > n_rows ? n_rows : b);
        }
        starts[w] = a;
        stops[w] = b;
        offsets[w] = off;
        off += (b - a) * n_channels;
    }
    offsets[n_windows] = off;
}

extern "C" int ms_plan_phase_windows(const int64_t* d_transitions, const int32_t* d_n_found, int32_t num_segments,
                                     int32_t cycles, int64_t divisor, int64_t n_rows, int32_t n_channels,
                                     int64_t* d_starts, int64_t* d_stops, int64_t* d_offsets, void* stream) {
    if (!d_transitions || !d_n_found || !d_starts || !d_stops || !d_offsets || divisor < 1 || n_rows < 0 ||
        n_channels < 0)
        return MS_E_INVALID;
    ms_plan_phase_windows_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_transitions, d_n_found, num_segments, cycles,
                                                                     divisor, n_rows, n_channels, d_starts, d_stops,
                                                                     d_offsets);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

