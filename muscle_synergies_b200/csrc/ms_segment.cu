// Trial windowing kernels (project/segment.py of the reference).
//
//   ms_segment_kernel       per sample: does a run of >= min_phase_size samples with exactly one / exactly two
//                           loaded force plates start here?  -> two bitmaps; then, in the block that finishes
//                           last, the alternating 1-leg / 2-leg search of _transition_indices
//                           (segment.py:667-755) over the bitmaps and the phase-window row ranges
//   ms_cut_windows_kernel   DeviceData.__getitem__(slice) (user_data.py:727-731) for a batch
//                           of windows over a channel-major array
#include <stdio.h>

#include <string.h>

#include "ms_common.cuh"

extern "C" int64_t ms_transitions_workspace_bytes(int64_t n) { return 2 * ((n + 31) / 32 + 1) * 4 + 16; }  // two bitmaps + a block counter

// loaded(i) = value != 0; NaN != 0 is true, as in numpy (segment.py:716-721)
// ---------------------------------------------------------------------------------------------------
// The whole transition search in ONE launch: every block writes its words of the two "a phase may start here"
// bitmaps; the block that finishes last (a counter in the workspace) then runs the 40 dependent searches with one
// warp - 1024 samples per step, each step one L2 round trip, the bitmaps read past L1 because other blocks wrote
// them - and turns the transitions into the row ranges of the phase (or cycle) windows of up to
// MS_MAX_WINDOW_PLANS devices, so that the gathers can be queued right behind without the host in between.
struct MsPlansArg {
    ms_window_plan p[MS_MAX_WINDOW_PLANS];
    int n;
};

__device__ void ms_plan_windows(const int64_t* __restrict__ transitions, int found, int num_segments, const ms_window_plan& pl) {
    const int per_trecho = pl.cycles ? 2 : 8, span = pl.cycles ? 4 : 1, n_windows = 4 * per_trecho;
    const bool ok = found >= num_segments && num_segments >= 40;
    int64_t off = 0;
    for (int w = 0; w < n_windows; w++) {
        int64_t a = 0, b = 0;
        if (ok) {
            const int j = 10 * (w / per_trecho) + 1 + (w % per_trecho) * span;
            a = transitions[j] / pl.divisor;
            b = (transitions[j + span] - 1) / pl.divisor;
            a = a < 0 ? 0 : (a > pl.n_rows ? pl.n_rows : a);
            b = b < a ? a : (b > pl.n_rows ? pl.n_rows : b);
        }
        pl.d_starts[w] = a;
        pl.d_stops[w] = b;
        pl.d_offsets[w] = off;
        off += (b - a) * pl.n_channels;
    }
    pl.d_offsets[n_windows] = off;
}

__global__ void __launch_bounds__(256)
    ms_segment_kernel(const double* __restrict__ left, const double* __restrict__ right, int64_t n, int min_phase,
                      uint32_t* __restrict__ valid1, uint32_t* __restrict__ valid2, unsigned int* __restrict__ counter,
                      int num_segments, int64_t* __restrict__ transitions, int32_t* __restrict__ loaded,
                      int32_t* __restrict__ n_found, const MsPlansArg plans) {
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31;
    // ---- part 1: this block's words of the bitmaps (segment.py:716-731)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + tid;
    bool v1 = false, v2 = false;
    if (i < n) {
        // correct_activation[ind : ind + min_phase_size].all() - the slice is cut at the end of the signal
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
    const int64_t n_words = (n + 31) / 32;
    if (lane == 0 && (i >> 5) < n_words) {
        valid1[i >> 5] = w1;
        valid2[i >> 5] = w2;
    }
    // ---- who is last?
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;  // wraps to 0: ready for the next call
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- part 2: the alternating search of _transition_indices (segment.py:738-755) by this whole block, 65536
    // samples per round (eight independent loads per thread, read past L1: other blocks wrote the bitmaps) - a search
    // across one gait phase ends in its first round, the rest between two walks over the plates takes a few
    __shared__ long long s_min;
    const long long none = 0x7fffffffffffffffll;
    int64_t cursor = 0;
    int found = 0;
    for (int s = 0; s < num_segments; s++) {
        const uint32_t* bm = (s & 1) ? valid2 : valid1;  // 1 leg, 2 legs, 1 leg, ...
        int64_t base = cursor >> 5;
        long long hit = -1;
        while (base < n_words) {
            if (tid == 0) s_min = none;
            __syncthreads();
            uint32_t word[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int64_t w = base + tid + k * 256;
                word[k] = w < n_words ? __ldcg(bm + w) : 0u;
                if (w == (cursor >> 5)) word[k] &= ~((1u << (cursor & 31)) - 1u);
            }
            long long best = none;
#pragma unroll
            for (int k = 7; k >= 0; k--)
                if (word[k]) best = (long long)((base + tid + k * 256) * 32 + __ffs(word[k]) - 1);  // the nearest wins
            if (best != none) atomicMin(&s_min, best);
            __syncthreads();
            const long long m = s_min;
            __syncthreads();
            if (m != none) {
                hit = m;
                break;
            }
            base += 8 * 256;
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
    __syncthreads();  // the transitions (thread 0's stores) before the plans read them
    // ---- part 3: window plans, one thread each
    if (tid < plans.n) ms_plan_windows(transitions, found, num_segments, plans.p[tid]);
}

extern "C" int ms_segment_trial(const double* d_left_fz, const double* d_right_fz, int64_t n, int32_t min_phase_size,
                                int32_t num_segments, void* d_work, int64_t* d_transitions, int32_t* d_loaded,
                                int32_t* d_n_found, const ms_window_plan* h_plans, int32_t n_plans, void* stream) {
    if (!d_left_fz || !d_right_fz || !d_work || !d_transitions || !d_n_found || n < 0 || num_segments < 0)
        return MS_E_INVALID;
    if (n_plans < 0 || n_plans > MS_MAX_WINDOW_PLANS || (n_plans > 0 && !h_plans)) return MS_E_INVALID;
    MsPlansArg plans;
    memset(&plans, 0, sizeof plans);
    for (int i = 0; i < n_plans; i++) {
        const ms_window_plan& p = h_plans[i];
        if (!p.d_starts || !p.d_stops || !p.d_offsets || p.divisor < 1 || p.n_rows < 0 || p.n_channels < 0) return MS_E_INVALID;
        plans.p[plans.n++] = p;
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* v1 = (uint32_t*)d_work;
    uint32_t* v2 = v1 + ((n + 31) / 32 + 1);
    unsigned int* counter = (unsigned int*)(v2 + ((n + 31) / 32 + 1));
    MS_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const int64_t padded = (n + 31) / 32 * 32;
    const unsigned blocks = (unsigned)(padded > 0 ? (padded + 255) / 256 : 1);
    ms_segment_kernel<<<blocks, 256, 0, st>>>(d_left_fz, d_right_fz, n, min_phase_size, v1, v2, counter, num_segments,
                                             d_transitions, d_loaded, d_n_found, plans);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_find_transitions(const double* d_left_fz, const double* d_right_fz, int64_t n, int32_t min_phase_size,
                                   int32_t num_segments, void* d_work, int64_t* d_transitions, int32_t* d_loaded,
                                   int32_t* d_n_found, void* stream) {
    return ms_segment_trial(d_left_fz, d_right_fz, n, min_phase_size, num_segments, d_work, d_transitions, d_loaded,
                            d_n_found, nullptr, 0, stream);
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

