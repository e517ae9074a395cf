// Trial windowing kernels (project/segment.py of the reference).
//
//   ms_segment_kernel       per sample: does a run of >= min_phase_size samples with exactly one / exactly two
//                           loaded force plates start here?  -> two bitmaps; then, in the block that finishes
//                           last, the alternating 1-leg / 2-leg search of _transition_indices
//                           (segment.py:667-755) as ONE parallel pass over the bitmaps, and the phase-window row ranges
//   ms_cut_windows_kernel   DeviceData.__getitem__(slice) (user_data.py:727-731) for a batch
//                           of windows over a channel-major array
#include <stdio.h>

#include <string.h>

#include "ms_common.cuh"

#define SEG_THREADS 1024
#define SEG_WARPS (SEG_THREADS / 32)
#define SEG_SPB 4                            // stretches of 32 bitmap words (1024 samples, one summary) per block
#define SEG_BLOCK_WORDS (32 * SEG_SPB)
static_assert(SEG_WARPS == 32, "a warp per bitmap word of a stretch");
#define SEG_MAX_HALO 2048     // words of "plates loaded" bits a block may stage past its own (min_phase_size <= 65506)

// two bitmaps, a block counter, one 8-byte summary per block of 1024 samples
extern "C" int64_t ms_transitions_workspace_bytes(int64_t n) { return 2 * ((n + 31) / 32 + 2) * 4 + 16 + ((n + 1023) / 1024 + SEG_SPB) * 8; }

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

// are the m bits that start at bit q of r[w] all set?  (m >= 1)
__device__ __forceinline__ bool ms_bits_all_set(const uint32_t* r, int w, int q, int m) {
    const int avail = 32 - q;
    const uint32_t first = ~(r[w] >> q);
    if (m <= avail) return (first & (m >= 32 ? 0xffffffffu : ((1u << m) - 1u))) == 0u;
    if (first & (avail >= 32 ? 0xffffffffu : ((1u << avail) - 1u))) return false;
    m -= avail;
    w++;
    while (m >= 32) {
        if (r[w] != 0xffffffffu) return false;
        m -= 32;
        w++;
    }
    return m == 0 || (~r[w] & ((1u << m) - 1u)) == 0u;
}

// What the search carries from one stretch of samples to the next, packed in 64 bits: the label (1 = one plate loaded,
// 2 = both) of the first and of the last sample of the stretch at which a phase may start, and the number of label
// changes inside it.  _transition_indices alternates "first start of a 1-leg phase at or after the cursor", "first
// start of a 2-leg phase ...": the two kinds of start exclude each other, so its result is every start whose label
// differs from the label of the start before it - with an imaginary 2-leg start in front of the signal, because the
// search opens with a 1-leg phase.  That is a scan, not a chase.
__device__ __forceinline__ unsigned long long seg_pack(unsigned long long cnt, uint32_t f, uint32_t l) { return (cnt << 4) | (f << 2) | l; }
__device__ __forceinline__ unsigned long long seg_combine(unsigned long long a, unsigned long long b) {
    const uint32_t af = (uint32_t)(a >> 2) & 3u, al = (uint32_t)a & 3u, bf = (uint32_t)(b >> 2) & 3u, bl = (uint32_t)b & 3u;
    const unsigned long long cnt = (a >> 4) + (b >> 4) + ((al && bf && al != bf) ? 1ull : 0ull);
    return seg_pack(cnt, af ? af : bf, bl ? bl : al);
}

// label changes of one bitmap word pair: x = starts of 1-leg phases, y = starts of 2-leg phases; cy = "the last start
// before this word was a 2-leg one" (updated).  The label of the last start before each bit is a carry chain
// (a 2-leg start generates, a sample without a start propagates, a 1-leg start kills): one 64-bit addition.
__device__ __forceinline__ uint32_t seg_changes(uint32_t x, uint32_t y, uint32_t& cy) {
    const uint32_t a = y | ~(x | y);
    const unsigned long long sum = (unsigned long long)a + y + cy;
    const uint32_t fy = (uint32_t)sum ^ a ^ y;  // bit b: the last start before sample b is a 2-leg one
    cy = (uint32_t)(sum >> 32);
    return (x & fy) | (y & ~fy);
}

// the summary {first label, last label, changes} of a run of bitmap words held one per lane, in lane order
__device__ __forceinline__ unsigned long long seg_warp_scan(uint32_t x, uint32_t y, int lane, unsigned long long* inclusive) {
    const uint32_t e = x | y;
    uint32_t f = 0, l = 0;
    if (e) {
        f = ((x >> (__ffs(e) - 1)) & 1u) ? 1u : 2u;
        l = ((x >> (31 - __clz(e))) & 1u) ? 1u : 2u;
    }
    uint32_t cy = f == 2u ? 1u : 0u;  // as if the start before the first one had its label: no change there
    unsigned long long inc = seg_pack((unsigned long long)__popc(seg_changes(x, y, cy)), f, l);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc = seg_combine(o, inc);
    }
    *inclusive = inc;
    return __shfl_sync(0xffffffffu, inc, 31);
}

__global__ void __launch_bounds__(SEG_THREADS, 2)
    ms_segment_kernel(const double* __restrict__ left, const double* __restrict__ right, int64_t n, int min_phase, int halo,
                      uint32_t* __restrict__ valid1, uint32_t* __restrict__ valid2, unsigned int* __restrict__ counter,
                      unsigned long long* __restrict__ summary, long long num_segments, int64_t* __restrict__ transitions,
                      int32_t* __restrict__ loaded, int32_t* __restrict__ n_found, const MsPlansArg plans) {
    extern __shared__ uint32_t s_r[];  // [2][SEG_BLOCK_WORDS + halo]: "exactly one plate loaded", "both loaded" per sample
    __shared__ bool s_last;
    __shared__ uint32_t s_v[2][SEG_BLOCK_WORDS];
    __shared__ unsigned long long s_scan[SEG_WARPS];
    __shared__ unsigned long long s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_words = (n + 31) / 32;
    // ---- part 1: this block's SEG_BLOCK_WORDS words of the bitmaps (segment.py:716-731).  loaded(i) = value != 0;
    // NaN != 0 is true, as in numpy.  correct_activation[ind : ind + min_phase_size].all(): the slice is cut at the end
    // of the signal, so samples past the end count as set.
    const int64_t w_blk = (int64_t)blockIdx.x * SEG_BLOCK_WORDS;  // first bitmap word of this block
    if (halo >= 0) {
        uint32_t* r1 = s_r;
        uint32_t* r2 = s_r + SEG_BLOCK_WORDS + halo;
        for (int wi0 = warp; wi0 < SEG_BLOCK_WORDS + halo; wi0 += SEG_WARPS * SEG_SPB) {
            // SEG_SPB words per warp and round: their loads are in flight together
            double lv[SEG_SPB], rv[SEG_SPB];
#pragma unroll
            for (int q = 0; q < SEG_SPB; q++) {
                const int64_t i = (w_blk + wi0 + q * SEG_WARPS) * 32 + lane;
                const bool in = wi0 + q * SEG_WARPS < SEG_BLOCK_WORDS + halo && i < n;
                lv[q] = in ? left[i] : 0.0;
                rv[q] = in ? right[i] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < SEG_SPB; q++) {
                const int wi = wi0 + q * SEG_WARPS;
                const int64_t i = (w_blk + wi) * 32 + lane;
                const bool a = lv[q] != 0.0, b = rv[q] != 0.0;
                const bool x1 = i < n ? a != b : true, x2 = i < n ? a && b : true;
                const uint32_t b1 = __ballot_sync(0xffffffffu, x1), b2 = __ballot_sync(0xffffffffu, x2);
                if (lane == 0 && wi < SEG_BLOCK_WORDS + halo) {
                    r1[wi] = b1;
                    r2[wi] = b2;
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < SEG_SPB; q++) {
        const int wl = warp + q * SEG_WARPS;  // word of the block
        const int64_t i_own = (w_blk + wl) * 32 + lane;
        bool v1 = false, v2 = false;
        if (i_own < n) {
            if (halo < 0) {
                // a minimum phase longer than a block can stage: every sample walks its own window
                const int64_t end = min(n, i_own + (int64_t)min_phase);
                v1 = v2 = true;
                for (int64_t j = i_own; j < end; j++) {
                    const bool a = left[j] != 0.0, b = right[j] != 0.0;
                    v1 = v1 && (a != b);
                    v2 = v2 && (a && b);
                    if (!v1 && !v2) break;
                }
            } else if (min_phase <= 0) {  // an empty slice is all-true, but look_for only visits true samples
                v1 = (s_r[wl] >> lane) & 1u;
                v2 = (s_r[SEG_BLOCK_WORDS + halo + wl] >> lane) & 1u;
            } else {
                v1 = ms_bits_all_set(s_r, wl, lane, min_phase);
                v2 = ms_bits_all_set(s_r + SEG_BLOCK_WORDS + halo, wl, lane, min_phase);
            }
        }
        const uint32_t w1 = __ballot_sync(0xffffffffu, v1), w2 = __ballot_sync(0xffffffffu, v2);
        if (lane == 0) {
            s_v[0][wl] = w1;
            s_v[1][wl] = w2;
            if (w_blk + wl < n_words) {
                valid1[w_blk + wl] = w1;
                valid2[w_blk + wl] = w2;
            }
        }
    }
    __syncthreads();
    // what the search needs to know about every 1024 samples (32 words): one word
    if (warp < SEG_SPB) {
        unsigned long long inc;
        const unsigned long long total = seg_warp_scan(s_v[0][warp * 32 + lane], s_v[1][warp * 32 + lane], lane, &inc);
        if (lane == 0) summary[(int64_t)blockIdx.x * SEG_SPB + warp] = total;
    }
#if defined(SEG_ABLATE) && SEG_ABLATE == 1
    return;
#endif
    // ---- who is last?
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;  // wraps to 0: ready for the next call
    __syncthreads();
    if (!s_last) return;
#if defined(SEG_ABLATE) && SEG_ABLATE == 2
    return;
#endif
    __threadfence();
    // ---- part 2: _transition_indices (segment.py:738-755) by this whole block, over the blocks' summaries (read past
    // L1: other blocks wrote them): a scan gives every 1024-sample stretch the label in front of it and the rank of
    // its first change; the few stretches that hold one of the first num_segments changes look at their 32 words again.
    unsigned long long before = seg_pack(0, 2, 2);  // the imaginary 2-leg start in front of the signal
    const int64_t n_blocks = (int64_t)gridDim.x * SEG_SPB;  // stretches
    for (int64_t base = 0; base < n_blocks; base += SEG_THREADS) {
        const int64_t blk = base + tid;
        const unsigned long long mine = blk < n_blocks ? __ldcg(summary + blk) : 0ull;
        unsigned long long inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc = seg_combine(o, inc);
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long v = s_scan[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v = seg_combine(o, v);
            }
            s_scan[lane] = v;
            if (lane == 31) s_total = seg_combine(before, v);
        }
        __syncthreads();
        // everything before this thread's stretch: earlier rounds, earlier warps, earlier lanes
        unsigned long long pre = before;
        if (warp > 0) pre = seg_combine(pre, s_scan[warp - 1]);
        {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane > 0) pre = seg_combine(pre, o);
        }
        // A stretch that holds one of the first num_segments changes is looked at again by its whole warp: one word
        // per lane (a thread walking its 32 words alone waited out an L2 round trip per word).
        const uint32_t f = (uint32_t)(mine >> 2) & 3u;
        const bool again = (mine >> 4) + ((f && (pre & 3u) != f) ? 1u : 0u) != 0 && (long long)(pre >> 4) < num_segments;
        uint32_t todo = __ballot_sync(0xffffffffu, again);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const unsigned long long pre_s = __shfl_sync(0xffffffffu, pre, src);
            const int64_t w = (base + (tid - lane) + src) * 32 + lane;
            const uint32_t x = w < n_words ? __ldcg(valid1 + w) : 0u, y = w < n_words ? __ldcg(valid2 + w) : 0u;
            unsigned long long winc;
            seg_warp_scan(x, y, lane, &winc);
            unsigned long long pre_w = pre_s;
            {
                const unsigned long long o = __shfl_up_sync(0xffffffffu, winc, 1);
                if (lane > 0) pre_w = seg_combine(pre_w, o);
            }
            long long rank = (long long)(pre_w >> 4);
            uint32_t cy = (pre_w & 3u) == 2u ? 1u : 0u;
            uint32_t ch = seg_changes(x, y, cy);
            while (ch && rank < num_segments) {
                const int b = __ffs(ch) - 1;
                ch &= ch - 1u;
                const int64_t hit = w * 32 + b;
                transitions[rank] = hit;
                rank++;
            }
        }
        before = s_total;
        __syncthreads();  // s_scan / s_total are rewritten by the next round
        if ((long long)(before >> 4) >= num_segments) break;
    }
    const long long total = (long long)(before >> 4);
    const int found = (int)(total < num_segments ? total : num_segments);
    if (tid == 0) *n_found = found;
    __syncthreads();  // the transitions (global stores of this block) before the plans read them
    if (loaded)
        for (int t = tid; t < found; t += SEG_THREADS) {
            const int64_t hit = transitions[t];
            loaded[t] = (left[hit] != 0.0 ? 1 : 0) | (right[hit] != 0.0 ? 2 : 0);
        }
    // ---- part 3: window plans, one thread each
    if (tid < plans.n) ms_plan_windows(transitions, found, (int)(num_segments > 0x7fffffff ? 0x7fffffff : num_segments), plans.p[tid]);
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
    unsigned long long* summary = (unsigned long long*)(((uintptr_t)(counter + 1) + 15) & ~(uintptr_t)15);
    MS_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const int64_t n_words = (n + 31) / 32;
    const unsigned blocks = (unsigned)(n_words > 0 ? (n_words + SEG_BLOCK_WORDS - 1) / SEG_BLOCK_WORDS : 1);
    // words of per-sample bits a block stages past its own: a window of m samples that starts at the last bit of a word
    // ends (m + 30) / 32 words further on
    int halo = min_phase_size > 1 ? (int)(((int64_t)min_phase_size + 30) / 32) : 0;
    if (halo > SEG_MAX_HALO) halo = -1;  // the kernel walks every window sample by sample instead
    const size_t smem = halo >= 0 ? sizeof(uint32_t) * 2 * (SEG_BLOCK_WORDS + halo) : 0;
    ms_segment_kernel<<<blocks, SEG_THREADS, smem, st>>>(d_left_fz, d_right_fz, n, min_phase_size, halo, v1, v2, counter, summary,
                                                         (long long)num_segments, d_transitions, d_loaded, d_n_found, plans);
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

