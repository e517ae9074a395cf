// Process-wide bookkeeping of the C ABI (include/ms_b200.h).
#include <emmintrin.h>
#include <stdio.h>
#include <string.h>

#include "ms_common.cuh"

int64_t g_ms_launches = 0;                      // bumped atomically: entry points may run on several host threads
thread_local char g_ms_last_error[256] = "";    // per calling thread, like errno

extern "C" const char* ms_last_cuda_error(void) { return g_ms_last_error; }
extern "C" const char* ms_version(void) { return "muscle_synergies_b200 0.2 (sm_100a)"; }

// `rows` runs of `width_bytes` each from device memory (pitch src_pitch_bytes) to host memory (pitch
// dst_pitch_bytes): one 2-D copy on the DMA engine - the used part of a channel-major block whose channels are
// farther apart than they are long.
extern "C" int ms_copy_rows_to_host(void* h_dst, int64_t dst_pitch_bytes, const void* d_src, int64_t src_pitch_bytes,
                                    int64_t width_bytes, int64_t rows, void* stream) {
    if (!h_dst || !d_src || width_bytes < 0 || rows < 0 || dst_pitch_bytes < width_bytes || src_pitch_bytes < width_bytes)
        return MS_E_INVALID;
    if (width_bytes == 0 || rows == 0) return MS_OK;
    MS_CUDA_CHECK(cudaMemcpy2DAsync(h_dst, (size_t)dst_pitch_bytes, d_src, (size_t)src_pitch_bytes, (size_t)width_bytes,
                                    (size_t)rows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return MS_OK;
}
extern "C" int64_t ms_launch_count(void) { return __atomic_load_n(&g_ms_launches, __ATOMIC_RELAXED); }

// Host memory to host memory with non-temporal stores: the destination's cache lines are written without being read
// first, so a copy out of the page cache into a pinned staging buffer moves two bytes of DRAM traffic per byte instead
// of three.  Several ranks reading files at once are bound by exactly that traffic (tools/nt_probe_mp.py: 4 processes x
// 3 threads 44 -> 58 GB/s, 5 x 3: 50 -> 66 GB/s against preadv).  SSE2 only: part of every x86-64.
extern "C" int ms_host_copy_stream(void* h_dst, const void* h_src, int64_t n_bytes) {
    if (n_bytes < 0 || (n_bytes > 0 && (!h_dst || !h_src))) return MS_E_INVALID;
    uint8_t* dst = (uint8_t*)h_dst;
    const uint8_t* src = (const uint8_t*)h_src;
    size_t n = (size_t)n_bytes;
    size_t head = (64 - ((uintptr_t)dst & 63)) & 63;  // up to the destination's next cache line
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head, src += head, n -= head;
    const size_t lines = n / 64;
    for (size_t i = 0; i < lines; i++) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + 64 * i));
        const __m128i b = _mm_loadu_si128((const __m128i*)(src + 64 * i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + 64 * i + 32));
        const __m128i d = _mm_loadu_si128((const __m128i*)(src + 64 * i + 48));
        _mm_stream_si128((__m128i*)(dst + 64 * i), a);
        _mm_stream_si128((__m128i*)(dst + 64 * i + 16), b);
        _mm_stream_si128((__m128i*)(dst + 64 * i + 32), c);
        _mm_stream_si128((__m128i*)(dst + 64 * i + 48), d);
    }
    _mm_sfence();
    memcpy(dst + 64 * lines, src + 64 * lines, n - 64 * lines);
    return MS_OK;
}
