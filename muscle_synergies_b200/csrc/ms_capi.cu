// Process-wide bookkeeping of the C ABI (include/ms_b200.h).
#include <stdio.h>

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
