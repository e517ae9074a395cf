// Shared device helpers for the loader kernels: byte classification of 16-byte vectors
// into ordered bit masks, launch bookkeeping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ms_b200.h"

extern int64_t g_ms_launches;
extern thread_local char g_ms_last_error[256];

#define MS_CUDA_CHECK(expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            snprintf(g_ms_last_error, sizeof g_ms_last_error, "%s: %s", #expr, cudaGetErrorString(_e)); \
            return MS_E_CUDA;                                                                  \
        }                                                                                      \
    } while (0)

#define MS_COUNT_LAUNCH() ((void)__atomic_fetch_add(&g_ms_launches, 1, __ATOMIC_RELAXED))

// ---- per-byte flags on a 32-bit word (4 bytes), result has 0x80 in each matching byte ------
__device__ __forceinline__ uint32_t ms_eq_flags(uint32_t w, uint32_t rep) {
    uint32_t y = w ^ rep;
    uint32_t t = (y & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | y | 0x7f7f7f7fu);
}
// byte >= n (n <= 0x80); bytes >= 0x80 count as >= n
__device__ __forceinline__ uint32_t ms_ge_flags(uint32_t w, uint32_t n_rep) {
    return (((w | 0x80808080u) - n_rep) | w) & 0x80808080u;
}
// gathers flag bits 7,15,23,31 into bits 0..3
__device__ __forceinline__ uint32_t ms_gather4(uint32_t flags) { return (flags * 0x00204081u) >> 28; }

// Four flag words (0x80 in each matching byte) -> ordered 16-bit mask.  Integer dot products do the gathering: the
// flags of a word times {1, 2, 4, 8} (or {16, 32, 64, 128}) add up to 128 x its four mask bits - six instructions
// where multiply-and-shift per word took a dozen.
__device__ __forceinline__ uint32_t ms_mask16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
#ifdef MS_NO_DP4A_MASKS
    return ms_gather4(f0) | (ms_gather4(f1) << 4) | (ms_gather4(f2) << 8) | (ms_gather4(f3) << 12);
#else
    const uint32_t lo = __dp4a(f1, 0x80402010u, __dp4a(f0, 0x08040201u, 0u));
    const uint32_t hi = __dp4a(f3, 0x80402010u, __dp4a(f2, 0x08040201u, 0u));
    return (hi * 256u + lo) >> 7;
#endif
}
// the same for three words: a 12-bit mask
__device__ __forceinline__ uint32_t ms_mask12(uint32_t f0, uint32_t f1, uint32_t f2) {
#ifdef MS_NO_DP4A_MASKS
    return ms_gather4(f0) | (ms_gather4(f1) << 4) | (ms_gather4(f2) << 8);
#else
    const uint32_t lo = __dp4a(f1, 0x80402010u, __dp4a(f0, 0x08040201u, 0u));
    const uint32_t hi = __dp4a(f2, 0x08040201u, 0u);
    return (hi * 256u + lo) >> 7;
#endif
}

// Ordered 16-bit masks (bit i = byte i of the vector) of the structural bytes.
struct MsDelims {
    uint32_t lf, cr, comma;
};

__device__ __forceinline__ MsDelims ms_delims16(uint4 v) {
    MsDelims d;
    d.lf = ms_mask16(ms_eq_flags(v.x, 0x0a0a0a0au), ms_eq_flags(v.y, 0x0a0a0a0au), ms_eq_flags(v.z, 0x0a0a0a0au),
                     ms_eq_flags(v.w, 0x0a0a0a0au));
    d.cr = ms_mask16(ms_eq_flags(v.x, 0x0d0d0d0du), ms_eq_flags(v.y, 0x0d0d0d0du), ms_eq_flags(v.z, 0x0d0d0d0du),
                     ms_eq_flags(v.w, 0x0d0d0d0du));
    d.comma = ms_mask16(ms_eq_flags(v.x, 0x2c2c2c2cu), ms_eq_flags(v.y, 0x2c2c2c2cu), ms_eq_flags(v.z, 0x2c2c2c2cu),
                        ms_eq_flags(v.w, 0x2c2c2c2cu));
    return d;
}

// Row terminator ends: '\n', or '\r' not followed by '\n' (universal newlines, as
// open(filename) in load_csv.py:29 delivers them).  next_is_lf: the byte after the vector.
__device__ __forceinline__ uint32_t ms_term16(uint32_t lf, uint32_t cr, uint32_t next_is_lf) {
    uint32_t lf_next = (lf >> 1) | (next_is_lf << 15);
    return lf | (cr & ~lf_next);
}

// str.strip() whitespace: 9..13, 28..32 (reader.py:126).  Flags for one word.
__device__ __forceinline__ uint32_t ms_strip_space_flags(uint32_t w) {
    uint32_t lo = ms_ge_flags(w, 0x09090909u) & ~ms_ge_flags(w, 0x0e0e0e0eu);
    uint32_t hi = ms_ge_flags(w, 0x1c1c1c1cu) & ~ms_ge_flags(w, 0x21212121u);
    return (lo | hi) & ~w & 0x80808080u;
}
