#include <immintrin.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
// dst 64-byte aligned chunks written with non-temporal stores (no read-for-ownership of the destination lines)
void copy_nt(uint8_t* dst, const uint8_t* src, size_t n) {
    size_t head = (64 - ((uintptr_t)dst & 63)) & 63;
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; i++) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + 64 * i));
        __m256i b = _mm256_loadu_si256((const __m256i*)(src + 64 * i + 32));
        _mm256_stream_si256((__m256i*)(dst + 64 * i), a);
        _mm256_stream_si256((__m256i*)(dst + 64 * i + 32), b);
    }
    _mm_sfence();
    memcpy(dst + 64 * blocks, src + 64 * blocks, n - 64 * blocks);
}
