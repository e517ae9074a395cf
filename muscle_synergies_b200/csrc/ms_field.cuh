// Field-level parsing shared by the loader kernels (ms_loader.cu: two-pass path, ms_fused.cu: single pass):
// the per-field float() of reader.py:940-948 on bytes staged in shared memory, and the column-chunk
// schedule of a row group.
#pragma once
#include "ms_common.cuh"
#include "ms_parse_double.cuh"

#ifndef PARSE_THREADS
#define PARSE_THREADS 512
#endif
#define PARSE_WARPS (PARSE_THREADS / 32)
#define PARSE_REGION (MS_TILE_BYTES + MS_MAX_ROW_BYTES)  // bytes staged per CTA
#define PARSE_CHUNK (PARSE_REGION / PARSE_THREADS)        // 112 bytes per thread
#define PARSE_SEGS (PARSE_CHUNK / 16)
static_assert(PARSE_CHUNK % 16 == 0 && PARSE_CHUNK * PARSE_THREADS == PARSE_REGION, "chunking");
#define PARSE_PAD 16  // bytes staged before and after the region
#define PARSE_BYTES_SMEM (PARSE_REGION + 2 * PARSE_PAD)
#define PARSE_NSEG (PARSE_REGION / 16)
#define PARSE_ROWS_CAP 512
#define PARSE_MAX_CHUNKS 32
#define PARSE_SMEM (PARSE_BYTES_SMEM + PARSE_NSEG * 2 + 32 + (PARSE_ROWS_CAP + 1) * 4)

// widest chunk = MS_EXP_UNUM / MS_EXP_UDEN of a warp's fair share of the tile.  A/B on the single-pass kernel (T10 / T127):
// 1/2 +2 %, 5/8 +0.7 %, 3/4 reference, 1/1 -1.3 % / -2.5 %, 5/4 -1.2 %, 3/2 -0.6 %: every chunk but a row's first costs a
// walk over the row's comma masks to its first column, which outweighs the better balance of finer chunks.
#ifndef MS_EXP_UNUM
#define MS_EXP_UNUM 1
#define MS_EXP_UDEN 1
#endif
// Column chunks of DECREASING width, handed out widest first (chunk-major), so that the last items a warp
// can draw are small and the warps finish the tile together (uniform chunks left ~30 % of the warps idle at
// the end of every tile).  The widest chunk is the END of the row: in a Devices row those are the EMG
// columns, the longest fields.  cols[] is descending: chunk j covers columns [cols[j+1], cols[j]).
template <typename T>
__host__ __device__ inline int ms_chunk_table(int groups, int ncols, T* cols, int warps = PARSE_WARPS) {
    const int ideal = (groups * ncols + warps - 1) / warps;  // column-groups per warp
    int u = (ideal * MS_EXP_UNUM + MS_EXP_UDEN - 1) / MS_EXP_UDEN;
    if (u < 2) u = 2;
    int k = 0, col = 0;
    while (col < ncols && k < PARSE_MAX_CHUNKS - 1) {
        cols[k++] = (T)(ncols - col);
        const int rest = ncols - col;
        int step = (rest + 2) / 3;
        if (step > u) step = u;
        if (step < 1) step = 1;
        col += step;
    }
    if (col < ncols) cols[k++] = (T)(ncols - col);  // cap reached: one last chunk takes the rest
    cols[k] = 0;
    return k;
}

// The table depends on the section (its column count) and on the number of row groups in the tile only, so the
// host fills it in for up to PARSE_TAB_GROUPS groups: computed by one thread per tile it was a serial chain of
// ~100 dependent instructions that every CTA waited out on a busy SM (15 % of a CTA's lifetime, measured).
#define PARSE_TAB_GROUPS 8
struct MsSectionsArg {
    ms_section s[MS_MAX_SECTIONS];
    int n;
    uint16_t chunk_tab[MS_MAX_SECTIONS][PARSE_TAB_GROUPS][PARSE_MAX_CHUNKS + 1];
    uint8_t chunk_cnt[MS_MAX_SECTIONS][PARSE_TAB_GROUPS];  // 0: not tabulated (too many columns for 16 bits)
};

__device__ __forceinline__ bool ms_is_delim(unsigned c) { return c == ',' || c == '\n' || c == '\r'; }

// Everything the inline path of ms_parse_next does not take: finds the extent of the field that
// starts at reg[fs], parses it with the general parser and records an error.  Handles fields that
// start with '"' (excel dialect of csv.reader, load_csv.py:30): the content runs to the closing
// quote, "" is a literal quote and - csv is not strict - text after the closing quote is appended
// up to the next delimiter; an empty content is an empty field (None -> NaN in the reference).
// Returns the bits; *pend = offset of the delimiter that ends the field.
#define MS_QUOTED_MAX 64
static __device__ __noinline__ uint64_t ms_parse_slow_call(const uint8_t* __restrict__ reg, int fs, int* pend,
                                                    unsigned long long* status, int64_t t0) {
    uint64_t bits = MS_NAN_BITS;
    int st = MS_PARSE_OK;
    int q = fs;
    if (reg[fs] == '"') {
        uint8_t buf[MS_QUOTED_MAX];
        int n = 0;
        bool too_long = false;
        q = fs + 1;
        for (;;) {
            if (q >= PARSE_REGION) break;  // unbalanced quote: stop at the end of the staged bytes
            unsigned c = reg[q];
            if (c == '"') {
                if (reg[q + 1] != '"') {
                    q++;
                    break;
                }
                q++;
            }
            if (n < MS_QUOTED_MAX)
                buf[n++] = (uint8_t)c;
            else
                too_long = true;
            q++;
        }
        while (q < PARSE_REGION && !ms_is_delim(reg[q])) {
            if (n < MS_QUOTED_MAX)
                buf[n++] = reg[q];
            else
                too_long = true;
            q++;
        }
        if (too_long)
            st = MS_PARSE_BAD;
        else if (n > 0)
            st = ms_parse_field(buf, buf + n, &bits);
    } else {
        while (!ms_is_delim(reg[q])) q++;
        st = ms_parse_field(reg + fs, reg + q, &bits);
    }
    if (st != MS_PARSE_OK) {
        bits = MS_NAN_BITS;
        atomicMin(status, ((unsigned long long)(t0 + fs) << 3) |
                              (st == MS_PARSE_NONASCII ? MS_ERR_KIND_NON_ASCII : MS_ERR_KIND_BAD_FLOAT));
    }
    *pend = q;
    return bits;
}

// Four bytes at an arbitrary offset of the staged region (little endian): two aligned word
// loads and a funnel shift.
__device__ __forceinline__ uint32_t ms_load4(const uint8_t* __restrict__ reg, int p) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(reg) + (p >> 2);
    return __funnelshift_r(w[0], w[1], (p & 3) << 3);
}

static __constant__ uint32_t ms_pow10_u32[5] = {1u, 10u, 100u, 1000u, 10000u};

// Value of four ASCII-digit bytes already reduced to 0..9 (first character most significant).
__device__ __forceinline__ uint32_t ms_digits4(uint32_t t) {
    const uint32_t v = (t * 10u + (t >> 8)) & 0x00ff00ffu;  // two-digit values in bytes 0 and 2
    return (v & 0xffu) * 100u + (v >> 16);
}

// Parses the field that starts at *pp and advances *pp past its delimiter.
//   inline path: [-]digits[.digits][(e|E)[+-]d{1,3}] with <= 19 digits, mantissa <= 2^53 and a
//                decimal exponent in Clinger's exact range -> one IEEE multiply or divide.
//                Digits are converted four at a time (SWAR on one 32-bit word).  (A 32-bit
//                accumulator variant was measured 3 % slower: more selects than it saves.)
//   anything else, quoted fields included: ms_parse_slow_call decides (and reports errors).
// Returns true when the delimiter ended the row.
__device__ __forceinline__ bool ms_parse_next(const uint8_t* __restrict__ reg, int* pp, uint64_t* bits_out,
                                              unsigned long long* status, int64_t t0) {
    const int fs = *pp;
    int p = fs;
    unsigned c = reg[p];
    uint64_t bits = MS_NAN_BITS;
    if (!ms_is_delim(c)) {
        // the sign costs no branch: one byte decides where the digits' first word starts
        const bool neg = c == '-';
        const uint64_t sign = neg ? 0x8000000000000000ull : 0ull;
        p += neg ? 1 : 0;
        uint32_t x = ms_load4(reg, p);
        uint64_t acc = 0;
        int ndig = 0, nfrac = 0;
        bool dot = false;
#ifndef MS_NO_SHAPE_SHORTCUTS
        // Straight-line paths for the two shapes that fill Vicon exports (same results as the loop below).  Lanes
        // of a warp parse the same column, so what matters is that they all stay on ONE path: a lane that leaves
        // it makes the whole warp execute the general loop as well (which is why the first path runs to nine
        // digits: stopping at eight left one lane in most warps behind and was slower than no shortcut at all).
        if ((x & 0xffffu) == 0x2e30u) {
            // "0." and 4 to 9 fraction digits, then the delimiter - every EMG sample ("0.0123456", "0.00123457",
            // "0.000123457"): both digit words at once, 32-bit arithmetic, one exact scaling
            const int pf = p + 2;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(reg) + (pf >> 2);
            const int sh = (pf & 3) << 3;
            const uint32_t w1 = w[1];
            const uint32_t y0 = __funnelshift_r(w[0], w1, sh), y1 = __funnelshift_r(w1, w[2], sh);
            const uint32_t u0 = y0 ^ 0x30303030u, u1 = y1 ^ 0x30303030u;
            const uint32_t n0 = ((u0 + 0x76767676u) | u0) & 0x80808080u;
            const uint32_t n1 = ((u1 + 0x76767676u) | u1) & 0x80808080u;
            if (n0 == 0) {
                const int j1 = n1 ? (__ffs(n1) - 1) >> 3 : 4;  // fraction digits in the second word
                uint32_t frac = ms_digits4(u0) * ms_pow10_u32[j1] + ms_digits4((uint32_t)((uint64_t)u1 << ((4 - j1) << 3)));
                int nf = 4 + j1;
                unsigned cj = (y1 >> (j1 << 3)) & 0xffu;  // meaningless when j1 == 4
                if (j1 == 4) {
                    // a ninth digit (%.6g just above 1e-4), then the delimiter
                    cj = reg[pf + 8];
                    if (cj - '0' <= 9u) {
                        frac = frac * 10u + (cj - '0');  // < 10^9
                        nf = 9;
                        cj = reg[pf + 9];
                    }
                }
                if (ms_is_delim(cj)) {
                    *bits_out = sign | ms_double_to_bits(ms_div_pow10_u32(frac, nf));
                    *pp = pf + nf + 1;
                    return cj != ',';
                }
            }
        } else {
            const uint32_t t = x ^ 0x30303030u;
            const uint32_t nd = ((t + 0x76767676u) | t) & 0x80808080u;
            const int j0 = (__ffs(nd) - 1) >> 3;  // digits before the first other byte; nd == 0 gives -1
            if (j0 > 0) {
                const unsigned c0 = (x >> (j0 << 3)) & 0xffu;
                const uint32_t ip = ms_digits4(t << ((4 - j0) << 3));  // the 1-3 leading digits
                if (ms_is_delim(c0)) {
                    // an integer - unloaded force plates ("0"), sub-frames, CoP values
                    *bits_out = sign | ms_double_to_bits((double)ip);
                    *pp = p + j0 + 1;
                    return c0 != ',';
                }
            }
        }
#endif
        // One loop body for whole and partial words (a full word is the j == 4 case), so the lanes of a
        // warp - same column, different digit counts - stay on one path and differ only in trip count.
        for (;;) {
            const uint32_t t = x ^ 0x30303030u;
            const uint32_t nd = ((t + 0x76767676u) | t) & 0x80808080u;  // bytes that are not digits
            const int j = nd ? (__ffs(nd) - 1) >> 3 : 4;                // leading digits in this word: 0..4
            acc = acc * ms_pow10_u32[j] + ms_digits4((uint32_t)((uint64_t)t << ((4 - j) << 3)));
            ndig += j;
            nfrac += dot ? j : 0;
            c = (uint32_t)((uint64_t)x >> (j << 3)) & 0xffu;  // 0 when j == 4
            const bool isdot = c == '.' && !dot;
            p += j + (isdot ? 1 : 0);
            if (j == 4 || isdot) {
                dot = dot || isdot;
                x = ms_load4(reg, p);
                continue;
            }
            break;
        }
        // c = reg[p]: the first byte that is neither a digit nor the (first) decimal point
        int ex = -nfrac;
        bool ok = (unsigned)(ndig - 1) <= 14u;  // 1..15 digits: acc < 10^15 < 2^53
        if (ok && (c | 0x20u) == 'e') {
            // exponent: at most three digits
            const uint8_t* r = reg + p + 1;
            unsigned cc = *r;
            bool eneg = false;
            if (cc == '-' || cc == '+') {
                eneg = cc == '-';
                cc = *++r;
            }
            int ev = 0, nd = 0;
            while (cc - '0' <= 9u && nd < 4) {
                ev = ev * 10 + (int)(cc - '0');
                nd++;
                cc = *++r;
            }
            if (nd >= 1 && nd <= 3 && ms_is_delim(cc)) {
                ex += eneg ? -ev : ev;
                p = (int)(r - reg);
                c = cc;
            } else {
                ok = false;
            }
        }
        if (ok && ms_is_delim(c) && (unsigned)(ex + 22) <= 44u) {
            double v;
            if (ex <= 0 && (acc >> 32) == 0) {
                // integers (ex == 0) take the same path as fractions: a / 1 is exact there too
                v = ms_div_pow10_u32((uint32_t)acc, -ex);  // exact, division-free (exhaustively verified)
            } else if (ex >= 0) {
                v = (double)acc * ms_pow10_double[ex];
            } else {
                v = (double)acc / ms_pow10_double[-ex];
            }
            bits = sign | ms_double_to_bits(v);
        } else {
            // everything else (quoted fields too): out of line
            bits = ms_parse_slow_call(reg, fs, &p, status, t0);
            c = reg[p];
        }
    }
    *bits_out = bits;
    *pp = p + 1;
    return c != ',';
}


