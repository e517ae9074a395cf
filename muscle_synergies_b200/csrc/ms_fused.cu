// Single-pass Vicon Nexus CSV loader for sm_100a: ms_load_kernel.
//
// One launch takes the CSV bytes to the channel-major float64 blocks of both sections.  Every byte is read from
// HBM once (one TMA bulk copy per tile into shared memory) and classified once; every kept double is written
// once.  What the two-pass path (ms_loader.cu: ms_scan -> host -> ms_parse) settled between its passes is
// settled inside the launch:
//
//   * the csv row index of a tile's first row: a decoupled look-back over per-tile words that carry
//     {blank rows so far, rows since the last blank row} - which is all a row needs to know about the rows
//     before it: its section (reader.py:971-987 switches section at a blank row) and its index in the section
//     (lines 0-4 are the header, reader.py:250-835; line i >= 5 is data row i - 5);
//   * the column count of a section: the tile that owns the section's coordinates line (line 3) counts its
//     fields the way CoordinatesState does (strip, drop trailing blanks: reader.py:772-783), carves the
//     section's block out of the caller's arena and publishes a descriptor the later tiles wait for;
//   * the header text of both sections is copied out for the host (names, units, frequencies are host work).
//
// The kernel only promises results for well-formed input - two sections, at most one trailing blank row, no
// quotes, rows that fit the tile overhang, outputs that fit the arena.  Anything else raises a flag in
// ms_load_result.flags and the caller runs the two-pass path, which reproduces the reference's behaviour
// (and its errors) in full.  Replaces, per data row, reader.py:886-948, aggregator.py:96-124, 229-241,
// user_data.py:391-396.
#include <stdio.h>

#include "ms_field.cuh"

#ifndef FUSED_THREADS
#define FUSED_THREADS 512
#endif
// Build-time variants kept for A/B runs (tools/ab_fused.py); the defaults are what measured best on T10 / T127:
//   FUSED_LATE_COMMAS  comma masks found after the tile's aggregate is published (P1c) instead of in the first sweep
//   FUSED_WALK64       the walk to a column chunk's first field reads four segments' delimiter masks per step
//   FUSED_PIN_LUT      the field LUT's shared-window address kept in one register across the field loop
//   FUSED_DYNAMIC_ITEMS  (off) the warps of a tile draw (row group, column chunk) items from a shared counter, chunk
//                      widths from a per-section table, instead of taking equal static shares of the tile's (row group,
//                      column) pairs; A/B: the static shares are 4 % faster (T10 0.580 -> 0.555 ms, T127 0.160 -> 0.154 ms)
#ifndef FUSED_AB_BASE
#ifndef FUSED_EARLY_COMMAS
#define FUSED_LATE_COMMAS
#endif
#define FUSED_WALK64
#define FUSED_PIN_LUT
#endif
#define FUSED_WARPS (FUSED_THREADS / 32)
#ifndef FUSED_MAX_TILE
#define FUSED_MAX_TILE MS_TILE_BYTES  // largest tile (and the default)
#endif
#ifndef FUSED_MIN_CTAS
#define FUSED_MIN_CTAS 3
#endif
#define FUSED_MAX_REGION (FUSED_MAX_TILE + MS_MAX_ROW_BYTES)
#define FUSED_MAX_NSEG (FUSED_MAX_REGION / 16)
#define FUSED_PAD 16
#ifndef FUSED_LB_WINDOWS
#define FUSED_LB_WINDOWS 1  // 32-tile windows of look-back words fetched per round trip (A/B: 2 and 4 were 1.5 % and 3.5 % slower)
#endif
#define FUSED_BYTES_SMEM (FUSED_MAX_REGION + 2 * FUSED_PAD)
#define FUSED_ROWS_CAP 1024  // rows that may start in one tile
#define FUSED_HIT_WORDS (FUSED_MAX_NSEG / 32 + 2)
// dynamic shared memory: staged bytes | comma masks | terminator masks | hit bitmap | row starts
#define FUSED_OFF_CMASK FUSED_BYTES_SMEM
#define FUSED_OFF_TMASK (FUSED_OFF_CMASK + FUSED_MAX_NSEG * 2 + 16)
#define FUSED_OFF_HIT (FUSED_OFF_TMASK + FUSED_MAX_NSEG * 2)
#define FUSED_OFF_ROWS (FUSED_OFF_HIT + FUSED_HIT_WORDS * 4)
#define FUSED_OFF_LUT ((FUSED_OFF_ROWS + (FUSED_ROWS_CAP + 2) * 2 + 15) / 16 * 16)
#define FUSED_SMEM (FUSED_OFF_LUT + 512 + 256)  // + 256: the LUT starts at the next 256-byte boundary of the shared window
static_assert(FUSED_OFF_CMASK % 16 == 0 && FUSED_OFF_TMASK % 16 == 0 && FUSED_OFF_HIT % 8 == 0 && FUSED_OFF_ROWS % 4 == 0,
              "shared memory layout");
static_assert(FUSED_MAX_REGION + FUSED_PAD < 65536, "row starts are 16-bit");
static_assert(sizeof(double2) * 16 + sizeof(uint4) * 16 == 512, "MsFieldLut is 512 bytes");

// ---- workspace: everything the tiles tell each other (zeroed before the launch) -----------------------------
struct MsSecDesc {
    uint32_t ready;    // 1 once the fields below are valid (written last, release)
    int32_t num_cols;  // fields parsed per row
    int32_t n_keep;    // channels stored: num_cols - 2
    int32_t pad;
    long long stride;      // elements between channels
    long long out_offset;  // element offset of the block in the arena
#ifdef FUSED_DYNAMIC_ITEMS
    uint16_t chunk_tab[PARSE_TAB_GROUPS][PARSE_MAX_CHUNKS + 1];
    uint8_t chunk_cnt[PARSE_TAB_GROUPS];
#endif
};
struct MsFusedWs {
    uint32_t ticket;  // tile ids are handed out in launch order: a tile's predecessors are running or done
    uint32_t pad[15];
    MsSecDesc desc[2];
    // look-back words follow (8 bytes per tile)
};
#define FUSED_LB_OFFSET ((int64_t)((sizeof(MsFusedWs) + 255) / 256 * 256))

// look-back word: [63:62] state, [61:60] blank rows (saturating at 3), [59:0] rows since the last blank row
#define LB_INVALID 0ull
#define LB_AGGREGATE 1ull
#define LB_INCLUSIVE 2ull
struct LbVal {
    uint32_t nb;
    unsigned long long dist;
};
__device__ __forceinline__ unsigned long long lb_pack(unsigned long long state, LbVal v) {
    return (state << 62) | ((unsigned long long)(v.nb > 3u ? 3u : v.nb) << 60) | (v.dist & ((1ull << 60) - 1ull));
}
__device__ __forceinline__ LbVal lb_unpack(unsigned long long w) {
    LbVal v;
    v.nb = (uint32_t)(w >> 60) & 3u;
    v.dist = w & ((1ull << 60) - 1ull);
    return v;
}
// state after `left` followed by `right`
__device__ __forceinline__ LbVal lb_combine(LbVal left, LbVal right) {
    LbVal r;
    if (right.nb) {
        r.nb = min(3u, left.nb + right.nb);
        r.dist = right.dist;
    } else {
        r.nb = left.nb;
        r.dist = left.dist + right.dist;
    }
    return r;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
// 0x80 in every byte of w that is a comma.  Exact for words of ASCII bytes; a byte >= 0x80 may spoil its neighbour's
// flag, which is harmless: such a segment raises MS_LOAD_HIGH_BYTES and the caller discards this kernel's result.
__device__ __forceinline__ uint32_t ms_comma_flags(uint32_t w) {
#ifdef FUSED_EXACT_COMMAS
    return ms_eq_flags(w, 0x2c2c2c2cu);
#else
    return ~((w ^ 0x2c2c2c2cu) + 0x7f7f7f7fu) & 0x80808080u;
#endif
}
// barrier among the worker warps only (warp 0 is away looking back)
__device__ __forceinline__ void ms_bar_workers(int n) { asm volatile("bar.sync 1, %0;\n" ::"r"(n) : "memory"); }
// hand-off of the tile's aggregate: the first worker warp arrives on named barrier 2 after it wrote the aggregate to shared
// memory, warp 0 waits on it (producer / consumer use of bar.arrive + bar.sync: the barrier orders the shared-memory writes)
__device__ __forceinline__ void ms_bar_agg_arrive() { asm volatile("bar.arrive 2, 64;\n" ::: "memory"); }
__device__ __forceinline__ void ms_bar_agg_wait() { asm volatile("bar.sync 2, 64;\n" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

struct MsFusedArgs {
    double* arena;
    long long arena_elems;
    long long cap_rows[2];
    int tile_bytes;
    int region_bytes;  // tile_bytes + the overhang staged past the tile
    long long n_tiles;
};

// ---- the field parser of the single-pass kernel ----------------------------------------------------------------
// A field's end is known before it is read (the delimiter masks of P1), so the common shapes - an optional '-',
// digits, at most one '.', 1 to 9 digits in all, at most 12 bytes: every number a Vicon export prints without an
// exponent - take ONE straight-line path whatever their length, and the lanes of a warp (32 rows, same column)
// do not diverge on digit counts:
//   * the 12 bytes that END at the delimiter are loaded; the bytes of earlier fields in front are masked off;
//   * chars before the '.' are moved one place towards the end (a second view of the same words, shifted by a
//     byte), which removes the '.': the digits are then the last `ndig` chars, zero padded in front;
//   * three 4-digit groups by integer dot products (IDP4A), one exact scaling by 10^-nfrac (ms_div_pow10_u32).
// Everything else - exponents, blanks, '+', inf/nan, long mantissas, quotes, errors - goes to ms_parse_next.
struct MsFieldLut {
    uint4 last[16];     // last[n]: byte masks (three words, first char in the low byte) of the last n chars of 12
    double2 pow10[16];  // {10^k, RN(10^-k)}
};
__device__ __forceinline__ void ms_field_lut_init(MsFieldLut* lut, int tid) {
    if (tid < 16) {
        uint32_t w[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint32_t m = 0;
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (4 * k + b >= 12 - tid) m |= 0xffu << (8 * b);
            w[k] = m;
        }
        lut->last[tid] = make_uint4(w[0], w[1], w[2], 0u);
        lut->pow10[tid] = make_double2(ms_pow10_double[tid], ms_rpow10_double[tid]);
    }
}
// four chars already reduced to 0..9, first char in the low byte -> their value
__device__ __forceinline__ uint32_t ms_digits4_dp(uint32_t g) {
    const uint32_t hi = __dp4a(g, 0x0000010au, 0u);  // 10 c0 + c1
    return __dp4a(g, 0x010a0000u, hi * 100u);         // + 10 c2 + c3
}

// Shared-memory loads by 32-bit shared-window address: one base register for the whole loop instead of a generic
// pointer that is converted again at every use.  volatile: they stay ordered with the barriers around them.
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_d2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}

// The field that starts at byte p of the staged region: its end (the delimiter's position) from the delimiter masks,
// and - when it has one of the common shapes - its value.  Returns false for every other field: the caller queues it
// for the general parser (ms_slow_fields), so that this path has no branches the lanes of a warp could part on.
// sreg / sdm / slut: shared-window addresses of reg[0], the delimiter masks and the MsFieldLut.
__device__ __forceinline__ bool ms_field_fast(uint32_t sreg, uint32_t sdm, uint32_t slut, int p, int* e_out, uint64_t* bits_out) {
    uint32_t sd = sdm + ((p >> 4) << 1);
    uint32_t dm = ((lds_u16(sd + 2) << 16) | lds_u16(sd)) >> (p & 15);
    int L = __ffs(dm) - 1;  // bytes before the field's delimiter
    if (dm == 0u) {
        // none within reach (a field of 17 .. 32 bytes or more): walk the masks; the row's line end stops the walk
        int seg = (p >> 4) + 2;
        uint32_t m;
        while ((m = lds_u16(sdm + (seg << 1))) == 0u) seg++;
        L = (seg << 4) + __ffs(m) - 1 - p;
    }
    const int e = p + L;
    *e_out = e;
    // the 12 bytes that end at the delimiter
    const int a = e - 12;
    const uint32_t sw = sreg + (a & ~3);
    const int sh = (a & 3) << 3;
    const uint32_t a0 = lds_u32(sw), a1 = lds_u32(sw + 4), a2 = lds_u32(sw + 8), a3 = lds_u32(sw + 12);
    const uint32_t t0w = __funnelshift_r(a0, a1, sh) ^ 0x30303030u, t1w = __funnelshift_r(a1, a2, sh) ^ 0x30303030u,
                   t2w = __funnelshift_r(a2, a3, sh) ^ 0x30303030u;
    const uint4 in = lds_v4(slut | ((L & 15) << 4));  // the chars of this field
    // chars of the field that are not digits -> bit j of M (j = 0: 12 bytes before the delimiter)
    const uint32_t n0 = ((t0w + 0x76767676u) | t0w) & in.x & 0x80808080u, n1 = ((t1w + 0x76767676u) | t1w) & in.y & 0x80808080u,
                   n2 = ((t2w + 0x76767676u) | t2w) & in.z & 0x80808080u;
    const uint32_t M = ms_mask12(n0, n1, n2);
    const unsigned c0 = lds_u8(sreg + p);
    const uint32_t neg = c0 == '-' ? 1u : 0u;
    const uint32_t Md = M & ~(neg << ((12 - L) & 31));  // what is left must be the point
    const int dotj = 31 - __clz(Md);                      // -1: no point
    const uint32_t hasdot = Md != 0 ? 1u : 0u;
    const unsigned cd = lds_u8(sreg + a + (dotj & 15));
    const int nfrac = hasdot ? 11 - dotj : 0;
    const int ndig = L - (int)neg - (int)hasdot;
    // the point taken out: chars before it from the view shifted by one byte
    const uint4 keep = lds_v4(slut | ((hasdot ? nfrac & 15 : 12) << 4));
    const uint4 dig = lds_v4(slut | ((ndig & 15) << 4));
    const uint32_t h0 = t0w << 8, h1 = __funnelshift_l(t0w, t1w, 8), h2 = __funnelshift_l(t1w, t2w, 8);
    const uint32_t g0 = ((t0w & keep.x) | (h0 & ~keep.x)) & dig.x, g1 = ((t1w & keep.y) | (h1 & ~keep.y)) & dig.y,
                   g2 = ((t2w & keep.z) | (h2 & ~keep.z)) & dig.z;
    const uint32_t v0 = ms_digits4_dp(g0);
    const uint32_t N = (v0 * 10000u + ms_digits4_dp(g1)) * 10000u + ms_digits4_dp(g2);
    // 1 to 12 digits whose value fits 32 bits (leading zeros are free: "-0.000944047" has ten digits)
    const bool fast = (unsigned)(L - 1) <= 11u && (Md & (Md - 1u)) == 0u && ndig >= 1 && v0 <= 41u && (!hasdot || cd == '.');
    const double2 pw = lds_d2((slut | ((nfrac & 15) << 4)) + 256);
    const double an = (double)N;
    const double q0 = __dmul_rn(an, pw.y);
    const double r = __fma_rn(-q0, pw.x, an);
    const uint64_t bits = ms_double_to_bits(__fma_rn(r, pw.y, q0)) | ((uint64_t)neg << 63);
    // an empty field is None -> NaN (reader.py:944-948, user_data.py:396)
    *bits_out = L == 0 ? MS_NAN_BITS : bits;
    return fast || L == 0;
}

// The fields ms_field_fast left: exponents, blanks around numbers, '+', inf / nan, long mantissas, Unicode digits,
// and everything float() rejects.  Queued by the lanes that met them ({p, e, arena index}), parsed here by the whole
// block after the tile's items - densely, instead of by one or two lanes of a warp in the middle of its column.
#define FUSED_SLOW_CAP (FUSED_MAX_NSEG * 2 / 8)  // the queue lives where the terminator masks were
__device__ __forceinline__ void ms_slow_fields(const uint8_t* __restrict__ reg, const uint2* __restrict__ queue, int n, double* __restrict__ arena,
                                               unsigned long long* status, long long t0, int tid) {
    for (int i = tid; i < n; i += FUSED_THREADS) {
        const uint2 q = queue[i];
        const int p = (int)(q.x & 0xffffu), e = (int)(q.x >> 16);
        uint64_t bits = MS_NAN_BITS;
        const int st = ms_parse_field(reg + p, reg + e, &bits);
        if (st != MS_PARSE_OK) {
            bits = MS_NAN_BITS;
            atomicMin(status, ((unsigned long long)(t0 + p) << 3) |
                                  (st == MS_PARSE_NONASCII ? MS_ERR_KIND_NON_ASCII : MS_ERR_KIND_BAD_FLOAT));
        }
        if (q.y != 0xffffffffu) arena[q.y] = ms_bits_to_double(bits);
    }
}

__global__ void __launch_bounds__(FUSED_THREADS, FUSED_MIN_CTAS)
    ms_load_kernel(const uint8_t* __restrict__ src, long long n, MsFusedWs* __restrict__ ws, const MsFusedArgs args,
                   ms_load_result* __restrict__ res, uint8_t* __restrict__ peek) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const reg = smem_raw + FUSED_PAD;  // reg[i] = src[t0 + i]
    uint16_t* const cmask = reinterpret_cast<uint16_t*>(smem_raw + FUSED_OFF_CMASK);  // commas per 16-byte segment
    uint16_t* const tmask = reinterpret_cast<uint16_t*>(smem_raw + FUSED_OFF_TMASK);  // terminator ends, hit segments only
    uint32_t* const hitmap = reinterpret_cast<uint32_t*>(smem_raw + FUSED_OFF_HIT);   // segments with a byte < 0x23 or >= 0x80
    uint16_t* const row_start = reinterpret_cast<uint16_t*>(smem_raw + FUSED_OFF_ROWS);  // [L]: first byte of local row L
    __shared__ int s_warp_terms[FUSED_WARPS];
    __shared__ int s_lt_end, s_nblank, s_blank_lo, s_blank_hi, s_quotes, s_stop;
#ifdef FUSED_DYNAMIC_ITEMS
    __shared__ int s_next_item, s_nchunks;
    __shared__ int s_chunk_col[PARSE_MAX_CHUNKS + 1];
    __shared__ uint32_t s_inv_groups;
#endif
    __shared__ uint32_t s_tile, s_flags, s_pre_nb, s_agg_nb, s_fatal;
    __shared__ unsigned long long s_agg_dist;
    __shared__ unsigned long long s_pre_dist;
    __shared__ int s_q, s_commas;
    __shared__ __align__(8) unsigned long long s_stage_bar;
    // the field LUT on a 256-byte boundary of the shared window: an entry's address is then `base | index << 4`, one LOP3
    const uint32_t lut_off = ((((uint32_t)__cvta_generic_to_shared(smem_raw) + FUSED_OFF_LUT + 255u) & ~255u) -
                              (uint32_t)__cvta_generic_to_shared(smem_raw));
    MsFieldLut* const lut_p = reinterpret_cast<MsFieldLut*>(smem_raw + lut_off);
    __shared__ int s_slow_n;
    __shared__ __align__(16) MsSecDesc s_desc;  // the descriptor of the section this tile starts in, fetched by warp 0
    __shared__ int s_desc_sec;                  // which section s_desc describes; -1: none (not published at the time)
    uint2* const slow_queue = reinterpret_cast<uint2*>(smem_raw + FUSED_OFF_TMASK);  // P4 only: the terminator masks are done with

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* const lb = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(ws) + FUSED_LB_OFFSET);

    ms_field_lut_init(lut_p, tid);
    if (tid == 0) {
        // Which tile this block gets is decided by a ticket (below), one L2 round trip away.  Meanwhile ask L2 for the
        // tile its block index names: tickets only permute tiles among blocks that start together, so this is the
        // tile some block is about to stage.
        const long long pf0 = (long long)blockIdx.x * args.tile_bytes;
        const long long pfb = min((long long)args.region_bytes, n - pf0) & ~15ll;
        if (pfb > 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src + pf0), "r"((uint32_t)pfb) : "memory");
        s_tile = atomicAdd(&ws->ticket, 1u);
        s_lt_end = -1;
        s_nblank = 0;
        s_blank_lo = 0x7fffffff;
        s_blank_hi = -1;
        s_quotes = 0;
        s_flags = 0;
        s_stop = 0;
        s_fatal = 0;
        s_slow_n = 0;
        s_desc_sec = -1;
    }
    __syncthreads();
    const long long tile = (long long)s_tile;
#if defined(FUSED_ABLATE) && FUSED_ABLATE == 1
    return;  // launch + ticket only
#endif
    const int tile_bytes = args.tile_bytes, region = args.region_bytes;
    const long long t0 = tile * (long long)tile_bytes;

    const int tile_len = (int)min((long long)tile_bytes, n - t0);
    const int nseg = region >> 4;

    // ---- P0. stage [t0 - 16, t0 + region + 16) in shared memory; beyond either end of the file: '\n'.
    // One bulk asynchronous copy (cp.async.bulk: the TMA engine, SASS UBLKCP) issued by one thread moves every whole
    // 16-byte chunk inside the buffer and completes on an mbarrier; chunks at the edges of the file are filled by hand.
    const long long off0 = t0 - FUSED_PAD;
    const int n_chunks_smem = (region + 2 * FUSED_PAD) >> 4;
    const int lo_chunk = off0 < 0 ? (int)((-off0) >> 4) : 0;
    const long long whole = (n >> 4) - (off0 >> 4);  // chunks that end at or before byte n
    const int hi_chunk = (int)max((long long)lo_chunk, min((long long)n_chunks_smem, whole));
    if (tid == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_stage_bar);
        const uint32_t bytes = (uint32_t)(hi_chunk - lo_chunk) * 16u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
        if (bytes) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_raw + lo_chunk * 16);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                         "l"(src + off0 + (long long)lo_chunk * 16), "r"(bytes), "r"(bar)
                         : "memory");
        }
    }
    const bool by_hand = !(lo_chunk == 0 && hi_chunk == n_chunks_smem);  // the usual tile: nothing by hand
    for (int i = tid; by_hand && i < n_chunks_smem; i += FUSED_THREADS) {
        if (i >= lo_chunk && i < hi_chunk) continue;
        const long long off = off0 + (long long)i * 16;
        uint4 v = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
        if (off >= 0 && off < n) {
            // the chunk that holds byte n: the allocation is readable up to n rounded up to 16 (ABI requirement)
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(src + off));
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
            uint32_t w[4];
            const int valid = (int)(n - off);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int have = valid - 4 * k;
                if (have >= 4)
                    w[k] = rw[k];
                else if (have > 0) {
                    const uint32_t m = (1u << (8 * have)) - 1u;
                    w[k] = (rw[k] & m) | (0x0a0a0a0au & ~m);
                } else
                    w[k] = 0x0a0a0a0au;
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        *reinterpret_cast<uint4*>(smem_raw + i * 16) = v;
    }
    if (tid < 2) hitmap[(nseg >> 5) + tid] = 0;  // the words a thread's hit window may reach past the last segment
    __syncthreads();  // mbarrier init and the hand-filled chunks are visible

    // From here to P4 the block works in two parts.  Warp 0 looks back: what it needs - the words of the tiles before
    // this one - does not depend on this tile's bytes, so it starts at once and its L2 round trips run under the
    // staging and P1/P2 of the other warps instead of after them (measured: a third of a block's lifetime went by
    // with all sixteen warps parked behind the look-back).  Warps 1.. (the "workers", synchronised among themselves
    // on named barrier 1) find this tile's rows and publish its aggregate as soon as it is known.
#if defined(FUSED_ABLATE) && FUSED_ABLATE == 2
    {  // + staging: wait for the bulk copy, nothing else
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_stage_bar);
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n.reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
                "selp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done)
                : "r"(bar)
                : "memory");
        }
        return;
    }
#endif
    constexpr int NW = FUSED_THREADS - 32, WW = FUSED_WARPS - 1;
    const int wtid = tid - 32;
    const int nseg_ = nseg;
    int lt_first = 0, lt_last = -1, n_own = 0, nb = 0, blank_lo = 0, blank_hi = -1;
    uint32_t fatal = 0;
    if (warp == 0) {
        LbVal pre;
        pre.nb = 0;
        pre.dist = 0;
        if (tile > 0) {
            // FUSED_LB_WINDOWS x 32 predecessors per L2 round trip: the nearest tile with an inclusive word is usually a
            // hundred or more tiles back (everything younger is still between its own P3 and its own look-back), and a
            // walk of one 32-tile window per round trip took four or five of them AFTER the last aggregate appeared
            long long j = tile - 1;
            for (;;) {
                unsigned long long wq[FUSED_LB_WINDOWS];
#pragma unroll
                for (int q = 0; q < FUSED_LB_WINDOWS; q++) {
                    const long long idx = j - 32 * q - lane;
                    wq[q] = LB_INCLUSIVE << 62;  // before the first tile: nothing (no blank rows, no rows)
                    if (idx >= 0) wq[q] = ld_relaxed_u64(&lb[idx]);
                }
                bool finished = false;
                int used = 0;
#pragma unroll
                for (int q = 0; q < FUSED_LB_WINDOWS; q++) {
                    if (finished || used < q) continue;
                    const unsigned long long w = wq[q];
                    const uint32_t state = (uint32_t)(w >> 62);
                    const uint32_t invalid = __ballot_sync(0xffffffffu, state == (uint32_t)LB_INVALID);
                    const uint32_t inclusive = __ballot_sync(0xffffffffu, state == (uint32_t)LB_INCLUSIVE);
                    const int count = inclusive ? __ffs(inclusive) : 32;  // lanes 0 .. count-1 are what is needed
                    const uint32_t need = count >= 32 ? 0xffffffffu : ((1u << count) - 1u);
                    if (invalid & need) continue;  // this window (and the older ones) again, after a pause
                    LbVal val = lb_unpack(w);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        LbVal o;
                        o.nb = __shfl_down_sync(0xffffffffu, val.nb, d);
                        o.dist = __shfl_down_sync(0xffffffffu, val.dist, d);
                        if (lane + d < count) val = lb_combine(o, val);
                    }
                    // lane 0: the tiles j - 32 q - count + 1 .. j - 32 q; they come before what was gathered so far
                    pre = lb_combine(val, pre);
                    used = q + 1;
                    if (inclusive) finished = true;
                }
                if (finished) break;
                j -= 32 * used;
                if (used < FUSED_LB_WINDOWS) __nanosleep(40);
            }
        }
        // The section this tile starts in is known now: fetch its descriptor (column count, block address, chunk
        // tables) if it is published - every tile but the few next to the header.  The loads are in flight while lane 0
        // waits for the workers; P4 then starts without the chain of L2 round trips it would spend asking for them.
        constexpr int DESC_WORDS = (int)(sizeof(MsSecDesc) / 4), DESC_PER_LANE = (DESC_WORDS + 31) / 32;
        uint32_t dv[DESC_PER_LANE];
        const int sec0 = (int)__shfl_sync(0xffffffffu, pre.nb, 0);
        bool have_desc = false;
#ifndef FUSED_NO_DESC_PREFETCH
        if (sec0 < 2) {
            const uint32_t* dsrc = reinterpret_cast<const uint32_t*>(&ws->desc[sec0]);
            have_desc = ld_acquire_u32(&ws->desc[sec0].ready) != 0u;
            if (have_desc) {
#pragma unroll
                for (int i = 0; i < DESC_PER_LANE; i++) {
                    const int wi = lane + 32 * i;
                    dv[i] = wi < DESC_WORDS ? __ldcg(dsrc + wi) : 0u;
                }
            }
        }
#endif
        ms_bar_agg_wait();  // this tile's own share comes from the workers (shared memory)
        if (lane == 0) {
            LbVal agg;
            agg.nb = s_agg_nb;
            agg.dist = s_agg_dist;
            const LbVal incl = lb_combine(pre, agg);
            st_relaxed_u64(&lb[tile], lb_pack(LB_INCLUSIVE, incl));
            s_pre_nb = pre.nb;
            s_pre_dist = pre.dist;
            if (tile == args.n_tiles - 1) {
                // the state at the end of the file
                res->n_blank_rows = incl.nb;
                res->tail_rows = (long long)incl.dist;
                if (incl.nb < 2u) {
                    res->data_rows[incl.nb] = (long long)incl.dist - 5;
                    atomicOr(&res->have, MS_LOAD_HAVE_ROWS0 << incl.nb);
                }
            }
        }
        if (have_desc) {
            uint32_t* ddst = reinterpret_cast<uint32_t*>(&s_desc);
#pragma unroll
            for (int i = 0; i < DESC_PER_LANE; i++) {
                const int wi = lane + 32 * i;
                if (wi < DESC_WORDS) ddst[wi] = dv[i];
            }
            if (lane == 0) s_desc_sec = sec0;
        }
        {
            // the staged bytes are read by this warp too from P4 on: observe the bulk copy's completion (long done)
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_stage_bar);
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n.reg .pred p;\n"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
                    "selp.u32 %0, 1, 0, p;\n}\n"
                    : "=r"(done)
                    : "r"(bar)
                    : "memory");
            }
        }
    } else {
        {
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_stage_bar);
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n.reg .pred p;\n"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
                    "selp.u32 %0, 1, 0, p;\n}\n"
                    : "=r"(done)
                    : "r"(bar)
                    : "memory");
            }
        }
        // ---- P1a. every 16-byte segment once, lanes on consecutive segments: a quick test for "some byte is below
        // 0x23 or above 0x7f" - line ends, quotes, blanks, control and non-ASCII bytes; a data row has one or two such
        // segments, the rest is digits, signs, points and commas.  (w - 0x23..) | w has bit 7 of a byte set for every
        // such byte; a borrow can only add a false hit on a '#' that follows one.
        for (int base = 0; base < nseg_; base += NW) {
            const int v = base + wtid;
            const bool in = v < nseg_;
            uint32_t ctrl = 0;
            if (in) {
                const uint4 x = *reinterpret_cast<const uint4*>(reg + (v << 4));
                ctrl = ((x.x - 0x23232323u) | x.x) | ((x.y - 0x23232323u) | x.y) | ((x.z - 0x23232323u) | x.z) |
                       ((x.w - 0x23232323u) | x.w);
#ifndef FUSED_LATE_COMMAS
                // the commas of the segment, exactly; P1b adds the line-end bytes of the segments the test hit
                cmask[v] = (uint16_t)ms_mask16(ms_comma_flags(x.x), ms_comma_flags(x.y), ms_comma_flags(x.z), ms_comma_flags(x.w));
#endif
            }
            const uint32_t hits = __ballot_sync(0xffffffffu, in && (ctrl & 0x80808080u));
            if (lane == 0 && v < nseg_) hitmap[v >> 5] = hits;
        }
        ms_bar_workers(NW);

        // ---- P1b. a worker owns a run of consecutive segments; the exact line ends of those the quick test hit
        // (universal newlines: '\n', "\r\n", lone '\r' - what open(filename) gives csv.reader, load_csv.py:29)
        const int vpt = (nseg_ + NW - 1) / NW;  // segments per worker, <= 8
        const int v0 = wtid * vpt;
        uint32_t my_hits = 0;
        if (v0 < nseg_) {
            const uint32_t h0 = hitmap[v0 >> 5], h1 = hitmap[(v0 >> 5) + 1];
            my_hits = __funnelshift_r(h0, h1, v0 & 31) & ((1u << vpt) - 1u);
        }
        int my_terms = 0, my_quotes = 0;
        uint32_t my_flags = 0;
        int lt_end_part = -1;  // terminators of my segments before position tile_len - 1, if that position is mine
        {
            const int q = tile_len - 1;
            const bool mine = q >= (v0 << 4) && q < ((v0 + vpt) << 4);
            if (mine) lt_end_part = 0;
            uint32_t hb = my_hits;
            while (hb) {
                const int k = __ffs(hb) - 1;
                hb &= hb - 1u;
                const int v = v0 + k;
                const uint4 x = *reinterpret_cast<const uint4*>(reg + (v << 4));
                const uint32_t lf = ms_mask16(ms_eq_flags(x.x, 0x0a0a0a0au), ms_eq_flags(x.y, 0x0a0a0a0au),
                                              ms_eq_flags(x.z, 0x0a0a0a0au), ms_eq_flags(x.w, 0x0a0a0a0au));
                const uint32_t cr = ms_mask16(ms_eq_flags(x.x, 0x0d0d0d0du), ms_eq_flags(x.y, 0x0d0d0d0du),
                                              ms_eq_flags(x.z, 0x0d0d0d0du), ms_eq_flags(x.w, 0x0d0d0d0du));
                my_quotes += __popc(ms_eq_flags(x.x, 0x22222222u)) + __popc(ms_eq_flags(x.y, 0x22222222u)) +
                             __popc(ms_eq_flags(x.z, 0x22222222u)) + __popc(ms_eq_flags(x.w, 0x22222222u));
                if ((x.x | x.y | x.z | x.w) & 0x80808080u) my_flags |= MS_LOAD_HIGH_BYTES;
                const uint32_t term = ms_term16(lf, cr, reg[(v << 4) + 16] == '\n');
                tmask[v] = (uint16_t)term;
#ifdef FUSED_LATE_COMMAS
                cmask[v] = (uint16_t)(lf | cr);  // a field also ends at a line end; P1c adds the commas
#else
                cmask[v] |= (uint16_t)(lf | cr);  // a field also ends at a line end: commas + line-end bytes = delimiters
#endif
                my_terms += __popc(term);
                if (mine) {
                    const int p0 = v << 4;
                    if (q >= p0 + 16)
                        lt_end_part += __popc(term);
                    else if (q > p0)
                        lt_end_part += __popc(term & ((1u << (q - p0)) - 1u));
                }
            }
        }
        // exclusive prefix sum of the terminator counts over the workers
        int inc = my_terms;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_warp_terms[warp - 1] = inc;
        if (my_quotes) atomicAdd(&s_quotes, my_quotes);
        if (my_flags) atomicOr(&s_flags, my_flags);
        ms_bar_workers(NW);
        int before, total_terms;
        {
            const int mine_w = lane < WW ? s_warp_terms[lane] : 0;
            int sc = mine_w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, sc, d);
                if (lane >= d) sc += o;
            }
            total_terms = __shfl_sync(0xffffffffu, sc, WW - 1);
            before = __shfl_sync(0xffffffffu, sc - mine_w, warp - 1);
        }
        const int lt0 = before + inc - my_terms;  // terminators before my segments
        if (lt_end_part >= 0) s_lt_end = lt0 + lt_end_part;
        // local row L starts right after the L-th terminator of the region (L >= 1); row 0 starts at byte 0
        {
            int lt = lt0;
            uint32_t hb = my_hits;
            while (hb) {
                const int k = __ffs(hb) - 1;
                hb &= hb - 1u;
                const int v = v0 + k;
                uint32_t term = tmask[v];
                while (term) {
                    const int b2 = __ffs(term) - 1;
                    term &= term - 1u;
                    lt++;
                    if (lt <= FUSED_ROWS_CAP + 1) row_start[lt] = (uint16_t)((v << 4) + b2 + 1);
                }
            }
            if (wtid == 0) row_start[0] = 0;
        }
        ms_bar_workers(NW);

        // ---- ownership: the rows that START in [t0, t0 + tile_len)
        const bool starts_at_t0_w = (t0 == 0) || reg[-1] == '\n' || (reg[-1] == '\r' && reg[0] != '\n');
        lt_first = starts_at_t0_w ? 0 : 1;
        lt_last = s_lt_end;  // inclusive; tile_len >= 1, so some worker set it
        n_own = max(0, lt_last - lt_first + 1);
        if (n_own > 0 && total_terms < lt_last + 1) fatal |= MS_LOAD_ROW_TOO_LONG;  // the last owned row does not end in the region
        if (lt_last + 1 > FUSED_ROWS_CAP) fatal |= MS_LOAD_DENSE_ROWS;

        // ---- P2. blank rows (section separators): every field empty after str.strip() (reader.py:886-901).  A
        // data row starts with its frame number, so only rows that start with a comma or a blank are looked at in full.
        if (!fatal) {
            for (int L = lt_first + wtid; L <= lt_last; L += NW) {
                const int p = row_start[L], e = row_start[L + 1];
                const unsigned c = reg[p];
                if (c == ',' || c <= 0x20u) {
                    bool blank = true;
                    for (int i = p; i < e; i++) {
                        const unsigned b2 = reg[i];
                        if (!(b2 == ',' || ms_is_strip_space(b2))) {
                            blank = false;
                            break;
                        }
                    }
                    if (blank) {
                        atomicAdd(&s_nblank, 1);
                        atomicMin(&s_blank_lo, L);
                        atomicMax(&s_blank_hi, L);
                    }
                }
            }
        }
        ms_bar_workers(NW);
        // ---- P3. tell the other tiles what this one holds
        if (wtid == 0) {
            const int nbw = s_nblank;
            LbVal agg;
            agg.nb = (uint32_t)min(nbw, 3);
            agg.dist = (unsigned long long)(nbw ? lt_last - s_blank_hi : n_own);
            if (tile > 0) st_relaxed_u64(&lb[tile], lb_pack(LB_AGGREGATE, agg));
            s_agg_nb = agg.nb;
            s_agg_dist = agg.dist;
            s_fatal = fatal;
            if (s_quotes) atomicAdd((unsigned long long*)&res->n_quotes, (unsigned long long)s_quotes);
            const uint32_t f = s_flags | fatal | (nbw > 2 ? MS_LOAD_MANY_BLANKS : 0u);
            if (f) atomicOr(&res->flags, f);
        }
        if (warp == 1) ms_bar_agg_arrive();
#ifdef FUSED_LATE_COMMAS
        // ---- P1c. the commas of every segment, exactly (delimiters = commas + the line-end bytes P1b left for the
        // segments the quick test hit).  Nothing the other tiles wait for depends on them, so they are found AFTER the
        // aggregate is out: the look-back of this tile and of its successors runs under this loop instead of after it.
        for (int base = 0; base < nseg_; base += NW) {
            const int v = base + wtid;
            if (v < nseg_) {
                const uint4 x = *reinterpret_cast<const uint4*>(reg + (v << 4));
                uint32_t m = ms_mask16(ms_comma_flags(x.x), ms_comma_flags(x.y), ms_comma_flags(x.z), ms_comma_flags(x.w));
                if ((hitmap[v >> 5] >> (v & 31)) & 1u) m |= cmask[v];
                cmask[v] = (uint16_t)m;
            }
        }
#endif
    }
    __syncthreads();
    fatal = s_fatal;
    if (fatal) return;
#if defined(FUSED_ABLATE) && FUSED_ABLATE == 3
    return;  // + rows, look-back, delimiter masks; no parsing
#endif
    {
        const bool starts_at_t0 = (t0 == 0) || reg[-1] == '\n' || (reg[-1] == '\r' && reg[0] != '\n');
        lt_first = starts_at_t0 ? 0 : 1;
        lt_last = s_lt_end;
        nb = s_nblank;
        blank_lo = s_blank_lo;
        blank_hi = s_blank_hi;
    }
    const int pre_nb = (int)s_pre_nb;
    const long long pre_dist = (long long)s_pre_dist;

    // ---- P4. the runs of rows between this tile's blank rows: header lines, then data rows
    const int n_cut = min(nb, 2);
    for (int run = 0; run <= n_cut; run++) {
        const int lo = run == 0 ? lt_first : (run == 1 ? blank_lo : blank_hi) + 1;
        const int hi = run < n_cut ? (run == 0 ? blank_lo : blank_hi) - 1 : lt_last;
        const int sec = pre_nb + run;
        const long long idx_lo = run == 0 ? pre_dist : 0;  // index in its section of local row `lo`
        if (run < n_cut && sec < 2 && tid == 0) {
            // the blank row that closes this run closes section `sec`
            const int Lb = run == 0 ? blank_lo : blank_hi;
            res->data_rows[sec] = idx_lo + (Lb - lo) - 5;
            res->blank_end[sec] = t0 + row_start[Lb + 1] - 1;
            atomicOr(&res->have, MS_LOAD_HAVE_ROWS0 << sec);
        }
        if (lo > hi) continue;
        if (sec >= 2) {
            if (tid == 0) atomicOr(&res->flags, MS_LOAD_TAIL_ROWS);  // rows after the second blank row (Appendix C)
            continue;
        }
        MsSecDesc* const desc = &ws->desc[sec];
        // header line 0: where the section's header text starts; copy it out for the host
        if (idx_lo == 0) {
            const long long off = t0 + row_start[lo];
            const long long cnt = min((long long)MS_LOAD_PEEK, n - off);
            for (long long i = tid; i < cnt; i += FUSED_THREADS) peek[(long long)sec * MS_LOAD_PEEK + i] = src[off + i];
            if (tid == 0) {
                res->header_offset[sec] = off;
                res->peek_bytes[sec] = cnt;
                atomicOr(&res->have, MS_LOAD_HAVE_HEADER0 << sec);
            }
        }
        // header line 3 (coordinates): its field count is the section's column count (reader.py:772-783)
        if (idx_lo <= 3 && lo + (3 - idx_lo) <= hi) {
            const int L3 = lo + (int)(3 - idx_lo);
            const int p = row_start[L3], e = row_start[L3 + 1];
            if (tid == 0) {
                s_q = -1;
                s_commas = 0;
            }
            __syncthreads();
            int q = -1;
            for (int i = p + tid; i < e; i += FUSED_THREADS) {
                const unsigned b = reg[i];
                if (!(b == ',' || ms_is_strip_space(b))) q = i;
            }
            if (q >= 0) atomicMax(&s_q, q);
            __syncthreads();
            const int last = s_q;  // last byte of the last non-blank field
            int commas = 0;
            for (int i = p + tid; i < last; i += FUSED_THREADS) commas += reg[i] == ',';
            if (commas) atomicAdd(&s_commas, commas);
            __syncthreads();
            const int ncols = last >= 0 ? s_commas + 1 : 0;
            const int keep = ncols - 2;
#ifdef FUSED_DYNAMIC_ITEMS
            if (tid < PARSE_TAB_GROUPS && ncols > 0 && ncols <= 65535)
                desc->chunk_cnt[tid] = (uint8_t)ms_chunk_table(tid + 1, ncols, desc->chunk_tab[tid], FUSED_WARPS);
            __syncthreads();
#endif
            if (tid == 0) {
                long long offset = 0, stride = args.cap_rows[sec];
                bool ok = keep >= 1 && ncols <= 65535;
                if (ok && sec == 1) {
                    // behind the first section's block; its descriptor was published by an earlier tile (or above)
                    const MsSecDesc* d0 = &ws->desc[0];
                    while (!ld_acquire_u32(&d0->ready)) {
                        if (*(volatile uint32_t*)&res->flags) {
                            ok = false;
                            break;
                        }
                        __nanosleep(100);
                    }
                    if (ok) {
                        offset = (__ldcg(&d0->stride) * __ldcg(&d0->n_keep) + 1) & ~1ll;
                        if (stride <= 0) stride = ((args.arena_elems - offset) / keep) & ~1ll;
                    }
                }
                if (ok && (stride <= 0 || offset + stride * keep > args.arena_elems)) {
                    atomicOr(&res->flags, MS_LOAD_OVERFLOW);
                    ok = false;
                }
                if (!ok && !*(volatile uint32_t*)&res->flags) atomicOr(&res->flags, MS_LOAD_BAD_HEADER);
                if (ok) {
                    desc->num_cols = ncols;
                    desc->n_keep = keep;
                    desc->stride = stride;
                    desc->out_offset = offset;
                    res->num_cols[sec] = ncols;
                    res->n_keep[sec] = keep;
                    res->stride[sec] = stride;
                    res->out_offset[sec] = offset;
                    atomicOr(&res->have, MS_LOAD_HAVE_DESC0 << sec);
                    __threadfence();
                    st_release_u32(&desc->ready, 1u);
                }
            }
        }
        // data rows: lines 5.. of the section
        const int Ld = lo + (int)max(0ll, 5 - idx_lo);
        if (Ld > hi) continue;
        const bool local_desc = s_desc_sec == sec;  // block-uniform; s_desc was written before the barrier above
        if (!local_desc) {
            if (tid == 0) {
                int stop = 0;
                while (!ld_acquire_u32(&desc->ready)) {
                    if (*(volatile uint32_t*)&res->flags) {  // the tile that owns the header gave up: so does the caller
                        stop = 1;
                        break;
                    }
                    __nanosleep(100);
                }
                s_stop = stop;
            }
            __syncthreads();
            if (s_stop) return;
        }
        const int ncols = local_desc ? s_desc.num_cols : __ldcg(&desc->num_cols);
        const long long out_stride = local_desc ? s_desc.stride : __ldcg(&desc->stride);
        const long long out_offset = local_desc ? s_desc.out_offset : __ldcg(&desc->out_offset);
        const int nrows = hi - Ld + 1;
        const long long out_row0 = idx_lo + (Ld - lo) - 5;  // output row of local row Ld
        if (out_row0 + nrows > out_stride) {
            if (tid == 0) atomicOr(&res->flags, MS_LOAD_OVERFLOW);
            continue;
        }
        const int groups = (nrows + 31) >> 5;
#ifdef FUSED_DYNAMIC_ITEMS
        const int tab_cnt = groups > PARSE_TAB_GROUPS ? 0 : (local_desc ? (int)s_desc.chunk_cnt[groups - 1] : (int)__ldcg(&desc->chunk_cnt[groups - 1]));
        const bool tabulated = tab_cnt != 0;
        if (tabulated && tid <= PARSE_MAX_CHUNKS)
            s_chunk_col[tid] = local_desc ? s_desc.chunk_tab[groups - 1][tid] : __ldcg(&desc->chunk_tab[groups - 1][tid]);
        if (tid == 0) {
            s_next_item = 0;
            s_inv_groups = (65536u + groups - 1) / groups;  // item / groups by multiply-shift (items < 2^10)
            if (tabulated)
                s_nchunks = tab_cnt;
            else
                s_nchunks = ms_chunk_table(groups, ncols, s_chunk_col, FUSED_WARPS);
        }
        __syncthreads();

        // lanes = rows, in lockstep over the columns of a chunk; warps draw (row group, column chunk) items
        const int nchunks = s_nchunks;
        const int items = groups * nchunks;
        const uint32_t inv_groups = s_inv_groups;
#endif
        // one register holds the shared-window address of the staged bytes for the whole loop (a plain value would
        // be recomputed from the special registers at every use under this kernel's register budget)
#ifdef FUSED_SMEM_SYM
        // the shared-window address of the staged bytes as a link-time constant: ptxas folds it into the loads' offsets
        uint32_t sreg;
        asm("mov.u32 %0, smem_raw;" : "=r"(sreg));
        sreg += FUSED_PAD;
#else
        uint32_t sreg = (uint32_t)__cvta_generic_to_shared(reg);
#ifndef FUSED_NO_PIN
        asm volatile("mov.u32 %0, %0;" : "+r"(sreg));
#endif
#endif
        const uint32_t sdm = sreg + (FUSED_OFF_CMASK - FUSED_PAD);
        uint32_t slut = sreg - FUSED_PAD + lut_off;
#ifdef FUSED_PIN_LUT
        asm volatile("mov.u32 %0, %0;" : "+r"(slut));
#endif
        double* const arena = args.arena;
        const uint32_t stride32 = (uint32_t)out_stride, out_idx0 = (uint32_t)(out_offset + out_row0);
#ifndef FUSED_DYNAMIC_ITEMS
        // lanes = rows, in lockstep over columns.  Every warp takes an equal share of the tile's (row group, column) pairs in row-group order: a straight-line
        // field costs the same whatever its length, so equal counts are equal work.  A warp's share starts in the middle
        // of one row group (one walk to its first column) and runs on from the row starts of the next: 16 walks a tile
        // instead of one per column chunk, and no shared counter.
        const int total = groups * ncols;
        int f = (total * warp) / FUSED_WARPS;
        const int f_end = (total * (warp + 1)) / FUSED_WARPS;
        while (f < f_end) {
            const int g = f / ncols;
            const int c_lo = f - g * ncols;
            const int c_hi = min(ncols, c_lo + (f_end - f));
            f += c_hi - c_lo;
#else
#ifdef FUSED_STATIC_ITEMS
        for (int item = warp; item < items; item += FUSED_WARPS) {
#else
        for (;;) {
            int item = 0;
            if (lane == 0) item = atomicAdd(&s_next_item, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= items) break;
#endif
            const int k = (int)(((uint32_t)item * inv_groups) >> 16), g = item - k * groups;  // chunk-major: wide chunks first
            const int c_lo = s_chunk_col[k + 1], c_hi = s_chunk_col[k];
#endif
            const int r = (g << 5) + lane;
            if (r < nrows) {
                int p = row_start[Ld + r];
                bool done = false;
                if (c_lo > 0) {
                    // first byte of column c_lo = one past the c_lo-th comma of the row, if the row has it
                    const int row_end = row_start[Ld + r + 1];  // one past the row's terminator
#ifdef FUSED_WALK64
                    // four segments' delimiter masks (64 bytes of the row) per step
                    const uint32_t sdm64 = (uint32_t)__cvta_generic_to_shared(cmask);
                    int j = p >> 6;
                    uint2 w = lds_v2(sdm64 + (j << 3));
                    {
                        const int b = p & 63;  // bits below the row's first byte belong to the row before
                        const uint32_t below = b >= 32 ? 0xffffffffu : ((1u << b) - 1u);
                        const uint32_t below_hi = b >= 32 ? ((1u << (b - 32)) - 1u) : 0u;
                        w.x &= ~below;
                        w.y &= ~below_hi;
                    }
                    int need = c_lo;
                    int cnt = __popc(w.x) + __popc(w.y);
                    while (cnt < need && (j << 6) < row_end) {
                        need -= cnt;
                        w = lds_v2(sdm64 + (++j << 3));
                        cnt = __popc(w.x) + __popc(w.y);
                    }
                    if (cnt < need) {
                        done = true;
                    } else {
                        const int clo = __popc(w.x);
                        const bool up = need > clo;
                        uint32_t m = up ? w.y : w.x;
                        need -= up ? clo : 0;
                        for (int i = 1; i < need; i++) m &= m - 1u;
                        p = (j << 6) + (up ? 32 : 0) + __ffs(m);  // position after that delimiter
                        if (p > row_end - 1) done = true;         // it belongs to a later row
                    }
#else
                    int seg = p >> 4;
                    uint32_t m = cmask[seg] & ~((1u << (p & 15)) - 1u);
                    int need = c_lo;
                    int cnt = __popc(m);
                    while (cnt < need && (seg << 4) < row_end) {
                        need -= cnt;
                        m = cmask[++seg];
                        cnt = __popc(m);
                    }
                    if (cnt < need) {
                        done = true;
                    } else {
                        for (int i = 1; i < need; i++) m &= m - 1u;
                        p = (seg << 4) + __ffs(m);  // position after that comma
                        if (p > row_end - 1) done = true;  // the comma belongs to a later row
                    }
#endif
                }
                // element index into the arena, 32 bits (the entry point refuses larger arenas)
                uint32_t oi = out_idx0 + (uint32_t)(c_lo - 2) * stride32 + (uint32_t)r;
                int c = c_lo;
#ifdef FUSED_OLD_LOOP
                for (; c < c_hi; c++, oi += stride32) {
                    uint64_t bits = MS_NAN_BITS;
                    if (!done) {
                        int e;
                        if (!ms_field_fast(sreg, sdm, slut, p, &e, &bits)) {
                            const int slot = atomicAdd(&s_slow_n, 1);
                            if (slot < FUSED_SLOW_CAP) slow_queue[slot] = make_uint2((uint32_t)p | ((uint32_t)e << 16), c >= 2 ? oi : 0xffffffffu);
                            bits = MS_NAN_BITS;
                        }
                        done = lds_u8(sreg + e) != ',';
                        p = e + 1;
                    }
                    if (c >= 2) arena[oi] = ms_bits_to_double(bits);  // Frame, Sub Frame are never stored
                }
#else
                // the fields the row has ...
                if (!done) {
                    while (c < c_hi) {
                        int e;
                        uint64_t bits;
                        if (!ms_field_fast(sreg, sdm, slut, p, &e, &bits)) {
                            // for the general parser, after the items (Frame / Sub Frame: parsed for their errors only)
                            const int slot = atomicAdd(&s_slow_n, 1);
                            if (slot < FUSED_SLOW_CAP) slow_queue[slot] = make_uint2((uint32_t)p | ((uint32_t)e << 16), c >= 2 ? oi : 0xffffffffu);
                            bits = MS_NAN_BITS;
                        }
                        const bool last = lds_u8(sreg + e) != ',';
                        p = e + 1;
                        if (c >= 2) arena[oi] = ms_bits_to_double(bits);  // Frame, Sub Frame are never stored
                        c++;
                        oi += stride32;
                        if (last) break;
                    }
                }
                // ... and None -> NaN for the ones a short row does not have (reader.py:944-948, user_data.py:396)
                for (; c < c_hi; c++, oi += stride32)
                    if (c >= 2) arena[oi] = ms_bits_to_double(MS_NAN_BITS);
#endif
            }
        }
        __syncthreads();
        {
            const int n_slow = s_slow_n;
            if (n_slow > FUSED_SLOW_CAP && tid == 0) atomicOr(&res->flags, MS_LOAD_DENSE_ROWS);  // cannot happen below 896 odd fields a tile
            ms_slow_fields(reg, slow_queue, min(n_slow, FUSED_SLOW_CAP), arena, reinterpret_cast<unsigned long long*>(&res->status), t0, tid);
        }
        __syncthreads();  // the queue is reused by the next run
        if (tid == 0) s_slow_n = 0;
    }
}

// ===================================================================================================
// C ABI
// ===================================================================================================
// bytes a tile stages past its end (the longest row it can finish), and the tile size that then fits shared memory
static int ms_fused_overhang(int32_t overhang_bytes) {
    if (overhang_bytes <= 0) return MS_MAX_ROW_BYTES;
    int o = (overhang_bytes + 15) / 16 * 16;
    if (o < 512) o = 512;
    if (o > MS_MAX_ROW_BYTES) o = MS_MAX_ROW_BYTES;
    return o;
}
static int ms_fused_tile_bytes(int32_t tile_bytes, int overhang) {
    const int largest = FUSED_MAX_REGION - overhang;
    if (tile_bytes <= 0) return largest;  // short rows (a small overhang) leave more of the staged region to the tile
    int t = tile_bytes / 16 * 16;
    if (t < 4096) t = 4096;
    if (t > largest) t = largest;
    return t;
}

extern "C" int64_t ms_load_workspace_bytes(int64_t n_bytes, int32_t tile_bytes) {
    (void)tile_bytes;  // sized for the smallest tile: a look-back word per tile is cheap, and any plan then fits
    const int64_t n_tiles = n_bytes <= 0 ? 0 : (n_bytes + 4095) / 4096;
    return FUSED_LB_OFFSET + (n_tiles + 1) * 8;
}

extern "C" int ms_load_fused(const uint8_t* d_bytes, int64_t n_bytes, const ms_load_plan* h_plan, void* d_workspace,
                             int64_t workspace_bytes, ms_load_result* d_result, uint8_t* d_peek, void* stream) {
    if (!d_bytes || !h_plan || !d_workspace || !d_result || !d_peek || n_bytes < 0) return MS_E_INVALID;
    if (((uintptr_t)d_bytes & 15) != 0 || ((uintptr_t)d_workspace & 15) != 0) return MS_E_INVALID;
    if (!h_plan->d_arena || h_plan->arena_elems < 0 || h_plan->cap_rows[0] <= 0 || h_plan->cap_rows[1] < 0) return MS_E_INVALID;
    const int overhang = ms_fused_overhang(h_plan->overhang_bytes);
    const int tile = ms_fused_tile_bytes(h_plan->tile_bytes, overhang);
    if (workspace_bytes < ms_load_workspace_bytes(n_bytes, tile)) return MS_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t need = FUSED_LB_OFFSET + ((n_bytes + tile - 1) / tile + 1) * 8;  // what this launch touches
    MS_CUDA_CHECK(cudaMemsetAsync(d_workspace, 0, (size_t)need, st));
    MS_CUDA_CHECK(cudaMemsetAsync(d_result, 0, sizeof(ms_load_result), st));
    MS_CUDA_CHECK(cudaMemsetAsync(&d_result->status, 0xFF, sizeof(uint64_t), st));
    const int64_t n_tiles = n_bytes == 0 ? 0 : (n_bytes + tile - 1) / tile;
    if (n_tiles == 0) return MS_OK;
    MsFusedArgs a;
    a.arena = h_plan->d_arena;
    a.arena_elems = h_plan->arena_elems;
    a.cap_rows[0] = h_plan->cap_rows[0];
    a.cap_rows[1] = h_plan->cap_rows[1];
    a.tile_bytes = tile;
    a.region_bytes = tile + overhang;
    a.n_tiles = n_tiles;
    MS_CUDA_CHECK(cudaFuncSetAttribute(ms_load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM));
    ms_load_kernel<<<(unsigned)n_tiles, FUSED_THREADS, FUSED_SMEM, st>>>(d_bytes, n_bytes, (MsFusedWs*)d_workspace, a, d_result,
                                                                          d_peek);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}
