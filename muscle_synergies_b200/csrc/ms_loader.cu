// Vicon Nexus CSV loader kernels for sm_100a: the two-pass path (the single-pass kernel is in ms_fused.cu).
//
//   ms_scan_kernel     pass 1, one CTA per 48 KiB tile: row terminators, quotes, blank rows; leaves the
//                      terminator / comma masks of every 16-byte segment for pass 2
//   ms_resolve_kernel  single CTA: terminator prefix per tile, blank rows that straddle tiles, their csv
//                      row indices -> ms_scan_summary
//   ms_parse_kernel    pass 2, one CTA per tile: row starts from the masks, then one LANE per row with the lanes
//                      of a warp in lockstep over the columns of a chunk - each lane walks its row and finds a
//                      field's end while parsing it: correctly rounded decimal -> double, coalesced stores
//                      into channel-major float64 arrays (no field-offset table, no transpose staging)
//   ms_row_index_kernel / ms_parse_rows_kernel   the same result for buffers with rows longer than a tile's overhang
//
// This path answers for EVERY input (quoted fields, any number of blank rows, errors in the reference's words);
// ms_load_fused hands over to it whatever it declines.  Replaces the per-row Python of the reference
// (reader.py:886-948, aggregator.py:96-124, 229-241, user_data.py:391-396); see include/ms_b200.h for the boundary.
#include <stdio.h>

#include "ms_common.cuh"
#include "ms_parse_double.cuh"

// ---- workspace layout ------------------------------------------------------------------------
#define MS_TILE_BLANK_SLOTS 8
struct MsTileInfo {
    uint32_t n_term;      // terminator ends inside the tile
    uint32_t flags;       // MS_TI_*
    int32_t first_term;   // tile-relative offset of the first terminator end, -1 if none
    int32_t last_term;    // tile-relative offset of the last terminator end, -1 if none
    uint32_t n_quotes;
    uint32_t n_blank;     // blank rows fully decided inside the tile
    int32_t blank_pos[MS_TILE_BLANK_SLOTS];  // tile-relative terminator-end offsets of the first of them, ascending
    int32_t pad[2];
};
static_assert(sizeof(MsTileInfo) == 64, "MsTileInfo layout");
#define MS_TI_HAS_TERM 1u
#define MS_TI_NB_TAIL 2u    // non-blank byte after the last terminator (or anywhere if none)
#define MS_TI_NB_HEAD 4u    // non-blank byte before the first terminator
#define MS_TI_HIGH 8u       // byte >= 0x80 present
#define MS_TI_OVERFLOW 16u  // more than MS_TILE_BLANK_CAP blank rows in the tile
#define MS_TI_HAS_CR 32u

#define MS_TILE_BLANK_CAP 16  // >= MS_TILE_BLANK_SLOTS

static inline int64_t ms_num_tiles(int64_t n) { return (n + MS_TILE_BYTES - 1) / MS_TILE_BYTES; }

struct MsWorkspaceView {
    MsTileInfo* tiles;
    unsigned long long* term_prefix;  // terminator ends before the tile
    uint32_t* masks;                  // per 16-byte segment: terminator-end bits | comma bits << 16
    uint8_t* tile_in_quote;           // quote-aware rescan only: csv in-quote state at the start of each tile
    uint2* brief;                     // per tile {n_term, flags | MS_TB_*}: what ms_resolve_kernel reads of most tiles
};
#define MS_TB_HAS_BLANK 0x100u   // the tile decided blank rows itself (blank_pos[] is not empty)
#define MS_TB_HAS_QUOTES 0x200u  // n_quotes != 0

__host__ __device__ static inline int64_t ms_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static MsWorkspaceView ms_view(void* ws, int64_t n_tiles, int64_t n_bytes_for_view) {
    MsWorkspaceView v;
    v.tiles = (MsTileInfo*)ws;
    char* p = (char*)ws + ms_align_up(n_tiles * (int64_t)sizeof(MsTileInfo), 256);
    v.term_prefix = (unsigned long long*)p;
    p += ms_align_up((n_tiles + 1) * 8, 256);
    v.masks = (uint32_t*)p;
    p += ms_align_up((n_bytes_for_view + 15) / 16 * 4 + 32, 256);
    v.tile_in_quote = (uint8_t*)p;
    p += ms_align_up(n_tiles, 256);
    v.brief = (uint2*)p;
    return v;
}

// Byte offset, inside the workspace, of the per-16-byte delimiter masks written by ms_scan
// (uint32: terminator-end bits | comma bits << 16) - used by the host to locate a row from a
// byte offset on the error path.
extern "C" int64_t ms_workspace_masks_offset(int64_t n_bytes) {
    int64_t t = ms_num_tiles(n_bytes < 1 ? 1 : n_bytes);
    return ms_align_up(t * (int64_t)sizeof(MsTileInfo), 256) + ms_align_up((t + 1) * 8, 256);
}

extern "C" int64_t ms_workspace_bytes(int64_t n_bytes) {
    int64_t t = ms_num_tiles(n_bytes < 1 ? 1 : n_bytes);
    int64_t nb = n_bytes < 1 ? 1 : n_bytes;
    return ms_align_up(t * (int64_t)sizeof(MsTileInfo), 256) + ms_align_up((t + 1) * 8, 256) +
           ms_align_up((nb + 15) / 16 * 4 + 32, 256) + ms_align_up(t, 256) + ms_align_up(t * 8, 256);
}

// ---- shared helpers --------------------------------------------------------------------------------
// 16-byte load with everything at or beyond n replaced by `fill`.
__device__ __forceinline__ uint4 ms_load16(const uint8_t* __restrict__ src, int64_t off, int64_t n, uint32_t fill_rep) {
    if (off + 16 <= n) return __ldg(reinterpret_cast<const uint4*>(src + off));
    uint32_t w[4] = {fill_rep, fill_rep, fill_rep, fill_rep};
    if (off < n) {
        // the allocation is readable up to n rounded up to 16 (ABI requirement)
        uint4 v = __ldg(reinterpret_cast<const uint4*>(src + off));
        uint32_t r[4] = {v.x, v.y, v.z, v.w};
        int valid = (int)(n - off);
        for (int i = 0; i < 4; i++) {
            int k = valid - 4 * i;
            if (k >= 4)
                w[i] = r[i];
            else if (k > 0) {
                uint32_t m = (1u << (8 * k)) - 1u;
                w[i] = (r[i] & m) | (fill_rep & ~m);
            }
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ===================================================================================================
// pass 1
// ===================================================================================================
#define SCAN_THREADS 256
#define SCAN_WARPS_MAX (SCAN_THREADS / 32)

struct MsRun {  // state of the row in progress
    uint32_t has_term;  // a terminator was seen (in the span this state summarises)
    uint32_t nb_tail;   // non-blank byte since the last terminator (or since the span start)
};
__device__ __forceinline__ MsRun ms_run_combine(MsRun a, MsRun b) {
    MsRun r;
    r.has_term = a.has_term | b.has_term;
    r.nb_tail = b.has_term ? b.nb_tail : (a.nb_tail | b.nb_tail);
    return r;
}

struct MsWarpSummary {
    uint32_t n_term, n_quotes, flags;
    uint32_t has_term, nb_tail, nb_head;  // nb_head: non-blank before the span's first terminator
    int first_term, last_term;            // tile-relative, -1 if none
};

// Each warp walks a contiguous 6 KiB span of the tile, 512 bytes (one coalesced 16-byte load per
// lane) per iteration, carrying the state of the row in progress in registers; the eight warp
// summaries are combined once at the end.  The delimiter masks are written out so that pass 2
// does not classify the bytes again.
//
// QUOTES variant (rare: the buffer has '"' in its data rows): one warp walks the whole tile and
// also carries the csv in-quote state (excel dialect: a quote toggles it, "" toggles twice), so
// commas and line ends inside quoted fields are not delimiters.  tile_in_quote[t] is the state
// at the first byte of tile t (parity of the quotes before it).
template <int SCAN_WARPS, bool QUOTES>
__global__ void __launch_bounds__(SCAN_WARPS * 32) ms_scan_kernel(const uint8_t* __restrict__ src, int64_t n,
                                                                   MsTileInfo* __restrict__ tiles,
                                                                   uint32_t* __restrict__ masks,
                                                                   const uint8_t* __restrict__ tile_in_quote,
                                                                   uint2* __restrict__ brief) {
    constexpr int SCAN_SPAN = MS_TILE_BYTES / SCAN_WARPS;  // bytes of the tile one warp walks
    constexpr int SCAN_ITERS = SCAN_SPAN / 512;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t t0 = tile * (int64_t)MS_TILE_BYTES;

    __shared__ MsWarpSummary s_warp[SCAN_WARPS];
    __shared__ int s_blank_count;
    __shared__ int s_blank[MS_TILE_BLANK_CAP];
    if (tid == 0) s_blank_count = 0;
    __syncthreads();

    MsRun run;  // warp-uniform: state at the start of the current iteration
    run.has_term = 0;
    run.nb_tail = 0;
    uint32_t in_quote = 0;  // warp-uniform (QUOTES only): inside a quoted field at the start of the iteration
    if (QUOTES) in_quote = tile_in_quote[tile];
    uint32_t my_nterm = 0, my_nquote = 0, my_flags = 0;
    int w_first = -1, w_last = -1;  // meaningful in the lane that sees them; reduced at the end
    uint32_t w_nb_head = 0;

    // software pipeline: the vector of iteration it+1 is in flight while iteration it is classified
    // (beyond the end: commas - blank, not a terminator)
    const int64_t span0 = t0 + warp * SCAN_SPAN;
    uint4 vcur = make_uint4(0x2c2c2c2cu, 0x2c2c2c2cu, 0x2c2c2c2cu, 0x2c2c2c2cu);
    if (span0 < n) vcur = ms_load16(src, span0 + lane * 16, n, 0x2c2c2c2cu);
#pragma unroll 4
    for (int it = 0; it < SCAN_ITERS; it++) {
        const int rel = warp * SCAN_SPAN + it * 512 + lane * 16;
        const int64_t off = t0 + rel;
        if (span0 + it * 512 >= n) break;  // warp-uniform: nothing left in this span
        const bool has_next = (it + 1 < SCAN_ITERS) && (span0 + (it + 1) * 512 < n);  // warp-uniform
        uint4 vnext = make_uint4(0x2c2c2c2cu, 0x2c2c2c2cu, 0x2c2c2c2cu, 0x2c2c2c2cu);
        if (has_next) vnext = ms_load16(src, off + 512, n, 0x2c2c2c2cu);
        const uint4 v = vcur;
        vcur = vnext;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t lf16 = 0, cr16 = 0, comma16 = 0, gt16 = 0, hib = 0, quote16 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            lf16 |= ms_gather4(ms_eq_flags(w[k], 0x0a0a0a0au)) << (4 * k);
            cr16 |= ms_gather4(ms_eq_flags(w[k], 0x0d0d0d0du)) << (4 * k);
            comma16 |= ms_gather4(ms_eq_flags(w[k], 0x2c2c2c2cu)) << (4 * k);
            gt16 |= ms_gather4(ms_ge_flags(w[k], 0x21212121u)) << (4 * k);
            const uint32_t qf = ms_eq_flags(w[k], 0x22222222u);
            my_nquote += __popc(qf);
            if (QUOTES) quote16 |= ms_gather4(qf) << (4 * k);
            hib |= w[k];
        }
        if (hib & 0x80808080u) my_flags |= MS_TI_HIGH;
        if (cr16) my_flags |= MS_TI_HAS_CR;
        if (QUOTES) {
            // in-quote state of every byte: prefix parity of the quote bits, carried across lanes
            uint32_t par = quote16;
            par ^= par << 1;
            par ^= par << 2;
            par ^= par << 4;
            par ^= par << 8;
            par &= 0xffffu;  // bit i: parity of the quotes in bytes 0..i of this vector
            const uint32_t lane_odd = __popc(quote16) & 1u;
            const uint32_t odd_lanes = __ballot_sync(0xffffffffu, lane_odd);
            const uint32_t before = (in_quote ^ (__popc(odd_lanes & ((1u << lane) - 1u)) & 1u)) ? 0xffffu : 0u;
            const uint32_t inq = par ^ before;  // delimiters at set bits are inside a quoted field
            in_quote ^= __popc(odd_lanes) & 1u;
            lf16 &= ~inq;
            cr16 &= ~inq;
            gt16 = (gt16 | (comma16 & inq)) & ~quote16;  // a quoted comma is content, the quotes are not
            comma16 &= ~inq;
        }
        uint32_t nb = gt16 & ~comma16;
        if (~gt16 & ~lf16 & ~cr16 & 0xffffu) {
            // rare: spaces, tabs or control bytes - the exact str.strip() whitespace set decides
            const uint32_t ws = ms_mask16(ms_strip_space_flags(v.x), ms_strip_space_flags(v.y),
                                          ms_strip_space_flags(v.z), ms_strip_space_flags(v.w));
            nb = ~(ws | comma16) & 0xffffu;
            if (QUOTES) nb &= ~quote16;
        }
        // byte after this vector: first byte of the next lane's vector
        uint32_t next_lf = __shfl_down_sync(0xffffffffu, lf16 & 1u, 1);
        {
            // lane 31: the byte after its vector is lane 0's first byte of the next iteration (already
            // loaded); at the end of the span it is read directly.  Raw byte: a '\r' that survives the
            // quote filter is outside quotes, and so is the byte after it.
            const uint32_t first_next = __shfl_sync(0xffffffffu, vnext.x & 0xffu, 0);
            if (lane == 31) {
                if (has_next)
                    next_lf = first_next == '\n';
                else
                    next_lf = (off + 16 < n) ? (uint32_t)(src[off + 16] == '\n') : 0u;
            }
        }
        const uint32_t term = ms_term16(lf16, cr16, next_lf);
        if (off < n) {
            uint32_t cm = comma16;
            if (off + 16 > n) cm &= (1u << (int)(n - off)) - 1u;  // not the fill bytes
            masks[off >> 4] = term | (cm << 16);
        }
        my_nterm += __popc(term);

        const uint32_t has_t = term != 0;
        uint32_t nb_head, nb_tail;
        if (has_t) {
            const uint32_t low = term & (0u - term);
            nb_head = (nb & (low - 1u)) != 0;
            nb_tail = (nb >> (32 - __clz(term))) != 0;
        } else {
            nb_head = nb_tail = nb != 0;
        }
        const uint32_t HT = __ballot_sync(0xffffffffu, has_t);
        const uint32_t NT = __ballot_sync(0xffffffffu, nb_tail);
        const uint32_t below = (1u << lane) - 1u;
        if (has_t) {
            const uint32_t ht = HT & below;
            uint32_t carry, known;
            if (ht) {
                const int j = 31 - __clz(ht);
                carry = ((NT & below) >> j) != 0;
                known = 1;
            } else {
                carry = run.nb_tail | ((NT & below) != 0);
                known = run.has_term;
            }
            const int first_bit = __ffs(term) - 1;
            if (known) {
                if (!(carry | nb_head)) {  // the row ending at my first terminator is blank
                    const int slot = atomicAdd(&s_blank_count, 1);
                    if (slot < MS_TILE_BLANK_CAP) s_blank[slot] = rel + first_bit;
                }
            } else {
                // first terminator of this warp's span: decided when the warps are combined
                w_first = rel + first_bit;
                w_nb_head = carry | nb_head;
            }
            // rows that begin and end inside this vector
            uint32_t rest = term & (term - 1u);
            int prev = first_bit;
            while (rest) {
                const int b = __ffs(rest) - 1;
                rest &= rest - 1u;
                const uint32_t between = (nb >> (prev + 1)) & ((1u << (b - prev - 1)) - 1u);
                if (!between) {
                    const int slot = atomicAdd(&s_blank_count, 1);
                    if (slot < MS_TILE_BLANK_CAP) s_blank[slot] = rel + b;
                }
                prev = b;
            }
            w_last = rel + 31 - __clz(term);
        }
        MsRun agg;
        agg.has_term = HT != 0;
        agg.nb_tail = HT ? ((NT >> (31 - __clz(HT))) != 0) : (NT != 0);
        run = ms_run_combine(run, agg);
    }

    // ---- warp summary
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_nterm += __shfl_xor_sync(0xffffffffu, my_nterm, o);
        my_nquote += __shfl_xor_sync(0xffffffffu, my_nquote, o);
        my_flags |= __shfl_xor_sync(0xffffffffu, my_flags, o);
        w_first = max(w_first, __shfl_xor_sync(0xffffffffu, w_first, o));  // set by exactly one lane
        w_last = max(w_last, __shfl_xor_sync(0xffffffffu, w_last, o));
        w_nb_head |= __shfl_xor_sync(0xffffffffu, w_nb_head, o);
    }
    if (lane == 0) {
        MsWarpSummary ws;
        ws.n_term = my_nterm;
        ws.n_quotes = my_nquote;
        ws.flags = my_flags;
        ws.has_term = run.has_term;
        ws.nb_tail = run.nb_tail;
        ws.nb_head = run.has_term ? w_nb_head : run.nb_tail;
        ws.first_term = w_first;
        ws.last_term = w_last;
        s_warp[warp] = ws;
    }
    __syncthreads();

    if (tid == 0) {
        MsTileInfo ti;
        ti.n_term = 0;
        ti.n_quotes = 0;
        ti.first_term = -1;
        ti.last_term = -1;
        uint32_t f = 0, tile_nb_head = 0;
        MsRun carry;
        carry.has_term = 0;
        carry.nb_tail = 0;
        int nbk = s_blank_count;
        for (int w = 0; w < SCAN_WARPS; w++) {
            const MsWarpSummary ws = s_warp[w];
            ti.n_term += ws.n_term;
            ti.n_quotes += ws.n_quotes;
            f |= ws.flags;
            if (ws.has_term) {
                if (carry.has_term) {
                    // the row ending at this span's first terminator began in an earlier span of the tile
                    if (!(carry.nb_tail | ws.nb_head)) {
                        if (nbk < MS_TILE_BLANK_CAP) s_blank[nbk] = ws.first_term;
                        nbk++;
                    }
                } else {
                    ti.first_term = ws.first_term;  // began in an earlier tile: ms_resolve_kernel decides
                    tile_nb_head = carry.nb_tail | ws.nb_head;
                }
                ti.last_term = ws.last_term;
            }
            MsRun r;
            r.has_term = ws.has_term;
            r.nb_tail = ws.nb_tail;
            carry = ms_run_combine(carry, r);
        }
        if (carry.has_term) f |= MS_TI_HAS_TERM;
        if (carry.nb_tail) f |= MS_TI_NB_TAIL;
        if (tile_nb_head) f |= MS_TI_NB_HEAD;
        const int n_blank = nbk;
        if (nbk > MS_TILE_BLANK_CAP) {
            f |= MS_TI_OVERFLOW;
            nbk = MS_TILE_BLANK_CAP;
        }
        // smallest positions first (insertion sort of at most MS_TILE_BLANK_CAP entries)
        for (int i = 1; i < nbk; i++) {
            int x = s_blank[i], j = i - 1;
            while (j >= 0 && s_blank[j] > x) {
                s_blank[j + 1] = s_blank[j];
                j--;
            }
            s_blank[j + 1] = x;
        }
        ti.n_blank = (uint32_t)n_blank;
        for (int i = 0; i < MS_TILE_BLANK_SLOTS; i++) ti.blank_pos[i] = i < nbk ? s_blank[i] : -1;
        ti.pad[0] = ti.pad[1] = 0;
        ti.flags = f;
        tiles[tile] = ti;
        brief[tile] = make_uint2(ti.n_term, f | (n_blank ? MS_TB_HAS_BLANK : 0u) | (ti.n_quotes ? MS_TB_HAS_QUOTES : 0u));
    }
}

__global__ void ms_quote_parity_kernel(const MsTileInfo* __restrict__ tiles, int64_t n_tiles,
                                       uint8_t* __restrict__ tile_in_quote) {
    uint32_t p = 0;
    for (int64_t t = 0; t < n_tiles; t++) {
        tile_in_quote[t] = (uint8_t)p;
        p ^= tiles[t].n_quotes & 1u;
    }
}

// ---------------------------------------------------------------------------------------------------
// resolve: single CTA
// ---------------------------------------------------------------------------------------------------
#define RESOLVE_THREADS 1024
#define RESOLVE_CAND_CAP 64

// terminator ends in [from, to) of the buffer, from the masks the scan kernel wrote (so that a
// quote-aware scan is honoured); cooperative over the whole block
__device__ unsigned long long ms_block_count_terms(const uint32_t* __restrict__ masks, int64_t from, int64_t to,
                                                   unsigned long long* s_acc) {
    if (threadIdx.x == 0) *s_acc = 0;
    __syncthreads();
    unsigned long long mine = 0;
    const int64_t first_vec = from / 16, last_vec = (to + 15) / 16;
    for (int64_t vi = first_vec + threadIdx.x; vi < last_vec; vi += blockDim.x) {
        const int64_t off = vi * 16;
        uint32_t term = masks[vi] & 0xffffu;
        // keep bytes in [from, to)
        if (off < from) term &= ~((1u << (from - off)) - 1u);
        if (off + 16 > to) term &= (1u << (to - off)) - 1u;
        mine += __popc(term);
    }
    if (mine) atomicAdd(s_acc, mine);
    __syncthreads();
    return *s_acc;
}

__global__ void __launch_bounds__(RESOLVE_THREADS)
    ms_resolve_kernel(const uint32_t* __restrict__ masks, int64_t n, const MsTileInfo* __restrict__ tiles,
                      const uint2* __restrict__ brief, unsigned long long* __restrict__ term_prefix, int64_t n_tiles,
                      ms_scan_summary* __restrict__ out) {
    const int tid = threadIdx.x;
    __shared__ unsigned long long s_scan[RESOLVE_THREADS / 32];
    __shared__ unsigned long long s_quotes, s_blank_total, s_acc;
    __shared__ uint32_t s_flags;
    __shared__ int s_ncand;
    __shared__ long long s_cand[RESOLVE_CAND_CAP];

    if (tid == 0) {
        s_quotes = 0;
        s_blank_total = 0;
        s_flags = 0;
        s_ncand = 0;
    }
    __syncthreads();

    // Every thread owns a run of consecutive tiles and reads only their 8-byte briefs (coalesced across the
    // block; the 64-byte records are touched for the rare tiles with blank rows or quotes): a thread sums
    // its run's terminator counts, one block-wide scan turns the sums into bases, and a second sweep over
    // the run writes the prefix and settles the rows that straddle tiles.
    const int64_t per = (n_tiles + RESOLVE_THREADS - 1) / RESOLVE_THREADS;
    const int64_t k0 = min(n_tiles, (int64_t)tid * per), k1 = min(n_tiles, k0 + per);
    unsigned long long v = 0;
    for (int64_t k = k0; k < k1; k++) v += brief[k].x;
    unsigned long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) s_scan[tid >> 5] = inc;
    __syncthreads();
    unsigned long long run = inc - v;
    for (int w = 0; w < (tid >> 5); w++) run += s_scan[w];
    if (tid == RESOLVE_THREADS - 1) term_prefix[n_tiles] = run + v;

    // ---- blank rows
    unsigned long long my_quotes = 0, my_blank = 0;
    uint32_t my_flags = 0;
    uint32_t prev_flags = k0 > 0 ? brief[k0 - 1].y : 0u;  // flags of tile k - 1
    for (int64_t k = k0; k < k1; k++) {
        const uint2 bf = brief[k];
        term_prefix[k] = run;
        run += bf.x;
        if (bf.y & MS_TI_HIGH) my_flags |= MS_SCAN_HAS_HIGH_BYTES;
        if (bf.y & MS_TI_OVERFLOW) my_flags |= MS_SCAN_BLANK_OVERFLOW;
        if (bf.y & MS_TI_HAS_CR) my_flags |= MS_SCAN_HAS_CR;
        const int64_t t0 = k * (int64_t)MS_TILE_BYTES;
        if (bf.y & (MS_TB_HAS_BLANK | MS_TB_HAS_QUOTES)) {
            const MsTileInfo ti = tiles[k];
            my_quotes += ti.n_quotes;
            my_blank += ti.n_blank;
            for (int i = 0; i < MS_TILE_BLANK_SLOTS; i++) {
                if (ti.blank_pos[i] >= 0) {
                    int slot = atomicAdd(&s_ncand, 1);
                    if (slot < RESOLVE_CAND_CAP) s_cand[slot] = t0 + ti.blank_pos[i];
                }
            }
        }
        if (bf.y & MS_TI_HAS_TERM) {
            // the row ending at this tile's first terminator began in an earlier tile (or at
            // the first byte of this one): walk back to the previous terminator
            uint32_t acc = prev_flags & MS_TI_NB_TAIL;
            if (k > 0 && !(prev_flags & MS_TI_HAS_TERM)) {
                for (int64_t j = k - 2; j >= 0; j--) {
                    const uint32_t f = brief[j].y;
                    acc |= f & MS_TI_NB_TAIL;
                    if (f & MS_TI_HAS_TERM) break;
                }
            }
            if (!acc && !(bf.y & MS_TI_NB_HEAD)) {
                my_blank += 1;
                int slot = atomicAdd(&s_ncand, 1);
                if (slot < RESOLVE_CAND_CAP) s_cand[slot] = t0 + tiles[k].first_term;
            }
        }
        prev_flags = bf.y;
    }
    // tiles with more than MS_TILE_BLANK_SLOTS decided blank rows: the rest is not recorded;
    // harmless unless they are among the first MS_MAX_BLANK_ROWS of the file (flagged below)
    if (my_quotes) atomicAdd(&s_quotes, my_quotes);
    if (my_blank) atomicAdd(&s_blank_total, my_blank);
    if (my_flags) atomicOr(&s_flags, my_flags);
    __syncthreads();

    // ---- unterminated last row
    __shared__ int s_eof_row;
    if (tid == 0) {
        s_eof_row = 0;
        if (n > 0) {
            const MsTileInfo last = tiles[n_tiles - 1];
            const int64_t t0 = (n_tiles - 1) * (int64_t)MS_TILE_BYTES;
            bool ends_with_term = (last.flags & MS_TI_HAS_TERM) && (t0 + last.last_term == n - 1);
            if (!ends_with_term) {
                s_eof_row = 1;
                uint32_t acc = 0;
                for (int64_t j = n_tiles - 1; j >= 0; j--) {
                    uint32_t f = tiles[j].flags;
                    acc |= f & MS_TI_NB_TAIL;
                    if (f & MS_TI_HAS_TERM) break;
                }
                if (!acc) {
                    s_blank_total += 1;
                    int slot = s_ncand++;
                    if (slot < RESOLVE_CAND_CAP) s_cand[slot] = n;  // virtual terminator
                }
            }
        }
    }
    __syncthreads();

    // ---- order the candidates, keep the first MS_MAX_BLANK_ROWS
    __shared__ int s_nrep;
    if (tid == 0) {
        int nc = s_ncand;
        if (nc > RESOLVE_CAND_CAP) {
            s_flags |= MS_SCAN_BLANK_OVERFLOW;
            nc = RESOLVE_CAND_CAP;
        }
        for (int i = 1; i < nc; i++) {  // insertion sort, nc is tiny
            long long x = s_cand[i];
            int j = i - 1;
            while (j >= 0 && s_cand[j] > x) {
                s_cand[j + 1] = s_cand[j];
                j--;
            }
            s_cand[j + 1] = x;
        }
        // a tile that decided more than MS_TILE_BLANK_SLOTS blank rows did not record all: if one of the
        // first MS_MAX_BLANK_ROWS candidates comes from such a tile the order is unreliable
        int nrep = nc < MS_MAX_BLANK_ROWS ? nc : MS_MAX_BLANK_ROWS;
        for (int i = 0; i < nrep; i++) {
            long long p = s_cand[i];
            int64_t k = p >= n ? n_tiles - 1 : p / MS_TILE_BYTES;
            if (tiles[k].n_blank > MS_TILE_BLANK_SLOTS) s_flags |= MS_SCAN_BLANK_OVERFLOW;
        }
        s_nrep = nrep;
    }
    __syncthreads();

    // ---- csv row index of each reported blank row
    const int nrep = s_nrep;
    for (int i = 0; i < nrep; i++) {
        const long long p = s_cand[i];
        unsigned long long row;
        if (p >= n) {
            row = term_prefix[n_tiles];
        } else {
            const int64_t k = p / MS_TILE_BYTES;
            unsigned long long c = ms_block_count_terms(masks, k * (int64_t)MS_TILE_BYTES, p, &s_acc);
            row = term_prefix[k] + c;
        }
        if (tid == 0) {
            out->blank_row[i] = (int64_t)row;
            out->blank_end[i] = p;
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int i = nrep; i < MS_MAX_BLANK_ROWS; i++) {
            out->blank_row[i] = -1;
            out->blank_end[i] = -1;
        }
        out->n_bytes = n;
        out->n_terminators = (int64_t)term_prefix[n_tiles];
        out->n_rows = (int64_t)term_prefix[n_tiles] + s_eof_row;
        out->n_quotes = (int64_t)s_quotes;
        out->n_blank_rows = (int64_t)s_blank_total;
        out->flags = s_flags;
        out->n_reported = (uint32_t)nrep;
    }
}

// ===================================================================================================
// pass 2
// ===================================================================================================
// One CTA per 48 KiB tile (three CTAs per SM); the CTA owns the rows that START inside the tile and reads up to
// MS_MAX_ROW_BYTES past it to finish the last one.
//
//   A. stage the bytes in shared memory (16-byte coalesced loads); per thread, comma and
//      terminator bit masks of a contiguous 112-byte chunk; block-wide prefix sum of the
//      terminator counts -> start offset of every owned row (shared array); comma masks are kept
//      in shared memory for step B's column lookups
//   B. one LANE per row, lanes in lockstep over the columns: the 32 lanes of a warp parse the
//      same column of 32 consecutive rows (same kind of text in every lane -> little
//      divergence), walking their row left to right and finding each field's end while parsing
//      it, and store 32 consecutive doubles of one channel (coalesced 256-byte stores, no
//      transpose staging).  The ignored tail of a row (fields beyond num_cols) is never touched.
//      To keep all warps busy a row group is split into column chunks handed out from a shared
//      counter; a chunk locates its first column with a popcount walk over the row's comma masks
//      (precomputing those starts in a separate step was measured slower: it serialises on a barrier).
//
// Rows are processed in batches of PARSE_ROWS_CAP per section, so pathological inputs
// (thousands of tiny rows in a tile) only cost more rounds.
#include "ms_field.cuh"

#ifndef PARSE_MIN_CTAS
#define PARSE_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(PARSE_THREADS, PARSE_MIN_CTAS)
    ms_parse_kernel(const uint8_t* __restrict__ src, int64_t n, const unsigned long long* __restrict__ term_prefix,
                    const uint32_t* __restrict__ masks, const MsSectionsArg secs,
                    unsigned long long* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const reg = smem_raw + PARSE_PAD;  // reg[i] = src[t0 + i]
    uint16_t* const cmask = reinterpret_cast<uint16_t*>(smem_raw + PARSE_BYTES_SMEM);  // commas per 16-byte segment
    int* const row_start = reinterpret_cast<int*>(smem_raw + PARSE_BYTES_SMEM + PARSE_NSEG * 2 + 32);
    __shared__ int s_warp_terms[PARSE_WARPS];
    __shared__ int s_lt_end, s_next_item, s_nchunks;
    __shared__ int s_chunk_col[PARSE_MAX_CHUNKS + 1];  // column chunk boundaries (descending)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t t0 = tile * (int64_t)MS_TILE_BYTES;
    const int tile_len = (int)min((int64_t)MS_TILE_BYTES, n - t0);

    // rows of this tile that can be data rows at all?  (cheap early exit for header-only tiles)
    const long long row_lo = (long long)term_prefix[tile], row_hi = (long long)term_prefix[tile + 1] + 1;
    bool any = false;
    for (int i = 0; i < secs.n; i++)
        if (row_hi >= secs.s[i].row_begin && row_lo < secs.s[i].row_end) any = true;
    if (!any) return;

    // ---- A1. stage [t0 - 16, t0 + REGION + 16) in shared memory; beyond the end: '\n'.
    // One bulk asynchronous copy (cp.async.bulk, the TMA engine: UBLKCP) issued by a single thread moves every
    // whole 16-byte chunk that lies inside the buffer and reports to an mbarrier; it stays in flight while the
    // masks are fetched and scanned below (the bytes are first needed for the ownership test).  The chunks at
    // the edges of the file - before byte 0, or holding / beyond byte n - are filled by hand.
    __shared__ __align__(8) unsigned long long s_stage_bar;
    const int64_t off0 = t0 - PARSE_PAD;                                  // file offset of chunk 0
    const int lo_chunk = off0 < 0 ? (int)((-off0) >> 4) : 0;              // first chunk inside the file
    const int64_t whole = (n >> 4) - (off0 >> 4);                         // chunks that end at or before byte n
    const int hi_chunk = (int)max((int64_t)lo_chunk, min((int64_t)(PARSE_BYTES_SMEM / 16), whole));
    if (tid == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_stage_bar);
        const uint32_t bytes = (uint32_t)(hi_chunk - lo_chunk) * 16u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
        if (bytes) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_raw + lo_chunk * 16);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                         "l"(src + off0 + (int64_t)lo_chunk * 16), "r"(bytes), "r"(bar)
                         : "memory");
        }
    }
    for (int i = tid; i < PARSE_BYTES_SMEM / 16; i += PARSE_THREADS) {
        if (i >= lo_chunk && i < hi_chunk) {
            if (lo_chunk == 0 && hi_chunk == PARSE_BYTES_SMEM / 16) break;  // the usual tile: nothing by hand
            continue;
        }
        const int64_t off = off0 + (int64_t)i * 16;
        uint4 v;
        if (off < 0)
            v = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
        else
            v = ms_load16(src, off, n, 0x0a0a0a0au);
        *reinterpret_cast<uint4*>(smem_raw + i * 16) = v;
    }

    // ---- A2. delimiter masks of my chunk: classified once, by ms_scan_kernel
    const int c0 = tid * PARSE_CHUNK;
    uint32_t mterm[PARSE_SEGS];
    int my_terms = 0;
    int lt_end_part = -1;  // terminators of my chunk before position tile_len - 1, if it is mine
    {
        const int64_t n_seg = (n + 15) >> 4;
        const int64_t seg0 = (t0 >> 4) + tid * PARSE_SEGS;
#pragma unroll
        for (int s = 0; s < PARSE_SEGS; s++) {
            const int64_t sa = seg0 + s;
            uint32_t term = 0xffffu, comma = 0;  // beyond the end: every (fill) byte ends a row
            if (sa < n_seg) {
                const uint32_t mk = __ldg(masks + sa);
                term = mk & 0xffffu;
                comma = mk >> 16;
                if (sa == n_seg - 1 && (n & 15)) term |= 0xffffu & ~((1u << (int)(n & 15)) - 1u);
            }
            mterm[s] = term;
            cmask[tid * PARSE_SEGS + s] = (uint16_t)comma;
            my_terms += __popc(term);
        }
    }
    {
        const int q = tile_len - 1;
        if (q >= c0 && q < c0 + PARSE_CHUNK) {
            int cnt = 0;
#pragma unroll
            for (int s = 0; s < PARSE_SEGS; s++) {
                const int p0 = c0 + s * 16;
                if (q >= p0 + 16)
                    cnt += __popc(mterm[s]);
                else if (q > p0)
                    cnt += __popc(mterm[s] & ((1u << (q - p0)) - 1u));
            }
            lt_end_part = cnt;
        }
    }
    // ---- A3. block-wide exclusive prefix sum of terminator counts
    int inc = my_terms;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp_terms[warp] = inc;
    __syncthreads();
    int before = 0, total_terms = 0;
    for (int w = 0; w < PARSE_WARPS; w++) {
        const int v = s_warp_terms[w];
        if (w < warp) before += v;
        total_terms += v;
    }
    const int lt0 = before + inc - my_terms;  // terminators before my chunk
    if (lt_end_part >= 0) s_lt_end = lt0 + lt_end_part;
    __syncthreads();

    // the staged bytes are needed from here on (the barrier above also made the mbarrier's init visible)
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
    __syncthreads();

    // ---- ownership: rows that START in [t0, t0 + tile_len)
    // local row index lt = terminators in [t0, position); row lt starts right after the lt-th one
    const bool starts_at_t0 = (t0 == 0) || reg[-1] == '\n' || (reg[-1] == '\r' && reg[0] != '\n');
    const int lt_first = starts_at_t0 ? 0 : 1;
    const int lt_last = s_lt_end;  // inclusive
    if (lt_last < lt_first) return;
    if (total_terms < lt_last + 1) {
        // the last owned row does not end inside the staged region
        if (tid == 0) atomicMin(status, ((unsigned long long)t0 << 3) | MS_ERR_KIND_ROW_TOO_LONG);
        return;
    }

    for (int si = 0; si < secs.n; si++) {
        // local rows of this section owned by this tile
        const long long a_ll = max((long long)lt_first, secs.s[si].row_begin - row_lo);
        const long long b_ll = min((long long)lt_last, secs.s[si].row_end - 1 - row_lo);
        if (a_ll > b_ll) continue;
        const int sec_a = (int)a_ll, sec_b = (int)b_ll;
        const int ncols = secs.s[si].num_cols;
        const int n_keep = secs.s[si].n_keep;
        double* const out_base = secs.s[si].d_out;
        const int64_t out_stride = secs.s[si].stride;
        const long long out_row0 = row_lo - secs.s[si].row_begin;  // output row of local row 0

        for (int ba = sec_a; ba <= sec_b; ba += PARSE_ROWS_CAP) {
            const int bb = min(sec_b, ba + PARSE_ROWS_CAP - 1);
            const int nrows = bb - ba + 1;
            // ---- A4. start offset of rows ba .. bb+1 (the last one closes row bb)
            const int groups_ = (nrows + 31) >> 5;
            const bool tabulated = groups_ <= PARSE_TAB_GROUPS && secs.chunk_cnt[si][groups_ - 1] != 0;
            if (tabulated && tid <= PARSE_MAX_CHUNKS) s_chunk_col[tid] = secs.chunk_tab[si][groups_ - 1][tid];
            if (tid == 0) {
                s_next_item = 0;
                if (starts_at_t0 && ba == 0) row_start[0] = 0;
                if (tabulated)
                    s_nchunks = secs.chunk_cnt[si][groups_ - 1];
                else
                    s_nchunks = ms_chunk_table(groups_, ncols, s_chunk_col);
            }
            {
                int lt = lt0;
#pragma unroll
                for (int s = 0; s < PARSE_SEGS; s++) {
                    uint32_t term = mterm[s];
                    while (term) {
                        const int b = __ffs(term) - 1;
                        term &= term - 1u;
                        lt++;  // the row that starts after this terminator
                        if (lt >= ba && lt <= bb + 1) row_start[lt - ba] = c0 + s * 16 + b + 1;
                    }
                }
            }
            __syncthreads();

            // ---- B. lanes = rows, lockstep over columns
            const int groups = (nrows + 31) >> 5;
            const int nchunks = s_nchunks;
            const int items = groups * nchunks;
            const uint32_t inv_groups = (65536u + groups - 1) / groups;  // item / groups by multiply-shift (items < 2^9)
            for (;;) {
                // warps take (row group, column chunk) items from a shared counter
                int item = 0;
                if (lane == 0) item = atomicAdd(&s_next_item, 1);
                item = __shfl_sync(0xffffffffu, item, 0);
                if (item >= items) break;
                const int k = (int)(((uint32_t)item * inv_groups) >> 16), g = item - k * groups;  // chunk-major: wide chunks first
                const int r = (g << 5) + lane;
                if (r < nrows) {
                const int c_lo = s_chunk_col[k + 1], c_hi = s_chunk_col[k];
                int p = row_start[r];
                bool done = false;
                if (c_lo > 0) {
                    // first byte of column c_lo = one past the c_lo-th comma of the row, if the row has it
                    const int row_end = row_start[r + 1];  // one past the row's terminator
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
                        // the comma must belong to this row (a terminator may come first)
                        if (p > row_end - 1) done = true;
                    }
                }
                double* out = out_base + (int64_t)(c_lo - 2) * out_stride + (out_row0 + ba + r);
                for (int c = c_lo; c < c_hi; c++, out += out_stride) {
                    uint64_t bits = MS_NAN_BITS;
                    if (!done) done = ms_parse_next(reg, &p, &bits, status, t0);
                    const int ch = c - 2;
                    if (ch >= 0 && ch < n_keep) *out = ms_bits_to_double(bits);
                }
                }
            }
            __syncthreads();  // row_start is reused by the next batch
        }
    }
}

// ===================================================================================================
// C ABI
// ===================================================================================================
static int ms_scan_impl(const uint8_t* d_bytes, int64_t n_bytes, void* d_workspace, int64_t workspace_bytes,
                        ms_scan_summary* d_summary, void* stream, bool quoted) {
    if (!d_bytes || !d_workspace || !d_summary || n_bytes < 0) return MS_E_INVALID;
    if (((uintptr_t)d_bytes & 15) != 0) return MS_E_INVALID;
    if (workspace_bytes < ms_workspace_bytes(n_bytes)) return MS_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_tiles = ms_num_tiles(n_bytes);
    const int64_t nb = n_bytes < 1 ? 1 : n_bytes;
    MsWorkspaceView v = ms_view(d_workspace, ms_num_tiles(nb), nb);
    if (n_tiles > 0) {
        if (!quoted) {
            ms_scan_kernel<SCAN_WARPS_MAX, false>
                <<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(d_bytes, n_bytes, v.tiles, v.masks, nullptr, v.brief);
        } else {
            // in-quote state at every tile start from the quote counts of the plain scan
            ms_quote_parity_kernel<<<1, 1, 0, st>>>(v.tiles, n_tiles, v.tile_in_quote);
            MS_COUNT_LAUNCH();
            ms_scan_kernel<1, true><<<(unsigned)n_tiles, 32, 0, st>>>(d_bytes, n_bytes, v.tiles, v.masks, v.tile_in_quote, v.brief);
        }
        MS_COUNT_LAUNCH();
        MS_CUDA_CHECK(cudaGetLastError());
    }
    ms_resolve_kernel<<<1, RESOLVE_THREADS, 0, st>>>(v.masks, n_bytes, v.tiles, v.brief, v.term_prefix, n_tiles, d_summary);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_scan(const uint8_t* d_bytes, int64_t n_bytes, void* d_workspace, int64_t workspace_bytes,
                       ms_scan_summary* d_summary, void* stream) {
    return ms_scan_impl(d_bytes, n_bytes, d_workspace, workspace_bytes, d_summary, stream, false);
}

// Second scan for buffers whose DATA rows contain '"' (ms_scan reported more quotes than the
// header lines hold): same outputs, with commas and line ends inside quoted fields not counted
// as delimiters.  Must follow ms_scan on the same buffer and workspace.
extern "C" int ms_scan_quoted(const uint8_t* d_bytes, int64_t n_bytes, void* d_workspace, int64_t workspace_bytes,
                              ms_scan_summary* d_summary, void* stream) {
    return ms_scan_impl(d_bytes, n_bytes, d_workspace, workspace_bytes, d_summary, stream, true);
}

// The header lines of the second section start right after the first blank row, a position only the
// device knows when ms_scan finishes: this copies the bytes that follow it next to the summary, so
// the host gets summary and header text in one device->host transfer instead of two round trips.
__global__ void ms_peek_kernel(const uint8_t* __restrict__ src, int64_t n, const ms_scan_summary* __restrict__ summary,
                               int which, uint8_t* __restrict__ out, int max_bytes, int64_t* __restrict__ info) {
    int64_t from = -1, count = 0;
    if ((uint32_t)which < summary->n_reported && summary->blank_end[which] >= 0) {
        from = summary->blank_end[which] + 1;
        if (from > n) from = n;
        count = n - from < (int64_t)max_bytes ? n - from : (int64_t)max_bytes;
    }
    for (int64_t i = threadIdx.x; i < count; i += blockDim.x) out[i] = src[from + i];
    if (threadIdx.x == 0) {
        info[0] = from;
        info[1] = count;
    }
}

extern "C" int ms_peek_after_blank(const uint8_t* d_bytes, int64_t n_bytes, const ms_scan_summary* d_summary,
                                   int32_t which, uint8_t* d_out, int32_t max_bytes, int64_t* d_info, void* stream) {
    if (!d_bytes || !d_summary || !d_out || !d_info || n_bytes < 0 || max_bytes < 0 || which < 0 ||
        which >= MS_MAX_BLANK_ROWS)
        return MS_E_INVALID;
    ms_peek_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_bytes, n_bytes, d_summary, which, d_out, max_bytes, d_info);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_parse(const uint8_t* d_bytes, int64_t n_bytes, const void* d_workspace, const ms_section* h_sections,
                        int32_t n_sections, uint64_t* d_status, void* stream) {
    if (!d_bytes || !d_workspace || !d_status || n_bytes < 0) return MS_E_INVALID;
    if (n_sections < 0 || n_sections > MS_MAX_SECTIONS || (n_sections > 0 && !h_sections)) return MS_E_INVALID;
    if (((uintptr_t)d_bytes & 15) != 0) return MS_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    MS_CUDA_CHECK(cudaMemsetAsync(d_status, 0xFF, sizeof(uint64_t), st));
    const int64_t n_tiles = ms_num_tiles(n_bytes);
    MsSectionsArg arg;
    memset(&arg, 0, sizeof arg);
    arg.n = 0;
    for (int i = 0; i < n_sections; i++) {
        const ms_section& s = h_sections[i];
        if (s.row_end <= s.row_begin || s.num_cols <= 0) continue;  // nothing to parse
        if (!s.d_out && s.n_keep > 0) return MS_E_INVALID;
        if (s.stride < s.row_end - s.row_begin || s.n_keep < 0 || s.n_keep > s.num_cols) return MS_E_INVALID;
        if (s.num_cols <= 65535) {
            for (int g = 1; g <= PARSE_TAB_GROUPS; g++)
                arg.chunk_cnt[arg.n][g - 1] = (uint8_t)ms_chunk_table(g, s.num_cols, arg.chunk_tab[arg.n][g - 1]);
        }
        arg.s[arg.n++] = s;
    }
    if (n_tiles == 0 || arg.n == 0) return MS_OK;
    MS_CUDA_CHECK(cudaFuncSetAttribute(ms_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PARSE_SMEM));
    MsWorkspaceView v = ms_view(const_cast<void*>(d_workspace), ms_num_tiles(n_bytes < 1 ? 1 : n_bytes), n_bytes < 1 ? 1 : n_bytes);
    ms_parse_kernel<<<(unsigned)n_tiles, PARSE_THREADS, PARSE_SMEM, st>>>(d_bytes, n_bytes, v.term_prefix, v.masks, arg,
                                                                           (unsigned long long*)d_status);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}

// ===================================================================================================
// rows of any length
// ===================================================================================================
// The tiled kernels stage a tile and MS_MAX_ROW_BYTES beyond it; a file with a longer row (a Trajectories section of
// more than ~300 markers, or 800-digit numbers) is parsed by this pair instead: exact, unhurried, straight from
// global memory.  ms_row_index_kernel turns the scan's terminator masks into the byte offset of every row;
// ms_parse_rows_kernel gives every data row to one thread, which walks its fields by the scan's comma masks (so a
// quote-aware scan is honoured) and parses each with the general parser.  Same outputs and status word as ms_parse.
__global__ void __launch_bounds__(256)
    ms_row_index_kernel(const uint32_t* __restrict__ masks, long long n, const unsigned long long* __restrict__ term_prefix,
                        long long* __restrict__ row_start) {
    __shared__ int s_warp[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile = blockIdx.x;
    const long long seg0 = tile * (MS_TILE_BYTES / 16), n_seg = (n + 15) >> 4;
    constexpr int PER = MS_TILE_BYTES / 16 / 256;  // mask words per thread, consecutive
    uint32_t term[PER];
    int mine = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const long long sa = seg0 + (long long)tid * PER + k;
        term[k] = sa < n_seg ? (masks[sa] & 0xffffu) : 0u;
        mine += __popc(term[k]);
    }
    int inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int before = inc - mine;
    for (int w = 0; w < warp; w++) before += s_warp[w];
    long long j = (long long)term_prefix[tile] + before;  // ordinal of my first terminator in the file
    if (tile == 0 && tid == 0) row_start[0] = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        uint32_t t = term[k];
        while (t) {
            const int b = __ffs(t) - 1;
            t &= t - 1u;
            row_start[++j] = ((seg0 + (long long)tid * PER + k) << 4) + b + 1;  // row j starts after terminator j - 1
        }
    }
}

// float() of the csv field src[s, e) (quotes handled as csv.reader's excel dialect does, load_csv.py:30)
__device__ int ms_parse_span(const uint8_t* __restrict__ src, long long s, long long e, uint64_t* bits) {
    *bits = MS_NAN_BITS;
    if (s >= e) return MS_PARSE_OK;  // empty: None -> NaN
    if (src[s] != '"') return ms_parse_field(src + s, src + e, bits);
    uint8_t buf[MS_QUOTED_MAX];
    int n = 0;
    bool open = true;
    for (long long q = s + 1; q < e; q++) {
        const unsigned c = src[q];
        if (open && c == '"') {
            if (q + 1 < e && src[q + 1] == '"')
                q++;  // "" inside quotes: one literal quote
            else {
                open = false;
                continue;
            }
        }
        if (n >= MS_QUOTED_MAX) return MS_PARSE_BAD;
        buf[n++] = (uint8_t)c;
    }
    if (n == 0) return MS_PARSE_OK;
    return ms_parse_field(buf, buf + n, bits);
}

__global__ void __launch_bounds__(128)
    ms_parse_rows_kernel(const uint8_t* __restrict__ src, long long n, const uint32_t* __restrict__ masks,
                         const long long* __restrict__ row_start, long long n_terminators, const MsSectionsArg secs,
                         unsigned long long* __restrict__ status) {
    const long long n_rows = n_terminators + ((n_terminators == 0 ? n > 0 : row_start[n_terminators] < n) ? 1 : 0);
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        int si = -1;
        for (int i = 0; i < secs.n; i++)
            if (r >= secs.s[i].row_begin && r < secs.s[i].row_end) si = i;
        if (si < 0) continue;
        const int ncols = secs.s[si].num_cols, n_keep = secs.s[si].n_keep;
        double* const out = secs.s[si].d_out + (r - secs.s[si].row_begin);
        const long long stride = secs.s[si].stride;
        long long p = row_start[r];
        // one past the row's last content byte: before its terminator ("\n", "\r\n" or "\r"), or the end of the file
        long long end = r < n_terminators ? row_start[r + 1] - 1 : n;
        if (r < n_terminators && src[end] == '\n' && end > p && src[end - 1] == '\r') end--;
        bool done = false;
        for (int c = 0; c < ncols; c++) {
            uint64_t bits = MS_NAN_BITS;
            if (!done) {
                // the field ends at the next comma the scan recorded, or with the row
                long long e = end;
                for (long long w = p >> 4; (w << 4) < end; w++) {
                    uint32_t cm = masks[w] >> 16;
                    if (w == (p >> 4)) cm &= ~((1u << (p & 15)) - 1u);
                    if (cm) {
                        const long long at = (w << 4) + __ffs(cm) - 1;
                        if (at < end) e = at;
                        break;
                    }
                }
                const int st = ms_parse_span(src, p, e, &bits);
                if (st != MS_PARSE_OK) {
                    bits = MS_NAN_BITS;
                    atomicMin(status, ((unsigned long long)p << 3) |
                                          (st == MS_PARSE_NONASCII ? MS_ERR_KIND_NON_ASCII : MS_ERR_KIND_BAD_FLOAT));
                }
                done = e >= end;
                p = e + 1;
            }
            const int ch = c - 2;
            if (ch >= 0 && ch < n_keep) out[(long long)ch * stride] = ms_bits_to_double(bits);
        }
    }
}

extern "C" int64_t ms_parse_long_workspace_bytes(int64_t n_terminators) { return (n_terminators + 2) * 8; }

extern "C" int ms_parse_long(const uint8_t* d_bytes, int64_t n_bytes, const void* d_workspace, const ms_section* h_sections,
                             int32_t n_sections, int64_t n_terminators, void* d_rows, uint64_t* d_status, void* stream) {
    if (!d_bytes || !d_workspace || !d_status || !d_rows || n_bytes < 0 || n_terminators < 0) return MS_E_INVALID;
    if (n_sections < 0 || n_sections > MS_MAX_SECTIONS || (n_sections > 0 && !h_sections)) return MS_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    MS_CUDA_CHECK(cudaMemsetAsync(d_status, 0xFF, sizeof(uint64_t), st));
    MsSectionsArg arg;
    memset(&arg, 0, sizeof arg);
    for (int i = 0; i < n_sections; i++) {
        const ms_section& s = h_sections[i];
        if (s.row_end <= s.row_begin || s.num_cols <= 0) continue;
        if (!s.d_out && s.n_keep > 0) return MS_E_INVALID;
        if (s.stride < s.row_end - s.row_begin || s.n_keep < 0 || s.n_keep > s.num_cols) return MS_E_INVALID;
        arg.s[arg.n++] = s;
    }
    const int64_t n_tiles = ms_num_tiles(n_bytes);
    if (n_tiles == 0 || arg.n == 0) return MS_OK;
    MsWorkspaceView v = ms_view(const_cast<void*>(d_workspace), n_tiles, n_bytes);
    ms_row_index_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(v.masks, n_bytes, v.term_prefix, (long long*)d_rows);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    const int64_t n_rows = n_terminators + 1;
    const unsigned blocks = (unsigned)((n_rows + 127) / 128 < 148 * 16 ? (n_rows + 127) / 128 : 148 * 16);
    ms_parse_rows_kernel<<<blocks, 128, 0, st>>>(d_bytes, n_bytes, v.masks, (const long long*)d_rows, n_terminators, arg,
                                                 (unsigned long long*)d_status);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}
