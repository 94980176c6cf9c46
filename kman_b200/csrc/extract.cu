// K1+K2: LUT encode + rolling-window key extraction.
//
// Replaces Sequence.yield_kmers (kmermaid/seq.py:284-328), its reverse-complement variant
// (seq.py:245-282) and the second alphabet filter (seq.py:500-509 via batcher.py:559-561).
//
// Narrow stream: a tile of 16-base words is staged in shared memory as 2-bit codes plus two
// bitmasks (bad-for-narrow, invalid-for-any-stream); every thread then produces the keys of
// its consecutive window starts with funnel shifts out of 64-bit big-endian code words, so a
// base is looked up in the LUT exactly once.  Valid keys are compacted in position order
// (block scan + decoupled look-back across tiles), staged in shared memory and written with
// fully coalesced stores.
#include "common.cuh"

namespace kmg {

constexpr int EX_BLOCK_DEFAULT = 256;  // threads of the position-ordered extraction (MODE 0)
constexpr int EX_HALO_WORDS = 8;  // 16-base words past the tile that masks/codes may touch

struct ExtractParams {
    const uint8_t* bases;
    uint64_t n_bases;
    uint64_t win_begin, win_end;
    uint64_t first_tile;  // absolute tile index (tile = EX_BLOCK*PPT window starts, buffer-aligned)
    int k;
    const uint8_t* lut;
    const uint8_t* comp16;
    void* keys_out;
    void* vals_out;
    uint64_t pos_offset;
    unsigned long long* counts;  // [0] emitted, [1] wide windows seen
    uint64_t* tile_state;
    uint32_t* ticket;
    uint32_t* err;
    // fused digit histograms (narrow stream): hs[o][x] counts the plain 4-mers x found at window
    // start + hist_off[o]; hs[n_off] is the unshifted 4-mer histogram G (see hist_finalize_kernel)
    unsigned long long* hs;
    int n_off;
    int hist_off[2 * MAX_PASSES];
    // MODE 2 (fused first prefix pass): every key goes straight to the region of its digit
    // (key >> digit_shift) & 255; cursors[256] hold the next free element of every region
    unsigned long long* cursors;
    int digit_shift;
};

constexpr int EX_MAX_OFF = 2 * MAX_PASSES;

// 4 consecutive bases starting at base e (0 <= e <= 92) of the 96-base big-endian string X0:X1:X2
__device__ __forceinline__ uint32_t mer4_at(uint64_t X0, uint64_t X1, uint64_t X2, int e) {
    const int w = e >> 5, off = 2 * (e & 31);
    const uint64_t a = w == 0 ? X0 : (w == 1 ? X1 : X2);
    const uint64_t b = w == 0 ? X1 : (w == 1 ? X2 : 0ull);
    const uint64_t v = off == 0 ? a : ((a << off) | (b >> (64 - off)));
    return (uint32_t)(v >> 56);
}
// are the 4 bases starting at base e (0 <= e <= 124) of the 128-base mask string B0:B1 all plain?
__device__ __forceinline__ bool plain4_at(uint64_t B0, uint64_t B1, int e) {
    uint64_t v;
    if (e == 0) v = B0;
    else if (e < 64) v = (B0 << e) | (B1 >> (64 - e));
    else v = B1 << (e - 64);
    return (v >> 60) == 0;
}

__device__ __forceinline__ uint64_t rev2_64(uint64_t x) {
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}
__device__ __forceinline__ uint64_t rc_key(uint64_t key, int k) { return rev2_64(~key) >> (64 - 2 * k); }
__device__ __forceinline__ u128 rc_key(const u128& key, int k) {
    uint64_t hi = rev2_64(~key.lo), lo = rev2_64(~key.hi);
    int s = 128 - 2 * k;  // 0 <= s <= 62 because k >= 33
    u128 r;
    if (s == 0) {
        r.lo = lo;
        r.hi = hi;
    } else {
        r.lo = (lo >> s) | (hi << (64 - s));
        r.hi = hi >> s;
    }
    return r;
}

template <typename ValT>
__device__ __forceinline__ ValT make_val(uint64_t pos, uint32_t strand) {
    return (ValT)((pos << 1) | strand);
}

// KeyT: uint64_t (k<=32) or u128 (33<=k<=64).  PPT: window starts per thread (8 or 16).
// MODE 0: keys (and payload) of the valid windows in position order (kmg_extract).
// MODE 1: nothing is emitted -- only the 4-mer histograms (HIST) and the window counts: the
//         pre-pass of the fused pipeline (pipeline.cu), which needs every prefix pass' histogram
//         before the first key is placed.
// MODE 2: extraction fused with the FIRST prefix pass of the hybrid sort: that pass may place the
//         keys of one digit in any order (the local sort orders everything below the prefix), so
//         a tile ranks its keys with the return values of shared-memory atomics, reserves its
//         slots in the 256 digit regions with one global atomic per digit and writes digit runs
//         -- no compaction in position order, no tile prefix, no waiting for other tiles, and the
//         keys never make the extra round trip through HBM in extraction order.
template <typename KeyT, int PPT, bool RC, int VAL_BYTES, bool HIST, int MODE, int EX_BLOCK>
__global__ void __launch_bounds__(EX_BLOCK) extract_narrow_kernel(const ExtractParams p) {
    static_assert(MODE != 1 || HIST, "the pre-pass exists for its histograms");
    static_assert(MODE != 2 || !HIST, "the fused pass takes its histograms from the pre-pass");
    constexpr int TILE = EX_BLOCK * PPT;
    constexpr int WORDS = TILE / 16 + EX_HALO_WORDS;
    constexpr int OUT_PER_WIN = RC ? 2 : 1;
    using ValT = typename std::conditional<VAL_BYTES == 4, uint32_t, uint64_t>::type;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t STAGE = TILE * OUT_PER_WIN;
    constexpr uint32_t STAGE_PAD = STAGE + STAGE / 8 + 8;  // room for pad_idx() of either width
    KeyT* s_keys = reinterpret_cast<KeyT*>(smem_raw);
    ValT* s_vals = reinterpret_cast<ValT*>(smem_raw + sizeof(KeyT) * STAGE_PAD);
    constexpr int KB = sizeof(KeyT);
    constexpr int VB = VAL_BYTES == 0 ? 8 : VAL_BYTES;
    __shared__ uint32_t s_codes[WORDS];
    __shared__ uint16_t s_bad[WORDS];
    __shared__ uint16_t s_inv[WORDS];
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_scan[EX_BLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    __shared__ uint32_t s_wide;
    __shared__ uint32_t s_g[HIST ? 256 : 1];                 // 4-mer histogram of the tile's window starts
    uint32_t* s_c = reinterpret_cast<uint32_t*>(smem_raw);  // [n_off][256] corrections for skipped windows;
                                                            // aliases the key staging buffer, used before it
    __shared__ uint32_t s_any_skipped;
    __shared__ uint32_t s_cnt[MODE == 2 ? 256 : 1];            // keys per digit in this tile
    __shared__ uint32_t s_off[MODE == 2 ? 256 : 1];            // first staging slot of every digit
    __shared__ unsigned long long s_gb[MODE == 2 ? 256 : 1];   // global slot of staging slot i = s_gb[digit] + i

    const int t = threadIdx.x;
    if (t == 0) {
        // (only the position-ordered mode waits for earlier tiles: it needs launch-ordered ids)
        s_tile = MODE == 0 ? atomicAdd(p.ticket, 1u) : blockIdx.x;
        s_wide = 0;
        s_any_skipped = 0;
    }
    if (t < 256) {
        if (HIST) s_g[t] = 0;
        if (MODE == 2) s_cnt[t] = 0;
        s_lut[t] = p.lut[t];
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_pos = (p.first_tile + tile) * (uint64_t)TILE;

    // ---- K1: encode 16 bases per word ----------------------------------------------------
    for (int w = t; w < WORDS; w += EX_BLOCK) {
        const uint64_t pos = tile_pos + (uint64_t)w * 16;
        uint32_t b[4];
        if (pos + 16 <= p.n_bases) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.bases + pos));
            b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t x = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint64_t pp = pos + q * 4 + r;
                    // 0xFF is never in an alphabet built by kmg_build_lut; the explicit
                    // range test below does not rely on that.
                    const uint32_t byte = (pp < p.n_bases) ? p.bases[pp] : 0xFFu;
                    x |= byte << (8 * r);
                }
                b[q] = x;
            }
        }
        const int nvalid = pos + 16 <= p.n_bases ? 16 : (pos < p.n_bases ? (int)(p.n_bases - pos) : 0);
        uint32_t codes = 0, bad = 0, inv = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t byte = (b[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            uint32_t e = s_lut[byte];
            if (j >= nvalid) e = KMG_LUT_INVALID;
            codes |= (e & 3u) << (30 - 2 * j);
            bad |= ((e >> 6) ? 1u : 0u) << (15 - j);
            inv |= (e >> 7) << (15 - j);
        }
        s_codes[w] = codes;
        s_bad[w] = (uint16_t)bad;
        s_inv[w] = (uint16_t)inv;
    }
    __syncthreads();

    // ---- K2: rolling windows ---------------------------------------------------------------
    const int k = p.k;
    const int wbase = (t * PPT) >> 4;
    const int joff = (t * PPT) & 15;
    const uint64_t X0 = ((uint64_t)s_codes[wbase] << 32) | s_codes[wbase + 1];
    const uint64_t X1 = ((uint64_t)s_codes[wbase + 2] << 32) | s_codes[wbase + 3];
    const uint64_t X2 = ((uint64_t)s_codes[wbase + 4] << 32) | s_codes[wbase + 5];
    const uint64_t B0 = ((uint64_t)s_bad[wbase] << 48) | ((uint64_t)s_bad[wbase + 1] << 32) |
                        ((uint64_t)s_bad[wbase + 2] << 16) | s_bad[wbase + 3];
    const uint64_t B1 = ((uint64_t)s_bad[wbase + 4] << 48) | ((uint64_t)s_bad[wbase + 5] << 32) |
                        ((uint64_t)s_bad[wbase + 6] << 16) | s_bad[wbase + 7];
    const uint64_t I0 = ((uint64_t)s_inv[wbase] << 48) | ((uint64_t)s_inv[wbase + 1] << 32) |
                        ((uint64_t)s_inv[wbase + 2] << 16) | s_inv[wbase + 3];
    const uint64_t I1 = ((uint64_t)s_inv[wbase + 4] << 48) | ((uint64_t)s_inv[wbase + 5] << 32) |
                        ((uint64_t)s_inv[wbase + 6] << 16) | s_inv[wbase + 7];

    // validity of my PPT windows first: the tile's count is published as early as possible and
    // the prefix over the earlier tiles is resolved only after the keys have been built
    const uint64_t first_pos = tile_pos + (uint64_t)t * PPT;
    uint32_t in_range_mask;
    {   // window starts [win_begin, win_end) among my PPT positions
        const uint64_t lo = p.win_begin > first_pos ? p.win_begin - first_pos : 0;
        const uint64_t hi = p.win_end > first_pos ? p.win_end - first_pos : 0;
        const uint32_t l = lo < PPT ? (uint32_t)lo : PPT, h = hi < PPT ? (uint32_t)hi : PPT;
        in_range_mask = h > l ? (((1u << (h - l)) - 1u) << l) : 0u;
    }
    // fast path: no base that is bad for the narrow stream among the PPT+k-1 bases my windows span
    bool span_clean;
    {
        const uint64_t W0 = joff == 0 ? B0 : ((B0 << joff) | (B1 >> (64 - joff)));
        const int span = PPT + k - 1;  // 17 .. 79
        if (span <= 64) span_clean = (W0 >> (64 - span)) == 0;
        else span_clean = W0 == 0 && ((B1 << joff) >> (128 - span)) == 0;
    }
    uint32_t vf = 0, nwide = 0;
    if (span_clean) {
        vf = in_range_mask;
    } else {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int je = joff + j;  // 0..15
            const uint64_t bm = (je == 0 ? B0 : ((B0 << je) | (B1 >> (64 - je)))) >> (64 - k);
            const uint64_t im = (je == 0 ? I0 : ((I0 << je) | (I1 >> (64 - je)))) >> (64 - k);
            const bool in_range = (in_range_mask >> j) & 1u;
            if (in_range && bm == 0) vf |= 1u << j;
            if (in_range && im == 0 && bm != 0) ++nwide;
        }
    }
    const uint32_t cnt = __popc(vf);
    uint32_t total = 0, excl = 0;
    if constexpr (MODE != 2) excl = block_excl_scan<EX_BLOCK, uint32_t>(cnt, s_scan, total);
    if constexpr (MODE == 0) {
        if (t == 0) tile_prefix_publish(p.tile_state, tile, (uint64_t)total * OUT_PER_WIN);
    }
    if (MODE != 2 && nwide) atomicAdd(&s_wide, nwide);

    if constexpr (HIST) {
        // G: one count per window start whose first 4 bases are plain.  Digit p of a valid
        // window's key IS the 4-mer at window start + (k-4-4p), so every pass' histogram is G
        // over a shifted range minus the contributions of the skipped windows.
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int je = joff + j;
            if (((in_range_mask >> j) & 1u) && (span_clean || plain4_at(B0, B1, je)))
                atomicAdd(&s_g[mer4_at(X0, X1, X2, je)], 1u);
        }
        const uint32_t skipped = in_range_mask & ~vf;
        if (skipped) s_any_skipped = 1;  // the correction table is zeroed lazily: most tiles never touch it
        __syncthreads();
        if (s_any_skipped) {
            for (int i = t; i < p.n_off * 256; i += EX_BLOCK) s_c[i] = 0;
            __syncthreads();
            if (skipped) {
                for (int j = 0; j < PPT; ++j) {
                    if (!((skipped >> j) & 1u)) continue;
                    for (int o = 0; o < p.n_off; ++o) {
                        const int e = joff + j + p.hist_off[o];
                        if (plain4_at(B0, B1, e)) atomicAdd(&s_c[o * 256 + mer4_at(X0, X1, X2, e)], 1u);
                    }
                }
            }
            __syncthreads();
            for (int i = t; i < p.n_off * 256; i += EX_BLOCK) {
                const uint32_t c = s_c[i];
                if (c) atomicAdd(&p.hs[i], 0ull - (unsigned long long)c);
            }
            __syncthreads();  // the staging buffer is about to be reused for the keys
        }
        if (t < 256) {
            const uint32_t c = s_g[t];
            if (c) atomicAdd(&p.hs[p.n_off * 256 + t], (unsigned long long)c);
        }
    }
    if constexpr (MODE == 1) {  // the pre-pass stops here: window counts only
        __syncthreads();
        if (t == 0) {
            if (total) atomicAdd(&p.counts[0], (unsigned long long)total * OUT_PER_WIN);
            if (s_wide) atomicAdd(&p.counts[1], (unsigned long long)s_wide);
        }
        return;
    }

    // key of my j-th window start
    auto build_key = [&](int j) -> KeyT {
        const int je = joff + j;
        const int s2 = 2 * je;
        const uint64_t hi = je == 0 ? X0 : ((X0 << s2) | (X1 >> (64 - s2)));
        KeyT key;
        if constexpr (sizeof(KeyT) == 8) {
            key = hi >> (64 - 2 * k);
        } else {
            const uint64_t lo = je == 0 ? X1 : ((X1 << s2) | (X2 >> (64 - s2)));
            const int s = 128 - 2 * k;
            if (s == 0) {
                key.lo = lo;
                key.hi = hi;
            } else {
                key.lo = (lo >> s) | (hi << (64 - s));
                key.hi = hi >> s;
            }
        }
        return key;
    };

    if constexpr (MODE == 2) {
        // ---- fused first prefix pass: rank inside the tile = what the digit counter returned ----------
        const int dsh = p.digit_shift;
        uint32_t wh[PPT * OUT_PER_WIN];  // digit | rank << 8
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            if (!(vf & (1u << j))) continue;
            const KeyT key = build_key(j);
            const uint32_t d0 = key_digit(key, dsh, 255u);
            wh[j * OUT_PER_WIN] = d0 | (atomicAdd(&s_cnt[d0], 1u) << 8);
            if constexpr (RC) {
                const uint32_t d1 = key_digit(rc_key(key, k), dsh, 255u);
                wh[j * 2 + 1] = d1 | (atomicAdd(&s_cnt[d1], 1u) << 8);
            }
        }
        __syncthreads();
        // one thread per digit: reserve the tile's slots in the digit's region (any order is fine,
        // so a plain atomic cursor does it -- no tile waits for another one)
        const uint32_t c = t < 256 ? s_cnt[t] : 0u;
        unsigned long long g = 0;
        if (c) g = atomicAdd(&p.cursors[t], (unsigned long long)c);
        uint32_t n_tile;
        const uint32_t off = block_excl_scan<EX_BLOCK, uint32_t>(c, s_scan, n_tile);
        if (t < 256) {
            s_off[t] = off;
            s_gb[t] = g - off;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            if (!(vf & (1u << j))) continue;
            const KeyT key = build_key(j);
            const uint64_t pos = tile_pos + (uint64_t)t * PPT + j + p.pos_offset;
            const uint32_t w0 = wh[j * OUT_PER_WIN];
            const uint32_t i0 = s_off[w0 & 255u] + (w0 >> 8);
            s_keys[i0] = key;
            if constexpr (VAL_BYTES != 0) s_vals[i0] = make_val<ValT>(pos, 0);
            if constexpr (RC) {
                const uint32_t w1 = wh[j * 2 + 1];
                const uint32_t i1 = s_off[w1 & 255u] + (w1 >> 8);
                s_keys[i1] = rc_key(key, k);
                if constexpr (VAL_BYTES != 0) s_vals[i1] = make_val<ValT>(pos, 1);
            }
        }
        __syncthreads();
        KeyT* keys_out = reinterpret_cast<KeyT*>(p.keys_out);
        ValT* vals_out = reinterpret_cast<ValT*>(p.vals_out);
        for (uint32_t i = t; i < n_tile; i += EX_BLOCK) {
            const KeyT key = s_keys[i];
            const unsigned long long at = s_gb[key_digit(key, dsh, 255u)] + i;
            keys_out[at] = key;
            if constexpr (VAL_BYTES != 0) vals_out[at] = s_vals[i];
        }
        return;
    }

    // ---- keys of the valid windows, compacted in position order into the staging buffer -----
    uint32_t r = excl;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        if (!(vf & (1u << j))) continue;
        const KeyT key = build_key(j);
        const uint64_t pos = tile_pos + (uint64_t)t * PPT + j + p.pos_offset;
        s_keys[pad_idx<KB>(r * OUT_PER_WIN)] = key;
        if constexpr (VAL_BYTES != 0) s_vals[pad_idx<VB>(r * OUT_PER_WIN)] = make_val<ValT>(pos, 0);
        if constexpr (RC) {
            s_keys[pad_idx<KB>(r * 2 + 1)] = rc_key(key, k);
            if constexpr (VAL_BYTES != 0) s_vals[pad_idx<VB>(r * 2 + 1)] = make_val<ValT>(pos, 1);
        }
        ++r;
    }
    __syncthreads();
    if (t < 32) {
        const uint64_t e = tile_prefix_resolve_warp(p.tile_state, gridDim.x, tile, (uint64_t)total * OUT_PER_WIN, p.err);
        if (t == 0) s_base = e;
    }
    __syncthreads();
    const uint64_t base = s_base;
    const uint32_t n_out = total * OUT_PER_WIN;
    KeyT* keys_out = reinterpret_cast<KeyT*>(p.keys_out) + base;
    for (uint32_t i = t; i < n_out; i += EX_BLOCK) keys_out[i] = s_keys[pad_idx<KB>(i)];
    if constexpr (VAL_BYTES != 0) {
        ValT* vals_out = reinterpret_cast<ValT*>(p.vals_out) + base;
        for (uint32_t i = t; i < n_out; i += EX_BLOCK) vals_out[i] = s_vals[pad_idx<VB>(i)];
    }
    if (t == 0) {
        if (s_wide) atomicAdd(&p.counts[1], (unsigned long long)s_wide);
        if (tile == gridDim.x - 1) p.counts[0] = base + n_out;
    }
}

// ---- wide stream: windows that pass the alphabet test but contain a non-plain symbol -----
// Rare by construction (N runs, IUPAC codes), so the kernel is simple: 4 consecutive window
// starts per thread, bytes staged in shared memory, 4-bit ASCII-rank codes, 128-bit keys.
constexpr int EXW_BLOCK = 256;
constexpr int EXW_PPT = 4;
constexpr int EXW_TILE = EXW_BLOCK * EXW_PPT;

// LIMBS: 64-bit limbs of a key, 2 (k <= 32, u128) or 4 (k <= 64, u256); limb 0 is the least significant
template <int LIMBS>
struct WideKey {
    uint64_t w[LIMBS];
    __device__ __forceinline__ void push(uint32_t code) {  // key = key << 4 | code
#pragma unroll
        for (int i = LIMBS - 1; i > 0; --i) w[i] = (w[i] << 4) | (w[i - 1] >> 60);
        w[0] = (w[0] << 4) | code;
    }
    __device__ __forceinline__ uint32_t pop() {  // lowest nibble; key >>= 4
        const uint32_t c = (uint32_t)w[0] & 0xFu;
#pragma unroll
        for (int i = 0; i < LIMBS - 1; ++i) w[i] = (w[i] >> 4) | (w[i + 1] << 60);
        w[LIMBS - 1] >>= 4;
        return c;
    }
    __device__ __forceinline__ void set_nibble(int q, uint32_t c) { w[q >> 4] |= (uint64_t)c << (4 * (q & 15)); }
};

template <bool RC, int VAL_BYTES, int LIMBS>
__global__ void __launch_bounds__(EXW_BLOCK) extract_wide_kernel(const ExtractParams p) {
    using ValT = typename std::conditional<VAL_BYTES == 4, uint32_t, uint64_t>::type;
    using KeyT = WideKey<LIMBS>;  // same layout as u128 / u256
    constexpr int OUT_PER_WIN = RC ? 2 : 1;
    __shared__ uint8_t s_e[EXW_TILE + 64];  // LUT entries of the tile bytes (+ halo: k <= 64)
    __shared__ uint8_t s_lut[256];
    __shared__ uint8_t s_comp[16];
    __shared__ uint32_t s_scan[EXW_BLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;

    const int t = threadIdx.x;
    if (t == 0) s_tile = atomicAdd(p.ticket, 1u);
    s_lut[t] = p.lut[t];
    if (t < 16) s_comp[t] = p.comp16[t];
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_pos = (p.first_tile + tile) * (uint64_t)EXW_TILE;
    for (int i = t; i < EXW_TILE + 64; i += EXW_BLOCK) {
        const uint64_t pos = tile_pos + i;
        s_e[i] = pos < p.n_bases ? s_lut[p.bases[pos]] : (uint8_t)KMG_LUT_INVALID;
    }
    __syncthreads();

    const int k = p.k;
    KeyT keys[EXW_PPT];
    uint32_t vf = 0;
#pragma unroll
    for (int j = 0; j < EXW_PPT; ++j) {
        const int o = t * EXW_PPT + j;
        const uint64_t pos = tile_pos + o;
        uint32_t any_inv = 0, any_nonplain = 0;
        KeyT key;
#pragma unroll
        for (int i = 0; i < LIMBS; ++i) key.w[i] = 0;
        for (int q = 0; q < k; ++q) {
            const uint32_t e = s_e[o + q];
            any_inv |= e & 0x80u;
            any_nonplain |= e & 0x40u;
            key.push((e >> 2) & 0xFu);
        }
        keys[j] = key;
        if (pos >= p.win_begin && pos < p.win_end && !any_inv && any_nonplain) vf |= 1u << j;
    }
    const uint32_t cnt = __popc(vf);
    uint32_t total;
    const uint32_t excl = block_excl_scan<EXW_BLOCK, uint32_t>(cnt, s_scan, total);
    if (t < 32) {
        const uint64_t e = tile_prefix_exclusive_warp(p.tile_state, gridDim.x, tile, (uint64_t)total * OUT_PER_WIN, p.err);
        if (t == 0) s_base = e;
    }
    __syncthreads();
    const uint64_t base = s_base;
    KeyT* keys_out = reinterpret_cast<KeyT*>(p.keys_out);
    ValT* vals_out = reinterpret_cast<ValT*>(p.vals_out);
    uint64_t r = base + (uint64_t)excl * OUT_PER_WIN;
#pragma unroll
    for (int j = 0; j < EXW_PPT; ++j) {
        if (!(vf & (1u << j))) continue;
        const uint64_t pos = tile_pos + t * EXW_PPT + j + p.pos_offset;
        keys_out[r] = keys[j];
        if constexpr (VAL_BYTES != 0) vals_out[r] = make_val<ValT>(pos, 0);
        ++r;
        if constexpr (RC) {
            // reverse complement: symbol q of the window (nibble k-1-q of the key) becomes nibble q
            KeyT f = keys[j], rcv;
#pragma unroll
            for (int i = 0; i < LIMBS; ++i) rcv.w[i] = 0;
            for (int q = k - 1; q >= 0; --q) rcv.set_nibble(q, s_comp[f.pop()]);
            keys_out[r] = rcv;
            if constexpr (VAL_BYTES != 0) vals_out[r] = make_val<ValT>(pos, 1);
            ++r;
        }
    }
    if (t == 0 && tile == gridDim.x - 1) p.counts[0] = base + (uint64_t)total * OUT_PER_WIN;
}

template <typename KeyT, int PPT, bool RC, int VB, bool HIST, int MODE = 0, int BLOCK = EX_BLOCK_DEFAULT>
static int launch_narrow(const ExtractParams& p, uint32_t n_tiles, cudaStream_t st) {
    constexpr int TILE = BLOCK * PPT;
    constexpr uint32_t STAGE = TILE * (RC ? 2 : 1);
    constexpr uint32_t STAGE_PAD = STAGE + STAGE / 8 + 8;
    // (the pre-pass stages no keys: its dynamic shared memory only holds the correction tables)
    const size_t smem = MODE == 1 ? (size_t)8 * 256 * sizeof(uint32_t)
                                  : (sizeof(KeyT) + (VB == 4 ? 4 : (VB == 8 ? 8 : 0))) * (size_t)STAGE_PAD;
    auto kern = extract_narrow_kernel<KeyT, PPT, RC, VB, HIST, MODE, BLOCK>;
    KMG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_tiles, BLOCK, smem, st>>>(p);
    KMG_LAUNCH_CHECK();
    return KMG_OK;
}

template <typename KeyT, int PPT, bool HIST>
static int dispatch_narrow2(const ExtractParams& p, uint32_t n_tiles, int rc, int vb, cudaStream_t st) {
    if (rc) {
        if (vb == 0) return launch_narrow<KeyT, PPT, true, 0, HIST>(p, n_tiles, st);
        if (vb == 4) return launch_narrow<KeyT, PPT, true, 4, HIST>(p, n_tiles, st);
        return launch_narrow<KeyT, PPT, true, 8, HIST>(p, n_tiles, st);
    }
    if (vb == 0) return launch_narrow<KeyT, PPT, false, 0, HIST>(p, n_tiles, st);
    if (vb == 4) return launch_narrow<KeyT, PPT, false, 4, HIST>(p, n_tiles, st);
    return launch_narrow<KeyT, PPT, false, 8, HIST>(p, n_tiles, st);
}
template <typename KeyT, int PPT>
static int dispatch_narrow(const ExtractParams& p, uint32_t n_tiles, int rc, int vb, bool hist, cudaStream_t st) {
    return hist ? dispatch_narrow2<KeyT, PPT, true>(p, n_tiles, rc, vb, st)
                : dispatch_narrow2<KeyT, PPT, false>(p, n_tiles, rc, vb, st);
}

// fused first prefix pass (MODE 2).  Key-only forward extraction stages 8192 keys per tile in 512
// threads (digit runs of ~32 keys = 256 bytes); with payload or reverse complements 256 threads
// keep the staging buffer at 64 KB.
template <typename KeyT, int PPT>
static int dispatch_scatter_digit(const ExtractParams& p, uint64_t n_win_tiles_256, int rc, int vb, cudaStream_t st) {
    const uint32_t t256 = (uint32_t)n_win_tiles_256;
    if (rc) {
        if (vb == 0) return launch_narrow<KeyT, PPT, true, 0, false, 2, 256>(p, t256, st);
        if (vb == 4) return launch_narrow<KeyT, PPT, true, 4, false, 2, 256>(p, t256, st);
        return launch_narrow<KeyT, PPT, true, 8, false, 2, 256>(p, t256, st);
    }
    if (vb == 4) return launch_narrow<KeyT, PPT, false, 4, false, 2, 256>(p, t256, st);
    if (vb == 8) return launch_narrow<KeyT, PPT, false, 8, false, 2, 256>(p, t256, st);
    return KMG_ERR_ARG;  // (key-only forward: 512 threads, see extract_digit_scatter)
}

// ---- fused digit histograms: edges and final assembly -----------------------------------------------
__device__ __forceinline__ uint32_t rc4(uint32_t x) {  // reverse complement of a packed 4-mer
    x = ~x & 0xFFu;
    return ((x & 3u) << 6) | ((x & 0xCu) << 2) | ((x >> 2) & 0xCu) | (x >> 6);
}

// hs[o] so far = G over [win_begin, win_end) minus the skipped windows' 4-mers at offset o.  The
// histogram over the SHIFTED range [win_begin+o, win_end+o) differs from G by the two edges.
__global__ void hist_edges_kernel(const ExtractParams p) {
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = p.lut[threadIdx.x];
    __syncthreads();
    for (int o = 0; o < p.n_off; ++o) {
        const int off = p.hist_off[o];
        for (int j = threadIdx.x; j < 2 * off; j += blockDim.x) {
            const bool tail = j >= off;  // + 4-mers at [win_end, win_end+off), - those at [win_begin, win_begin+off)
            const uint64_t q = (tail ? p.win_end : p.win_begin) + (uint64_t)(tail ? j - off : j);
            if (q + 4 > p.n_bases) continue;
            uint32_t x = 0;
            bool plain = true;
            for (int b = 0; b < 4; ++b) {
                const uint32_t e = s_lut[p.bases[q + b]];
                plain = plain && !(e & 0xC0u);
                x = (x << 2) | (e & 3u);
            }
            if (plain) atomicAdd(&p.hs[o * 256 + x], tail ? 1ull : 0ull - 1ull);
        }
    }
}

// digit histograms of the sort plan for 2k key bits from the per-offset 4-mer histograms
//   forward key, full digit p : 4-mer at window offset k-4-4p;   top digit (b bits): first b/2 bases
//   rc key,      full digit p : rc4 of the 4-mer at offset 4p;    top digit: rc of the last b/2 bases
// Rows 13..15 (u64 keys, k >= 12): histograms of the three TOP key bytes, i.e. the 4-mers at window
// offsets 8, 4, 0 (rc: rc4 of those at k-12, k-8, k-4), for the prefix passes of the hybrid sort;
// top_idx = index of the first of those extra offsets in hs (fwd 4, fwd 8, [rc k-8, rc k-12]) or -1.
__global__ void hist_finalize_kernel(const unsigned long long* __restrict__ hs, int n_off, int k, int rc, int top_idx,
                                     unsigned long long* __restrict__ hist_out) {
    const int P = (2 * k + 7) / 8;
    const int b = 2 * k - 8 * (P - 1);  // bits of the top digit: 2, 4, 6 or 8
    const unsigned long long* G = hs + (size_t)n_off * 256;
    const int x = threadIdx.x;  // 256 threads
    for (int pss = 0; pss < P; ++pss) hist_out[pss * 256 + x] = 0;
    __syncthreads();
    // offsets were laid out by the host as: [0 .. P-2] forward full digits, [P-1] forward top (offset 0),
    // then, if rc: [P .. 2P-2] rc full digits (offset 4p), [2P-1] rc top (offset k-4)
    for (int pss = 0; pss < P - 1; ++pss) {
        unsigned long long v = G[x] + hs[(size_t)pss * 256 + x];
        if (rc) v += G[rc4(x)] + hs[(size_t)(P + pss) * 256 + rc4(x)];
        hist_out[pss * 256 + x] = v;
    }
    // top digit: marginalise
    {
        const unsigned long long v = G[x] + hs[(size_t)(P - 1) * 256 + x];
        atomicAdd(&hist_out[(P - 1) * 256 + (x >> (8 - b))], v);
        if (rc) {
            const unsigned long long w = G[x] + hs[(size_t)(2 * P - 1) * 256 + x];
            // last b/2 bases of the window = low b bits of x; their reverse complement is the rc key's top digit
            const uint32_t low = x & ((1u << b) - 1u);
            const uint32_t r = rc4(low << (8 - b)) & ((1u << b) - 1u);
            atomicAdd(&hist_out[(P - 1) * 256 + r], w);
        }
    }
    if (top_idx >= 0) {
        // byte 15: offset 0 (hs row P-1) / rc k-4 (row 2P-1); bytes 14, 13: the extra offsets
        for (int j = 0; j < 3; ++j) {
            const int fi = j == 0 ? P - 1 : top_idx + (j - 1);
            unsigned long long v = G[x] + hs[(size_t)fi * 256 + x];
            if (rc) {
                const int ri = j == 0 ? 2 * P - 1 : top_idx + 2 + (j - 1);
                v += G[rc4(x)] + hs[(size_t)ri * 256 + rc4(x)];
            }
            hist_out[(15 - j) * 256 + x] = v;
        }
    }
}

// Pre-pass of the fused pipeline: histograms of the three TOP key bytes only (rows 0..2 = bits
// [2k-24,2k-16), [2k-16,2k-8), [2k-8,2k)), i.e. of the 4-mers at window offsets 8, 4, 0 (rc keys:
// rc4 of those at k-12, k-8, k-4) -- valid for 8- and 16-byte keys alike.  hs rows as laid out by
// extract_top_hist: [0..2] forward offsets 8, 4, 0, then (rc) [3..5] offsets k-12, k-8, k-4.
// plan[0] = number of keys, plan[1] = largest top-byte count (how skewed the keys are).
__global__ void hist_top_finalize_kernel(const unsigned long long* __restrict__ hs, int n_off, int rc,
                                         unsigned long long* __restrict__ top, unsigned long long* __restrict__ plan) {
    __shared__ unsigned long long s_sum[8], s_max[8];
    const unsigned long long* G = hs + (size_t)n_off * 256;
    const int x = threadIdx.x;  // 256 threads
    unsigned long long v_top = 0;
    for (int j = 0; j < 3; ++j) {
        unsigned long long v = G[x] + hs[(size_t)j * 256 + x];
        if (rc) v += G[rc4(x)] + hs[(size_t)(3 + j) * 256 + rc4(x)];
        top[j * 256 + x] = v;
        v_top = v;
    }
    unsigned long long sum = v_top, mx = v_top;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((x & 31) == 0) {
        s_sum[x >> 5] = sum;
        s_max[x >> 5] = mx;
    }
    __syncthreads();
    if (x == 0) {
        sum = 0;
        mx = 0;
        for (int i = 0; i < 8; ++i) {
            sum += s_sum[i];
            mx = max(mx, s_max[i]);
        }
        plan[0] = sum;
        plan[1] = mx;
    }
}

template <bool RC, int LIMBS>
static int dispatch_wide(const ExtractParams& p, uint32_t n_tiles, int vb, cudaStream_t st) {
    if (vb == 0) extract_wide_kernel<RC, 0, LIMBS><<<n_tiles, EXW_BLOCK, 0, st>>>(p);
    else if (vb == 4) extract_wide_kernel<RC, 4, LIMBS><<<n_tiles, EXW_BLOCK, 0, st>>>(p);
    else extract_wide_kernel<RC, 8, LIMBS><<<n_tiles, EXW_BLOCK, 0, st>>>(p);
    KMG_LAUNCH_CHECK();
    return KMG_OK;
}

static void fill_common(ExtractParams& p, const uint8_t* d_bases, uint64_t n_bases, uint64_t win_begin, uint64_t win_end,
                        uint64_t tile, int k, const uint8_t* d_lut256) {
    memset(&p, 0, sizeof(p));
    p.bases = d_bases;
    p.n_bases = n_bases;
    p.win_begin = win_begin;
    p.win_end = win_end;
    p.first_tile = win_begin / tile;
    p.k = k;
    p.lut = d_lut256;
}

size_t top_hist_workspace_bytes() { return sizeof(WsHeader) + (size_t)(6 + 1) * 256 * sizeof(uint64_t); }

// see common.cuh
int extract_top_hist(const uint8_t* d_bases, uint64_t n_bases, uint64_t win_begin, uint64_t win_end, int k, int rc,
                     const uint8_t* d_lut256, uint64_t* d_counts, unsigned long long* d_top, unsigned long long* d_plan,
                     void* d_ws, size_t ws_bytes, cudaStream_t st) {
    KMG_REQUIRE(k >= 12 && k <= 64, KMG_ERR_ARG, "the top-byte histograms need 12 <= k <= 64");
    KMG_REQUIRE(ws_bytes >= top_hist_workspace_bytes(), KMG_ERR_WS, "top-hist workspace too small");
    KMG_CUDA(cudaMemsetAsync(d_counts, 0, 2 * sizeof(uint64_t), st));
    KMG_CUDA(cudaMemsetAsync(d_ws, 0, top_hist_workspace_bytes(), st));
    constexpr uint64_t TILE = (uint64_t)EX_BLOCK_DEFAULT * 16;
    ExtractParams p;
    fill_common(p, d_bases, n_bases, win_begin, win_end, TILE, k, d_lut256);
    p.counts = reinterpret_cast<unsigned long long*>(d_counts);
    p.hs = reinterpret_cast<unsigned long long*>((char*)d_ws + sizeof(WsHeader));
    p.hist_off[0] = 8;
    p.hist_off[1] = 4;
    p.hist_off[2] = 0;
    p.n_off = 3;
    if (rc) {
        p.hist_off[3] = k - 12;
        p.hist_off[4] = k - 8;
        p.hist_off[5] = k - 4;
        p.n_off = 6;
    }
    if (win_end > win_begin) {
        const uint64_t n_tiles = (win_end + TILE - 1) / TILE - p.first_tile;
        KMG_REQUIRE(n_tiles < (1ull << 31), KMG_ERR_RANGE, "too many tiles");
        // (the histograms do not depend on the key width; one window per lane and 16 per thread
        // span at most 79 bases, which the 96-base code words cover for every k <= 64)
        const int rcode = launch_narrow<uint64_t, 16, false, 0, true, 1, EX_BLOCK_DEFAULT>(p, (uint32_t)n_tiles, st);
        if (rcode != KMG_OK) return rcode;
        hist_edges_kernel<<<1, 256, 0, st>>>(p);
        KMG_LAUNCH_CHECK();
    }
    hist_top_finalize_kernel<<<1, 256, 0, st>>>(p.hs, p.n_off, rc, d_top, d_plan);
    KMG_LAUNCH_CHECK();
    return KMG_OK;
}

int extract_digit_scatter(const uint8_t* d_bases, uint64_t n_bases, uint64_t win_begin, uint64_t win_end, int k, int rc,
                          const uint8_t* d_lut256, void* d_keys_out, int key_bytes, void* d_vals_out, int val_bytes,
                          uint64_t pos_offset, unsigned long long* d_cursors, int digit_shift, cudaStream_t st) {
    if (win_end == win_begin) return KMG_OK;
    const int ppt = key_bytes == 8 ? 16 : 8;
    const bool big = !rc && val_bytes == 0;  // 512-thread tiles
    const uint64_t tile = (uint64_t)(big ? 512 : 256) * ppt;
    ExtractParams p;
    fill_common(p, d_bases, n_bases, win_begin, win_end, tile, k, d_lut256);
    p.keys_out = d_keys_out;
    p.vals_out = d_vals_out;
    p.pos_offset = pos_offset;
    p.cursors = d_cursors;
    p.digit_shift = digit_shift;
    const uint64_t n_tiles = (win_end + tile - 1) / tile - p.first_tile;
    KMG_REQUIRE(n_tiles < (1ull << 31), KMG_ERR_RANGE, "too many tiles");
    if (big) {
        if (key_bytes == 8) return launch_narrow<uint64_t, 16, false, 0, false, 2, 512>(p, (uint32_t)n_tiles, st);
        return launch_narrow<u128, 8, false, 0, false, 2, 512>(p, (uint32_t)n_tiles, st);
    }
    if (key_bytes == 8) return dispatch_scatter_digit<uint64_t, 16>(p, n_tiles, rc, val_bytes, st);
    return dispatch_scatter_digit<u128, 8>(p, n_tiles, rc, val_bytes, st);
}

}  // namespace kmg

using namespace kmg;

// smallest tile any extract kernel uses -> upper bound on the number of tiles
static constexpr uint64_t EX_MIN_TILE = 1024;

extern "C" size_t kmg_extract_workspace_bytes(uint64_t n_windows) {
    // header + tile states (+2 tiles for the unaligned head/tail) + per-offset 4-mer histograms
    return sizeof(WsHeader) + align_up(sc_state_words(n_windows / EX_MIN_TILE + 3) * sizeof(uint64_t), 256) +
           (size_t)(EX_MAX_OFF + 1) * 256 * sizeof(uint64_t);
}

extern "C" int kmg_extract(const uint8_t* d_bases, uint64_t n_bases, uint64_t win_begin, uint64_t win_end, int k,
                           int rc, int wide, const uint8_t* d_lut256, const uint8_t* d_comp16, void* d_keys_out,
                           int key_bytes, void* d_vals_out, int val_bytes, uint64_t pos_offset,
                           uint64_t* d_counts, uint64_t* d_hist_out, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KMG_REQUIRE(k >= 2, KMG_ERR_ARG, "k must be >= 2, got %d", k);  // batcher.py:477-478
    KMG_REQUIRE(k <= 64, KMG_ERR_RANGE, "k=%d: this build supports k <= 64 (no CPU fallback)", k);
    KMG_REQUIRE(win_begin <= win_end, KMG_ERR_ARG, "win_begin > win_end");
    KMG_REQUIRE(val_bytes == 0 || val_bytes == 4 || val_bytes == 8, KMG_ERR_ARG, "val_bytes must be 0, 4 or 8");
    KMG_REQUIRE((val_bytes == 0) == (d_vals_out == nullptr), KMG_ERR_ARG, "d_vals_out / val_bytes mismatch");
    const int want_kb = wide ? (k <= 32 ? 16 : 32) : (k <= 32 ? 8 : 16);
    KMG_REQUIRE(key_bytes == want_kb, KMG_ERR_ARG, "key_bytes must be %d for k=%d wide=%d", want_kb, k, wide);
    KMG_REQUIRE(((uintptr_t)d_bases & 15) == 0, KMG_ERR_ARG, "d_bases must be 16-byte aligned");
    KMG_REQUIRE(((uintptr_t)d_keys_out & (key_bytes - 1)) == 0, KMG_ERR_ARG, "d_keys_out misaligned");
    KMG_REQUIRE(d_lut256 && d_counts && d_ws, KMG_ERR_ARG, "null pointer argument");
    KMG_REQUIRE(!(wide && rc) || d_comp16, KMG_ERR_ARG, "d_comp16 required for wide rc");
    if (val_bytes == 4)
        KMG_REQUIRE(((pos_offset + n_bases) << 1) < (1ull << 32), KMG_ERR_RANGE, "val_bytes=4 needs < 2^31 positions");
    KMG_REQUIRE(ws_bytes >= kmg_extract_workspace_bytes(win_end - win_begin), KMG_ERR_WS, "extract workspace too small");
    KMG_REQUIRE(!d_hist_out || (!wide && k >= 4), KMG_ERR_ARG, "fused digit histograms need the narrow stream and k >= 4");

    KMG_CUDA(cudaMemsetAsync(d_counts, 0, 2 * sizeof(uint64_t), st));
    KMG_CUDA(cudaMemsetAsync(d_ws, 0, sizeof(WsHeader), st));
    if (d_hist_out) KMG_CUDA(cudaMemsetAsync(d_hist_out, 0, sizeof(uint64_t) * MAX_PASSES * 256, st));
    if (win_end == win_begin) return KMG_OK;

    const uint64_t tile = wide ? (uint64_t)EXW_TILE : (uint64_t)(EX_BLOCK_DEFAULT * (key_bytes == 8 ? 16 : 8));
    const uint64_t first_tile = win_begin / tile;
    const uint64_t n_tiles = (win_end + tile - 1) / tile - first_tile;
    KMG_REQUIRE(n_tiles < (1ull << 31), KMG_ERR_RANGE, "too many tiles");
    const size_t state_bytes = align_up(sc_state_words(n_tiles) * sizeof(uint64_t), 256);
    const size_t hs_bytes = (size_t)(EX_MAX_OFF + 1) * 256 * sizeof(uint64_t);
    KMG_REQUIRE(ws_bytes >= sizeof(WsHeader) + state_bytes + hs_bytes, KMG_ERR_WS, "extract workspace too small");
    KMG_CUDA(cudaMemsetAsync(d_ws, 0, sizeof(WsHeader) + state_bytes + (d_hist_out ? hs_bytes : 0), st));
    WsHeader* hdr = reinterpret_cast<WsHeader*>(d_ws);

    ExtractParams p;
    p.bases = d_bases;
    p.n_bases = n_bases;
    p.win_begin = win_begin;
    p.win_end = win_end;
    p.first_tile = first_tile;
    p.k = k;
    p.lut = d_lut256;
    p.comp16 = d_comp16;
    p.keys_out = d_keys_out;
    p.vals_out = d_vals_out;
    p.pos_offset = pos_offset;
    p.counts = reinterpret_cast<unsigned long long*>(d_counts);
    p.tile_state = reinterpret_cast<uint64_t*>(hdr + 1);
    p.ticket = &hdr->ticket;
    p.err = &hdr->err;
    p.hs = reinterpret_cast<unsigned long long*>((char*)d_ws + sizeof(WsHeader) + state_bytes);
    p.n_off = 0;
    int top_idx = -1;
    if (d_hist_out) {
        const int P = (2 * k + 7) / 8;
        for (int q = 0; q < P - 1; ++q) p.hist_off[q] = k - 4 - 4 * q;
        p.hist_off[P - 1] = 0;
        p.n_off = P;
        if (rc) {
            for (int q = 0; q < P - 1; ++q) p.hist_off[P + q] = 4 * q;
            p.hist_off[2 * P - 1] = k - 4;
            p.n_off = 2 * P;
        }
        if (key_bytes == 8 && k >= 12) {
            top_idx = p.n_off;
            p.hist_off[p.n_off++] = 4;
            p.hist_off[p.n_off++] = 8;
            if (rc) {
                p.hist_off[p.n_off++] = k - 8;
                p.hist_off[p.n_off++] = k - 12;
            }
        }
    }

    if (wide && k <= 32) return rc ? dispatch_wide<true, 2>(p, (uint32_t)n_tiles, val_bytes, st)
                                   : dispatch_wide<false, 2>(p, (uint32_t)n_tiles, val_bytes, st);
    if (wide) return rc ? dispatch_wide<true, 4>(p, (uint32_t)n_tiles, val_bytes, st)
                        : dispatch_wide<false, 4>(p, (uint32_t)n_tiles, val_bytes, st);
    int rcode;
    if (key_bytes == 8) rcode = dispatch_narrow<uint64_t, 16>(p, (uint32_t)n_tiles, rc, val_bytes, d_hist_out != nullptr, st);
    else rcode = dispatch_narrow<u128, 8>(p, (uint32_t)n_tiles, rc, val_bytes, d_hist_out != nullptr, st);
    if (rcode != KMG_OK || !d_hist_out) return rcode;
    hist_edges_kernel<<<1, 256, 0, st>>>(p);
    KMG_LAUNCH_CHECK();
    hist_finalize_kernel<<<1, 256, 0, st>>>(p.hs, p.n_off, k, rc, top_idx, reinterpret_cast<unsigned long long*>(d_hist_out));
    KMG_LAUNCH_CHECK();
    return KMG_OK;
}
