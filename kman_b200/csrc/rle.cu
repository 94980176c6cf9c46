// K4: run-length / singleton selection on sorted keys.
//
// Replaces the grouping loop of Crawler.do_batch (kmermaid/join.py:95-130) and the two emit
// rules: KJoiner.join_sequence_count (join.py:265-285: every group, with its size) and
// KJoiner.join_unique (join.py:243-263: only groups of size exactly one).
//
// A warp walks its contiguous chunk 32 keys at a time; "is the key different from its
// predecessor" becomes one ballot per step, so ranks are popcounts and no shared-memory
// transposition of the keys is needed.  Tiles are chained with a decoupled look-back that
// carries (number of run heads, position of the last run head); a run's length is written
// by the tile that sees the run END, which is what makes the pass single-sweep.
#include <type_traits>

#include "common.cuh"

namespace kmg {

constexpr int RLE_BLOCK = 256;
constexpr int RLE_WARPS = RLE_BLOCK / 32;

template <typename KeyT>
__device__ __forceinline__ KeyT shfl_key(const KeyT& k, int src);
template <>
__device__ __forceinline__ uint64_t shfl_key<uint64_t>(const uint64_t& k, int src) {
    return __shfl_sync(0xffffffffu, k, src);
}
template <>
__device__ __forceinline__ u128 shfl_key<u128>(const u128& k, int src) {
    return u128{__shfl_sync(0xffffffffu, k.lo, src), __shfl_sync(0xffffffffu, k.hi, src)};
}

struct RleParams {
    const void* keys_in;
    const void* vals_in;
    uint64_t n;
    void* keys_out;
    void* vals_out;
    uint32_t* counts_out;
    unsigned long long* n_out;
    uint64_t* state_a;  // flag | heads
    uint64_t* state_b;  // flag | last head position + 1 (0 = none)
    uint32_t* ticket;
    uint32_t* err;
};

// lanes of step j whose element exists (tile-local index < n_local)
__device__ __forceinline__ uint32_t valid_mask(uint32_t step_first, uint32_t n_local) {
    if (step_first >= n_local) return 0u;
    const uint32_t left = n_local - step_first;
    return left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
}

// Loads the warp's chunk (IPT steps of 32 consecutive keys starting at tile-local index
// `wfirst`) and returns the extended head ballots: bit l of hb[j] is set iff element
// wfirst + 32*j + l starts a run, where the element one past the end of the whole array counts
// as a head (end sentinel) and anything beyond does not.  hnext = head flag of the element
// right after the chunk.  FULL: every element of the tile exists.
template <typename KeyT, int IPT, bool FULL>
__device__ __forceinline__ void load_heads(const KeyT* __restrict__ kin, uint64_t tile_base, uint32_t n_local,
                                           bool more_after_tile, uint32_t wfirst, KeyT (&keys)[IPT],
                                           uint32_t (&hb)[IPT], uint32_t& hnext) {
    const uint32_t lane = lane_id();
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t li = wfirst + 32 * j + lane;
        keys[j] = (FULL || li < n_local) ? kin[li] : KeyT{};
    }
    const bool first_of_all = tile_base == 0 && wfirst == 0;
    KeyT before = KeyT{};
    if (lane == 0 && !first_of_all && (FULL || wfirst <= n_local)) before = *(kin + wfirst - 1);
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t li = wfirst + 32 * j + lane;
        KeyT prev = shfl_key(keys[j], (int)lane - 1);
        if (lane == 0) prev = before;
        bool head = keys[j] != prev || (first_of_all && j == 0 && lane == 0);
        if (!FULL) head = (li < n_local && head) || li == n_local;
        hb[j] = __ballot_sync(0xffffffffu, head);
        before = shfl_key(keys[j], 31);  // only lane 0 uses it
    }
    const uint32_t nli = wfirst + 32 * IPT;
    uint32_t h = 0;
    if (lane == 0) {
        if (nli < n_local || (nli == n_local && more_after_tile)) h = kin[nli] != before ? 1u : 0u;
        else if (nli == n_local) h = 1;  // end of the array
    }
    hnext = __shfl_sync(0xffffffffu, h, 0);
}

template <typename KeyT, int IPT, bool FULL>
__device__ __forceinline__ void rle_count_tile(const RleParams& p, const uint32_t tile, uint32_t* s_hpos,
                                               uint32_t* s_wheads, uint64_t* s_bcast) {
    constexpr int TILE = RLE_BLOCK * IPT;
    const int t = threadIdx.x;
    const uint32_t lane = t & 31, warp = t >> 5;
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const uint32_t n_local = FULL ? (uint32_t)TILE : (uint32_t)(p.n - tile_base);
    const bool more_after = tile_base + TILE < p.n;
    const uint32_t wfirst = warp * 32 * IPT;
    const KeyT* kin = reinterpret_cast<const KeyT*>(p.keys_in) + tile_base;

    KeyT keys[IPT];
    uint32_t hb[IPT];
    uint32_t hnext;
    load_heads<KeyT, IPT, FULL>(kin, tile_base, n_local, more_after, wfirst, keys, hb, hnext);

    // real heads per warp -> tile-local ordinals
    uint32_t wheads = 0;
#pragma unroll
    for (int j = 0; j < IPT; ++j) wheads += __popc(FULL ? hb[j] : (hb[j] & valid_mask(wfirst + 32 * j, n_local)));
    if (lane == 0) s_wheads[warp] = wheads;
    __syncthreads();
    uint32_t wexcl = 0, theads = 0;
#pragma unroll
    for (int w = 0; w < RLE_WARPS; ++w) {
        const uint32_t c = s_wheads[w];
        if (w < (int)warp) wexcl += c;
        theads += c;
    }
    const uint32_t lt = lanemask_lt();
    {
        uint32_t run = wexcl;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t b = FULL ? hb[j] : (hb[j] & valid_mask(wfirst + 32 * j, n_local));
            if ((b >> lane) & 1u) s_hpos[run + __popc(b & lt)] = wfirst + 32 * j + lane;
            run += __popc(b);
        }
    }
    __syncthreads();
    if (t < 32) {
        // Two-level prefix (common.cuh) over the pair (number of heads: SUM, position+1 of the
        // last head: MAX -- positions grow with the tile index, so the nearest preceding head is
        // the maximum).  Both words carry their own flag; a slot counts when both are set.
        const uint32_t n_tiles = gridDim.x;
        const uint32_t g = tile / SC_GROUP, r = tile % SC_GROUP;
        uint64_t* ginc_a = p.state_a + n_tiles;
        uint64_t* ginc_b = p.state_b + n_tiles;
        const uint64_t last_plus1 = theads ? tile_base + s_hpos[theads - 1] + 1 : 0;
        if (lane == 0) {
            st_relaxed_u64(&p.state_b[tile], SC_FLAG | last_plus1);
            st_relaxed_u64(&p.state_a[tile], SC_FLAG | theads);
        }
        uint32_t spins = 0;
        uint64_t ga = SC_FLAG, gb = SC_FLAG;
        if (lane == 0 && g > 0) {
            ga = ld_relaxed_u64(&ginc_a[g - 1]);
            gb = ld_relaxed_u64(&ginc_b[g - 1]);
        }
        uint64_t wa[SC_GROUP / 32], wb[SC_GROUP / 32];
#pragma unroll
        for (int q = 0; q < (int)(SC_GROUP / 32); ++q) {
            const uint32_t j = lane + 1 + 32 * q;
            wa[q] = j <= r ? ld_relaxed_u64(&p.state_a[tile - j]) : SC_FLAG;
            wb[q] = j <= r ? ld_relaxed_u64(&p.state_b[tile - j]) : SC_FLAG;
        }
        uint64_t excl = 0, carry = 0;
#pragma unroll
        for (int q = 0; q < (int)(SC_GROUP / 32); ++q) {
            const uint32_t j = lane + 1 + 32 * q;
            if (j <= r) {
                wa[q] = sc_wait(&p.state_a[tile - j], wa[q], spins, p.err);
                wb[q] = sc_wait(&p.state_b[tile - j], wb[q], spins, p.err);
            }
            excl += wa[q] & SC_VALUE_MASK;
            carry = max(carry, wb[q] & SC_VALUE_MASK);
        }
        if (lane == 0 && g > 0) {
            ga = sc_wait(&ginc_a[g - 1], ga, spins, p.err);
            gb = sc_wait(&ginc_b[g - 1], gb, spins, p.err);
        }
        excl += ga & SC_VALUE_MASK;
        carry = max(carry, gb & SC_VALUE_MASK);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            excl += __shfl_xor_sync(0xffffffffu, excl, o);
            carry = max(carry, __shfl_xor_sync(0xffffffffu, carry, o));
        }
        if (lane == 0) {
            if (r == SC_GROUP - 1) {
                st_relaxed_u64(&ginc_b[g], SC_FLAG | max(carry, last_plus1));
                st_relaxed_u64(&ginc_a[g], SC_FLAG | (excl + theads));
            }
            s_bcast[0] = excl;
            s_bcast[1] = carry;  // position+1 of the head of the run that is open when the tile starts
            if (tile == gridDim.x - 1) *p.n_out = excl + theads;
        }
    }
    __syncthreads();
    const uint64_t excl_heads = s_bcast[0];
    const uint64_t carry = s_bcast[1];
    KeyT* keys_out = reinterpret_cast<KeyT*>(p.keys_out) + excl_heads;
    uint32_t* counts_out = p.counts_out + excl_heads;  // slot -1 (the run open at tile start) is valid when used

    uint32_t run = wexcl;  // heads of the tile before the current step
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t li = wfirst + 32 * j + lane;
        const uint32_t vm = FULL ? 0xffffffffu : valid_mask(wfirst + 32 * j, n_local);
        const uint32_t b = hb[j];
        const uint32_t nb = (j + 1 < IPT) ? hb[j + 1 < IPT ? j + 1 : j] : hnext;
        const uint32_t tb = (b >> 1) | ((nb & 1u) << 31);  // tail: the next element is a head
        const uint32_t h_before = run + __popc(b & vm & lt);
        const bool is_head = (b >> lane) & 1u;
        if (FULL || li < n_local) {
            if (is_head) keys_out[h_before] = keys[j];
            if ((tb >> lane) & 1u) {
                const uint32_t h_incl = h_before + (is_head ? 1u : 0u);
                uint64_t len;
                if (h_incl) len = (uint64_t)(li - s_hpos[h_incl - 1]) + 1;
                else len = tile_base + li - (carry - 1) + 1;
                if (len > 0xffffffffull) atomicExch(p.err, 2u);
                *(counts_out + (int64_t)h_incl - 1) = (uint32_t)len;
            }
        }
        run += __popc(b & vm);
    }
}

template <typename KeyT, int IPT>
__global__ void __launch_bounds__(RLE_BLOCK, 3) rle_count_kernel(const RleParams p) {
    constexpr int TILE = RLE_BLOCK * IPT;
    __shared__ uint32_t s_hpos[TILE];  // local position of the i-th head of the tile
    __shared__ uint32_t s_wheads[RLE_WARPS];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_bcast[2];
    if (threadIdx.x == 0) s_tile = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if ((uint64_t)(tile + 1) * TILE <= p.n) rle_count_tile<KeyT, IPT, true>(p, tile, s_hpos, s_wheads, s_bcast);
    else rle_count_tile<KeyT, IPT, false>(p, tile, s_hpos, s_wheads, s_bcast);
}

// singletons: head && tail, compacted in order, with payload
template <typename KeyT, int VAL_BYTES, int IPT, bool FULL>
__device__ __forceinline__ void select_tile(const RleParams& p, const uint32_t tile, uint32_t* s_wcnt, uint64_t* s_bcast) {
    constexpr int TILE = RLE_BLOCK * IPT;
    using ValT = typename std::conditional<VAL_BYTES == 4, uint32_t, uint64_t>::type;
    const int t = threadIdx.x;
    const uint32_t lane = t & 31, warp = t >> 5;
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const uint32_t n_local = FULL ? (uint32_t)TILE : (uint32_t)(p.n - tile_base);
    const bool more_after = tile_base + TILE < p.n;
    const uint32_t wfirst = warp * 32 * IPT;
    const KeyT* kin = reinterpret_cast<const KeyT*>(p.keys_in) + tile_base;

    KeyT keys[IPT];
    uint32_t hb[IPT];
    uint32_t hnext;
    load_heads<KeyT, IPT, FULL>(kin, tile_base, n_local, more_after, wfirst, keys, hb, hnext);

    uint32_t sb[IPT];  // singleton ballots
    uint32_t wcnt = 0;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t nb = (j + 1 < IPT) ? hb[j + 1 < IPT ? j + 1 : j] : hnext;
        const uint32_t tb = (hb[j] >> 1) | ((nb & 1u) << 31);
        sb[j] = hb[j] & tb & (FULL ? 0xffffffffu : valid_mask(wfirst + 32 * j, n_local));
        wcnt += __popc(sb[j]);
    }
    if (lane == 0) s_wcnt[warp] = wcnt;
    __syncthreads();
    uint32_t wexcl = 0, total = 0;
#pragma unroll
    for (int w = 0; w < RLE_WARPS; ++w) {
        const uint32_t c = s_wcnt[w];
        if (w < (int)warp) wexcl += c;
        total += c;
    }
    if (t < 32) {
        const uint64_t excl = tile_prefix_exclusive_warp(p.state_a, gridDim.x, tile, total, p.err);
        if (t == 0) {
            s_bcast[0] = excl;
            if (tile == gridDim.x - 1) *p.n_out = excl + total;
        }
    }
    __syncthreads();
    const uint64_t base = s_bcast[0] + wexcl;
    KeyT* keys_out = reinterpret_cast<KeyT*>(p.keys_out) + base;
    const ValT* vals_in = reinterpret_cast<const ValT*>(p.vals_in) + tile_base;
    ValT* vals_out = reinterpret_cast<ValT*>(p.vals_out) + base;
    const uint32_t lt = lanemask_lt();
    uint32_t run = 0;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        if ((sb[j] >> lane) & 1u) {
            const uint32_t o = run + __popc(sb[j] & lt);
            keys_out[o] = keys[j];
            if constexpr (VAL_BYTES != 0) vals_out[o] = vals_in[wfirst + 32 * j + lane];
        }
        run += __popc(sb[j]);
    }
}

template <typename KeyT, int VAL_BYTES, int IPT>
__global__ void __launch_bounds__(RLE_BLOCK, 3) select_singletons_kernel(const RleParams p) {
    constexpr int TILE = RLE_BLOCK * IPT;
    __shared__ uint32_t s_wcnt[RLE_WARPS];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_bcast[2];
    if (threadIdx.x == 0) s_tile = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if ((uint64_t)(tile + 1) * TILE <= p.n) select_tile<KeyT, VAL_BYTES, IPT, true>(p, tile, s_wcnt, s_bcast);
    else select_tile<KeyT, VAL_BYTES, IPT, false>(p, tile, s_wcnt, s_bcast);
}

constexpr int RLE_IPT8 = 16;   // 8-byte keys: 4096-key tiles
constexpr int RLE_IPT16 = 8;   // 16-byte keys: 2048-key tiles

}  // namespace kmg

using namespace kmg;

extern "C" size_t kmg_rle_workspace_bytes(uint64_t n) {
    const uint64_t tiles = n / 1024 + 2;
    return sizeof(WsHeader) + 2 * align_up(sc_state_words(tiles) * sizeof(uint64_t), 256);
}

static int rle_setup(RleParams& p, uint64_t n, int key_bytes, void* d_ws, size_t ws_bytes, uint32_t& tiles,
                     cudaStream_t st) {
    KMG_REQUIRE(key_bytes == 8 || key_bytes == 16, KMG_ERR_ARG, "key_bytes must be 8 or 16");
    KMG_REQUIRE(d_ws, KMG_ERR_ARG, "null workspace");
    KMG_REQUIRE(ws_bytes >= kmg_rle_workspace_bytes(n), KMG_ERR_WS, "rle workspace too small");
    const uint64_t tile = key_bytes == 8 ? RLE_BLOCK * RLE_IPT8 : RLE_BLOCK * RLE_IPT16;
    const uint64_t nt = (n + tile - 1) / tile;
    KMG_REQUIRE(nt < (1ull << 31), KMG_ERR_RANGE, "too many tiles");
    tiles = (uint32_t)nt;
    const size_t arr = align_up(sc_state_words(n / 1024 + 2) * sizeof(uint64_t), 256);
    KMG_CUDA(cudaMemsetAsync(d_ws, 0, sizeof(WsHeader) + 2 * arr, st));
    WsHeader* hdr = reinterpret_cast<WsHeader*>(d_ws);
    p.n = n;
    p.state_a = reinterpret_cast<uint64_t*>(hdr + 1);
    p.state_b = reinterpret_cast<uint64_t*>((char*)(hdr + 1) + arr);
    p.ticket = &hdr->ticket;
    p.err = &hdr->err;
    return KMG_OK;
}

extern "C" int kmg_rle_count(const void* d_sorted_keys, uint64_t n, int key_bytes, void* d_uniq_keys_out,
                             uint32_t* d_counts_out, uint64_t* d_n_out, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KMG_REQUIRE(d_n_out, KMG_ERR_ARG, "d_n_out is null");
    KMG_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(uint64_t), st));
    if (n == 0) return KMG_OK;  // join.py:107-111: nothing to crawl -> empty output
    KMG_REQUIRE(d_sorted_keys && d_uniq_keys_out && d_counts_out, KMG_ERR_ARG, "null pointer argument");
    RleParams p;
    memset(&p, 0, sizeof(p));
    uint32_t tiles = 0;
    int rcode = rle_setup(p, n, key_bytes, d_ws, ws_bytes, tiles, st);
    if (rcode != KMG_OK) return rcode;
    p.keys_in = d_sorted_keys;
    p.keys_out = d_uniq_keys_out;
    p.counts_out = d_counts_out;
    p.n_out = reinterpret_cast<unsigned long long*>(d_n_out);
    if (key_bytes == 8) rle_count_kernel<uint64_t, RLE_IPT8><<<tiles, RLE_BLOCK, 0, st>>>(p);
    else rle_count_kernel<u128, RLE_IPT16><<<tiles, RLE_BLOCK, 0, st>>>(p);
    KMG_LAUNCH_CHECK();
    return KMG_OK;
}

extern "C" int kmg_select_singletons(const void* d_sorted_keys, const void* d_vals, uint64_t n, int key_bytes,
                                     int val_bytes, void* d_keys_out, void* d_vals_out, uint64_t* d_n_out,
                                     void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KMG_REQUIRE(d_n_out, KMG_ERR_ARG, "d_n_out is null");
    KMG_REQUIRE(val_bytes == 0 || val_bytes == 4 || val_bytes == 8, KMG_ERR_ARG, "val_bytes must be 0, 4 or 8");
    KMG_REQUIRE((val_bytes == 0) == (d_vals == nullptr), KMG_ERR_ARG, "d_vals / val_bytes mismatch");
    KMG_REQUIRE(val_bytes == 0 || d_vals_out, KMG_ERR_ARG, "d_vals_out is null");
    KMG_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(uint64_t), st));
    if (n == 0) return KMG_OK;
    KMG_REQUIRE(d_sorted_keys && d_keys_out, KMG_ERR_ARG, "null pointer argument");
    RleParams p;
    memset(&p, 0, sizeof(p));
    uint32_t tiles = 0;
    int rcode = rle_setup(p, n, key_bytes, d_ws, ws_bytes, tiles, st);
    if (rcode != KMG_OK) return rcode;
    p.keys_in = d_sorted_keys;
    p.vals_in = d_vals;
    p.keys_out = d_keys_out;
    p.vals_out = d_vals_out;
    p.n_out = reinterpret_cast<unsigned long long*>(d_n_out);
    if (key_bytes == 8) {
        if (val_bytes == 0) select_singletons_kernel<uint64_t, 0, RLE_IPT8><<<tiles, RLE_BLOCK, 0, st>>>(p);
        else if (val_bytes == 4) select_singletons_kernel<uint64_t, 4, RLE_IPT8><<<tiles, RLE_BLOCK, 0, st>>>(p);
        else select_singletons_kernel<uint64_t, 8, RLE_IPT8><<<tiles, RLE_BLOCK, 0, st>>>(p);
    } else {
        if (val_bytes == 0) select_singletons_kernel<u128, 0, RLE_IPT16><<<tiles, RLE_BLOCK, 0, st>>>(p);
        else if (val_bytes == 4) select_singletons_kernel<u128, 4, RLE_IPT16><<<tiles, RLE_BLOCK, 0, st>>>(p);
        else select_singletons_kernel<u128, 8, RLE_IPT16><<<tiles, RLE_BLOCK, 0, st>>>(p);
    }
    KMG_LAUNCH_CHECK();
    return KMG_OK;
}
