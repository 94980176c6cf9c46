// Hybrid finish, second generation of the shared-memory local sort (same contract, geometry and
// HybridParams as local_sort_kernel in local_sort.cuh; included right after it by radix_sort.cu).
//
// What changed, and why (profiles/r01_ncu_local_sort_v5.txt: 41 % issue-active, barrier stalls
// dominant, 7 warp instructions per key): the first kernel orders the cells with THREE serial
// per-thread walks over runs of Poisson-distributed length (insertion sort, head count, head
// positions), so every barrier waits for the thread with the longest run.  Here all per-key work is
// done by the thread that LOADED the key, 16 keys per thread whatever the data looks like:
//   * a key's final place is  cell start + rank inside its cell , and the rank is found by reading
//     the cell's other members (one on average): no serial walk, no data-dependent thread load;
//   * run heads (distinct keys) / singletons are found position-parallel -- thread t looks at
//     positions t, t + 512, ... -- with warp ballots; a 256-word scan of their popcounts ranks them
//     and the next set ballot bit gives a run's length;
//   * cells too crowded for rank-by-reading (> 16 keys: near-copies of a repeat) are sorted by one
//     warp each (<= 128 keys) or by the whole block (bitonic), as before.
#pragma once

constexpr int LS2_RANK_MAX = 16;                       // largest cell whose keys rank themselves by reading it
constexpr int LS2_MAX_WLIST = 8192 / (LS2_RANK_MAX + 1) + 1;   // cells a tile can hold above that size
constexpr int LS2_MAX_BIG = 8192 / (LS_WARP_RUN_MAX + 1) + 1;  // ... above the warp limit

template <typename KeyT, int EMIT, int VB>
__global__ void __launch_bounds__(LS_BLOCK, 2) local_sort2_kernel(const HybridParams p) {
    constexpr bool PAIRS = VB != 0;
    constexpr bool COUNT = EMIT == 1, UNIQ = EMIT == 2, FUSED = EMIT != 0;
    static_assert(!(PAIRS && COUNT), "the fused count is key-only");
    static_assert(!UNIQ || PAIRS, "singletons carry their payload");
    constexpr int IPT = LS<KeyT, PAIRS>::IPT, CPT = LS<KeyT, PAIRS>::CPT;
    constexpr int CAP = ls_cap<KeyT, PAIRS>(), CELLS = ls_cells<KeyT, PAIRS>(), CELL_WORDS = ls_cell_words<KeyT, PAIRS>();
    constexpr int CELL_BITS = 13;
    constexpr int NW = CAP / 32;  // ballot words of a tile
    static_assert((1 << CELL_BITS) >= CELLS && CAP <= (1 << CELL_BITS), "cell / position widths");
    static_assert(CAP % LS_BLOCK == 0 && LS_BLOCK % 32 == 0 && 2 * NW + 2 <= CELL_WORDS, "ballot words live in the cell array");
    using ValT = typename ValType<VB == 0 ? 8 : VB>::type;
    extern __shared__ __align__(16) unsigned char ls_smem[];
    KeyT* s_stage = reinterpret_cast<KeyT*>(ls_smem);                                 // [CAP]
    uint32_t* s_cell = reinterpret_cast<uint32_t*>(ls_smem + sizeof(KeyT) * CAP);     // [CELL_WORDS]
    uint16_t* s_idx = reinterpret_cast<uint16_t*>(s_cell + CELL_WORDS);               // [CAP] (PAIRS)
    __shared__ uint32_t s_scan[LS_BLOCK / 32 + 1];
    __shared__ int s_bad;
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_wl_n, s_big_n;
    __shared__ uint32_t s_wl[LS2_MAX_WLIST];   // start | length << 13 of cells a warp sorts
    __shared__ uint32_t s_big[LS2_MAX_BIG][2]; // [start, end) of cells the block sorts
    const int t = threadIdx.x;
    const uint32_t lane = t & 31u, warp = t >> 5;
    const int sh_pref = p.key_bits - p.pb;
    // a fused table cannot be patched up afterwards: with oversize tiles around the host falls back to
    // sort + run-length stage, and this launch has nothing to do (every CTA takes the same exit)
    if (FUSED && p.over[0] != 0) return;

    if (t == 0) {
        s_bad = 0;
        s_wl_n = 0;
        s_big_n = 0;
        if (FUSED) s_tile = atomicAdd(p.ticket, 1u);
    }
    {
        uint4* z = reinterpret_cast<uint4*>(s_cell);
        for (uint32_t i = t; i < (uint32_t)CELL_WORDS / 4; i += LS_BLOCK) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (FUSED) __syncthreads();
    const uint32_t tile = FUSED ? s_tile : blockIdx.x;
    // a tile without output still takes part in the tile prefix (and the last one reports the total)
    auto finish_without_output = [&]() {
        if constexpr (FUSED) {
            if (t < 32) {
                if (t == 0) tile_prefix_publish(p.tile_state, tile, 0);
                const uint64_t base = tile_prefix_resolve_warp(p.tile_state, p.n_tiles, tile, 0, p.err);
                if (t == 0 && tile == p.n_tiles - 1) *p.n_out = *p.n_out_copy = base;
            }
        }
    };
    const uint64_t s = p.bounds[tile], e = p.bounds[tile + 1];
    if (s >= min((uint64_t)(tile + 1) * p.tile_t, p.n) || e <= s) {  // no bucket starts in this tile
        finish_without_output();
        return;
    }
    const uint64_t m64 = e - s;
    if (m64 > (uint64_t)CAP) {
        if (t == 0) {
            atomicAdd(p.irregular, 1ull);
            p.flag[tile] = 1;
        }
        finish_without_output();
        return;
    }
    const uint32_t m = (uint32_t)m64;
    const KeyT* kin = reinterpret_cast<const KeyT*>(p.keys_in) + s;
    const uint64_t p_first = shr64(kin[0], sh_pref), p_last = shr64(kin[m - 1], sh_pref);
    KeyT keys[IPT];
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t idx = t + j * LS_BLOCK;
        keys[j] = idx < m ? kin[idx] : KeyT{};
    }
    // Counting sort into <= CELLS cells through a monotone map of the key (see local_sort_kernel)
    CellMap<KeyT, CELLS> cm;
    {
        const uint64_t R = p_last - p_first + 1;  // <= 2^24
        int w = min(CELL_BITS, sh_pref);
        w = max(0, min(w, 30 - (63 - __clzll((long long)R))));  // (R << w) < 2^31
        cm.sh = sh_pref - w;
        cm.base = p_first << w;
        const uint64_t range = R << w;
        cm.inv = range <= (uint64_t)CELLS ? 0u : (uint32_t)((((uint64_t)CELLS) << 32) / range);
    }
    __syncthreads();     // cells are zero
    uint32_t meta[IPT];  // cell | slot inside the cell << CELL_BITS ; later: place | cell size << 13 | slot << 19
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t idx = t + j * LS_BLOCK;
        if (idx < m) {
            const uint32_t c = cm(keys[j]);
            meta[j] = c | (atomicAdd(&s_cell[pc<KeyT, PAIRS>(c)], 1u) << CELL_BITS);
        }
    }
    __syncthreads();
    // exclusive prefix over the cells: CPT consecutive cells per thread, 128 bits at a time; cells
    // too crowded to rank by reading are listed for a warp / the block on the way
    {
        uint4* cv = reinterpret_cast<uint4*>(s_cell + pc<KeyT, PAIRS>(t * CPT));
        uint4 q[CPT / 4];
        uint32_t sum = 0, mx = 0;
#pragma unroll
        for (int i = 0; i < CPT / 4; ++i) {
            q[i] = cv[i];
            sum += q[i].x + q[i].y + q[i].z + q[i].w;
            mx = max(max(mx, max(q[i].x, q[i].y)), max(q[i].z, q[i].w));
        }
        uint32_t total;
        uint32_t run = block_excl_scan<LS_BLOCK, uint32_t>(sum, s_scan, total);
#pragma unroll
        for (int i = 0; i < CPT / 4; ++i) {
            uint32_t v;
#define KMG_LS_STEP(f)                                                                  \
    v = f;                                                                              \
    f = run;                                                                            \
    if (mx > (uint32_t)LS2_RANK_MAX && v > (uint32_t)LS2_RANK_MAX) {                    \
        if (v > (uint32_t)LS_WARP_RUN_MAX) {                                            \
            const uint32_t slot = atomicAdd(&s_big_n, 1u);                              \
            s_big[slot][0] = run;                                                       \
            s_big[slot][1] = run + v;                                                   \
        } else {                                                                        \
            s_wl[atomicAdd(&s_wl_n, 1u)] = run | (v << 13);                             \
        }                                                                               \
    }                                                                                   \
    run += v;
            KMG_LS_STEP(q[i].x) KMG_LS_STEP(q[i].y) KMG_LS_STEP(q[i].z) KMG_LS_STEP(q[i].w)
#undef KMG_LS_STEP
            cv[i] = q[i];
        }
        asm volatile("" ::: "memory");
        if (t == LS_BLOCK - 1) s_cell[pc<KeyT, PAIRS>(CELLS)] = run;  // sentinel: end of the last cell
    }
    __syncthreads();
    // placement in arrival order: cell start + slot
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t idx = t + j * LS_BLOCK;
        if (idx < m) {
            const uint32_t c = meta[j] & ((1u << CELL_BITS) - 1u), slot = meta[j] >> CELL_BITS;
            const uint32_t cs = s_cell[pc<KeyT, PAIRS>(c)], n = s_cell[pc<KeyT, PAIRS>(c + 1)] - cs;
            const uint32_t at = cs + slot;
            s_stage[at] = keys[j];
            if constexpr (PAIRS) s_idx[at] = (uint16_t)idx;
            meta[j] = at | (min(n, 63u) << 13) | (min(slot, 63u) << 19);
        }
    }
    __syncthreads();
    // rank inside the cell = members that sort before me (equal keys: in arrival order)
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t idx = t + j * LS_BLOCK;
        const uint32_t n = idx < m ? (meta[j] >> 13) & 63u : 0u;
        if (n > 1 && n <= (uint32_t)LS2_RANK_MAX) {
            const uint32_t slot = meta[j] >> 19, cs = (meta[j] & 8191u) - slot;
            const KeyT key = keys[j];
            uint32_t rank = 0;
            for (uint32_t i = 0; i < n; ++i) {
                const KeyT o = s_stage[cs + i];
                rank += (o < key || (o == key && i < slot)) ? 1u : 0u;
            }
            meta[j] = (cs + rank) | (rank != slot ? 1u << 31 : 0u);
        } else {
            meta[j] = 0;  // alone in its cell, or the cell is sorted by a warp / the block
        }
    }
    __syncthreads();  // every cell has been read
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        if (meta[j] >> 31) {
            const uint32_t at = meta[j] & 8191u;
            s_stage[at] = keys[j];
            if constexpr (PAIRS) s_idx[at] = (uint16_t)(t + j * LS_BLOCK);
        }
    }
    __syncthreads();
    const uint32_t n_wl = s_wl_n, n_big = s_big_n;
    if (n_wl | n_big) {  // (block-uniform) crowded cells: near-copies of a repeat
        // cells of 17..128 keys: one warp each, rank sort with the keys in registers
        for (uint32_t r = warp; r < n_wl; r += LS_BLOCK / 32) {
            const uint32_t rlo = s_wl[r] & 8191u, rlen = s_wl[r] >> 13;
            constexpr int Q = LS_WARP_RUN_MAX / 32;
            KeyT kq[Q];
            uint32_t rk[Q];
            uint16_t iq[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const uint32_t x = lane + 32u * q;
                kq[q] = x < rlen ? s_stage[rlo + x] : KeyT{};
                if constexpr (PAIRS) iq[q] = x < rlen ? s_idx[rlo + x] : (uint16_t)0;
                rk[q] = 0;
            }
            for (uint32_t i = 0; i < rlen; ++i) {
                const KeyT o = s_stage[rlo + i];  // one address for the whole warp: a broadcast
#pragma unroll
                for (int q = 0; q < Q; ++q) rk[q] += (o < kq[q] || (o == kq[q] && i < lane + 32u * q)) ? 1u : 0u;
            }
            __syncwarp();  // every lane has read the cell
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                if (lane + 32u * q < rlen) {
                    s_stage[rlo + rk[q]] = kq[q];
                    if constexpr (PAIRS) s_idx[rlo + rk[q]] = iq[q];
                }
            }
        }
        if (n_big) {
            // bitonic sort of every bigger cell by the whole block, in the (now dead) cell array, padded
            // to a power of two with all-ones keys
            constexpr uint32_t TEMP_CAP = (uint32_t)(CELL_WORDS * sizeof(uint32_t) / (sizeof(KeyT) + (PAIRS ? 2 : 0)));
            constexpr uint32_t TEMP_POW2 = TEMP_CAP >= 4096 ? 4096 : (TEMP_CAP >= 2048 ? 2048 : 1024);
            static_assert(TEMP_POW2 <= TEMP_CAP, "bitonic buffer");
            KeyT* tmp_k = reinterpret_cast<KeyT*>(s_cell);
            uint16_t* tmp_i = reinterpret_cast<uint16_t*>(tmp_k + TEMP_POW2);
            if (t == 0) atomicAdd(&p.irregular[1], (unsigned long long)n_big);
            for (uint32_t b = 0; b < n_big; ++b) {
                const uint32_t blo = s_big[b][0], sz = s_big[b][1] - blo;
                if (sz > TEMP_POW2) {
                    s_bad = 1;  // (every thread takes the same branch)
                    break;
                }
                uint32_t N = 64;
                while (N < sz) N <<= 1;
                for (uint32_t x = t; x < N; x += LS_BLOCK) {
                    tmp_k[x] = x < sz ? s_stage[blo + x] : key_all_ones(KeyT{});
                    if constexpr (PAIRS) tmp_i[x] = x < sz ? s_idx[blo + x] : (uint16_t)0xFFFF;  // padding sorts last
                }
                __syncthreads();
                for (uint32_t kk = 2; kk <= N; kk <<= 1) {
                    for (uint32_t jj = kk >> 1; jj > 0; jj >>= 1) {
                        for (uint32_t x = t; x < N / 2; x += LS_BLOCK) {
                            // x-th compare-exchange of this stage: partner indices i < l = i ^ jj
                            const uint32_t i = ((x & ~(jj - 1)) << 1) | (x & (jj - 1));
                            const uint32_t l = i | jj;
                            const bool up = (i & kk) == 0;
                            const KeyT a0 = tmp_k[i], a1 = tmp_k[l];
                            bool less = a1 < a0;
                            if constexpr (PAIRS) {  // total order (key, index): the padding can never displace a real pair
                                const uint16_t i0 = tmp_i[i], i1 = tmp_i[l];
                                less = less || (a1 == a0 && i1 < i0);
                                if (less == up) {
                                    tmp_i[i] = i1;
                                    tmp_i[l] = i0;
                                }
                            }
                            if (less == up) {
                                tmp_k[i] = a1;
                                tmp_k[l] = a0;
                            }
                        }
                        __syncthreads();
                    }
                }
                for (uint32_t x = t; x < sz; x += LS_BLOCK) {
                    s_stage[blo + x] = tmp_k[x];
                    if constexpr (PAIRS) s_idx[blo + x] = tmp_i[x];
                }
                __syncthreads();
            }
        }
        __syncthreads();
        if (s_bad) {  // a cell beyond the bitonic buffer: leave the tile to the fallback
            if (t == 0) {
                atomicAdd(p.irregular, 1ull);
                p.flag[tile] = 1;
            }
            finish_without_output();
            return;
        }
    }
    if constexpr (!FUSED) {
        KeyT* kout = reinterpret_cast<KeyT*>(p.keys_out) + s;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t idx = t + j * LS_BLOCK;
            if (idx < m) kout[idx] = s_stage[idx];
        }
        if constexpr (PAIRS) {
            const ValT* vin = reinterpret_cast<const ValT*>(p.vals_in) + s;
            ValT* vout = reinterpret_cast<ValT*>(p.vals_out) + s;
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const uint32_t idx = t + j * LS_BLOCK;
                if (idx < m) vout[idx] = vin[s_idx[idx]];
            }
        }
    } else {
        // Position-parallel emission.  COUNT: a position is a run head when its key differs from the one
        // before; UNIQ: a singleton when it also differs from the one after.  Ballot word (j, warp) covers
        // the 32 consecutive positions j * 512 + warp * 32 ..; the words' popcounts are scanned by warp 0.
        uint32_t* s_mask = s_cell;          // [NW] ballots (the cell array is dead by now)
        uint32_t* s_woff = s_cell + NW;     // [NW + 1] emitted items before each ballot word
        uint32_t flags = 0;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t pos = t + j * LS_BLOCK;
            bool f = false;
            if (pos < m) {
                const KeyT key = s_stage[pos];
                f = pos == 0 || s_stage[pos - 1] != key;
                if constexpr (UNIQ) f = f && (pos + 1 == m || s_stage[pos + 1] != key);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_mask[j * (LS_BLOCK / 32) + warp] = bal;
            flags |= (f ? 1u : 0u) << j;
        }
        __syncthreads();
        if (warp == 0) {
            // ballot word w (position order) = s_mask[(w % 16) ... ]: position = j * 512 + warp * 32, so
            // word index in position order is j * 16 + warp -- the storage order already
            constexpr int PER = NW / 32;
            uint32_t c[PER], sum = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                c[i] = __popc(s_mask[lane * PER + i]);
                sum += c[i];
            }
            const uint32_t incl = warp_incl_scan(sum);
            uint32_t run = incl - sum;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                s_woff[lane * PER + i] = run;
                run += c[i];
            }
            const uint32_t H = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) {
                s_total = H;
                tile_prefix_publish(p.tile_state, tile, H);
            }
            const uint64_t base = tile_prefix_resolve_warp(p.tile_state, p.n_tiles, tile, H, p.err);
            if (lane == 0) {
                s_base = base;
                if (tile == p.n_tiles - 1) *p.n_out = *p.n_out_copy = base + H;
            }
        }
        __syncthreads();
        const uint64_t base = s_base;
        KeyT* kout = reinterpret_cast<KeyT*>(p.keys_out);
        constexpr uint32_t WPR = LS_BLOCK / 32;  // ballot words per round j
        const uint32_t n_words = (m + 31) / 32;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t pos = t + j * LS_BLOCK;
            const bool f = (flags >> j) & 1u;
            const uint32_t wi = j * WPR + warp;
            const uint32_t bal = __ballot_sync(0xffffffffu, f);
            if (!f) continue;
            const uint64_t h = base + s_woff[wi] + __popc(bal & lanemask_lt());
            kout[h] = s_stage[pos];
            if constexpr (COUNT) {
                // run length = distance to the next head (ballot bits above mine, then the following words)
                const uint32_t above = (bal >> lane) >> 1;
                uint32_t nxt;
                if (above) {
                    nxt = pos + (uint32_t)__ffs(above);
                } else {
                    uint32_t w2 = wi + 1;
                    while (w2 < n_words && s_mask[w2] == 0) ++w2;
                    nxt = w2 < n_words ? w2 * 32 + (uint32_t)__ffs(s_mask[w2]) - 1 : m;
                }
                p.counts_out[h] = nxt - pos;
            } else {
                const ValT* vin = reinterpret_cast<const ValT*>(p.vals_in) + s;
                reinterpret_cast<ValT*>(p.vals_out)[h] = vin[s_idx[pos]];
            }
        }
    }
}
