// FASTA text -> flat base buffer on the GPU (the step before the hot path).
//
// Text rules of the reference's SmartFastaParser (kmermaid/parsers.py:53-128) + the record
// naming input of FastaRecordBatcher.do (kmermaid/batcher.py:551):
//   * everything before the first line starting with '>' is skipped (parsers.py:53-66)
//   * a line starting with '>' is a header; the following lines, up to the next header, are the
//     record's sequence with line terminators and ' ' removed (parsers.py:82,124)
// Output layout = kman_b200/fasta.py's: record bytes, one '\n' after every record.
// Bytes the reference treats specially only at line ends (TAB, VT, FF, FS..US: str.rstrip()) and
// non-ASCII bytes are not handled here: the kernel raises a flag and the host loader takes over.
//
// A byte's fate depends on the kind of the line it is on, i.e. on a 3-state machine driven by
// line starts (PRE: before the first header, HDR: header line, SEQ: sequence line).  Segments of
// text compose as (state function, sequence-byte count per entry state, header count), which is a
// monoid, so: 1. every 4 KB tile summarises itself, 2. one block scans the tile summaries,
// 3. every tile re-walks its bytes with its entry state and output offsets and writes.
#include "common.cuh"

namespace kmg {

constexpr int FA_BLOCK = 256;
constexpr int FA_BPT = 16;  // bytes per thread
constexpr int FA_TILE = FA_BLOCK * FA_BPT;
enum { ST_PRE = 0, ST_HDR = 1, ST_SEQ = 2 };

// state function packed as 2 bits per entry state; counts of kept bytes per entry state
struct Seg {
    uint32_t f;      // out state for in = PRE | HDR << 2 | SEQ << 4
    uint32_t c[3];   // sequence bytes if the segment is entered in state s
    uint32_t nh;     // header lines starting in the segment
};
__device__ __forceinline__ uint32_t seg_apply(uint32_t f, uint32_t s) { return (f >> (2 * s)) & 3u; }
constexpr uint32_t F_ID = ST_PRE | (ST_HDR << 2) | (ST_SEQ << 4);
__device__ __forceinline__ Seg seg_identity() { return Seg{F_ID, {0, 0, 0}, 0}; }
// a then b
__device__ __forceinline__ Seg seg_compose(const Seg& a, const Seg& b) {
    Seg r;
    r.f = seg_apply(b.f, seg_apply(a.f, 0)) | (seg_apply(b.f, seg_apply(a.f, 1)) << 2) | (seg_apply(b.f, seg_apply(a.f, 2)) << 4);
#pragma unroll
    for (int s = 0; s < 3; ++s) r.c[s] = a.c[s] + b.c[seg_apply(a.f, s)];
    r.nh = a.nh + b.nh;
    return r;
}
__device__ __forceinline__ Seg seg_shfl_up(const Seg& v, int o) {
    Seg r;
    r.f = __shfl_up_sync(0xffffffffu, v.f, o);
    r.c[0] = __shfl_up_sync(0xffffffffu, v.c[0], o);
    r.c[1] = __shfl_up_sync(0xffffffffu, v.c[1], o);
    r.c[2] = __shfl_up_sync(0xffffffffu, v.c[2], o);
    r.nh = __shfl_up_sync(0xffffffffu, v.nh, o);
    return r;
}

__device__ __forceinline__ bool is_term(uint32_t c) { return c == 10u || c == 13u; }
__device__ __forceinline__ bool is_special(uint32_t c) { return c == 9u || c == 11u || c == 12u || (c >= 28u && c <= 31u) || c >= 128u; }

// summary of one thread's FA_BPT bytes; `prev` = byte before the first one (a terminator at file start)
__device__ __forceinline__ Seg seg_of_bytes(const uint8_t (&b)[FA_BPT], int n, uint32_t prev, uint32_t& special) {
    Seg s = seg_identity();
    uint32_t st[3] = {ST_PRE, ST_HDR, ST_SEQ};  // current state per entry state
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        if (i < n) {
            const uint32_t c = b[i];
            special |= is_special(c) ? 1u : 0u;
            if (is_term(prev)) {  // line start
                if (c == '>') {
                    st[0] = st[1] = st[2] = ST_HDR;
                    ++s.nh;
                } else {
#pragma unroll
                    for (int e = 0; e < 3; ++e)
                        if (st[e] != ST_PRE) st[e] = ST_SEQ;  // PRE stays PRE until the first header
                }
            }
            const bool keep = !is_term(c) && c != ' ';
#pragma unroll
            for (int e = 0; e < 3; ++e) s.c[e] += (keep && st[e] == ST_SEQ) ? 1u : 0u;
            prev = c;
        }
    }
    s.f = st[0] | (st[1] << 2) | (st[2] << 4);
    return s;
}

struct FastaParams {
    const uint8_t* raw;
    uint64_t n_raw;
    uint32_t n_tiles;
    Seg* tile_seg;           // [tiles] summaries
    uint32_t* tile_state;    // [tiles] entry state
    uint64_t* tile_seq;      // [tiles] sequence bytes before the tile
    uint64_t* tile_hdr;      // [tiles] headers before the tile
    uint8_t* bases_out;
    uint64_t* rec_starts;    // [max_rec + 1]
    uint64_t* hdr_begin;     // [max_rec] raw offset of the first title character
    uint64_t max_rec;
    unsigned long long* totals;  // [0] n_bases_out, [1] n_rec, [2] special-character flag
};

__device__ __forceinline__ void load_thread_bytes(const FastaParams& p, uint64_t pos, uint8_t (&b)[FA_BPT], int& n, uint32_t& prev) {
    n = pos >= p.n_raw ? 0 : (p.n_raw - pos >= FA_BPT ? FA_BPT : (int)(p.n_raw - pos));
    if (n == FA_BPT && ((uintptr_t)(p.raw + pos) & 15) == 0) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.raw + pos));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < FA_BPT; ++i) b[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
    } else {
#pragma unroll
        for (int i = 0; i < FA_BPT; ++i) b[i] = i < n ? p.raw[pos + i] : 0;
    }
    prev = pos == 0 ? 10u : (pos <= p.n_raw ? p.raw[pos - 1] : 0u);
}

// block-wide exclusive scan of segments in thread order; s_w needs FA_BLOCK/32 entries
__device__ __forceinline__ Seg block_excl_scan_seg(const Seg& mine, Seg* s_w, Seg& total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Seg incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Seg t = seg_shfl_up(incl, o);
        if (lane >= (uint32_t)o) incl = seg_compose(t, incl);
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    Seg wpre = seg_identity();
    Seg tot = seg_identity();
    for (int w = 0; w < FA_BLOCK / 32; ++w) {
        if (w == (int)warp) wpre = tot;
        tot = seg_compose(tot, s_w[w]);
    }
    total = tot;
    Seg excl = seg_shfl_up(incl, 1);
    if (lane == 0) excl = seg_identity();
    return seg_compose(wpre, excl);
}

__global__ void __launch_bounds__(FA_BLOCK) fasta_summaries_kernel(const FastaParams p) {
    __shared__ Seg s_w[FA_BLOCK / 32];
    const uint64_t pos = (uint64_t)blockIdx.x * FA_TILE + (uint64_t)threadIdx.x * FA_BPT;
    uint8_t b[FA_BPT];
    int n;
    uint32_t prev, special = 0;
    load_thread_bytes(p, pos, b, n, prev);
    const Seg mine = seg_of_bytes(b, n, prev, special);
    Seg total;
    block_excl_scan_seg(mine, s_w, total);
    if (threadIdx.x == 0) p.tile_seg[blockIdx.x] = total;
    if (special) atomicExch(&p.totals[2], 1ull);
}

// one block: exclusive scan over the tile summaries, starting in state PRE
__global__ void __launch_bounds__(1024) fasta_scan_kernel(const FastaParams p) {
    __shared__ uint32_t s_f[32];
    __shared__ uint64_t s_c[32][3], s_h[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per = ((p.n_tiles + 31) / 32 + 31) / 32 * 32;
    const uint32_t b = min(warp * per, p.n_tiles), e = min(b + per, p.n_tiles);
    // pass 1: the warp's segment as one summary (64-bit counts), lanes walk it 32 tiles at a time
    uint32_t f = F_ID;
    uint64_t c[3] = {0, 0, 0}, nh = 0;
    for (uint32_t i0 = b; i0 < e; i0 += 32) {
        const uint32_t i = i0 + lane;
        Seg v = i < e ? p.tile_seg[i] : seg_identity();
        Seg incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const Seg t = seg_shfl_up(incl, o);
            if (lane >= (uint32_t)o) incl = seg_compose(t, incl);
        }
        Seg step;  // the 32 tiles together (lane 31's inclusive value)
        step.f = __shfl_sync(0xffffffffu, incl.f, 31);
        step.c[0] = __shfl_sync(0xffffffffu, incl.c[0], 31);
        step.c[1] = __shfl_sync(0xffffffffu, incl.c[1], 31);
        step.c[2] = __shfl_sync(0xffffffffu, incl.c[2], 31);
        step.nh = __shfl_sync(0xffffffffu, incl.nh, 31);
        uint64_t nc[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) nc[s] = c[s] + step.c[seg_apply(f, s)];
        f = seg_apply(step.f, seg_apply(f, 0)) | (seg_apply(step.f, seg_apply(f, 1)) << 2) | (seg_apply(step.f, seg_apply(f, 2)) << 4);
        c[0] = nc[0]; c[1] = nc[1]; c[2] = nc[2];
        nh += step.nh;
    }
    if (lane == 0) {
        s_f[warp] = f;
        s_c[warp][0] = c[0]; s_c[warp][1] = c[1]; s_c[warp][2] = c[2];
        s_h[warp] = nh;
    }
    __syncthreads();
    // entry state / offsets of every warp segment (serial over 32 summaries), the file starts in PRE
    uint32_t st = ST_PRE;
    uint64_t seq = 0, hdr = 0;
    for (uint32_t w = 0; w < warp; ++w) {
        seq += s_c[w][st];
        hdr += s_h[w];
        st = seg_apply(s_f[w], st);
    }
    // pass 2: per-tile entry state and offsets
    for (uint32_t i0 = b; i0 < e; i0 += 32) {
        const uint32_t i = i0 + lane;
        Seg v = i < e ? p.tile_seg[i] : seg_identity();
        Seg incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const Seg t = seg_shfl_up(incl, o);
            if (lane >= (uint32_t)o) incl = seg_compose(t, incl);
        }
        Seg excl = seg_shfl_up(incl, 1);
        if (lane == 0) excl = seg_identity();
        if (i < e) {
            p.tile_state[i] = seg_apply(excl.f, st);
            p.tile_seq[i] = seq + excl.c[st];
            p.tile_hdr[i] = hdr + excl.nh;
        }
        const uint32_t sf = __shfl_sync(0xffffffffu, incl.f, 31);
        const uint32_t sc = __shfl_sync(0xffffffffu, incl.c[st], 31);
        const uint32_t sh = __shfl_sync(0xffffffffu, incl.nh, 31);
        seq += sc;
        hdr += sh;
        st = seg_apply(sf, st);
    }
    if (threadIdx.x == 1023) {  // the last warp's running totals are the file's
        p.totals[0] = seq + hdr;  // one separator per record
        p.totals[1] = hdr;
    }
}

__global__ void __launch_bounds__(FA_BLOCK) fasta_emit_kernel(const FastaParams p) {
    __shared__ Seg s_w[FA_BLOCK / 32];
    const uint32_t tile = blockIdx.x;
    const uint64_t pos = (uint64_t)tile * FA_TILE + (uint64_t)threadIdx.x * FA_BPT;
    uint8_t b[FA_BPT];
    int n;
    uint32_t prev, special = 0;
    load_thread_bytes(p, pos, b, n, prev);
    const Seg mine = seg_of_bytes(b, n, prev, special);
    Seg total;
    const Seg excl = block_excl_scan_seg(mine, s_w, total);
    const uint32_t st_tile = p.tile_state[tile];
    uint32_t st = seg_apply(excl.f, st_tile);
    uint64_t seq = p.tile_seq[tile] + excl.c[st_tile];
    uint64_t hdr = p.tile_hdr[tile] + excl.nh;
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i) {
        if (i < n) {
            const uint32_t c = b[i];
            if (is_term(prev)) {
                if (c == '>') {
                    st = ST_HDR;
                    if (hdr < p.max_rec) {
                        p.rec_starts[hdr] = seq + hdr;
                        p.hdr_begin[hdr] = pos + i + 1;
                    }
                    if (hdr > 0) p.bases_out[seq + hdr - 1] = 10;  // separator closing the previous record
                    ++hdr;
                } else if (st != ST_PRE) {
                    st = ST_SEQ;
                }
            }
            if (st == ST_SEQ && !is_term(c) && c != ' ') {
                p.bases_out[seq + hdr - 1] = (uint8_t)c;
                ++seq;
            }
            prev = c;
        }
    }
    if (tile == gridDim.x - 1 && threadIdx.x == FA_BLOCK - 1 && hdr > 0) {
        p.bases_out[seq + hdr - 1] = 10;  // separator after the last record
        if (hdr <= p.max_rec) p.rec_starts[hdr] = seq + hdr;
    }
}

}  // namespace kmg

using namespace kmg;

extern "C" size_t kmg_fasta_workspace_bytes(uint64_t n_raw) {
    const uint64_t tiles = n_raw / FA_TILE + 2;
    return align_up(tiles * sizeof(Seg), 256) + align_up(tiles * 4, 256) + 2 * align_up(tiles * 8, 256);
}

extern "C" int kmg_fasta_flatten(const uint8_t* d_raw, uint64_t n_raw, uint8_t* d_bases_out, uint64_t* d_rec_starts,
                                 uint64_t* d_hdr_begin, uint64_t max_rec, uint64_t* d_totals, void* d_ws, size_t ws_bytes,
                                 void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KMG_REQUIRE(d_totals && d_ws, KMG_ERR_ARG, "null pointer argument");
    KMG_CUDA(cudaMemsetAsync(d_totals, 0, 3 * sizeof(uint64_t), st));
    if (n_raw == 0) return KMG_OK;
    KMG_REQUIRE(d_raw && d_bases_out && d_rec_starts && d_hdr_begin, KMG_ERR_ARG, "null pointer argument");
    KMG_REQUIRE(ws_bytes >= kmg_fasta_workspace_bytes(n_raw), KMG_ERR_WS, "fasta workspace too small");
    const uint64_t tiles = (n_raw + FA_TILE - 1) / FA_TILE;
    KMG_REQUIRE(tiles < (1ull << 31), KMG_ERR_RANGE, "too many tiles");
    const uint64_t tmax = n_raw / FA_TILE + 2;
    FastaParams p;
    p.raw = d_raw;
    p.n_raw = n_raw;
    p.n_tiles = (uint32_t)tiles;
    char* at = (char*)d_ws;
    p.tile_seg = (Seg*)at; at += align_up(tmax * sizeof(Seg), 256);
    p.tile_state = (uint32_t*)at; at += align_up(tmax * 4, 256);
    p.tile_seq = (uint64_t*)at; at += align_up(tmax * 8, 256);
    p.tile_hdr = (uint64_t*)at;
    p.bases_out = d_bases_out;
    p.rec_starts = d_rec_starts;
    p.hdr_begin = d_hdr_begin;
    p.max_rec = max_rec;
    p.totals = reinterpret_cast<unsigned long long*>(d_totals);
    fasta_summaries_kernel<<<(unsigned)tiles, FA_BLOCK, 0, st>>>(p);
    KMG_LAUNCH_CHECK();
    fasta_scan_kernel<<<1, 1024, 0, st>>>(p);
    KMG_LAUNCH_CHECK();
    fasta_emit_kernel<<<(unsigned)tiles, FA_BLOCK, 0, st>>>(p);
    KMG_LAUNCH_CHECK();
    return KMG_OK;
}
