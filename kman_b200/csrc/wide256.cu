// 256-bit keys: the WIDE stream (windows holding N / IUPAC symbols, 4-bit ASCII-rank codes,
// kmermaid/seq.py:317-318 has no bound on k for them) at 33 <= k <= 64.  Such windows are the
// exception (N runs, isolated ambiguity codes), so this path is built from the 128-bit machinery
// instead of a fourth instantiation of every sort kernel: a stable LSD sort of the LOW halves
// carrying the element index, then of the HIGH halves in that order, then one gather --
// batch.py:156-168's stable order on the full key.  The run-length / singleton stage (rle.cu) and
// the text emission (emit.cu) take 32-byte keys directly.
#include "common.cuh"

namespace kmg {

__global__ void split256_kernel(const u256* __restrict__ keys, uint64_t n, u128* __restrict__ lo, uint64_t* __restrict__ idx) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lo[i] = keys[i].lo;
    idx[i] = i;
}
__global__ void gather_hi256_kernel(const u256* __restrict__ keys, const uint64_t* __restrict__ idx, uint64_t n,
                                    u128* __restrict__ hi) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hi[i] = keys[idx[i]].hi;
}
template <typename ValT>
__global__ void gather256_kernel(const u256* __restrict__ keys, const ValT* __restrict__ vals, const uint64_t* __restrict__ idx,
                                 uint64_t n, u256* __restrict__ keys_out, ValT* __restrict__ vals_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t j = idx[i];
    keys_out[i] = keys[j];
    if (vals) vals_out[i] = vals[j];
}

struct Ws256 {
    char *k0, *k1, *v0, *v1, *sort_ws;
    size_t sort_ws_bytes, total;
};
static Ws256 carve256(void* ws, uint64_t n) {
    Ws256 w;
    char* p = (char*)ws;
    auto take = [&](size_t b) {
        char* r = p;
        p += align_up(b, 256);
        return r;
    };
    w.k0 = take(n * 16);
    w.k1 = take(n * 16);
    w.v0 = take(n * 8);
    w.v1 = take(n * 8);
    w.sort_ws_bytes = kmg_radix_sort_workspace_bytes(n, 16, 8, 0, 128);
    w.sort_ws = take(w.sort_ws_bytes);
    w.total = (size_t)(p - (char*)ws);
    return w;
}

}  // namespace kmg

using namespace kmg;

extern "C" size_t kmg_sort256_workspace_bytes(uint64_t n) { return carve256(nullptr, n ? n : 1).total; }

extern "C" int kmg_sort256(const void* d_keys, void* d_keys_out, const void* d_vals, void* d_vals_out, uint64_t n,
                           int val_bytes, int end_bit, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KMG_REQUIRE(val_bytes == 0 || val_bytes == 4 || val_bytes == 8, KMG_ERR_ARG, "val_bytes must be 0, 4 or 8");
    KMG_REQUIRE(end_bit >= 0 && end_bit <= 256, KMG_ERR_ARG, "bad end_bit %d", end_bit);
    if (n == 0) return KMG_OK;
    KMG_REQUIRE(d_keys && d_keys_out && d_ws && (val_bytes == 0 || (d_vals && d_vals_out)), KMG_ERR_ARG, "null pointer argument");
    KMG_REQUIRE(((uintptr_t)d_keys & 15) == 0 && ((uintptr_t)d_keys_out & 15) == 0, KMG_ERR_ARG, "key buffers misaligned");
    const Ws256 w = carve256(d_ws, n);
    KMG_REQUIRE(ws_bytes >= w.total, KMG_ERR_WS, "sort256 workspace too small: %zu < %zu", ws_bytes, w.total);
    const unsigned grid = (unsigned)((n + 255) / 256);
    const u256* keys = reinterpret_cast<const u256*>(d_keys);
    split256_kernel<<<grid, 256, 0, st>>>(keys, n, (u128*)w.k0, (uint64_t*)w.v0);
    KMG_LAUNCH_CHECK();
    char *kc = w.k0, *ka = w.k1, *vc = w.v0, *va = w.v1;  // current / alternate
    int sel = 0;
    int rcode = kmg_radix_sort(kc, ka, vc, va, n, 16, 8, 0, end_bit < 128 ? end_bit : 128, nullptr, &sel, w.sort_ws,
                               w.sort_ws_bytes, stream);
    if (rcode != KMG_OK) return rcode;
    if (sel) {
        std::swap(kc, ka);
        std::swap(vc, va);
    }
    if (end_bit > 128) {
        gather_hi256_kernel<<<grid, 256, 0, st>>>(keys, (const uint64_t*)vc, n, (u128*)kc);
        KMG_LAUNCH_CHECK();
        rcode = kmg_radix_sort(kc, ka, vc, va, n, 16, 8, 0, end_bit - 128, nullptr, &sel, w.sort_ws, w.sort_ws_bytes, stream);
        if (rcode != KMG_OK) return rcode;
        if (sel) std::swap(vc, va);
    }
    if (val_bytes == 4)
        gather256_kernel<uint32_t><<<grid, 256, 0, st>>>(keys, (const uint32_t*)d_vals, (const uint64_t*)vc, n, (u256*)d_keys_out,
                                                       (uint32_t*)d_vals_out);
    else
        gather256_kernel<uint64_t><<<grid, 256, 0, st>>>(keys, (const uint64_t*)d_vals, (const uint64_t*)vc, n, (u256*)d_keys_out,
                                                       (uint64_t*)d_vals_out);
    KMG_LAUNCH_CHECK();
    return KMG_OK;
}
