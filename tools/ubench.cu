// Micro-benchmarks of warp-level digit ranking variants on sm_100a (development tool).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int BLOCK = 256, IPT = 16, RADIX = 256, WARPS = BLOCK / 32;

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int VARIANT>
__global__ void __launch_bounds__(BLOCK) rank_kernel(uint32_t* out, int reps, uint32_t digit_mask) {
    __shared__ uint32_t s_hist[WARPS * RADIX];
    __shared__ uint32_t s_tbl[WARPS * RADIX];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* my_hist = s_hist + warp * RADIX;
    uint32_t* my_tbl = s_tbl + warp * RADIX;
    for (int i = lane; i < RADIX; i += 32) { my_hist[i] = 0; my_tbl[i] = 0; }
    __syncwarp();
    uint32_t acc = 0;
    const uint32_t lt = lanemask_lt();
    for (int r = 0; r < reps; ++r) {
        uint32_t dg[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) dg[i] = hash32((blockIdx.x * BLOCK + threadIdx.x) * 977 + i * 131 + r * 7919) & digit_mask;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint32_t d = dg[i];
            uint32_t m;
            if (VARIANT == 0) {  // 8 ballots, xor form
                m = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const uint32_t bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                    m &= bal ^ (((d >> b) & 1u) - 1u);
                }
            } else if (VARIANT == 1) {  // hardware match
                m = __match_any_sync(0xffffffffu, d);
            } else if (VARIANT == 2) {  // shared-memory or-table
                atomicOr(&my_tbl[d], 1u << lane);
                __syncwarp();
                m = my_tbl[d];
                __syncwarp();
                if ((m & lt) == 0) my_tbl[d] = 0;
            } else if (VARIANT == 3) {  // 8 ballots, predicate + select form
                m = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const bool bit = d & (1u << b);
                    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                    m &= bit ? bal : ~bal;
                }
            } else {  // 4: no matching at all (floor: counter update only, WRONG ranks)
                m = 1u << lane;
            }
            const uint32_t lower = __popc(m & lt);
            uint32_t cur = 0;
            if (lower == 0) { cur = my_hist[d]; my_hist[d] = cur + __popc(m); }
            __syncwarp();
            cur = __shfl_sync(0xffffffffu, cur, __ffs(m) - 1);
            acc += cur + lower;
        }
    }
    out[blockIdx.x * BLOCK + threadIdx.x] = acc;
}

template <int V>
void run(const char* name, uint32_t* d_out, int grid, int reps, uint32_t mask) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rank_kernel<V><<<grid, BLOCK>>>(d_out, 2, mask);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    rank_kernel<V><<<grid, BLOCK>>>(d_out, reps, mask);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double items = (double)grid * BLOCK * IPT * reps;
    printf("%-28s mask=0x%02x: %8.3f ms  %8.2f G keys/s  (%.2f ns per warp-item per SM)\n", name, mask, ms, items / ms / 1e6,
           ms * 1e6 / (items / 32 / 148));
}

int main() {
    uint32_t* d_out; cudaMalloc(&d_out, 148 * 8 * BLOCK * 4);
    const int grid = 148 * 4, reps = 200;
    for (uint32_t mask : {0xffu, 0x3fu, 0x03u, 0x00u}) {
        run<0>("ballot xor (current)", d_out, grid, reps, mask);
        run<3>("ballot select", d_out, grid, reps, mask);
        run<1>("match.any", d_out, grid, reps, mask);
        run<2>("smem or-table", d_out, grid, reps, mask);
        run<4>("no match (floor)", d_out, grid, reps, mask);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
