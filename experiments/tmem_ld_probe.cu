// Probe: tcgen05.ld throughput per SM (bytes / clock) for 4, 8, 16 warps reading 32-column (4 KB per warp-instruction) or
// 16-column slices back to back.  One CTA; each warp reads its own lane quadrant.
#include <cstdio>
#include <cstdlib>
#include "../cwfa_b200/csrc/tc_common.cuh"
using namespace cwfa::tcx;

template <int COLS>
__global__ void probe(long long* cyc, int iters, uint32_t* sink) {
    __shared__ uint32_t slot;
    __shared__ long long t0s, t1s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if constexpr (COLS == 32) {
            uint32_t r[32];
            tmem_ld32_nowait(tm + ((i * 32) & 511), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j];
        } else {
            uint32_t r[16];
            tmem_ld16(tm + ((i * 16) & 511), r);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= r[j];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    sink[threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4096);
    const int iters = 2048;
    for (int cols : {32, 16})
        for (int warps : {1, 4, 8, 16}) {
            if (cols == 32) probe<32><<<1, warps * 32>>>(cyc, iters, sink); else probe<16><<<1, warps * 32>>>(cyc, iters, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * warps * 32 * cols * 4;
            printf("tcgen05.ld 32x32b.x%d, %2d warps: %.1f cycles per load per warp, %.1f B/clk per SM\n", cols, warps, (double)h / iters, bytes / h);
        }
    return 0;
}
