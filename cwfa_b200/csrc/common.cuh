// Shared helpers for the cwfa_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/cwfa_b200.h"

namespace cwfa {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of two values; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ void block_sum2(float& a, float& b) {
    __shared__ float sa[32], sb[32];
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        a = lane < nw ? sa[lane] : 0.f;
        b = lane < nw ? sb[lane] : 0.f;
        a = warp_sum(a);
        b = warp_sum(b);
    }
    __syncthreads();
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    switch (act) {
        case CWFA_ACT_ELU: return v > 0.f ? v : expm1f(v);
        case CWFA_ACT_PRELU: return v >= 0.f ? v : slope * v;
        case CWFA_ACT_RELU: return fmaxf(v, 0.f);
        case CWFA_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case CWFA_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

}  // namespace cwfa
