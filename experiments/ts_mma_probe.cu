// Probe (not part of the product): does an A operand staged in tensor memory (tcgen05.cp smem->TMEM, then TS-mode tcgen05.mma)
// lift the ~59-cycle cost of an SS-mode M=128 x K=16 MMA seen by the 64-channel kernels?  Checks the copy layout and the TS result
// against the SS result, then times instruction streams with clock64 (one CTA per SM, one issuing thread).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <type_traits>
#include <cuda_bf16.h>
#include "../cwfa_b200/csrc/tc_common.cuh"
using namespace cwfa::tcx;

constexpr uint32_t kABytes = 8 * 128 * 16;      // [8 chunks][128 rows][8 bf16]
constexpr uint32_t kBBytes = 8 * 256 * 16;      // [8 chunks][256 n][8 bf16]
constexpr uint32_t kHdr = 1024;
constexpr uint32_t kSmem = kHdr + kABytes + kBBytes;

__device__ __forceinline__ void tc_cp_128x256b(uint32_t taddr, uint32_t lo, uint32_t hi) {
    asm volatile("{\n.reg .b64 d;\nmov.b64 d, {%1, %2};\ntcgen05.cp.cta_group::1.128x256b [%0], d;\n}" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(128) probe(uint32_t* dump_a, float* out_ts, float* out_ss, long long* cyc, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t s0 = smem_u32(smem);
    const uint32_t sA = s0 + kHdr, sB = sA + kABytes, bar = s0 + 8, slot = s0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem + kHdr);
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + kHdr + kABytes);
    for (int i = tid; i < 128 * 64; i += 128) {
        const int m = i / 64, k = i % 64;
        A[(k / 8) * 128 * 8 + m * 8 + (k % 8)] = __float2bfloat16((float)(((m * 3 + k * 5) % 7) - 3));
    }
    for (int i = tid; i < 256 * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        B[(k / 8) * 256 * 8 + n * 8 + (k % 8)] = __float2bfloat16((float)(((n * 2 + k) % 5) - 2));
    }
    if (tid == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem);
    const uint32_t a_lbo = 128 * 16, a_sbo = 128, b_lbo = 256 * 16, b_sbo = 128;
    const uint32_t tmA = tm + 256;              // A staging: 4 K-steps x 8 columns
    uint32_t phase = 0;
    // ---- 1. copy A to TMEM, dump it
    if (tid == 0) {
        for (int ks = 0; ks < 4; ++ks) tc_cp_128x256b(tmA + 8 * ks, desc_lo(sA + ks * 2 * a_lbo, a_lbo), desc_hi(a_sbo));
        tc_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    {
        uint32_t r[32];
        tmem_ld32_nowait(tmA + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait();
        if (blockIdx.x == 0) for (int j = 0; j < 32; ++j) dump_a[(warp * 32 + lane) * 32 + j] = r[j];
    }
    tc_fence_before();
    __syncthreads();
    // ---- 2. TS result into cols 0..63, SS result into cols 64..127
    const uint32_t id64 = idesc_f16(64, 1);
    if (tid == 0) {
        tc_fence_after();
        for (int ks = 0; ks < 4; ++ks) tc_mma_ts(tm, tmA + 8 * ks, desc_lo(sB + ks * 2 * b_lbo, b_lbo), desc_hi(b_sbo), id64, ks > 0);
        for (int ks = 0; ks < 4; ++ks)
            tc_mma_f16_split(tm + 64, desc_lo(sA + ks * 2 * a_lbo, a_lbo), desc_hi(a_sbo), desc_lo(sB + ks * 2 * b_lbo, b_lbo), desc_hi(b_sbo), id64, ks > 0);
        tc_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    for (int h = 0; h < 4; ++h) {
        uint32_t r[32];
        tmem_ld32_nowait(tm + ((uint32_t)(warp * 32) << 16) + 32 * h, r);
        tmem_ld_wait();
        if (blockIdx.x == 0) {
            float* o = (h < 2 ? out_ts : out_ss) + (warp * 32 + lane) * 64 + 32 * (h & 1);
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    // ---- 3. timings
    uint32_t alo[4], blo[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) { alo[ks] = desc_lo(sA + ks * 2 * a_lbo, a_lbo); blo[ks] = desc_lo(sB + ks * 2 * b_lbo, b_lbo); }
    const uint32_t ahi = desc_hi(a_sbo), bhi = desc_hi(b_sbo);
    auto run = [&](auto vtag) {
        constexpr int variant = decltype(vtag)::value;
        long long t0 = 0;
        if (tid == 0) {
            tc_fence_after();
            constexpr int n = (variant == 4 || variant == 6) ? 128 : (variant == 5 || variant == 7) ? 256 : (variant == 9 ? 96 : (variant == 10 ? 32 : 64));
            const uint32_t id = idesc_f16(n, 1);
            t0 = clock64();
            for (int i = 0; i < iters; i += 12) {
#pragma unroll
                for (int u = 0; u < 12; ++u) {
                    const int ks = u & 3;
                    if constexpr (variant == 0 || variant == 4 || variant == 5 || variant == 8) tc_mma_f16_split(tm, alo[ks], ahi, blo[ks], bhi, id, 1);
                    if constexpr (variant == 1 || variant == 2 || variant == 6 || variant == 7 || variant == 9 || variant == 10 || variant == 11)
                        tc_mma_ts(tm, tmA + 8 * ks, blo[ks], bhi, id, 1);
                    if constexpr (variant == 2 || variant == 8) { if (u % 3 == 0) tc_cp_128x256b(tm + 384 + 8 * (u / 3), alo[ks], ahi); }
                    if constexpr (variant == 3 || variant == 11) tc_cp_128x256b(tm + 384 + 8 * (u & 7), alo[ks], ahi);
                }
            }
            tc_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1;
        if (tid == 0) {
            const long long t1 = clock64();
            if (blockIdx.x == 0) cyc[variant] = t1 - t0;
        }
        tc_fence_before();
        __syncthreads();
    };
    run(std::integral_constant<int, 0>{}); run(std::integral_constant<int, 1>{}); run(std::integral_constant<int, 2>{});
    run(std::integral_constant<int, 3>{}); run(std::integral_constant<int, 4>{}); run(std::integral_constant<int, 5>{});
    run(std::integral_constant<int, 6>{}); run(std::integral_constant<int, 7>{}); run(std::integral_constant<int, 8>{});
    run(std::integral_constant<int, 9>{}); run(std::integral_constant<int, 10>{}); run(std::integral_constant<int, 11>{});
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 3072;
    const int grid = argc > 2 ? atoi(argv[2]) : 1;
    uint32_t* dump; float *ots, *oss; long long* cyc;
    cudaMalloc(&dump, 128 * 32 * 4); cudaMalloc(&ots, 128 * 64 * 4); cudaMalloc(&oss, 128 * 64 * 4); cudaMalloc(&cyc, 16 * 8);
    cudaMemset(cyc, 0, 16 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    probe<<<grid, 128, kSmem>>>(dump, ots, oss, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> hd(128 * 32); std::vector<float> ht(128 * 64), hs(128 * 64); long long hc[16];
    cudaMemcpy(hd.data(), dump, hd.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ht.data(), ots, ht.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs.data(), oss, hs.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    // expected TMEM image of A: lane m, column j holds bf16 pair (k = 2j, 2j+1)
    int bad_copy = 0;
    for (int m = 0; m < 128; ++m)
        for (int j = 0; j < 32; ++j) {
            auto bf = [](float f) { __nv_bfloat16 b = __float2bfloat16(f); return (uint32_t)*reinterpret_cast<uint16_t*>(&b); };
            const uint32_t exp = bf((float)(((m * 3 + (2 * j) * 5) % 7) - 3)) | (bf((float)(((m * 3 + (2 * j + 1) * 5) % 7) - 3)) << 16);
            if (hd[m * 32 + j] != exp) { if (bad_copy < 6) printf("copy mismatch lane %d col %d: got %08x want %08x\n", m, j, hd[m * 32 + j], exp); ++bad_copy; }
        }
    int bad_ss = 0, bad_ts = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            float ref = 0;
            for (int k = 0; k < 64; ++k) ref += (float)(((m * 3 + k * 5) % 7) - 3) * (float)(((n * 2 + k) % 5) - 2);
            if (hs[m * 64 + n] != ref) ++bad_ss;
            if (ht[m * 64 + n] != ref) { if (bad_ts < 4) printf("TS mismatch m %d n %d: got %g want %g (ss %g)\n", m, n, ht[m * 64 + n], ref, hs[m * 64 + n]); ++bad_ts; }
        }
    printf("copy mismatches %d / 4096; SS result mismatches %d; TS result mismatches %d (of 8192)\n", bad_copy, bad_ss, bad_ts);
    const char* names[12] = {"SS N=64", "TS N=64", "TS N=64 + 1 cp(128x256b) per 3 MMAs", "cp 128x256b alone", "SS N=128", "SS N=256", "TS N=128", "TS N=256",
                             "SS N=64 + 1 cp per 3 MMAs", "TS N=96", "TS N=32", "TS N=64 + 1 cp per MMA"};
    for (int v = 0; v < 12; ++v) printf("%-40s %8.1f cycles / iteration\n", names[v], (double)hc[v] / iters);
    return 0;
}
