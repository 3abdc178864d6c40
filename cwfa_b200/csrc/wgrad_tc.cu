// Convolution WEIGHT GRADIENT on the tcgen05 tensor cores (sm_100a), training path.
//
//   dW[co, ci, kh, kw] = sum over pixels of dy[co, pix] * x[ci, pix + (kh - p, kw - p)]        (stride 1, 'same', 1x1 / 3x3)
//
// is a GEMM whose contraction runs over PIXELS.  Both operands are C8 tensors ([N][C/8][H][W][8], the layout every
// tensor-core kernel of this repo uses): a pixel's 8-channel chunk is one 16-byte unit and consecutive pixels of a row
// are consecutive units, which is exactly the canonical no-swizzle **MN-major** UMMA operand layout
// (cute/atom/mma_traits_sm100.hpp: ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), T = 8 halves): 8 channels contiguous,
// 8 pixels at a 16-byte stride form one 128-byte core matrix, the next 8 pixels of the row are LBO = 128 bytes away and
// the next channel chunk SBO = one chunk plane away.  So a TMA box of the tensor is an MMA operand as it lands, and a
// filter tap is the same x tile at a different descriptor start address (+ (kh * XW + kw) * 16 bytes) -- no im2col.
//
// One CTA per SM walks pixel tiles (8 rows x 16 columns) with stride gridDim.x:
//   warp 0  TMA producer: dy tile (all Cout chunks) and haloed x tile (the CTA's block of NB input channels) into a
//           4-stage ring (mbarrier expect-tx);
//   warp 1  TMEM allocator + MMA issuer: per tile row one K = 16 step, per tap one
//           tcgen05.mma.cta_group::1.kind::f16  M = 128 (output channels, zero padded) x N = NB x K = 16,
//           accumulating ALL tiles of the CTA into KS*KS accumulators of NB TMEM columns each;
//   warps 2-5 epilogue: TMEM -> one fp32 partial dW per CTA in global memory.
// A second kernel sums the per-CTA partials in a fixed order (bit-reproducible).
#include "tc_common.cuh"

using namespace cwfa;
using namespace cwfa::tcx;

namespace {

constexpr int WT_BH = 8, WT_BW = 16, WT_STAGES = 4, WT_THREADS = 192;
constexpr int WT_A_CHUNK = WT_BH * WT_BW * 16;       // bytes of one 8-channel chunk of the dy tile
constexpr int WT_A_BYTES = 16 * WT_A_CHUNK;          // M = 128 rows = 16 chunks (chunks >= Cout/8 stay zero)

struct WtParams {
    int N, H, W, Cout, Cin, co_chunks /* per M block: min(16, Cout_p/8) */, nb_chunks, tiles_x, tiles_y, num_tiles;
    float* part;                                      // [gridDim.x][KS*KS][Cin][Cout]: lanes (= co) write consecutive floats
};

template <int KS>
struct WtGeom {
    static constexpr int XH = WT_BH + KS - 1, XW = WT_BW + KS - 1;
    static constexpr int B_CHUNK = XH * XW * 16;
    static __host__ __device__ int stage_bytes(int nb_chunks) { return (WT_A_BYTES + nb_chunks * B_CHUNK + 127) & ~127; }
};

template <int KS>
__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dy,
                                                                 const __grid_constant__ CUtensorMap tm_x, const WtParams p,
                                                                 const int is_bf16) {
    using G = WtGeom<KS>;
    constexpr int PAD = KS / 2, KK = KS * KS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t s0 = smem_u32(smem);
    auto full = [&](int s) { return s0 + 8u * s; };
    auto empty = [&](int s) { return s0 + 8u * (WT_STAGES + s); };
    const uint32_t acc_done = s0 + 8u * (2 * WT_STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
    const uint32_t ring = s0 + 1024;
    const int stage_bytes = G::stage_bytes(p.nb_chunks);
    const int NB = p.nb_chunks * 8;
    const int ncols = KK * NB <= 32 ? 32 : KK * NB <= 64 ? 64 : KK * NB <= 128 ? 128 : KK * NB <= 256 ? 256 : 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < WT_STAGES; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // rows of the M = 128 operand beyond the real output channels are zero in every stage (TMA never writes them)
    {
        const int zero_units = (16 - p.co_chunks) * (WT_A_CHUNK / 16);
        for (int s = 0; s < WT_STAGES; ++s) {
            uint4* dst = reinterpret_cast<uint4*>(smem + 1024 + (size_t)s * stage_bytes + (size_t)p.co_chunks * WT_A_CHUNK);
            for (int i = threadIdx.x; i < zero_units; i += WT_THREADS) dst[i] = make_uint4(0, 0, 0, 0);
        }
    }
    fence_proxy_async();
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int ci_chunk0 = blockIdx.y * p.nb_chunks;
    const int co_chunk0 = blockIdx.z * 16;               // M block of 128 output channels (TMA zero-fills chunks past the tensor)

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (lane == 0) {
            const uint32_t tx = (uint32_t)(p.co_chunks * WT_A_CHUNK + p.nb_chunks * G::B_CHUNK);
            for (int i = 0; i < my_tiles; ++i) {
                const int t = blockIdx.x + i * gridDim.x;
                const int n = t / tiles_per_img, r = t % tiles_per_img;
                const int h0 = (r / p.tiles_x) * WT_BH, w0 = (r % p.tiles_x) * WT_BW;
                const int s = i % WT_STAGES;
                mbar_wait(empty(s), ((i / WT_STAGES) & 1) ^ 1);
                mbar_expect_tx(full(s), tx);
                const uint32_t base = ring + (uint32_t)(s * stage_bytes);
                tma_load_4d(base, &tm_dy, full(s), w0 * 8, h0, co_chunk0, n);
                tma_load_4d(base + WT_A_BYTES, &tm_x, full(s), (w0 - PAD) * 8, h0 - PAD, ci_chunk0, n);
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        // kind::f16, fp32 accumulate, A and B both MN-major (bits 15, 16), M = 128, N = NB
        const uint32_t idesc = idesc_f16(NB, is_bf16) | (1u << 15) | (1u << 16);
        const uint32_t a_hi = desc_hi(WT_A_CHUNK), b_hi = desc_hi(G::B_CHUNK);
        for (int i = 0; i < my_tiles; ++i) {
            const int s = i % WT_STAGES;
            mbar_wait(full(s), (i / WT_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = ring + (uint32_t)(s * stage_bytes), b0 = a0 + WT_A_BYTES;
#pragma unroll 1
                for (int r = 0; r < WT_BH; ++r) {
                    const uint32_t a_lo = desc_lo(a0 + (uint32_t)(r * WT_BW * 16), 128);
                    const uint32_t acc = (i > 0 || r > 0) ? 1u : 0u;
#pragma unroll
                    for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                        for (int kw = 0; kw < KS; ++kw) {
                            const uint32_t b_lo = desc_lo(b0 + (uint32_t)(((r + kh) * G::XW + kw) * 16), 128);
                            tc_mma_f16_split(tmem + (uint32_t)((kh * KS + kw) * NB), a_lo, a_hi, b_lo, b_hi, idesc, acc);
                        }
                }
                tc_commit(empty(s));                     // the stage is free once these MMAs have read it
                if (i == my_tiles - 1) tc_commit(acc_done);
            }
            __syncwarp();
        }
    } else {
        // ============================ epilogue: TMEM -> per-CTA partial dW ============================
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int co = co_chunk0 * 8 + q * 32 + lane;
        mbar_wait(acc_done, 0);
        tc_fence_after();
        float* dst = p.part + (size_t)blockIdx.x * p.Cout * p.Cin * KK;   // [t][ci][co]
        for (int t = 0; t < KK; ++t)
            for (int c0 = 0; c0 < NB; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NB + c0), v);
                if (co < p.Cout) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int ci = ci_chunk0 * 8 + c0 + j;
                        if (ci < p.Cin) dst[((size_t)t * p.Cin + ci) * p.Cout + co] = __uint_as_float(v[j]);
                    }
                }
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, ncols);
    }
}

// dW[co][ci][t] = sum_k part[k][t][ci][co]  (fixed order, double accumulation; reads follow the partial layout).
// One block = 256 / S consecutive elements x S slices of the partial index: slice s sums the partials k = s, s + S, ... with four
// loads in flight (a warp load = 128 contiguous bytes), the slices are then added in a fixed order through shared memory.
// S = 8 when there are many partials (64-channel convs: 74 or 148): one thread per element walking all of them was latency-bound
// (18 us per 64 -> 64 3x3 gradient); S = 1 for the wide U-Net convs (1-4 partials of up to 38 MB: bandwidth-bound as they are).
__global__ void __launch_bounds__(256) wgrad_tc_finalize_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                                int Cout, int Cin, int KK, int nparts, int S) {
    __shared__ double red[256];
    const int64_t n = (int64_t)Cout * Cin * KK;
    const int E = 256 / S, e = threadIdx.x % E, slice = threadIdx.x / E;
    const int64_t i = blockIdx.x * (int64_t)E + e;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (i < n) {
        int k = slice;
        for (; k + 3 * S < nparts; k += 4 * S) {
            const float a = __ldg(part + (int64_t)k * n + i), b = __ldg(part + (int64_t)(k + S) * n + i);
            const float c = __ldg(part + (int64_t)(k + 2 * S) * n + i), d = __ldg(part + (int64_t)(k + 3 * S) * n + i);
            s0 += (double)a; s1 += (double)b; s2 += (double)c; s3 += (double)d;
        }
        for (; k < nparts; k += S) s0 += (double)__ldg(part + (int64_t)k * n + i);
    }
    double s = (s0 + s1) + (s2 + s3);
    if (S > 1) {                                   // block-uniform
        red[threadIdx.x] = s;
        __syncthreads();
        if (slice != 0) return;
        s = 0.0;
        for (int j = 0; j < S; ++j) s += red[j * E + e];
    }
    if (i >= n) return;
    const int co = (int)(i % Cout), ci = (int)((i / Cout) % Cin), t = (int)(i / ((int64_t)Cout * Cin));
    out[((int64_t)co * Cin + ci) * KK + t] = (float)s;
}

struct WtPlan { int nb_chunks, n_ci_blk, n_m_blk, chunks; };
static WtPlan wt_plan(int N, int H, int W, int Cin_p, int Cout_p, int KS) {
    WtPlan pl;
    const int max_nb = (KS == 3) ? 48 : 64;                 // KS*KS*NB <= 512 TMEM columns
    int nb = Cin_p < max_nb ? Cin_p : max_nb;
    while (Cin_p % nb) nb -= 16;
    pl.nb_chunks = nb / 8;
    pl.n_ci_blk = Cin_p / nb;
    pl.n_m_blk = ceil_div(Cout_p, 128);
    const int64_t tiles = (int64_t)N * ceil_div(H, WT_BH) * ceil_div(W, WT_BW);
    int chunks = kNumSMs / (pl.n_ci_blk * pl.n_m_blk);
    if (chunks < 1) chunks = 1;
    if (chunks > tiles) chunks = (int)tiles;
    pl.chunks = chunks;
    return pl;
}

}  // namespace

extern "C" int64_t cwfa_wgrad_tc_workspace_floats(int N, int H, int W, int Cin, int Cin_p, int Cout, int Cout_p, int KH) {
    if (KH != 1 && KH != 3) return -1;
    return (int64_t)wt_plan(N, H, W, Cin_p, Cout_p, KH).chunks * Cout * Cin * KH * KH;
}

extern "C" int cwfa_wgrad_tc(const void* x_c8, const void* dy_c8, float* dw, float* workspace, int N, int H, int W, int Cin,
                             int Cin_p, int Cout, int Cout_p, int KH, int KW, int is_bf16, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || KH != KW || (KH != 1 && KH != 3) || (Cin_p % 16) || (Cout_p % 16) ||
        Cin_p < Cin || Cout_p < Cout || !x_c8 || !dy_c8 || !dw || !workspace) {
        set_error("wgrad_tc: unsupported configuration (Cin_p=%d Cout_p=%d K=%dx%d; needs 1x1 or 3x3)", Cin_p, Cout_p, KH, KW);
        return CWFA_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(x_c8) & 15) || (reinterpret_cast<uintptr_t>(dy_c8) & 15)) {
        set_error("wgrad_tc: pointers must be 16-byte aligned");
        return CWFA_EINVAL;
    }
    const WtPlan pl = wt_plan(N, H, W, Cin_p, Cout_p, KH);
    WtParams p{};
    p.N = N; p.H = H; p.W = W; p.Cout = Cout; p.Cin = Cin;
    p.co_chunks = Cout_p / 8 < 16 ? Cout_p / 8 : 16; p.nb_chunks = pl.nb_chunks;
    p.tiles_x = ceil_div(W, WT_BW); p.tiles_y = ceil_div(H, WT_BH);
    const int64_t nt = (int64_t)p.tiles_x * p.tiles_y * N;
    if (nt > 0x7fffffff) { set_error("wgrad_tc: too many tiles"); return CWFA_EINVAL; }
    p.num_tiles = (int)nt;
    p.part = workspace;
    CUtensorMap tm_dy, tm_x;
    int rc = make_c8_tensor_map(&tm_dy, dy_c8, N, Cout_p / 8, H, W, WT_BW, WT_BH, p.co_chunks, is_bf16);
    if (rc) return rc;
    rc = make_c8_tensor_map(&tm_x, x_c8, N, Cin_p / 8, H, W, WT_BW + KH - 1, WT_BH + KH - 1, pl.nb_chunks, is_bf16);
    if (rc) return rc;
    dim3 grid(pl.chunks, pl.n_ci_blk, pl.n_m_blk);
    if (KH == 3) {
        const size_t smem = 2048 + (size_t)WT_STAGES * WtGeom<3>::stage_bytes(pl.nb_chunks);
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = true; }
        wgrad_tc_kernel<3><<<grid, WT_THREADS, smem, st>>>(tm_dy, tm_x, p, is_bf16);
    } else {
        const size_t smem = 2048 + (size_t)WT_STAGES * WtGeom<1>::stage_bytes(pl.nb_chunks);
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = true; }
        wgrad_tc_kernel<1><<<grid, WT_THREADS, smem, st>>>(tm_dy, tm_x, p, is_bf16);
    }
    rc = check_launch("wgrad_tc");
    if (rc) return rc;
    const int64_t n = (int64_t)Cout * Cin * KH * KW;
    const int slices = pl.chunks >= 32 ? 8 : (pl.chunks >= 8 ? 4 : 1);
    wgrad_tc_finalize_kernel<<<(unsigned)ceil_div(n, 256 / slices), 256, 0, st>>>(workspace, dw, Cout, Cin, KH * KW, pl.chunks, slices);
    return check_launch("wgrad_tc_finalize");
}
