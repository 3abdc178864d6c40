// Conditioning-net depth stencil Conv3d(1 -> 32, 3x3x3) -> PReLU -> Conv3d(32 -> 1, 3x3x3) over the (H, W, depth) volume
// (reference networks.py:221-225 and :239) as ONE fused tcgen05 kernel in its true 3-D form: GEMM rows are VOXELS, so the
// 32-channel hidden volume (805 MB per frame at level 0 in the banded 2-D form of conv_tc.cu) never leaves the SM.
//
// Per CTA: a strip of TW x TH output pixels x all D depths, walked one pixel-row of the hidden volume at a time.  Rows of a
// pixel-row are r = px * (D + 1) + dq for px in [0, TW + 2) (one halo pixel each side), dq in [0, D]; dq = D is a ZERO row
// that serves as the depth padding of its own pixel (above) and of the next pixel (below).
//   1. im2col: A1[r][k] = x[(y + kh - 1), (x + kw - 1), (d + kd - 1)], k = (kh * 3 + kw) * 3 + kd < 27; k = 27, 28 are 1.0 and
//      multiply the [hi | lo] bf16 split of the first conv's bias (fp32-accurate bias through the MMA); K = 32.
//   2. GEMM1 (SS, M = 128, N = 32 hidden channels, 2 K-steps) -> TMEM; epilogue: PReLU, zero outside the volume, pack,
//      store h[r][32] over the A1 buffer (K-major chunks of 8 channels, 16 B per row and chunk).
//   3. GEMM2: depth taps by ADDRESS SHIFT of the A operand (row r +- 1 is the depth neighbour: start address +- 16 B), spatial
//      taps as N: Q[r][kh, kw] = sum_kd sum_c w2[c, kh, kw, kd] h[c][r + kd - 1]   (6 MMAs of N = 16 per 128 rows).
//   4. Q (fp16) goes to a 3-pixel-row ring in shared memory; the output row y - 1 is the 9-term gather
//      out[y, x, d] = b2 + sum_{kh, kw} Q[(y + kh - 1), (x + kw - 1), d][kh, kw]  (the col2im of the scatter form), written as C8.
// Phases are separated by block barriers; two CTAs per SM overlap one CTA's tensor phases with the other's epilogues.
#include "tc_common.cuh"

namespace {
using namespace cwfa;
using namespace cwfa::tcx;

constexpr int kThreads = 256;
constexpr int kMaxMT = 6;                      // 128-row M tiles per pixel-row
constexpr int kRowsMax = kMaxMT * 128;         // 768
constexpr int kRA = kRowsMax + 8;              // allocated rows per chunk plane (1 leading zero row + slack for the +1 shift)
constexpr int kXsHalves = 1280;                // one x pixel-row in shared memory: (TW + 4) x (D + 3) halves
constexpr uint32_t kOffW = 256;                // W1 [4][32][8] (2 KB) then W2[kd] [4][16][8] (1 KB each)
constexpr uint32_t kWBytes = 2048 + 3 * 1024;
constexpr uint32_t kOffAH = kOffW + kWBytes;
constexpr uint32_t kAHBytes = 4 * kRA * 16;
constexpr uint32_t kOffQ = kOffAH + kAHBytes;
constexpr uint32_t kQBytes = 3 * 9 * kRowsMax * 2;
constexpr uint32_t kOffXs = kOffQ + kQBytes;
constexpr uint32_t kXsBytes = 4 * kXsHalves * 2;
constexpr uint32_t kOffStage = kOffXs + kXsBytes;
constexpr uint32_t kStageBytes = 4096;
constexpr uint32_t kSmemBytes = kOffStage + kStageBytes;
constexpr uint32_t kTmemCols = 256;

struct StParams {
    const uint4* x;
    uint4* y;
    const uint4* w;
    const float* b2;
    const float* slope;
    int N, H, W, D, cin_chunks, cout_chunks;
    int TW, TH, strips;
};

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 2) stencil3d_tc_kernel(const StParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int strip = blockIdx.x % p.strips, seg = blockIdx.x / p.strips, n = blockIdx.y;
    const int x0 = strip * p.TW, y0 = seg * p.TH;
    const int tw = min(p.TW, p.W - x0), th = min(p.TH, p.H - y0);
    if (tw <= 0 || th <= 0) return;
    const int D = p.D, Dq = D + 1, XD = D + 3, XW = tw + 4;
    const int npx = tw + 2, rows = npx * Dq, n_mt = (rows + 127) >> 7;
    const uint32_t div_magic = 65536u / (uint32_t)Dq + 1u;          // r / Dq exact for r * Dq < 65536

    const uint32_t s0 = smem_u32(smem);
    const uint32_t bar = s0 + 8, sW = s0 + kOffW, sAH = s0 + kOffAH;
    unsigned short* xs = reinterpret_cast<unsigned short*>(smem + kOffXs);
    __half* q = reinterpret_cast<__half*>(smem + kOffQ);
    unsigned short* stage = reinterpret_cast<unsigned short*>(smem + kOffStage);

    // ---- set-up: weights, zeroed operand buffer / x ring / Q ring, barrier, tensor memory
    for (int i = tid; i < (int)(kWBytes / 16); i += kThreads) reinterpret_cast<uint4*>(smem + kOffW)[i] = p.w[i];
    for (int i = tid; i < (int)((kAHBytes + kQBytes + kXsBytes) / 16); i += kThreads)
        reinterpret_cast<uint4*>(smem + kOffAH)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) tmem_alloc(s0, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem);
    const float slope = p.slope[0], b2 = p.b2[0];
    const bool slope_le1 = slope <= 1.f;
    const uint32_t id32 = idesc_f16(32, BF16 ? 1 : 0), id16 = idesc_f16(16, BF16 ? 1 : 0);
    const uint32_t ah_lbo = kRA * 16, hi128 = desc_hi(128);
    const unsigned short one = BF16 ? 0x3F80 : 0x3C00;
    uint32_t phase = 0;

    // x pixel-row yy of the input volume -> ring slot: xs[slot][px][1 + d], px <-> gx = x0 - 2 + px; zeros outside the image
    auto load_xrow = [&](int xr) {
        const int yy = y0 - 2 + xr;
        unsigned short* dst = xs + (xr & 3) * kXsHalves;
        const bool row_ok = yy >= 0 && yy < p.H;
        const int nch = (D + 7) >> 3;
        for (int i = tid; i < XW * nch; i += kThreads) {
            const int ch = i / XW, px = i - ch * XW, gx = x0 - 2 + px;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (row_ok && gx >= 0 && gx < p.W) v = __ldg(p.x + ((size_t)(n * p.cin_chunks + ch) * p.H + yy) * p.W + gx);
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
            unsigned short* o = dst + px * XD + 1 + ch * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (ch * 8 + j < D) o[j] = (unsigned short)(u[j >> 1] >> ((j & 1) * 16));
        }
    };
    load_xrow(0);
    load_xrow(1);

    for (int py = -1; py <= th; ++py) {
        load_xrow(py + 3);
        __syncthreads();
        // ---- 1. im2col of the 27 input taps (+ the two bias ones) for every row of this pixel-row
        {
            const unsigned short* x_m = xs + ((py + 1) & 3) * kXsHalves;
            const unsigned short* x_c = xs + ((py + 2) & 3) * kXsHalves;
            const unsigned short* x_p = xs + ((py + 3) & 3) * kXsHalves;
            for (int r = tid; r < rows; r += kThreads) {
                const int px = (int)(((uint32_t)r * div_magic) >> 16), dq = r - px * Dq;
                const int o = px * XD + dq;
                unsigned short v[32];
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                        v[(0 * 3 + kw) * 3 + kd] = x_m[o + kw * XD + kd];
                        v[(1 * 3 + kw) * 3 + kd] = x_c[o + kw * XD + kd];
                        v[(2 * 3 + kw) * 3 + kd] = x_p[o + kw * XD + kd];
                    }
                v[27] = one; v[28] = one; v[29] = 0; v[30] = 0; v[31] = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 w4;
                    w4.x = v[8 * c + 0] | ((uint32_t)v[8 * c + 1] << 16);
                    w4.y = v[8 * c + 2] | ((uint32_t)v[8 * c + 3] << 16);
                    w4.z = v[8 * c + 4] | ((uint32_t)v[8 * c + 5] << 16);
                    w4.w = v[8 * c + 6] | ((uint32_t)v[8 * c + 7] << 16);
                    sts128(sAH + (uint32_t)(c * kRA + 1 + r) * 16, w4);
                }
            }
        }
        fence_proxy_async();
        __syncthreads();
        // ---- 2. GEMM1: hidden = A1 * W1^T  (N = 32, K = 32)
        if (tid == 0) {
            tc_fence_after();
            for (int m = 0; m < n_mt; ++m)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    tc_mma_f16_split(tm + 32 * m, desc_lo(sAH + (uint32_t)(2 * ks * kRA + 1 + 128 * m) * 16, ah_lbo), hi128,
                                     desc_lo(sW + ks * 2 * 512, 512), hi128, id32, ks);
            tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue 1: PReLU, zero outside the volume, pack, h over the A1 rows
        {
            const int gy = y0 + py;
            const bool row_in = gy >= 0 && gy < p.H;
            const int qd = warp & 3;
            for (int m = warp >> 2; m < n_mt; m += 2) {
                if (128 * m + 32 * qd >= rows) continue;
                const int r = 128 * m + 32 * qd + lane;
                const int px = (int)(((uint32_t)r * div_magic) >> 16), dq = r - px * Dq, gx = x0 - 1 + px;
                const bool valid = row_in && r < rows && dq < D && gx >= 0 && gx < p.W;
                uint32_t a[32];
                tmem_ld32_nowait(tm + ((uint32_t)(32 * qd) << 16) + 32 * m, a);
                tmem_ld_wait();
                uint32_t h[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float u0 = __uint_as_float(a[2 * j]), u1 = __uint_as_float(a[2 * j + 1]);
                    u0 = slope_le1 ? fmaxf(u0, slope * u0) : fminf(u0, slope * u0);
                    u1 = slope_le1 ? fmaxf(u1, slope * u1) : fminf(u1, slope * u1);
                    h[j] = valid ? pack2<BF16>(u0, u1) : 0u;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    sts128(sAH + (uint32_t)(c * kRA + 1 + r) * 16, make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]));
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- 3. GEMM2: Q[r][kh, kw] = sum_kd h[r + kd - 1] * W2[kd]^T  (N = 16, K = 32 per depth tap)
        if (tid == 0) {
            tc_fence_after();
            for (int m = 0; m < n_mt; ++m)
#pragma unroll
                for (int kd = 0; kd < 3; ++kd)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        tc_mma_f16_split(tm + 32 * m, desc_lo(sAH + (uint32_t)(2 * ks * kRA + 128 * m + kd) * 16, ah_lbo), hi128,
                                         desc_lo(sW + 2048 + kd * 1024 + ks * 2 * 256, 256), hi128, id16, kd | ks);
            tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue 2: Q of this pixel-row into the ring (fp16, one plane per spatial tap)
        {
            __half* qs = q + ((py + 1) % 3) * 9 * kRowsMax;
            const int qd = warp & 3;
            for (int m = warp >> 2; m < n_mt; m += 2) {
                if (128 * m + 32 * qd >= rows) continue;
                const int r = 128 * m + 32 * qd + lane;
                uint32_t a[16];
                tmem_ld16(tm + ((uint32_t)(32 * qd) << 16) + 32 * m, a);
                if (r < rows) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) qs[t * kRowsMax + r] = __float2half_rn(__uint_as_float(a[t]));
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- 4. output row y0 + py - 1: gather of the nine shifted partials
        const int yo = py - 1;
        if (yo >= 0 && yo < th) {
            const int Dp = p.cout_chunks * 8;
            const __half* q_m = q + ((yo + 0) % 3) * 9 * kRowsMax;       // pixel-row yo - 1 -> slot (yo - 1 + 1) % 3
            const __half* q_c = q + ((yo + 1) % 3) * 9 * kRowsMax;
            const __half* q_p = q + ((yo + 2) % 3) * 9 * kRowsMax;
            for (int i = tid; i < tw * Dp; i += kThreads) {
                const int px = i / Dp, d = i - px * Dp;
                float s = 0.f;
                if (d < D) {
                    const int o = px * Dq + d;
                    s = b2;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        s += __half2float(q_m[(0 * 3 + kw) * kRowsMax + o + kw * Dq]);
                        s += __half2float(q_c[(1 * 3 + kw) * kRowsMax + o + kw * Dq]);
                        s += __half2float(q_p[(2 * 3 + kw) * kRowsMax + o + kw * Dq]);
                    }
                }
                if constexpr (BF16) {
                    const __nv_bfloat16 b = __float2bfloat16_rn(s);
                    stage[i] = *reinterpret_cast<const unsigned short*>(&b);
                } else {
                    const __half b = __float2half_rn(s);
                    stage[i] = *reinterpret_cast<const unsigned short*>(&b);
                }
            }
            __syncthreads();
            for (int i = tid; i < tw * p.cout_chunks; i += kThreads) {
                const int ch = i / tw, px = i - ch * tw;
                p.y[((size_t)(n * p.cout_chunks + ch) * p.H + y0 + yo) * p.W + x0 + px] =
                    *reinterpret_cast<const uint4*>(stage + px * Dp + ch * 8);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, kTmemCols);
}

}  // namespace

// Strip geometry for a volume of D depths: pixels per strip row so that (TW + 2) * (D + 1) <= rows_max.
static int stencil_geometry(int H, int W, int D, int rows_max, int* TW, int* TH, int* strips, int* ysegs) {
    if (rows_max <= 0 || rows_max > kRowsMax) rows_max = kRowsMax;
    int tw_max = rows_max / (D + 1) - 2;
    if (tw_max > 128) tw_max = 128;
    if (tw_max < 1) return 0;
    *strips = ceil_div(W, tw_max);
    *TW = ceil_div(W, *strips);
    int ys = (2 * kNumSMs) / *strips;              // one unit per resident CTA (2 per SM)
    if (ys < 1) ys = 1;
    if (ys > H) ys = H;
    *TH = ceil_div(H, ys);
    *ysegs = ceil_div(H, *TH);
    return 1;
}

extern "C" int cwfa_stencil3d_tc(const void* x, void* y, const void* wpack, const float* b2, const float* slope, int N, int H,
                                 int W, int D, int cin_chunks, int cout_chunks, int rows_max, int is_bf16, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !y || !wpack || !b2 || !slope || N <= 0 || H <= 0 || W <= 0 || D <= 0 || D > 64 || cin_chunks * 8 < D ||
        cout_chunks * 8 < D || N > 65535) {
        set_error("stencil3d_tc: bad arguments (1 <= D <= 64, chunk counts must cover D)");
        return CWFA_EINVAL;
    }
    StParams p;
    int ysegs = 0;
    if (!stencil_geometry(H, W, D, rows_max, &p.TW, &p.TH, &p.strips, &ysegs) || (p.TW + 4) * (D + 3) > kXsHalves ||
        p.TW * cout_chunks * 8 * 2 > (int)kStageBytes) {
        set_error("stencil3d_tc: no strip geometry for D = %d", D);
        return CWFA_EINVAL;
    }
    p.x = (const uint4*)x; p.y = (uint4*)y; p.w = (const uint4*)wpack; p.b2 = b2; p.slope = slope;
    p.N = N; p.H = H; p.W = W; p.D = D; p.cin_chunks = cin_chunks; p.cout_chunks = cout_chunks;
    static bool attr_done[2] = {false, false};
    auto kern = is_bf16 ? stencil3d_tc_kernel<true> : stencil3d_tc_kernel<false>;
    if (!attr_done[is_bf16 ? 1 : 0]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess)
            return check_launch("stencil3d_tc (attributes)");
        attr_done[is_bf16 ? 1 : 0] = true;
    }
    kern<<<dim3(p.strips * ysegs, N), kThreads, kSmemBytes, st>>>(p);
    return check_launch("stencil3d_tc");
}
