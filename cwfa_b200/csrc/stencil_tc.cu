// Conditioning-net depth stencil Conv3d(1 -> 32, 3x3x3) -> PReLU -> Conv3d(32 -> 1, 3x3x3) over the (H, W, depth) volume
// (reference networks.py:221-225 and :239) as ONE fused tcgen05 kernel in its true 3-D form: GEMM rows are VOXELS, so the
// 32-channel hidden volume (805 MB per frame at level 0 in the banded 2-D form of conv_tc.cu) never leaves the SM.
//
// Per CTA: a strip of TW x TH output pixels x all D depths, walked one pixel-row of the hidden volume at a time.  Rows of a
// pixel-row are r = px * Dq + dq for px in [0, TW + 2) (one halo pixel each side), dq in [0, Dq), Dq = D + 1 or D + 2 (even);
// dq >= D are ZERO rows that serve as the depth padding of their own pixel (above) and of the next pixel (below).
//   1. im2col of the SPATIAL taps only: A1[r][t] = x[(y + kh - 1), (x + kw - 1), d], t = kh * 3 + kw < 9; columns 9, 10 are 1.0
//      and multiply the [hi | lo] split of the first conv's bias (fp32-accurate bias through the MMA); K = 16.
//   2. GEMM1 (SS, M = 128, N = 32 hidden channels): the DEPTH taps are an ADDRESS SHIFT of the A operand -- row r +- 1 is the
//      depth neighbour, so hidden[r] = sum_kd A1[r + kd - 1] W1[kd]^T is three MMAs whose A start differs by 16 bytes.
//      Epilogue: PReLU, zero outside the volume, pack, store h[r][32] over the A1 buffer (K-major chunks of 8 channels).
//   3. GEMM2, same trick: Q[r][kh, kw] = sum_kd sum_c w2[c, kh, kw, kd] h[c][r + kd - 1]  (6 MMAs of N = 16 per 128 rows;
//      the spatial taps are the N dimension).
//   4. Q (fp16) goes to a 3-pixel-row ring in shared memory; an output row is the 9-term gather
//      out[y, x, d] = b2 + sum_{kh, kw} Q[(y + kh - 1), (x + kw - 1), d][kh, kw]  (the col2im of the scatter form), written as C8;
//      the gather of row y - 2 runs while the tensor pipe works on GEMM1 of pixel-row y.
// Phases are separated by block barriers; two CTAs per SM overlap one CTA's tensor phases with the other's epilogues.
#include "tc_common.cuh"
#include "cwfa_b200_debug.h"

namespace {
using namespace cwfa;
using namespace cwfa::tcx;

constexpr int kThreads = 256;
constexpr int kRowsMax = 768;                  // GEMM rows per pixel-row of a strip (6 M tiles of 128)
constexpr uint32_t kOffW = 256;                // W1[kd] [2][32][8] (1 KB each) then W2[kd] [4][16][8] (1 KB each)
constexpr uint32_t kWBytes = 3 * 1024 + 3 * 1024;
constexpr uint32_t kOffAH = kOffW + kWBytes;

struct StParams {
    const uint4* x;
    uint4* y;
    const uint4* w;
    const float* b2;
    const float* slope;
    int N, H, W, D, cin_chunks, cout_chunks;
    int TW, TH, strips;
    int rows_cap, xs_halves;                    // smem geometry: rows per Q plane / operand plane, halves per x-ring slot
    uint32_t off_q, off_xs, off_stage, zero_bytes, tmem_cols;
    unsigned long long* prof;      // debug: 8 phase-cycle sums of CTA (0, 0) (NULL = off)
};

template <bool BF16, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) stencil3d_tc_kernel(const StParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int strip = blockIdx.x % p.strips, seg = blockIdx.x / p.strips, n = blockIdx.y;
    const int x0 = strip * p.TW, y0 = seg * p.TH;
    const int tw = min(p.TW, p.W - x0), th = min(p.TH, p.H - y0);
    if (tw <= 0 || th <= 0) return;
    const int D = p.D, Dq = (D + 2) & ~1, XW = tw + 4;           // Dq even: 1 or 2 zero rows after the D depths of a pixel
    const int npx = tw + 2, rows = npx * Dq, n_mt = (rows + 127) >> 7;
    const int RA = p.rows_cap + 8, QP = p.rows_cap;
    const uint32_t div_magic = 65536u / (uint32_t)Dq + 1u;          // r / Dq exact for r * Dq < 65536

    const uint32_t s0 = smem_u32(smem);
    const uint32_t bar = s0 + 8, sW = s0 + kOffW, sAH = s0 + kOffAH;
    unsigned short* xs = reinterpret_cast<unsigned short*>(smem + p.off_xs);
    __half* q = reinterpret_cast<__half*>(smem + p.off_q);
    unsigned short* stage = reinterpret_cast<unsigned short*>(smem + p.off_stage);

    // ---- set-up: weights, zeroed operand buffer / Q ring / x ring, barrier, tensor memory
    for (int i = tid; i < (int)(kWBytes / 16); i += kThreads) reinterpret_cast<uint4*>(smem + kOffW)[i] = p.w[i];
    for (int i = tid; i < (int)(p.zero_bytes / 16); i += kThreads) reinterpret_cast<uint4*>(smem + kOffAH)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) tmem_alloc(s0, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem);
    const float slope = p.slope[0], b2 = p.b2[0];
    const bool slope_le1 = slope <= 1.f;
    const uint32_t id32 = idesc_f16(32, BF16 ? 1 : 0), id16 = idesc_f16(16, BF16 ? 1 : 0);
    const uint32_t ah_lbo = (uint32_t)RA * 16, hi128 = desc_hi(128);
    const uint32_t ones = BF16 ? 0x3F803F80u : 0x3C003C00u;
    uint32_t phase = 0;
    long long t_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_last = clock64();
    auto stamp = [&](int k) {
        if (p.prof) { const long long t = clock64(); t_ph[k] += t - t_last; t_last = t; }
    };

    // x pixel-row of the input volume <-> ring slot xs[slot][px][d], px <-> gx = x0 - 2 + px; zeros outside the image and for
    // d in [D, Dq).  One 16-byte chunk (8 depths of one pixel) per thread: fetched early, stored to the ring later.
    const int nch = (D + 7) >> 3;
    const bool x_thread = tid < XW * nch;
    const int x_ch = tid / XW, x_px = tid - x_ch * XW;
    auto fetch_xrow = [&](int xr) -> uint4 {
        const int yy = y0 - 2 + xr, gx = x0 - 2 + x_px;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (x_thread && yy >= 0 && yy < p.H && gx >= 0 && gx < p.W)
            v = __ldg(p.x + ((size_t)(n * p.cin_chunks + x_ch) * p.H + yy) * p.W + gx);
        return v;
    };
    auto store_xrow = [&](int xr, const uint4& v) {
        if (!x_thread) return;
        unsigned short* o = xs + (xr & 3) * p.xs_halves + x_px * Dq + x_ch * 8;
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            if (x_ch * 8 + j + 1 < D) *reinterpret_cast<uint32_t*>(o + j) = u[j >> 1];       // Dq even, chunk start even: 4-byte aligned
            else if (x_ch * 8 + j < D) o[j] = (unsigned short)u[j >> 1];
        }
    };
    store_xrow(0, fetch_xrow(0));
    store_xrow(1, fetch_xrow(1));
    store_xrow(2, fetch_xrow(2));
    __syncthreads();

    for (int py = -1; py <= th + 1; ++py) {
        const bool cur = py <= th;                                  // the last iteration only flushes the final output row
        uint4 x_next = make_uint4(0, 0, 0, 0);
        if (cur) x_next = fetch_xrow(py + 4);                       // lands in the ring while GEMM2 runs
        // ---- 1. im2col of the nine spatial taps (+ the two bias ones) for every row of this pixel-row (chunk planes 0, 1)
        if (cur) {
            const unsigned short* x_m = xs + ((py + 1) & 3) * p.xs_halves;
            const unsigned short* x_c = xs + ((py + 2) & 3) * p.xs_halves;
            const unsigned short* x_p = xs + ((py + 3) & 3) * p.xs_halves;
            for (int r = tid; r < rows; r += kThreads) {
                const int o = r;                                    // (px + kw) * Dq + dq = r + kw * Dq
                uint32_t v[9];
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    v[0 * 3 + kw] = x_m[o + kw * Dq];
                    v[1 * 3 + kw] = x_c[o + kw * Dq];
                    v[2 * 3 + kw] = x_p[o + kw * Dq];
                }
                sts128(sAH + (uint32_t)(1 + r) * 16, make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16)));
                sts128(sAH + (uint32_t)(RA + 1 + r) * 16, make_uint4(v[8] | (ones << 16), ones & 0xFFFFu, 0u, 0u));
            }
        }
        fence_proxy_async();
        __syncthreads();
        stamp(1);
        // ---- 2. GEMM1: hidden[r] = sum_kd A1[r + kd - 1] * W1[kd]^T  (N = 32, K = 16 per depth tap)
        if (cur && tid == 0) {
            tc_fence_after();
            for (int m = 0; m < n_mt; ++m)
#pragma unroll
                for (int kd = 0; kd < 3; ++kd)
                    tc_mma_f16_split(tm + 32 * m, desc_lo(sAH + (uint32_t)(128 * m + kd) * 16, ah_lbo), hi128,
                                     desc_lo(sW + kd * 1024, 512), hi128, id32, kd);
            tc_commit(bar);
        }
        // ---- output row y0 + py - 2: gather of the nine shifted partials (two depths per thread) -- its pixel-rows py - 3 .. py - 1
        // are complete, so it runs UNDER GEMM1 of pixel-row py instead of after epilogue 2
        const int yo = py - 2;
        if (yo >= 0 && yo < th) {
            const int Dp = p.cout_chunks * 8, Dh = Dp >> 1;
            const uint32_t dh_magic = 65536u / (uint32_t)Dh + 1u;
            const __half* q_m = q + ((yo + 0) % 3) * 9 * QP;       // pixel-row yo - 1 -> slot (yo - 1 + 1) % 3
            const __half* q_c = q + ((yo + 1) % 3) * 9 * QP;
            const __half* q_p = q + ((yo + 2) % 3) * 9 * QP;
            for (int i = tid; i < tw * Dh; i += kThreads) {
                const int px = (int)(((uint32_t)i * dh_magic) >> 16), d = 2 * (i - px * Dh);
                float s0v = 0.f, s1v = 0.f;
                if (d < D) {
                    const int o = px * Dq + d;
                    s0v = b2;
                    s1v = b2;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(q_m + (0 * 3 + kw) * QP + o + kw * Dq));
                        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(q_c + (1 * 3 + kw) * QP + o + kw * Dq));
                        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(q_p + (2 * 3 + kw) * QP + o + kw * Dq));
                        s0v += a.x + b.x + c.x;
                        s1v += a.y + b.y + c.y;
                    }
                    if (d + 1 >= D) s1v = 0.f;
                }
                reinterpret_cast<uint32_t*>(stage)[i] = pack2<BF16>(s0v, s1v);
            }
            __syncthreads();
            for (int i = tid; i < tw * p.cout_chunks; i += kThreads) {
                const int ch = i / tw, px = i - ch * tw;
                p.y[((size_t)(n * p.cout_chunks + ch) * p.H + y0 + yo) * p.W + x0 + px] =
                    *reinterpret_cast<const uint4*>(stage + px * Dp + ch * 8);
            }
        }
        stamp(6);
        if (!cur) break;
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        stamp(2);
        // ---- epilogue 1: PReLU, zero outside the volume, pack, h over the A1 rows
        {
            const int gy = y0 + py;
            const bool row_in = gy >= 0 && gy < p.H;
            const int qd = warp & 3;
            for (int m = warp >> 2; m < n_mt; m += kThreads / 128) {
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
                    sts128(sAH + (uint32_t)(c * RA + 1 + r) * 16, make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]));
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        stamp(3);
        // ---- 3. GEMM2: Q[r][kh, kw] = sum_kd h[r + kd - 1] * W2[kd]^T  (N = 16, K = 32 per depth tap)
        if (tid == 0) {
            tc_fence_after();
            for (int m = 0; m < n_mt; ++m)
#pragma unroll
                for (int kd = 0; kd < 3; ++kd)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        tc_mma_f16_split(tm + 32 * m, desc_lo(sAH + (uint32_t)(2 * ks * RA + 128 * m + kd) * 16, ah_lbo), hi128,
                                         desc_lo(sW + 3072 + kd * 1024 + ks * 2 * 256, 256), hi128, id16, kd | ks);
            tc_commit(bar);
        }
        store_xrow(py + 4, x_next);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        stamp(4);
        // ---- epilogue 2: Q of this pixel-row into the ring (fp16, one plane per spatial tap)
        {
            __half* qs = q + ((py + 1) % 3) * 9 * QP;
            const int qd = warp & 3;
            for (int m = warp >> 2; m < n_mt; m += kThreads / 128) {
                if (128 * m + 32 * qd >= rows) continue;
                const int r = 128 * m + 32 * qd + lane;
                uint32_t a[16];
                tmem_ld16(tm + ((uint32_t)(32 * qd) << 16) + 32 * m, a);
                if (r < rows) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) qs[t * QP + r] = __float2half_rn(__uint_as_float(a[t]));
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        stamp(5);
    }
    if (p.prof && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        for (int k = 0; k < 7; ++k) p.prof[k] = (unsigned long long)t_ph[k];
        p.prof[7] = (unsigned long long)(th + 2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, p.tmem_cols);
}

}  // namespace

static unsigned long long* g_st_prof = nullptr;
extern "C" int cwfa_stencil_set_debug_buffer(void* buf) { g_st_prof = (unsigned long long*)buf; return CWFA_OK; }

extern "C" int cwfa_stencil3d_tc(const void* x, void* y, const void* wpack, const float* b2, const float* slope, int N, int H,
                                 int W, int D, int cin_chunks, int cout_chunks, int rows_max, int is_bf16, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !y || !wpack || !b2 || !slope || N <= 0 || H <= 0 || W <= 0 || D <= 0 || D > 64 || cin_chunks * 8 < D ||
        cout_chunks * 8 < D || (cout_chunks & 1) || N > 65535) {
        set_error("stencil3d_tc: bad arguments (1 <= D <= 64, chunk counts must cover D, even number of output chunks)");
        return CWFA_EINVAL;
    }
    // strip geometry: (TW + 2) * Dq rows per pixel-row <= rows_cap (a multiple of 128)
    const int Dq = (D + 2) & ~1;
    int rows_cap = rows_max <= 0 ? kRowsMax : (rows_max + 127) / 128 * 128;
    if (rows_cap > kRowsMax) rows_cap = kRowsMax;
    int tw_max = rows_cap / Dq - 2;
    if (tw_max > 128) tw_max = 128;
    while (tw_max >= 1 && (tw_max + 4) * ((D + 7) / 8) > kThreads) --tw_max;       // one 16-byte x chunk per thread and row
    if (tw_max < 1) {
        set_error("stencil3d_tc: no strip geometry for D = %d with %d rows", D, rows_cap);
        return CWFA_EINVAL;
    }
    StParams p;
    p.strips = ceil_div(W, tw_max);
    p.TW = ceil_div(W, p.strips);
    p.rows_cap = rows_cap;
    p.xs_halves = ((p.TW + 4) * Dq + 7) / 8 * 8;
    const uint32_t ah_bytes = 4u * (rows_cap + 8) * 16, q_bytes = 27u * rows_cap * 2, xs_bytes = 4u * p.xs_halves * 2;
    const uint32_t stage_bytes = ((uint32_t)p.TW * cout_chunks * 16 + 15) / 16 * 16;
    p.off_q = kOffAH + ah_bytes;
    p.off_xs = p.off_q + q_bytes;
    p.off_stage = p.off_xs + xs_bytes;
    p.zero_bytes = ah_bytes + q_bytes + xs_bytes;
    const uint32_t smem_bytes = p.off_stage + stage_bytes;
    const int n_mt = rows_cap / 128;
    p.tmem_cols = n_mt * 32 <= 32 ? 32 : n_mt * 32 <= 64 ? 64 : n_mt * 32 <= 128 ? 128 : 256;
    const int minb = 2;
    int ys = (minb * kNumSMs) / p.strips;          // one unit per resident CTA
    if (ys < 1) ys = 1;
    if (ys > H) ys = H;
    p.TH = ceil_div(H, ys);
    const int ysegs = ceil_div(H, p.TH);
    p.x = (const uint4*)x; p.y = (uint4*)y; p.w = (const uint4*)wpack; p.b2 = b2; p.slope = slope;
    p.prof = g_st_prof;
    p.N = N; p.H = H; p.W = W; p.D = D; p.cin_chunks = cin_chunks; p.cout_chunks = cout_chunks;
    void (*kern)(const StParams) = is_bf16 ? stencil3d_tc_kernel<true, 2> : stencil3d_tc_kernel<false, 2>;
    static bool attr_done[2] = {false, false};
    const int ki = is_bf16 ? 1 : 0;
    if (!attr_done[ki]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess)
            return check_launch("stencil3d_tc (attributes)");
        attr_done[ki] = true;
    }
    if (smem_bytes > 113 * 1024) {
        set_error("stencil3d_tc: %u bytes of shared memory", smem_bytes);
        return CWFA_EINVAL;
    }
    kern<<<dim3(p.strips * ysegs, N), kThreads, smem_bytes, st>>>(p);
    return check_launch("stencil3d_tc");
}
