// Reference-precision (fp32, CUDA-core) convolution kernels in the reference's NCHW layout.
// These serve (a) every small-channel convolution of the path (6-channel ConvNeXt / attention
// layers, 29->6 input projection) and (b) the fp32 mode of the module API, which the parity
// tests use with a tight tolerance.  The throughput path for the 64..1024-channel
// convolutions is the tcgen05 implicit-GEMM kernel in conv_tc.cu.
#include "common.cuh"
using namespace cwfa;

// ------------------------------------------------------------------------------------------
// Direct conv2d: 16x16 output pixels per block, CO_T output channels per block,
// input channels staged through shared memory CI_T at a time.
// ------------------------------------------------------------------------------------------
constexpr int TS = 16;      // spatial tile
constexpr int CO_T = 8;     // output channels per block
constexpr int CI_T = 8;     // input channels per smem stage

__global__ void __launch_bounds__(TS * TS) conv2d_f32_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ res, const float* __restrict__ slope_p, float* __restrict__ y,
    int Cin, int H, int W, int Cout, int KH, int KW, int act, int res_mode, int tiles_x) {
    extern __shared__ float smem[];
    const int TH = TS + KH - 1, TW = TS + KW - 1;
    float* xs = smem;                               // [CI_T][TH][TW]
    float* ws = smem + CI_T * TH * TW;              // [CI_T][KH*KW][CO_T]
    const int tx = threadIdx.x % TS, ty = threadIdx.x / TS;
    const int tile = blockIdx.x;
    const int w0 = (tile % tiles_x) * TS, h0 = (tile / tiles_x) * TS;
    const int co0 = blockIdx.y * CO_T;
    const int n = blockIdx.z;
    const int ph = KH / 2, pw = KW / 2;
    const int KK = KH * KW;
    float acc[CO_T];
#pragma unroll
    for (int i = 0; i < CO_T; ++i) acc[i] = 0.f;

    for (int c0 = 0; c0 < Cin; c0 += CI_T) {
        const int nci = min(CI_T, Cin - c0);
        // stage input tile (zero padded)
        for (int i = threadIdx.x; i < nci * TH * TW; i += TS * TS) {
            const int ci = i / (TH * TW);
            const int r = (i / TW) % TH, c = i % TW;
            const int gh = h0 + r - ph, gw = w0 + c - pw;
            float v = 0.f;
            if (gh >= 0 && gh < H && gw >= 0 && gw < W)
                v = __ldg(x + (((int64_t)n * Cin + c0 + ci) * H + gh) * W + gw);
            xs[i] = v;
        }
        // stage weights as [ci][tap][co]
        for (int i = threadIdx.x; i < nci * KK * CO_T; i += TS * TS) {
            const int co = i % CO_T;
            const int tap = (i / CO_T) % KK;
            const int ci = i / (CO_T * KK);
            float v = 0.f;
            if (co0 + co < Cout) v = __ldg(w + (((int64_t)(co0 + co) * Cin + c0 + ci) * KK + tap));
            ws[i] = v;
        }
        __syncthreads();
        for (int ci = 0; ci < nci; ++ci) {
            const float* xt = xs + ci * TH * TW + ty * TW + tx;
            const float* wt = ws + ci * KK * CO_T;
            for (int kh = 0; kh < KH; ++kh) {
                for (int kw = 0; kw < KW; ++kw) {
                    const float v = xt[kh * TW + kw];
                    const float4 wa = *reinterpret_cast<const float4*>(wt + (kh * KW + kw) * CO_T);
                    const float4 wb = *reinterpret_cast<const float4*>(wt + (kh * KW + kw) * CO_T + 4);
                    acc[0] = fmaf(v, wa.x, acc[0]); acc[1] = fmaf(v, wa.y, acc[1]);
                    acc[2] = fmaf(v, wa.z, acc[2]); acc[3] = fmaf(v, wa.w, acc[3]);
                    acc[4] = fmaf(v, wb.x, acc[4]); acc[5] = fmaf(v, wb.y, acc[5]);
                    acc[6] = fmaf(v, wb.z, acc[6]); acc[7] = fmaf(v, wb.w, acc[7]);
                }
            }
        }
        __syncthreads();
    }
    const int oh = h0 + ty, ow = w0 + tx;
    if (oh >= H || ow >= W) return;
    const float slope = (act == CWFA_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
#pragma unroll
    for (int i = 0; i < CO_T; ++i) {
        const int co = co0 + i;
        if (co >= Cout) break;
        const int64_t o = (((int64_t)n * Cout + co) * H + oh) * W + ow;
        float v = acc[i] + (bias ? __ldg(bias + co) : 0.f);
        if (res_mode == 1) v += __ldg(res + o);
        v = apply_act(v, act, slope);
        if (res_mode == 2) v += __ldg(res + o);
        y[o] = v;
    }
}

extern "C" int cwfa_conv2d_f32(const float* x, const float* w, const float* bias, const float* res,
                               const float* slope, float* y, int N, int Cin, int H, int W, int Cout, int KH, int KW,
                               int act, int res_mode, void* stream) {
    if (N <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || !(KH & 1) || !(KW & 1) || KH > 7 || KW > 7) {
        set_error("conv2d_f32: unsupported shape (KH=%d KW=%d must be odd <= 7)", KH, KW);
        return CWFA_EINVAL;
    }
    if (res_mode != 0 && !res) { set_error("conv2d_f32: res_mode set but res is NULL"); return CWFA_EINVAL; }
    if (N > 65535) { set_error("conv2d_f32: N too large"); return CWFA_EINVAL; }
    const int tiles_x = ceil_div(W, TS), tiles_y = ceil_div(H, TS);
    const size_t smem = sizeof(float) * (CI_T * (TS + KH - 1) * (TS + KW - 1) + CI_T * KH * KW * CO_T);
    dim3 grid(tiles_x * tiles_y, ceil_div(Cout, CO_T), N);
    conv2d_f32_kernel<<<grid, TS * TS, smem, (cudaStream_t)stream>>>(x, w, bias, res, slope, y, Cin, H, W, Cout, KH,
                                                                     KW, act, res_mode, tiles_x);
    return check_launch("conv2d_f32");
}

// ------------------------------------------------------------------------------------------
// ConvTranspose2d k=2 s=2 (+ optional skip add): one thread per INPUT pixel, 4 output channels.
// ------------------------------------------------------------------------------------------
constexpr int CT_CO = 4;
constexpr int CT_CI = 32;
__global__ void __launch_bounds__(256) convT2x2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* __restrict__ skip,
                                                       float* __restrict__ y, int Cin, int H, int W, int Cout) {
    __shared__ float ws[CT_CI][CT_CO][4];
    const int64_t P = (int64_t)H * W;
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int co0 = blockIdx.y * CT_CO;
    const int n = blockIdx.z;
    float acc[CT_CO][4];
#pragma unroll
    for (int i = 0; i < CT_CO; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int c0 = 0; c0 < Cin; c0 += CT_CI) {
        const int nci = min(CT_CI, Cin - c0);
        for (int i = threadIdx.x; i < nci * CT_CO * 4; i += blockDim.x) {
            const int k = i % 4, co = (i / 4) % CT_CO, ci = i / (4 * CT_CO);
            ws[ci][co][k] = (co0 + co < Cout) ? __ldg(w + ((int64_t)(c0 + ci) * Cout + co0 + co) * 4 + k) : 0.f;
        }
        __syncthreads();
        if (p < P) {
            for (int ci = 0; ci < nci; ++ci) {
                const float v = __ldg(x + ((int64_t)n * Cin + c0 + ci) * P + p);
#pragma unroll
                for (int co = 0; co < CT_CO; ++co) {
                    const float4 wv = *reinterpret_cast<const float4*>(&ws[ci][co][0]);
                    acc[co][0] = fmaf(v, wv.x, acc[co][0]);
                    acc[co][1] = fmaf(v, wv.y, acc[co][1]);
                    acc[co][2] = fmaf(v, wv.z, acc[co][2]);
                    acc[co][3] = fmaf(v, wv.w, acc[co][3]);
                }
            }
        }
        __syncthreads();
    }
    if (p >= P) return;
    const int h = (int)(p / W), ww = (int)(p % W);
    const int W2 = 2 * W;
#pragma unroll
    for (int co = 0; co < CT_CO; ++co) {
        if (co0 + co >= Cout) break;
        const float b = bias ? __ldg(bias + co0 + co) : 0.f;
        const int64_t base = (((int64_t)n * Cout + co0 + co) * (2 * H) + 2 * h) * W2 + 2 * ww;
        float2 r0 = make_float2(acc[co][0] + b, acc[co][1] + b);
        float2 r1 = make_float2(acc[co][2] + b, acc[co][3] + b);
        if (skip) {
            const float2 s0 = __ldg(reinterpret_cast<const float2*>(skip + base));
            const float2 s1 = __ldg(reinterpret_cast<const float2*>(skip + base + W2));
            r0.x += s0.x; r0.y += s0.y; r1.x += s1.x; r1.y += s1.y;
        }
        *reinterpret_cast<float2*>(y + base) = r0;
        *reinterpret_cast<float2*>(y + base + W2) = r1;
    }
}
extern "C" int cwfa_convT2x2_f32(const float* x, const float* w, const float* bias, const float* skip, float* y, int N,
                                 int Cin, int H, int W, int Cout, void* stream) {
    if (N <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || N > 65535) { set_error("convT2x2: bad shape"); return CWFA_EINVAL; }
    dim3 grid(ceil_div((int64_t)H * W, 256), ceil_div(Cout, CT_CO), N);
    convT2x2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, skip, y, Cin, H, W, Cout);
    return check_launch("convT2x2");
}

// ------------------------------------------------------------------------------------------
// Conditioning net depth stencil: Conv3d(1->Cm) + PReLU + Conv3d(Cm->1) over (H,W,depth),
// fused per 8x8 spatial tile and depth slab; the hidden volume lives in shared memory only.
// Volume axes of the reference's view (B,1,H,W,ch): kernel index order (kh,kw,kd).
// ------------------------------------------------------------------------------------------
constexpr int ST = 8;        // spatial tile
constexpr int SD_MAX = 48;   // depth slab
constexpr int ST_THREADS = 256;

__global__ void __launch_bounds__(ST_THREADS) depth_stencil_kernel(
    const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
    const float* __restrict__ slope_p, const float* __restrict__ w2, const float* __restrict__ b2,
    float* __restrict__ y, int D, int H, int W, int Cm, int tiles_x, int nslabs, int SD) {
    extern __shared__ float smem[];
    const int XD = SD + 4, HD = SD + 2;
    float* xs = smem;                                  // [ST+4][ST+4][XD]  (depth fastest)
    float* hs = xs + (ST + 4) * (ST + 4) * XD;         // [ST+2][ST+2][HD]
    float* wsm = hs + (ST + 2) * (ST + 2) * HD;        // [Cm][27] w1, [Cm][27] w2, [Cm] b1
    const int tile = blockIdx.x;
    const int w0 = (tile % tiles_x) * ST, h0 = (tile / tiles_x) * ST;
    const int slab = blockIdx.y % nslabs;
    const int b = blockIdx.y / nslabs;
    const int d0 = slab * SD;
    const float slope = __ldg(slope_p);
    const float bias2 = __ldg(b2);
    const int64_t P = (int64_t)H * W;

    for (int i = threadIdx.x; i < Cm * 27; i += ST_THREADS) {
        wsm[i] = __ldg(w1 + i);
        wsm[Cm * 27 + i] = __ldg(w2 + i);
    }
    for (int i = threadIdx.x; i < Cm; i += ST_THREADS) wsm[2 * Cm * 27 + i] = __ldg(b1 + i);
    // input tile with halo 2 in all three axes, zero outside the volume
    const int nx = (ST + 4) * (ST + 4) * XD;
    for (int i = threadIdx.x; i < nx; i += ST_THREADS) {
        const int dd = i % XD;
        const int c = (i / XD) % (ST + 4);
        const int r = i / (XD * (ST + 4));
        const int gd = d0 + dd - 2, gh = h0 + r - 2, gw = w0 + c - 2;
        float v = 0.f;
        if (gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W)
            v = __ldg(x + ((int64_t)b * D + gd) * P + (int64_t)gh * W + gw);
        xs[i] = v;
    }
    __syncthreads();

    constexpr int NOUT = (ST * ST * SD_MAX) / ST_THREADS;   // 12 outputs per thread at SD = 48
    float acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = 0.f;
    const int nh = (ST + 2) * (ST + 2) * HD;
    const int nout = ST * ST * SD;

    for (int c = 0; c < Cm; ++c) {
        float wr[27];
#pragma unroll
        for (int t = 0; t < 27; ++t) wr[t] = wsm[c * 27 + t];
        const float bb = wsm[2 * Cm * 27 + c];
        // phase A: hidden channel c on the (ST+2)^2 x (SD+2) halo-1 tile
        for (int i = threadIdx.x; i < nh; i += ST_THREADS) {
            const int dd = i % HD;
            const int cc = (i / HD) % (ST + 2);
            const int r = i / (HD * (ST + 2));
            const int gd = d0 + dd - 1, gh = h0 + r - 1, gw = w0 + cc - 1;
            float v = 0.f;
            if (gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
                v = bb;
                const float* xp = xs + (r * (ST + 4) + cc) * XD + dd;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                        for (int kd = 0; kd < 3; ++kd)
                            v = fmaf(wr[(kh * 3 + kw) * 3 + kd], xp[(kh * (ST + 4) + kw) * XD + kd], v);
                v = v >= 0.f ? v : slope * v;
            }
            hs[i] = v;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 27; ++t) wr[t] = wsm[Cm * 27 + c * 27 + t];
        // phase B: accumulate conv2 contribution of channel c
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            const int o = threadIdx.x + k * ST_THREADS;
            if (o < nout) {
                const int dd = o % SD;
                const int pix = o / SD;
                const int cc = pix % ST, r = pix / ST;
                const float* hp = hs + (r * (ST + 2) + cc) * HD + dd;
                float a = acc[k];
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                        for (int kd = 0; kd < 3; ++kd)
                            a = fmaf(wr[(kh * 3 + kw) * 3 + kd], hp[(kh * (ST + 2) + kw) * HD + kd], a);
                acc[k] = a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
        const int o = threadIdx.x + k * ST_THREADS;
        if (o < nout) {
            const int dd = o % SD;
            const int pix = o / SD;
            const int cc = pix % ST, r = pix / ST;
            const int gd = d0 + dd, gh = h0 + r, gw = w0 + cc;
            if (gd < D && gh < H && gw < W) y[((int64_t)b * D + gd) * P + (int64_t)gh * W + gw] = acc[k] + bias2;
        }
    }
}

extern "C" int cwfa_depth_stencil3d_f32(const float* x, const float* w1, const float* b1, const float* slope,
                                        const float* w2, const float* b2, float* y, int B, int ch, int H, int W,
                                        int Cm, void* stream) {
    if (B <= 0 || ch <= 0 || H <= 0 || W <= 0 || Cm <= 0 || Cm > 64) { set_error("depth_stencil3d: bad shape"); return CWFA_EINVAL; }
    const int SD = ch < SD_MAX ? ch : SD_MAX;
    const int nslabs = ceil_div(ch, SD);
    const int tiles_x = ceil_div(W, ST), tiles_y = ceil_div(H, ST);
    if ((int64_t)B * nslabs > 65535) { set_error("depth_stencil3d: batch too large"); return CWFA_EINVAL; }
    const size_t smem = sizeof(float) * ((ST + 4) * (ST + 4) * (SD + 4) + (ST + 2) * (ST + 2) * (SD + 2) + 2 * Cm * 27 + Cm);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(depth_stencil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_set = true;
    }
    dim3 grid(tiles_x * tiles_y, B * nslabs);
    depth_stencil_kernel<<<grid, ST_THREADS, smem, (cudaStream_t)stream>>>(x, w1, b1, slope, w2, b2, y, ch, H, W, Cm,
                                                                            tiles_x, nslabs, SD);
    return check_launch("depth_stencil3d");
}
