// Training-time (backward) kernels of the fp32 module path, sm_100a.
//
// The reference trains one flow level at a time with autograd (CWFA.py:928-1015: inverse pass with
// gradients for the MSE term, forward pass for the NLL term, Lion update).  These kernels are the
// hand-written adjoints of the forward kernels in elementwise.cu / conv_f32.cu:
//   * conv2d weight gradient (register-tiled, split over pixel tiles, deterministic two-stage sum);
//     the data gradient is the forward direct convolution with flipped/transposed weights;
//   * affine coupling adjoint (dx, d a_s through the atan clamp, d a_t, log-det cotangent folded in);
//   * ELU / PReLU adjoints, the 3-D depth-stencil convolutions of the conditioning net and their
//     weight gradients, the Lion update on a flat parameter buffer.
// All reductions are two-stage with a fixed summation order (bit-reproducible run to run).
#include "common.cuh"
#include "tc_common.cuh"

using namespace cwfa;

// ------------------------------------------------------------------------------------------
// conv2d weight gradient: dW[co,ci,kh,kw] = sum_{n,h,w} dy[n,co,h,w] * x[n,ci,h+kh-p,w+kw-p]
// CTA = 32 output channels x 32 input channels; a thread owns a 2x2 channel block x KS*KS taps and
// walks the 256 pixels of a tile with a sliding register window over x.
// ------------------------------------------------------------------------------------------
constexpr int WG_TH = 8, WG_TW = 32;
constexpr int WG_DPLANE = WG_TH * WG_TW + 1;

// CB = channels per thread in each of the two channel axes: 2 (1x1 / 3x3: 2x2xKS^2 accumulators) or 1 (7x7: 49 accumulators)
template <int KS, int CB>
struct WgradGeom {
    static constexpr int C = 16 * CB;                // channel block of the CTA (both Cout and Cin)
    static constexpr int XH = WG_TH + KS - 1, XW = WG_TW + KS - 1;
    static constexpr int XPLANE = (XH * XW) | 1;     // odd plane stride: 16 channels hit 16 banks
    static constexpr size_t smem = sizeof(float) * (C * XPLANE + C * WG_DPLANE);
};

template <int KS, int CB>
__global__ void __launch_bounds__(256) conv2d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           float* __restrict__ part, int N, int Cin, int H, int W,
                                                           int Cout, int tiles_x, int tiles_y, int n_ci_blk) {
    using G = WgradGeom<KS, CB>;
    constexpr int PAD = KS / 2, KK = KS * KS, WG_C = G::C;
    extern __shared__ float smem[];
    float* xs = smem;                       // [WG_C][XPLANE]
    float* ds = smem + WG_C * G::XPLANE;    // [WG_C][WG_DPLANE]
    const int tid = threadIdx.x;
    const int tco = tid >> 4, tci = tid & 15;          // channels tco (+16) / tci (+16)
    const int co0 = (blockIdx.y / n_ci_blk) * WG_C, ci0 = (blockIdx.y % n_ci_blk) * WG_C;
    const int tiles = tiles_x * tiles_y;
    const int64_t P = (int64_t)H * W;

    float acc[CB][CB][KK];
#pragma unroll
    for (int i = 0; i < CB; ++i)
#pragma unroll
        for (int j = 0; j < CB; ++j)
#pragma unroll
            for (int t = 0; t < KK; ++t) acc[i][j][t] = 0.f;

    for (int item = blockIdx.x; item < N * tiles; item += gridDim.x) {
        const int n = item / tiles, t = item % tiles;
        const int h0 = (t / tiles_x) * WG_TH, w0 = (t % tiles_x) * WG_TW;
        __syncthreads();
        for (int i = tid; i < WG_C * G::XH * G::XW; i += 256) {
            const int cl = i / (G::XH * G::XW), rem = i % (G::XH * G::XW);
            const int r = rem / G::XW, c = rem % G::XW;
            const int gh = h0 + r - PAD, gw = w0 + c - PAD, ci = ci0 + cl;
            float v = 0.f;
            if (ci < Cin && gh >= 0 && gh < H && gw >= 0 && gw < W) v = __ldg(x + ((int64_t)n * Cin + ci) * P + (int64_t)gh * W + gw);
            xs[cl * G::XPLANE + rem] = v;
        }
        for (int i = tid; i < WG_C * WG_TH * WG_TW; i += 256) {
            const int cl = i / (WG_TH * WG_TW), pix = i % (WG_TH * WG_TW);
            const int gh = h0 + pix / WG_TW, gw = w0 + pix % WG_TW, co = co0 + cl;
            float v = 0.f;
            if (co < Cout && gh < H && gw < W) v = __ldg(dy + ((int64_t)n * Cout + co) * P + (int64_t)gh * W + gw);
            ds[cl * WG_DPLANE + pix] = v;
        }
        __syncthreads();
        const float* xq[CB];
        const float* dq[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            xq[j] = xs + (tci + 16 * j) * G::XPLANE;
            dq[j] = ds + (tco + 16 * j) * WG_DPLANE;
        }
        for (int r = 0; r < WG_TH; ++r) {
            float win[CB][KS][KS];
#pragma unroll
            for (int j = 0; j < CB; ++j)
#pragma unroll
                for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                    for (int kw = 0; kw < KS - 1; ++kw) win[j][kh][kw + 1] = xq[j][(r + kh) * G::XW + kw];
#pragma unroll
            for (int c = 0; c < WG_TW; ++c) {
#pragma unroll
                for (int j = 0; j < CB; ++j)
#pragma unroll
                    for (int kh = 0; kh < KS; ++kh) {
#pragma unroll
                        for (int kw = 0; kw < KS - 1; ++kw) win[j][kh][kw] = win[j][kh][kw + 1];
                        win[j][kh][KS - 1] = xq[j][(r + kh) * G::XW + c + KS - 1];
                    }
                float g[CB];
#pragma unroll
                for (int i = 0; i < CB; ++i) g[i] = dq[i][r * WG_TW + c];
#pragma unroll
                for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                    for (int kw = 0; kw < KS; ++kw)
#pragma unroll
                        for (int i = 0; i < CB; ++i)
#pragma unroll
                            for (int j = 0; j < CB; ++j)
                                acc[i][j][kh * KS + kw] = fmaf(g[i], win[j][kh][kw], acc[i][j][kh * KS + kw]);
            }
        }
    }
    float* dst = part + (int64_t)blockIdx.x * Cout * Cin * KK;
#pragma unroll
    for (int i = 0; i < CB; ++i)
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            const int co = co0 + tco + 16 * i, ci = ci0 + tci + 16 * j;
            if (co < Cout && ci < Cin) {
#pragma unroll
                for (int t = 0; t < KK; ++t) dst[((int64_t)co * Cin + ci) * KK + t] = acc[i][j][t];
            }
        }
}

// out[i] = sum_k part[k*n + i] (fixed order, double accumulation), optionally accumulated into out.
__global__ void __launch_bounds__(256) partial_sum_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n,
                                                          int nparts, int accumulate) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < nparts; ++k) s += (double)part[(int64_t)k * n + i];
    out[i] = accumulate ? out[i] + (float)s : (float)s;
}

static int wgrad_chunks(int N, int H, int W, int Cout, int Cin, int KS) {
    const int cblk = KS == 7 ? 16 : 32;
    const int tiles = ceil_div(W, WG_TW) * ceil_div(H, WG_TH);
    const int combos = ceil_div(Cout, cblk) * ceil_div(Cin, cblk);
    int chunks = (kNumSMs * 4) / combos;
    if (chunks < 1) chunks = 1;
    const int64_t items = (int64_t)N * tiles;
    if (chunks > items) chunks = (int)items;
    return chunks;
}

extern "C" int64_t cwfa_conv2d_wgrad_workspace_floats(int N, int Cin, int H, int W, int Cout, int KH, int KW) {
    return (int64_t)wgrad_chunks(N, H, W, Cout, Cin, KH) * Cout * Cin * KH * KW;
}

template <int KS, int CB>
static void wgrad_launch(dim3 grid, cudaStream_t st, const float* x, const float* dy, float* ws, int N, int Cin, int H, int W, int Cout,
                         int tiles_x, int tiles_y, int n_ci) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(conv2d_wgrad_kernel<KS, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgradGeom<KS, CB>::smem); attr = true; }
    conv2d_wgrad_kernel<KS, CB><<<grid, 256, WgradGeom<KS, CB>::smem, st>>>(x, dy, ws, N, Cin, H, W, Cout, tiles_x, tiles_y, n_ci);
}

extern "C" int cwfa_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, float* workspace, int N, int Cin, int H,
                                     int W, int Cout, int KH, int KW, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || KH != KW || (KH != 1 && KH != 3 && KH != 7)) {
        set_error("conv2d_wgrad: unsupported shape (square kernels of size 1, 3 or 7, got %dx%d)", KH, KW);
        return CWFA_EINVAL;
    }
    const int cblk = KH == 7 ? 16 : 32;
    const int tiles_x = ceil_div(W, WG_TW), tiles_y = ceil_div(H, WG_TH);
    const int n_ci = ceil_div(Cin, cblk), n_co = ceil_div(Cout, cblk);
    const int chunks = wgrad_chunks(N, H, W, Cout, Cin, KH);
    dim3 grid(chunks, n_co * n_ci);
    if (KH == 3) wgrad_launch<3, 2>(grid, st, x, dy, workspace, N, Cin, H, W, Cout, tiles_x, tiles_y, n_ci);
    else if (KH == 1) wgrad_launch<1, 2>(grid, st, x, dy, workspace, N, Cin, H, W, Cout, tiles_x, tiles_y, n_ci);
    else wgrad_launch<7, 1>(grid, st, x, dy, workspace, N, Cin, H, W, Cout, tiles_x, tiles_y, n_ci);
    int rc = check_launch("conv2d_wgrad");
    if (rc) return rc;
    const int64_t n = (int64_t)Cout * Cin * KH * KW;
    partial_sum_kernel<<<ceil_div(n, 256), 256, 0, st>>>(workspace, dw, n, chunks, accumulate);
    return check_launch("conv2d_wgrad_finalize");
}

// Weights of the data-gradient convolution: wt[ci,co,kh,kw] = w[co,ci,KH-1-kh,KW-1-kw]
// (dx = conv_same(dy, wt) for stride 1, odd kernels).
__global__ void dgrad_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin, int KH, int KW) {
    const int64_t n = (int64_t)Cout * Cin * KH * KW;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int kw = (int)(i % KW), kh = (int)((i / KW) % KH);
    const int co = (int)((i / ((int64_t)KW * KH)) % Cout), ci = (int)(i / ((int64_t)KW * KH * Cout));
    wt[i] = __ldg(w + (((int64_t)co * Cin + ci) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw));
}
extern "C" int cwfa_conv2d_dgrad_weights_f32(const float* w, float* wt, int Cout, int Cin, int KH, int KW, void* stream) {
    const int64_t n = (int64_t)Cout * Cin * KH * KW;
    if (n <= 0) { set_error("dgrad_weights: bad shape"); return CWFA_EINVAL; }
    dgrad_weights_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w, wt, Cout, Cin, KH, KW);
    return check_launch("dgrad_weights");
}

// ------------------------------------------------------------------------------------------
// activation adjoints / small element-wise helpers
// ------------------------------------------------------------------------------------------
static inline int ew_blocks(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > kNumSMs * 16) b = kNumSMs * 16;
    return b < 1 ? 1 : (int)b;
}

// ELU(alpha=1) adjoint from the OUTPUT y: dv = dy * (y > 0 ? 1 : y + 1)
__global__ void __launch_bounds__(256) elu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                      float* __restrict__ dv, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float yy = __ldg(y + i);
        dv[i] = __ldg(dy + i) * (yy > 0.f ? 1.f : yy + 1.f);
    }
}
extern "C" int cwfa_elu_bwd_f32(const float* dy, const float* y, float* dv, int64_t n, void* stream) {
    if (n <= 0) { set_error("elu_bwd: bad size"); return CWFA_EINVAL; }
    elu_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(dy, y, dv, n);
    return check_launch("elu_bwd");
}

// out = alpha * a + beta * b   (b may be NULL)
__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                    float* __restrict__ out, float alpha, float beta, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = alpha * a[i] + (b ? beta * b[i] : 0.f);
}
extern "C" int cwfa_axpby_f32(const float* a, const float* b, float* out, float alpha, float beta, int64_t n, void* stream) {
    if (n <= 0 || !a || !out) { set_error("axpby: bad args"); return CWFA_EINVAL; }
    axpby_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, out, alpha, beta, n);
    return check_launch("axpby");
}

// PReLU with one shared slope (nn.PReLU(), networks.py:209): y = v > 0 ? v : a*v
__global__ void __launch_bounds__(256) prelu_fwd_kernel(const float* __restrict__ v, const float* __restrict__ slope,
                                                        float* __restrict__ y, int64_t n) {
    const float a = __ldg(slope);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = v[i];
        y[i] = x > 0.f ? x : a * x;
    }
}
extern "C" int cwfa_prelu_f32(const float* v, const float* slope, float* y, int64_t n, void* stream) {
    if (n <= 0) { set_error("prelu: bad size"); return CWFA_EINVAL; }
    prelu_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(v, slope, y, n);
    return check_launch("prelu");
}

constexpr int kReduceBlocks = kNumSMs * 2;
// dv = v > 0 ? dy : a*dy ; dslope = sum_{v<=0} v*dy  (torch's PReLU adjoint convention at v = 0)
__global__ void __launch_bounds__(256) prelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ v,
                                                        const float* __restrict__ slope, float* __restrict__ dv,
                                                        float* __restrict__ ws, int64_t n) {
    const float a = __ldg(slope);
    float s = 0.f, dummy = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = v[i], g = dy[i];
        if (x > 0.f) {
            dv[i] = g;
        } else {
            dv[i] = a * g;
            s = fmaf(x, g, s);
        }
    }
    block_sum2(s, dummy);
    if (threadIdx.x == 0) ws[blockIdx.x] = s;
}
extern "C" int cwfa_reduce_workspace_blocks(void) { return kReduceBlocks; }
extern "C" int cwfa_prelu_bwd_f32(const float* dy, const float* v, const float* slope, float* dv, float* dslope,
                                  float* workspace, int64_t n, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) { set_error("prelu_bwd: bad size"); return CWFA_EINVAL; }
    prelu_bwd_kernel<<<kReduceBlocks, 256, 0, st>>>(dy, v, slope, dv, workspace, n);
    int rc = check_launch("prelu_bwd");
    if (rc) return rc;
    partial_sum_kernel<<<1, 32, 0, st>>>(workspace, dslope, 1, kReduceBlocks, 0);
    return check_launch("prelu_bwd_finalize");
}

// ------------------------------------------------------------------------------------------
// affine coupling adjoint (forward: elementwise.cu affine_kernel; coupling_layers.py:490-500)
//   fwd  y = e^s x + t        : dx = dy e^s ; dt = dy ; ds = dy x e^s + gJ
//   inv  y = (x - t) e^{-s}   : dx = dy e^{-s} ; dt = -dx ; ds = -dy y - gJ
//   s = kk atan(a_s) -> d a_s = ds kk / (1 + a_s^2)  (raw: d a_s = ds);  t = t_scale a_t
// ------------------------------------------------------------------------------------------
template <bool INV>
__device__ __forceinline__ void affine_bwd_one(float as, float tv, float g, float xv, float gj, float kk, float t_scale, int raw,
                                               float k_in, float& gx, float& gt, float& gs) {
    const float th = raw == 2 ? tanhf(k_in * as) : 0.f;
    const float s = raw == 1 ? as : (raw == 2 ? kk * th : kk * atanf(as));
    float gsv;
    if (INV) {
        const float e = expf(-s);
        const float y = (xv - t_scale * tv) * e;
        gx = g * e;
        gt = -gx;
        gsv = -g * y - gj;
    } else {
        const float e = expf(s);
        gx = g * e;
        gt = g;
        gsv = gx * xv + gj;
    }
    gt *= t_scale;
    gs = raw == 1 ? gsv : (raw == 2 ? gsv * kk * k_in * (1.f - th * th) : gsv * kk / fmaf(as, as, 1.f));
}
// VEC = 4: 16-byte accesses on all seven streams (needs n, the leading dimensions and the pointers 16-byte aligned)
template <bool INV, int VEC>
__global__ void __launch_bounds__(256) affine_bwd_kernel(const float* __restrict__ x, const float* __restrict__ a_s,
                                                         const float* __restrict__ a_t, const float* __restrict__ dy,
                                                         const float* __restrict__ g_logdet, float* __restrict__ dx,
                                                         float* __restrict__ da_s, float* __restrict__ da_t, int64_t n,
                                                         int64_t ld_s, int64_t ld_t, int64_t ld_ds, int64_t ld_dt,
                                                         float kk, float t_scale, int raw, float k_in) {
    const int b = blockIdx.y;
    const float* xs = x ? x + (int64_t)b * n : nullptr;
    const float* ss = a_s + (int64_t)b * ld_s;
    const float* ts = a_t + (int64_t)b * ld_t;
    const float* gs = dy + (int64_t)b * n;
    const float gj = g_logdet ? __ldg(g_logdet + b) : 0.f;
    const int64_t nv = n / VEC;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        if constexpr (VEC == 4) {
            const float4 as = __ldg(reinterpret_cast<const float4*>(ss) + i), tv = __ldg(reinterpret_cast<const float4*>(ts) + i);
            const float4 g = __ldg(reinterpret_cast<const float4*>(gs) + i);
            const float4 xv = xs ? __ldg(reinterpret_cast<const float4*>(xs) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 ox, ot, os;
            affine_bwd_one<INV>(as.x, tv.x, g.x, xv.x, gj, kk, t_scale, raw, k_in, ox.x, ot.x, os.x);
            affine_bwd_one<INV>(as.y, tv.y, g.y, xv.y, gj, kk, t_scale, raw, k_in, ox.y, ot.y, os.y);
            affine_bwd_one<INV>(as.z, tv.z, g.z, xv.z, gj, kk, t_scale, raw, k_in, ox.z, ot.z, os.z);
            affine_bwd_one<INV>(as.w, tv.w, g.w, xv.w, gj, kk, t_scale, raw, k_in, ox.w, ot.w, os.w);
            if (dx) reinterpret_cast<float4*>(dx + (int64_t)b * n)[i] = ox;
            if (da_t) reinterpret_cast<float4*>(da_t + (int64_t)b * ld_dt)[i] = ot;
            if (da_s) reinterpret_cast<float4*>(da_s + (int64_t)b * ld_ds)[i] = os;
        } else {
            float ox, ot, os;
            affine_bwd_one<INV>(ss[i], ts[i], gs[i], xs ? xs[i] : 0.f, gj, kk, t_scale, raw, k_in, ox, ot, os);
            if (dx) dx[(int64_t)b * n + i] = ox;
            if (da_t) da_t[(int64_t)b * ld_dt + i] = ot;
            if (da_s) da_s[(int64_t)b * ld_ds + i] = os;
        }
    }
}
extern "C" int cwfa_affine_bwd(const float* x, const float* a_s, const float* a_t, const float* dy, const float* g_logdet,
                               float* dx, float* da_s, float* da_t, int B, int ch, int64_t P, int64_t ld_s, int64_t ld_t,
                               int64_t ld_ds, int64_t ld_dt, float clamp, float k_atan, float t_scale, int flags,
                               void* stream) {
    const int inverse = flags & 1, raw = (flags & 4) ? 2 : ((flags >> 1) & 1);
    if (B <= 0 || ch <= 0 || P <= 0 || !a_s || !a_t || !dy) { set_error("affine_bwd: bad args"); return CWFA_EINVAL; }
    if (!x && !inverse) { set_error("affine_bwd: x may be NULL only in inverse mode"); return CWFA_EINVAL; }
    const int64_t n = (int64_t)ch * P;
    const float kk = raw == 2 ? clamp : clamp * k_atan;
    const bool vec = (n % 4 == 0) && (ld_s % 4 == 0) && (ld_t % 4 == 0) && (ld_ds % 4 == 0) && (ld_dt % 4 == 0) && aligned16(a_s) &&
                     aligned16(a_t) && aligned16(dy) && (!x || aligned16(x)) && (!dx || aligned16(dx)) && (!da_s || aligned16(da_s)) &&
                     (!da_t || aligned16(da_t));
    const int64_t items = vec ? n / 4 : n;
    dim3 grid(ew_blocks(items) < kNumSMs * 8 ? ew_blocks(items) : kNumSMs * 8, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) {
        if (inverse) affine_bwd_kernel<true, 4><<<grid, 256, 0, st>>>(x, a_s, a_t, dy, g_logdet, dx, da_s, da_t, n, ld_s, ld_t, ld_ds, ld_dt, kk, t_scale, raw, k_atan);
        else affine_bwd_kernel<false, 4><<<grid, 256, 0, st>>>(x, a_s, a_t, dy, g_logdet, dx, da_s, da_t, n, ld_s, ld_t, ld_ds, ld_dt, kk, t_scale, raw, k_atan);
    } else {
        if (inverse) affine_bwd_kernel<true, 1><<<grid, 256, 0, st>>>(x, a_s, a_t, dy, g_logdet, dx, da_s, da_t, n, ld_s, ld_t, ld_ds, ld_dt, kk, t_scale, raw, k_atan);
        else affine_bwd_kernel<false, 1><<<grid, 256, 0, st>>>(x, a_s, a_t, dy, g_logdet, dx, da_s, da_t, n, ld_s, ld_t, ld_ds, ld_dt, kk, t_scale, raw, k_atan);
    }
    return check_launch("affine_bwd");
}

// ------------------------------------------------------------------------------------------
// Conditioning-net depth stencil, unfused (training): Conv3d(1,Cm,3,p1) and Conv3d(Cm,1,3,p1) over
// the (H,W,depth) volume of a (B,D,H,W) tensor (networks.py:221-225,239), plus weight gradients.
// Hidden tensors are (B,Cm,D,H,W).  Tap index = (kh*3+kw)*3+kd, as in the Conv3d weight (Cm,1,3,3,3)
// whose spatial axes are (H,W,depth).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_taps27(const float* __restrict__ vol, int D, int H, int W, int d, int h, int w, float* v) {
    const int64_t P = (int64_t)H * W;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                const int gd = d + kd - 1, gh = h + kh - 1, gw = w + kw - 1;
                v[(kh * 3 + kw) * 3 + kd] = (gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W)
                                                ? __ldg(vol + (int64_t)gd * P + (int64_t)gh * W + gw) : 0.f;
            }
}

// Row walker shared by the stencil kernels: a block covers `rpb` rows (b,d,h) per iteration, a thread a fixed column
// slot -- index arithmetic is a few 32-bit divisions per ROW, none per element.
struct RowWalk {
    int tpr, rpb, trow, tcol, rows;
    __device__ RowWalk(int B, int D, int H, int slots) {
        tpr = slots < 256 ? slots : 256;
        rpb = 256 / tpr;
        trow = threadIdx.x / tpr;
        tcol = threadIdx.x % tpr;
        rows = B * D * H;
    }
};

// out[b,c,pos] = bias[c] + sum_t w[c,t] * x[b,pos+t]   (flip: use w[c,26-t], the adjoint of Cm->1)
__global__ void __launch_bounds__(256) stencil_1toC_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out, int B, int D,
                                                           int H, int W, int Cm, int flip) {
    extern __shared__ float wsm[];      // [Cm][27] + [Cm]
    for (int i = threadIdx.x; i < Cm * 27; i += blockDim.x) {
        const int c = i / 27, t = i % 27;
        wsm[i] = __ldg(w + c * 27 + (flip ? 26 - t : t));
    }
    for (int i = threadIdx.x; i < Cm; i += blockDim.x) wsm[Cm * 27 + i] = bias ? __ldg(bias + i) : 0.f;
    __syncthreads();
    const int64_t V = (int64_t)D * H * W;
    const RowWalk rw(B, D, H, W);
    if (rw.trow >= rw.rpb) return;
    for (int row = blockIdx.x * rw.rpb + rw.trow; row < rw.rows; row += gridDim.x * rw.rpb) {
        const int h = row % H, d = (row / H) % D, b = row / (H * D);
        const int64_t rowoff = ((int64_t)d * H + h) * W;
        for (int ww = rw.tcol; ww < W; ww += rw.tpr) {
            float v[27];
            load_taps27(x + (int64_t)b * V, D, H, W, d, h, ww, v);
            float* o = out + (int64_t)b * Cm * V + rowoff + ww;
            for (int c = 0; c < Cm; ++c) {
                float a = wsm[Cm * 27 + c];
#pragma unroll
                for (int t = 0; t < 27; ++t) a = fmaf(wsm[c * 27 + t], v[t], a);
                o[(int64_t)c * V] = a;
            }
        }
    }
}

// out[b,pos] = bias + sum_c sum_t w[c,t] * hid[b,c,pos+t]   (flip: w[c,26-t], the adjoint of 1->Cm)
// A thread owns 4 consecutive w positions: every (kd,kh) row segment of 6 values is loaded once for 12 FMAs.
__global__ void __launch_bounds__(256) stencil_Cto1_kernel(const float* __restrict__ hid, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out, int B, int D,
                                                           int H, int W, int Cm, int flip) {
    extern __shared__ float wsm[];
    for (int i = threadIdx.x; i < Cm * 27; i += blockDim.x) {
        const int c = i / 27, t = i % 27;
        wsm[i] = __ldg(w + c * 27 + (flip ? 26 - t : t));
    }
    __syncthreads();
    const int W4 = (W + 3) / 4;
    const int64_t V = (int64_t)D * H * W, P = (int64_t)H * W;
    const float b0 = bias ? __ldg(bias) : 0.f;
    const RowWalk rw(B, D, H, W4);
    if (rw.trow >= rw.rpb) return;
    for (int row = blockIdx.x * rw.rpb + rw.trow; row < rw.rows; row += gridDim.x * rw.rpb) {
        const int h = row % H, d = (row / H) % D, b = row / (H * D);
        for (int wc = rw.tcol; wc < W4; wc += rw.tpr) {
            const int w0 = wc * 4;
            float a[4] = {b0, b0, b0, b0};
            for (int c = 0; c < Cm; ++c) {
                const float* vol = hid + ((int64_t)b * Cm + c) * V;
                const float* wc_ = wsm + c * 27;
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const int gd = d + kd - 1;
                    if (gd < 0 || gd >= D) continue;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const int gh = h + kh - 1;
                        if (gh < 0 || gh >= H) continue;
                        const float* rowp = vol + (int64_t)gd * P + (int64_t)gh * W;
                        float v[6];
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            const int gw = w0 + j - 1;
                            v[j] = (gw >= 0 && gw < W) ? __ldg(rowp + gw) : 0.f;
                        }
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const float ww = wc_[(kh * 3 + kw) * 3 + kd];
#pragma unroll
                            for (int j = 0; j < 4; ++j) a[j] = fmaf(ww, v[j + kw], a[j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (w0 + j < W) out[(int64_t)b * V + (int64_t)d * P + (int64_t)h * W + w0 + j] = a[j];
        }
    }
}

// part[blk][c][t] = sum_{b,pos in blk} multi[b,c,pos] * single[b,pos+t]  (flip: stored at 26-t)
//   1->Cm weights: single = x, multi = d(hidden pre-activation), flip = 0
//   Cm->1 weights: single = dy, multi = hidden,                 flip = 1
constexpr int kStencilWgBlocks = kNumSMs;
constexpr int SW_CG = 4;            // hidden channels per thread: the 27 taps of `single` are loaded once for 4*27 FMAs
__global__ void __launch_bounds__(256) stencil_wgrad_kernel(const float* __restrict__ single, const float* __restrict__ multi,
                                                            float* __restrict__ part, int B, int D, int H, int W, int Cm,
                                                            int flip) {
    __shared__ float red[8][SW_CG * 27];
    const int c0 = blockIdx.y * SW_CG;
    const int64_t V = (int64_t)D * H * W;
    float acc[SW_CG][27];
#pragma unroll
    for (int j = 0; j < SW_CG; ++j)
#pragma unroll
        for (int t = 0; t < 27; ++t) acc[j][t] = 0.f;
    const RowWalk rw(B, D, H, W);
    if (rw.trow < rw.rpb) {
        for (int row = blockIdx.x * rw.rpb + rw.trow; row < rw.rows; row += gridDim.x * rw.rpb) {
            const int h = row % H, d = (row / H) % D, b = row / (H * D);
            const int64_t rowoff = ((int64_t)d * H + h) * W;
            for (int ww = rw.tcol; ww < W; ww += rw.tpr) {
                float v[27];
                load_taps27(single + (int64_t)b * V, D, H, W, d, h, ww, v);
                const float* mp = multi + ((int64_t)b * Cm + c0) * V + rowoff + ww;
#pragma unroll
                for (int j = 0; j < SW_CG; ++j) {
                    const float g = (c0 + j < Cm) ? __ldg(mp + (int64_t)j * V) : 0.f;
#pragma unroll
                    for (int t = 0; t < 27; ++t) acc[j][t] = fmaf(g, v[t], acc[j][t]);
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < SW_CG; ++j)
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            const float s = warp_sum(acc[j][t]);
            if (lane == 0) red[wp][j * 27 + t] = s;
        }
    __syncthreads();
    if (threadIdx.x < SW_CG * 27) {
        const int j = threadIdx.x / 27, t0 = threadIdx.x % 27;
        if (c0 + j < Cm) {
            float s = 0.f;
            for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
            const int t = flip ? 26 - t0 : t0;
            part[((int64_t)blockIdx.x * Cm + c0 + j) * 27 + t] = s;
        }
    }
}

extern "C" int cwfa_stencil3d_1toC_f32(const float* x, const float* w, const float* bias, float* out, int B, int D, int H,
                                       int W, int Cm, int flip, void* stream) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || Cm <= 0 || Cm > 256 || (int64_t)B * D * H > 0x7fffffff) { set_error("stencil3d_1toC: bad shape"); return CWFA_EINVAL; }
    const int rows = B * D * H;
    const int grid = rows < kNumSMs * 8 ? rows : kNumSMs * 8;
    stencil_1toC_kernel<<<grid, 256, sizeof(float) * Cm * 28, (cudaStream_t)stream>>>(x, w, bias, out, B, D, H, W, Cm, flip);
    return check_launch("stencil3d_1toC");
}
extern "C" int cwfa_stencil3d_Cto1_f32(const float* hid, const float* w, const float* bias, float* out, int B, int D, int H,
                                       int W, int Cm, int flip, void* stream) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || Cm <= 0 || Cm > 256 || (int64_t)B * D * H > 0x7fffffff) { set_error("stencil3d_Cto1: bad shape"); return CWFA_EINVAL; }
    const int rows = B * D * H;
    const int grid = rows < kNumSMs * 8 ? rows : kNumSMs * 8;
    stencil_Cto1_kernel<<<grid, 256, sizeof(float) * Cm * 27, (cudaStream_t)stream>>>(hid, w, bias, out, B, D, H, W, Cm, flip);
    return check_launch("stencil3d_Cto1");
}
extern "C" int cwfa_stencil3d_wgrad_workspace_floats(int Cm) { return kStencilWgBlocks * Cm * 27; }
extern "C" int cwfa_stencil3d_wgrad_f32(const float* single, const float* multi, float* dw, float* workspace, int B, int D,
                                        int H, int W, int Cm, int flip, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || Cm <= 0 || Cm > 65535) { set_error("stencil3d_wgrad: bad shape"); return CWFA_EINVAL; }
    stencil_wgrad_kernel<<<dim3(kStencilWgBlocks, ceil_div(Cm, SW_CG)), 256, 0, st>>>(single, multi, workspace, B, D, H, W, Cm, flip);
    int rc = check_launch("stencil3d_wgrad");
    if (rc) return rc;
    const int64_t n = (int64_t)Cm * 27;
    partial_sum_kernel<<<ceil_div(n, 256), 256, 0, st>>>(workspace, dw, n, kStencilWgBlocks, 0);
    return check_launch("stencil3d_wgrad_finalize");
}

// ------------------------------------------------------------------------------------------
// Lion update on a flat buffer (lion_pytorch 0.0.7 `update_fn`, the optimiser of CWFA.py:381,608-610):
//   p *= 1 - lr*wd ; p -= lr * sign(b1*m + (1-b1)*g) ; m = b2*m + (1-b2)*g      (g pre-scaled by gscale)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lion_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   const uint8_t* __restrict__ mask, int64_t n, float lr, float b1, float b2,
                                                   float wd, float gscale) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;                 // parameter received no gradient this step: lion_pytorch skips it entirely
        const float gg = g[i] * gscale, mm = m[i];
        const float u = b1 * mm + (1.f - b1) * gg;
        const float sg = (u > 0.f) ? 1.f : (u < 0.f ? -1.f : 0.f);
        p[i] = p[i] * (1.f - lr * wd) - lr * sg;
        m[i] = b2 * mm + (1.f - b2) * gg;
    }
}
extern "C" int cwfa_lion_step_masked_f32(float* p, const float* g, float* m, const uint8_t* mask, int64_t n, float lr, float beta1,
                                         float beta2, float weight_decay, float grad_scale, void* stream) {
    if (n <= 0 || !p || !g || !m) { set_error("lion_step: bad args"); return CWFA_EINVAL; }
    lion_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, mask, n, lr, beta1, beta2, weight_decay, grad_scale);
    return check_launch("lion_step");
}
extern "C" int cwfa_lion_step_f32(float* p, const float* g, float* m, int64_t n, float lr, float beta1, float beta2,
                                  float weight_decay, float grad_scale, void* stream) {
    return cwfa_lion_step_masked_f32(p, g, m, nullptr, n, lr, beta1, beta2, weight_decay, grad_scale, stream);
}

// ------------------------------------------------------------------------------------------
// LRNN U-Net adjoints (unet.py:72-113,161-195): BatchNorm2d, 2x2 max-pool, ConvTranspose2d(k=2,s=2) as a 1x1 convolution
// followed by a pixel shuffle (+ skip add)
// ------------------------------------------------------------------------------------------
// per channel: (sum dy, sum dy*x) over (N,H,W); same two-stage scheme / workspace as channel_stats (32 blocks per channel)
constexpr int kDotBlocks = 32;
__global__ void __launch_bounds__(256) channel_dot_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ ws, int N, int C, int64_t P) {
    const int c = blockIdx.y;
    float s = 0.f, q = 0.f;
    for (int n = 0; n < N; ++n) {
        const float* xp = x + ((int64_t)n * C + c) * P;
        const float* gp = dy + ((int64_t)n * C + c) * P;
        for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
            const float g = __ldg(gp + p);
            s += g;
            q = fmaf(g, __ldg(xp + p), q);
        }
    }
    block_sum2(s, q);
    if (threadIdx.x == 0) {
        ws[((int64_t)c * gridDim.x + blockIdx.x) * 2 + 0] = s;
        ws[((int64_t)c * gridDim.x + blockIdx.x) * 2 + 1] = q;
    }
}
__global__ void channel_dot_finalize_kernel(const float* __restrict__ ws, float* __restrict__ out, int C, int nblocks) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < nblocks; ++i) {
        s += (double)ws[((int64_t)c * nblocks + i) * 2 + 0];
        q += (double)ws[((int64_t)c * nblocks + i) * 2 + 1];
    }
    out[c] = (float)s;
    out[C + c] = (float)q;
}
extern "C" int cwfa_channel_dot_workspace_blocks(void) { return kDotBlocks; }
extern "C" int cwfa_channel_dot_stats_f32(const float* x, const float* dy, float* out, float* workspace, int N, int C, int64_t P,
                                          void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0 || C <= 0 || C > 65535 || P <= 0) { set_error("channel_dot_stats: bad shape"); return CWFA_EINVAL; }
    channel_dot_kernel<<<dim3(kDotBlocks, C), 256, 0, st>>>(x, dy, workspace, N, C, P);
    int rc = check_launch("channel_dot_stats");
    if (rc) return rc;
    channel_dot_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(workspace, out, C, kDotBlocks);
    return check_launch("channel_dot_finalize");
}

// dx = a[c]*dy + b[c]*x + c0[c]   (BatchNorm adjoint with the per-channel coefficients computed from the two sums)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ a, const float* __restrict__ b,
                                                           const float* __restrict__ c0, float* __restrict__ dx, int64_t NC,
                                                           int C, int64_t P) {
    for (int64_t row = blockIdx.y; row < NC; row += gridDim.y) {
        const int c = (int)(row % C);
        const float aa = __ldg(a + c), bb = __ldg(b + c), cc = __ldg(c0 + c);
        const float* gp = dy + row * P;
        const float* xp = x + row * P;
        float* op = dx + row * P;
        for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x)
            op[p] = fmaf(aa, gp[p], fmaf(bb, xp[p], cc));
    }
}
extern "C" int cwfa_bn_bwd_apply_f32(const float* dy, const float* x, const float* a, const float* b, const float* c0, float* dx,
                                     int N, int C, int64_t P, void* stream) {
    if (N <= 0 || C <= 0 || P <= 0) { set_error("bn_bwd_apply: bad shape"); return CWFA_EINVAL; }
    const int64_t NC = (int64_t)N * C;
    int gx = (int)((P + 255) / 256);
    if (gx > 64) gx = 64;
    dim3 grid(gx, NC < 4096 ? (int)NC : 4096);
    bn_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, x, a, b, c0, dx, NC, C, P);
    return check_launch("bn_bwd_apply");
}

// 2x2 max-pool adjoint: the gradient goes to the FIRST maximum of the window in row-major scan order (torch's rule)
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           float* __restrict__ dx, int64_t NC, int H2, int W2) {
    const int W = 2 * W2;
    for (int64_t row = blockIdx.y; row < NC * H2; row += gridDim.y) {
        const int64_t nc = row / H2;
        const int h2 = (int)(row % H2);
        const float* xp = x + (nc * 2 * H2 + 2 * h2) * W;
        float* op = dx + (nc * 2 * H2 + 2 * h2) * W;
        const float* gp = dy + row * W2;
        for (int w2 = blockIdx.x * blockDim.x + threadIdx.x; w2 < W2; w2 += gridDim.x * blockDim.x) {
            const float v00 = xp[2 * w2], v01 = xp[2 * w2 + 1], v10 = xp[W + 2 * w2], v11 = xp[W + 2 * w2 + 1];
            int k = 0;
            float m = v00;
            if (v01 > m) { m = v01; k = 1; }
            if (v10 > m) { m = v10; k = 2; }
            if (v11 > m) { m = v11; k = 3; }
            const float g = gp[w2];
            op[2 * w2] = k == 0 ? g : 0.f;
            op[2 * w2 + 1] = k == 1 ? g : 0.f;
            op[W + 2 * w2] = k == 2 ? g : 0.f;
            op[W + 2 * w2 + 1] = k == 3 ? g : 0.f;
        }
    }
}
extern "C" int cwfa_maxpool2_bwd_f32(const float* x, const float* dy, float* dx, int N, int C, int H, int W, void* stream) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) { set_error("maxpool2_bwd: bad shape"); return CWFA_EINVAL; }
    const int64_t rows = (int64_t)N * C * (H / 2);
    dim3 grid(ceil_div(W / 2, 256), rows < 65535 ? (int)rows : 65535);
    maxpool2_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, dy, dx, (int64_t)N * C, H / 2, W / 2);
    return check_launch("maxpool2_bwd");
}

// pixel shuffle r = 2: y[n,c,2h+i,2w+j] = z[n,4c+2i+j,h,w] (+ skip);  inverse: z = unshuffle(y)
__global__ void __launch_bounds__(256) pixel_shuffle2_kernel(const float* __restrict__ src, const float* __restrict__ skip,
                                                             float* __restrict__ dst, int64_t NC, int H, int W, int inverse) {
    // one thread per LOW-resolution pixel (h,w) of one (n,c): moves the 2x2 block
    const int64_t P = (int64_t)H * W;
    for (int64_t row = blockIdx.y; row < NC * H; row += gridDim.y) {
        const int64_t nc = row / H;
        const int h = (int)(row % H);
        for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int64_t lo = (nc * 4 + 2 * i + j) * P + (int64_t)h * W + w;
                    const int64_t hi = (nc * 2 * H + 2 * h + i) * (2 * (int64_t)W) + 2 * w + j;
                    if (inverse) dst[lo] = src[hi];
                    else dst[hi] = src[lo] + (skip ? skip[hi] : 0.f);
                }
        }
    }
}
extern "C" int cwfa_pixel_shuffle2_f32(const float* src, const float* skip, float* dst, int N, int C, int H, int W, int inverse,
                                       void* stream) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) { set_error("pixel_shuffle2: bad shape"); return CWFA_EINVAL; }
    const int64_t rows = (int64_t)N * C * H;
    dim3 grid(ceil_div(W, 256), rows < 65535 ? (int)rows : 65535);
    pixel_shuffle2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, skip, dst, (int64_t)N * C, H, W, inverse);
    return check_launch("pixel_shuffle2");
}

// ------------------------------------------------------------------------------------------
// LRNN mean-volume branch adjoints (networks.py:244-262, 486-503, 554): activation adjoints from the layer OUTPUT,
// GELU + skip, LayerNorm([C,H,W]) with element-wise affine, attention gate
// ------------------------------------------------------------------------------------------
// kind: CWFA_ACT_ELU dv = dy*(y>0 ? 1 : y+1); CWFA_ACT_RELU dv = y>0 ? dy : 0; CWFA_ACT_SIGMOID dv = dy*y*(1-y)
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                      float* __restrict__ dv, int64_t n, int kind) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float yy = __ldg(y + i), g = __ldg(dy + i);
        float d;
        if (kind == CWFA_ACT_ELU) d = yy > 0.f ? 1.f : yy + 1.f;
        else if (kind == CWFA_ACT_RELU) d = yy > 0.f ? 1.f : 0.f;
        else d = yy * (1.f - yy);
        dv[i] = g * d;
    }
}
extern "C" int cwfa_act_bwd_f32(const float* dy, const float* y, float* dv, int64_t n, int kind, void* stream) {
    if (n <= 0 || (kind != CWFA_ACT_ELU && kind != CWFA_ACT_RELU && kind != CWFA_ACT_SIGMOID)) { set_error("act_bwd: bad args"); return CWFA_EINVAL; }
    act_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(dy, y, dv, n, kind);
    return check_launch("act_bwd");
}

// forward: y = gelu(v) + r (exact erf GELU, networks.py:492,503); backward (dy != NULL): dv = dy * (Phi(v) + v phi(v))
__global__ void __launch_bounds__(256) gelu_add_kernel(const float* __restrict__ v, const float* __restrict__ r,
                                                       const float* __restrict__ dy, float* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = v[i];
        const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
        if (dy) out[i] = dy[i] * (cdf + x * 0.39894228040143267794f * expf(-0.5f * x * x));
        else out[i] = x * cdf + (r ? r[i] : 0.f);
    }
}
extern "C" int cwfa_gelu_add_f32(const float* v, const float* r, const float* dy, float* out, int64_t n, void* stream) {
    if (n <= 0 || !v || !out) { set_error("gelu_add: bad args"); return CWFA_EINVAL; }
    gelu_add_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(v, r, dy, out, n);
    return check_launch("gelu_add");
}

// LayerNorm adjoint, stage 1: per sample (sum g, sum g*x) with g = dy*gamma (two-stage, kDotBlocks blocks per sample)
__global__ void __launch_bounds__(256) ln_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float* __restrict__ gamma, float* __restrict__ ws, int64_t n) {
    const int b = blockIdx.y;
    const float* xp = x + (int64_t)b * n;
    const float* gp = dy + (int64_t)b * n;
    float s = 0.f, q = 0.f;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const float g = gp[p] * __ldg(gamma + p);
        s += g;
        q = fmaf(g, xp[p], q);
    }
    block_sum2(s, q);
    if (threadIdx.x == 0) {
        ws[((int64_t)b * gridDim.x + blockIdx.x) * 2 + 0] = s;
        ws[((int64_t)b * gridDim.x + blockIdx.x) * 2 + 1] = q;
    }
}
extern "C" int cwfa_ln_bwd_stats_f32(const float* x, const float* dy, const float* gamma, float* out, float* workspace, int B,
                                     int64_t n, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || B > 65535 || n <= 0) { set_error("ln_bwd_stats: bad shape"); return CWFA_EINVAL; }
    ln_bwd_stats_kernel<<<dim3(kDotBlocks, B), 256, 0, st>>>(x, dy, gamma, workspace, n);
    int rc = check_launch("ln_bwd_stats");
    if (rc) return rc;
    channel_dot_finalize_kernel<<<ceil_div(B, 128), 128, 0, st>>>(workspace, out, B, kDotBlocks);     // out[b] = sum g, out[B+b] = sum g*x
    return check_launch("ln_bwd_stats_finalize");
}
// stage 2: coef[b] = (mean, rstd, mean(g), mean(g*xhat));  dx = rstd*(gamma*dy - mean(g) - xhat*mean(g*xhat));
// dgamma[p] = sum_b dy*xhat, dbeta[p] = sum_b dy  (a thread owns one element position p and walks the batch)
__global__ void __launch_bounds__(256) ln_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float* __restrict__ gamma, const float* __restrict__ coef,
                                                           float* __restrict__ dx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int B, int64_t n) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const float gm = __ldg(gamma + p);
        float dg = 0.f, db = 0.f;
        for (int b = 0; b < B; ++b) {
            const float mean = __ldg(coef + 4 * b), rstd = __ldg(coef + 4 * b + 1), mg = __ldg(coef + 4 * b + 2), mgx = __ldg(coef + 4 * b + 3);
            const float g = dy[(int64_t)b * n + p];
            const float xh = (x[(int64_t)b * n + p] - mean) * rstd;
            if (dx) dx[(int64_t)b * n + p] = rstd * (gm * g - mg - xh * mgx);
            dg = fmaf(g, xh, dg);
            db += g;
        }
        if (dgamma) dgamma[p] = dg;
        if (dbeta) dbeta[p] = db;
    }
}
extern "C" int cwfa_ln_bwd_apply_f32(const float* x, const float* dy, const float* gamma, const float* coef, float* dx,
                                     float* dgamma, float* dbeta, int B, int64_t n, void* stream) {
    if (B <= 0 || n <= 0) { set_error("ln_bwd_apply: bad shape"); return CWFA_EINVAL; }
    ln_bwd_apply_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, dy, gamma, coef, dx, dgamma, dbeta, B, n);
    return check_launch("ln_bwd_apply");
}

// attention gate y = x + m*2*(g-0.5) (networks.py:554): forward when dy == NULL (out0 = y), else the adjoint
// out0 = dm = dy*2*(g-0.5), out1 = dg = dy*2*m   (dx = dy needs no kernel)
__global__ void __launch_bounds__(256) gate_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                   const float* __restrict__ g, const float* __restrict__ dy,
                                                   float* __restrict__ out0, float* __restrict__ out1, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gg = 2.f * (g[i] - 0.5f);
        if (dy) {
            const float d = dy[i];
            if (out0) out0[i] = d * gg;
            if (out1) out1[i] = d * 2.f * m[i];
        } else {
            out0[i] = fmaf(m[i], gg, x[i]);
        }
    }
}
extern "C" int cwfa_gate_f32(const float* x, const float* m, const float* g, const float* dy, float* out0, float* out1, int64_t n,
                             void* stream) {
    if (n <= 0 || !m || !g || (!dy && (!x || !out0))) { set_error("gate: bad args"); return CWFA_EINVAL; }
    gate_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, m, g, dy, out0, out1, n);
    return check_launch("gate");
}

// ------------------------------------------------------------------------------------------
// Cotangent preparation of a tensor-core convolution's backward pass, ONE pass over dy (autograd.py:_Conv2dTC.backward):
//   g = dy * ELU'(y) (if y != NULL)  ->  g8 (C8 half layout for the data / weight gradient MMAs), optionally g as fp32 NCHW
//   (residual branch / fp32 weight-gradient path), and the per-channel sums of g = the bias gradient (deterministic: one
//   partial per block, fixed-order final sum).  Replaces elu_bwd + nchw_to_c8 + channel_stats (three passes, 97 us -> 35 us at
//   64 channels x 512 x 512).
constexpr int kPrepBlocks = 256;                // pixel blocks per (sample, 8-channel chunk)
template <int VEC> struct PrepVec;
template <> struct PrepVec<1> { using T = float; };
template <> struct PrepVec<2> { using T = float2; };
template <bool BF16, int VEC>
__global__ void __launch_bounds__(256, 3) dy_prep_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                         uint4* __restrict__ g8, float* __restrict__ g32, float* __restrict__ ws,
                                                         int C, int chunks, int64_t P) {
    using VT = typename PrepVec<VEC>::T;
    const int n = blockIdx.y / chunks, ch = blockIdx.y - n * chunks;
    const int64_t Pv = P / VEC;
    const int nc = min(8, C - ch * 8);                          // real channels of this chunk (<= 0: the whole chunk is padding)
    const int jclamp = nc > 0 ? 0 : (C - 1) - ch * 8;            // a plane inside the tensor for the discarded loads
    const int64_t plane0 = ((int64_t)n * C + ch * 8) * P;
    float sum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Pv; q += (int64_t)gridDim.x * blockDim.x) {
        // all loads first (predicated, no control flow in between): 8 (+ 8) independent requests in flight per thread
        VT d[8], yy[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int jj = j < nc ? j : jclamp;                 // padded channels re-read a real plane (value discarded below)
            d[j] = __ldg(reinterpret_cast<const VT*>(dy + plane0 + (int64_t)jj * P) + q);
        }
        if (y) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int jj = j < nc ? j : jclamp;
                yy[j] = __ldg(reinterpret_cast<const VT*>(y + plane0 + (int64_t)jj * P) + q);
            }
        }
        float v[8][VEC];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float dk[VEC], yk[VEC];
            if constexpr (VEC == 2) { dk[0] = d[j].x; dk[1] = d[j].y; yk[0] = yy[j].x; yk[1] = yy[j].y; }
            else { dk[0] = d[j]; yk[0] = yy[j]; }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float g = dk[k];
                if (y) g *= yk[k] > 0.f ? 1.f : yk[k] + 1.f;
                v[j][k] = j < nc ? g : 0.f;
                sum[j] += v[j][k];
            }
        }
        if (g32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nc) {
                    if constexpr (VEC == 2) *reinterpret_cast<float2*>(g32 + plane0 + (int64_t)j * P + q * 2) = make_float2(v[j][0], v[j][1]);
                    else g32[plane0 + (int64_t)j * P + q] = v[j][0];
                }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            uint4 o;
            o.x = cwfa::tcx::pack2<BF16>(v[0][k], v[1][k]); o.y = cwfa::tcx::pack2<BF16>(v[2][k], v[3][k]);
            o.z = cwfa::tcx::pack2<BF16>(v[4][k], v[5][k]); o.w = cwfa::tcx::pack2<BF16>(v[6][k], v[7][k]);
            g8[((int64_t)n * chunks + ch) * P + q * VEC + k] = o;
        }
    }
    if (ws) {
        __shared__ float red[8][8];
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float s = warp_sum(sum[j]);
            if (lane == 0) red[w][j] = s;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            float s = 0.f;
            for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
            ws[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = s;
        }
    }
}
// one block per 8-channel chunk: thread (lane = tid / 8, j = tid % 8) sums the partials of pixel blocks lane, lane + 32, ... over
// the samples, then a fixed-order tree over the 32 lanes (deterministic)
__global__ void __launch_bounds__(256) dy_prep_finalize_kernel(const float* __restrict__ ws, float* __restrict__ db, int N, int C,
                                                               int chunks, int nblk) {
    __shared__ double red[32][8];
    const int ch = blockIdx.x, j = threadIdx.x & 7, lane = threadIdx.x >> 3;
    double s = 0.0;
    for (int n = 0; n < N; ++n)
        for (int b = lane; b < nblk; b += 32) s += (double)__ldg(ws + (((int64_t)n * chunks + ch) * nblk + b) * 8 + j);
    red[lane][j] = s;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) {
        if (lane < o) red[lane][j] += red[lane + o][j];
        __syncthreads();
    }
    const int c = ch * 8 + j;
    if (lane == 0 && c < C) db[c] = (float)red[0][j];
}
extern "C" int cwfa_dy_prep_workspace_floats(int N, int Cp) { return N * (Cp / 8) * kPrepBlocks * 8; }
extern "C" int cwfa_dy_prep(const float* dy, const float* y, void* g8, float* g32, float* db, float* workspace, int N, int C,
                            int Cp, int64_t P, int is_bf16, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!dy || !g8 || N <= 0 || C <= 0 || Cp < C || (Cp % 8) || P <= 0 || (db && !workspace) || (int64_t)N * (Cp / 8) > 65535) {
        set_error("dy_prep: bad arguments");
        return CWFA_EINVAL;
    }
    const int chunks = Cp / 8;
    // two pixels per thread: 8-byte loads, and the thread's two 16-byte C8 stores fill one 32-byte sector
    const bool vec = (P % 2 == 0) && aligned16(dy) && (!y || aligned16(y)) && (!g32 || aligned16(g32));   // 8-byte accesses: plane starts stay 8-byte aligned
    dim3 grid(kPrepBlocks, N * chunks);
    float* ws = db ? workspace : nullptr;
    if (is_bf16) {
        if (vec) dy_prep_kernel<true, 2><<<grid, 256, 0, st>>>(dy, y, (uint4*)g8, g32, ws, C, chunks, P);
        else dy_prep_kernel<true, 1><<<grid, 256, 0, st>>>(dy, y, (uint4*)g8, g32, ws, C, chunks, P);
    } else {
        if (vec) dy_prep_kernel<false, 2><<<grid, 256, 0, st>>>(dy, y, (uint4*)g8, g32, ws, C, chunks, P);
        else dy_prep_kernel<false, 1><<<grid, 256, 0, st>>>(dy, y, (uint4*)g8, g32, ws, C, chunks, P);
    }
    int rc = check_launch("dy_prep");
    if (rc || !db) return rc;
    dy_prep_finalize_kernel<<<chunks, 256, 0, st>>>(workspace, db, N, C, chunks, kPrepBlocks);
    return check_launch("dy_prep_finalize");
}
