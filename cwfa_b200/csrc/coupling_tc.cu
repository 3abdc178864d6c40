// Persistent last-conv + affine-coupling kernel of a coupling sub-network (networks.py:635-638 +
// FrEIA/modules/coupling_layers.py:490-500 + the preceding permutation as a gather):
//     a = W3x3 (*) b + bias;  s = clamp*0.636*atan(a[:, :ch]);  t = a[:, ch:]  (or t = t_scale * t_ext)
//     fwd: y = exp(s)*x[gather] + t        inv: y = (x[gather] - t)*exp(-s)        logdet += +-sum s, sumsq = sum y^2
// One persistent CTA per SM walks 16x16-pixel tiles: the 3x3 weights (64 -> Cout_p <= 96) stay resident in shared
// memory, the accumulators are double buffered in TMEM so the MMAs of tile i+1 run under the epilogue of tile i, and
// SIXTEEN epilogue warps (4 per TMEM lane quadrant) share each tile: the coupling epilogue is a latency chain per
// thread (gather -> TMEM -> atan/exp -> store), so halving the chain length per thread matters more than anything.
// Warp roles (576 threads): warp0 TMA producer, warp1 MMA issuer, warps 2-17 epilogue.
#include "tc_common.cuh"
using namespace cwfa;
using namespace cwfa::tcx;

namespace {
constexpr int kChunks = 8;                                   // 64 hidden channels = 8 chunks
constexpr int kTH = 16, kTW = 16, kBH = 18, kBW = 18;
constexpr uint32_t kA1Bytes = kChunks * kBH * kBW * 16;      // 41472
constexpr int kMaxBN = 96;
constexpr uint32_t kHeader = 2048;
constexpr uint32_t kOffW = kHeader;       // weights (9*8*BN*16 bytes), then the A ring (2..4 stages, sized by the launcher)
constexpr int kMaxAStages = 4;   // barrier slots; the launcher uses at most 3 stages (measured: 2..4 are within noise)
constexpr int kThreads = 576, kEpiThreads = 512;
constexpr int kMaxG = 3;                                      // 8-channel groups per epilogue thread (ch <= 48, 2 M-blocks, 4-way split)

struct CpParams {
    int N, H, W, tiles_x, tiles_y, num_tiles;
    int BN, ch, axis, a_stages, in_chunk_off;
    uint32_t off_a;
    const uint8_t* w;            // packed [9][8][BN][8]
    const float* bias;           // BN floats or NULL
    const float* x;              // (N,ch,H,W) or NULL
    float* y;
    const float* t_ext;          // external shift or NULL
    const int* perm;
    float* ws;                   // [num_tiles][16][2]
    float kk, tscale;
    float* logdet;               // in-kernel finalize (ticket != NULL): logdet[n] (+)= sum s, sumsq[n] = sum y^2
    float* sumsq;
    int* ticket;                 // zero on entry; the last CTA to finish reduces the partials in a fixed order and resets it
    int accumulate;
};

template <bool BF16, bool INV, bool EXT>
__global__ void __launch_bounds__(kThreads, 1) coupling_tc_kernel(const __grid_constant__ CUtensorMap tmap, const CpParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t s0 = smem_u32(smem);
    const uint32_t w_full = s0;
    auto a_full = [&](int b) { return s0 + 8u * (1 + b); };          // up to 4
    auto a_empty = [&](int b) { return s0 + 8u * (5 + b); };         // up to 4
    auto acc_full = [&](int b) { return s0 + 8u * (9 + b); };
    auto acc_empty = [&](int b) { return s0 + 8u * (11 + b); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
    // smem + 256: red[2][16][2] floats (256 B); + 512: s_perm[64] ints; + 1024: s-bias[96]; + 1536: t-bias[64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = p.ch;
    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (int b = 0; b < kMaxAStages; ++b) {
            mbar_init(a_full(b), 1);
            mbar_init(a_empty(b), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full(b), 1);
            mbar_init(acc_empty(b), kEpiThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 128) {
        const int i = threadIdx.x - 64;
        if (i < kMaxBN) reinterpret_cast<float*>(smem + 1024)[i] = (p.bias && i < p.BN) ? __ldg(p.bias + i) : 0.f;
        if (i < 64) {
            reinterpret_cast<int*>(smem + 512)[i] = (i < ch && p.perm && p.axis == 1) ? __ldg(p.perm + i) : i;
            reinterpret_cast<float*>(smem + 1536)[i] = (!EXT && p.bias && i < ch) ? __ldg(p.bias + ch + i) : 0.f;
        }
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const size_t plane = (size_t)p.H * p.W;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t wbytes = 9u * kChunks * p.BN * 16u;
            mbar_expect_tx(w_full, wbytes);
            bulk_load(s0 + kOffW, p.w, wbytes, w_full);
            int sa = 0;
            uint32_t pa = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int t = blockIdx.x + i * gridDim.x;
                const int n = t / tiles_per_img, r = t % tiles_per_img;
                const int h0 = (r / p.tiles_x) * kTH, w0 = (r % p.tiles_x) * kTW;
                mbar_wait(a_empty(sa), pa ^ 1);
                mbar_expect_tx(a_full(sa), kA1Bytes);
                tma_load_4d(s0 + p.off_a + sa * kA1Bytes, &tmap, a_full(sa), (w0 - 1) * 8, h0 - 1, p.in_chunk_off, n);
                if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = idesc_f16(p.BN, BF16 ? 1 : 0);
        constexpr uint32_t a_lbo = kBH * kBW * 16, a_sbo = kBW * 16;
        const uint32_t w_lbo = p.BN * 16, tap_units = (kChunks * p.BN * 16) >> 4;
        const uint32_t a_hi = desc_hi(a_sbo), w_hi = desc_hi(128);
        const uint32_t w_lo0 = desc_lo(s0 + kOffW, w_lbo);
        const uint32_t leader = elect_one();
        mbar_wait(w_full, 0);
        int sa = 0;
        uint32_t pa = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1, ph = (i >> 1) & 1;
            mbar_wait(a_full(sa), pa);
            mbar_wait(acc_empty(b), ph ^ 1);
            tc_fence_after();
            if (leader) {
                const uint32_t a_base = s0 + p.off_a + sa * kA1Bytes;
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap) {
                    const int kh = tap / 3, kw = tap - kh * 3;
                    const uint32_t w_lo = w_lo0 + tap * tap_units;
                    const uint32_t a_lo0 = desc_lo(a_base + (uint32_t)((kh * kBW + kw) * 16), a_lbo);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                        for (int mb = 0; mb < 2; ++mb)
                            tc_mma_f16_split(tmem + b * 256 + mb * p.BN, a_lo0 + mb * 8 + kk * ((2 * a_lbo) >> 4), a_hi,
                                             w_lo + kk * ((2 * w_lbo) >> 4), w_hi, idesc, (tap | kk) ? 1u : 0u);
                    }
                }
                tc_commit(a_empty(sa));
                tc_commit(acc_full(b));
            }
            __syncwarp();
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
    } else {
        // ============================ epilogue: 16 warps ============================
        const int q = warp & 3;
        const int sub = (warp - 2) >> 2;             // 0..3
        const int m = q * 32 + lane;
        const int gpc = (ch + 7) >> 3;
        const int ngroups = 2 * gpc;
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1, ph = (i >> 1) & 1;
            const int t = blockIdx.x + i * gridDim.x;
            const int n = t / tiles_per_img, rr = t % tiles_per_img;
            const int h0 = (rr / p.tiles_x) * kTH, w0 = (rr % p.tiles_x) * kTW;
            const int orow = h0 + (m >> 3);
            const bool row_ok = orow < p.H;
            // gather the coupling inputs of all of this thread's groups BEFORE the accumulator is ready
            float xv[kMaxG][8], tx[kMaxG][8];
#pragma unroll
            for (int k = 0; k < kMaxG; ++k) {
                const int g = sub + 4 * k;
                const int mb = g >= gpc ? 1 : 0, c0 = (g - (mb ? gpc : 0)) << 3;      // two M-blocks: no division
                const int ocol = w0 + mb * 8 + (m & 7);
                const bool ok = g < ngroups && row_ok && ocol < p.W;
                int srow = orow, scol = ocol;
                if (ok && p.perm && p.axis == 2) srow = __ldg(p.perm + orow);
                if (ok && p.perm && p.axis == 3) scol = __ldg(p.perm + ocol);
                // 32-bit offsets inside one sample (ch * H * W < 2^31 is checked by the launcher), 64-bit base per sample
                const float* xs = p.x ? p.x + (size_t)n * ch * plane : nullptr;
                const float* ts = EXT ? p.t_ext + (size_t)n * ch * plane : nullptr;
                const uint32_t plane32 = (uint32_t)plane;
                const uint32_t spix = (uint32_t)srow * (uint32_t)p.W + (uint32_t)scol;
                const uint32_t opix = (uint32_t)orow * (uint32_t)p.W + (uint32_t)ocol;
                const bool okx = ok && xs != nullptr;
                // the 8 source channels of this group: two 16-byte shared loads (identity / permutation table, 64 entries)
                const float4 pl = lds128(s0 + 512u + 4u * (uint32_t)c0), ph4 = lds128(s0 + 512u + 4u * (uint32_t)c0 + 16u);
                const uint32_t src_c[8] = {__float_as_uint(pl.x), __float_as_uint(pl.y), __float_as_uint(pl.z), __float_as_uint(pl.w),
                                           __float_as_uint(ph4.x), __float_as_uint(ph4.y), __float_as_uint(ph4.z), __float_as_uint(ph4.w)};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + j;
                    xv[k][j] = (okx && c < ch) ? __ldg(xs + (src_c[j] * plane32 + spix)) : 0.f;
                    if constexpr (EXT) tx[k][j] = (ok && c < ch) ? __ldg(ts + ((uint32_t)c * plane32 + opix)) : 0.f;
                }
            }
            mbar_wait(acc_full(b), ph);
            tc_fence_after();
            float sum_s = 0.f, sum_q = 0.f;
#pragma unroll
            for (int k = 0; k < kMaxG; ++k) {
                const int g = sub + 4 * k;
                if (g < ngroups) {                          // warp-uniform
                    const int mb = g >= gpc ? 1 : 0, c0 = (g - (mb ? gpc : 0)) << 3;      // two M-blocks: no division
                    const int ocol = w0 + mb * 8 + (m & 7);
                    const bool ok = row_ok && ocol < p.W;
                    uint32_t rs[8], rt[8];
                    __syncwarp();
                    const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256 + mb * p.BN + c0);
                    tmem_ld8_nowait(ta, rs);
                    if constexpr (!EXT) tmem_ld8_nowait(ta + ch, rt);
                    tmem_ld_wait();
                    if (ok) {
                        const float4 a0 = lds128(s0 + 1024u + 4u * c0), a1 = lds128(s0 + 1024u + 4u * c0 + 16u);
                        const float bs[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float4 t0 = lds128(s0 + 1536u + 4u * c0), t1 = lds128(s0 + 1536u + 4u * c0 + 16u);
                        const float bt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                        float* yp = p.y + (size_t)n * ch * plane + ((uint32_t)c0 * (uint32_t)plane + (uint32_t)orow * (uint32_t)p.W + (uint32_t)ocol);
                        const uint32_t plane32 = (uint32_t)plane;
                        // straight-line over the 8 channels of the group (8 independent dependency chains in flight);
                        // only the store and the sums are predicated on the channel being real
                        float sv[8], yv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) sv[j] = p.kk * atan_fast(__uint_as_float(rs[j]) + bs[j]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float tv;
                            if constexpr (EXT) tv = p.tscale * tx[k][j];
                            else tv = __uint_as_float(rt[j]) + bt[j];
                            if constexpr (INV) yv[j] = (xv[k][j] - tv) * exp_fast(-sv[j]);
                            else yv[j] = fmaf(exp_fast(sv[j]), xv[k][j], tv);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (c0 + j < ch) {
                                yp[(uint32_t)j * plane32] = yv[j];
                                sum_s += sv[j];
                                sum_q = fmaf(yv[j], yv[j], sum_q);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty(b));
            // one partial per (tile, epilogue warp), summed in a fixed order by cwfa_coupling_finalize (reproducible);
            // no CTA-wide barrier: the 16 epilogue warps stay decoupled
            sum_s = warp_sum(sum_s);
            sum_q = warp_sum(sum_q);
            if (lane == 0) {
                float* w = p.ws + ((size_t)t * 16 + (warp - 2)) * 2;
                w[0] = INV ? -sum_s : sum_s;
                w[1] = sum_q;
            }
        }
        if (p.ticket) __threadfence();          // this warp's partials are visible device-wide before the CTA takes its ticket
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
    if (p.ticket == nullptr) return;
    // ---- in-kernel finalize: the LAST CTA to finish sums all (tile, warp) partials of every sample in a FIXED order
    // (thread <-> partial assignment and tree shape do not depend on which CTA is last => bit-reproducible), so the
    // separate cwfa_coupling_finalize launch disappears.
    // (scratch lives in the dynamic header: the kernel already uses the full 227 KB carve-out, so no static shared memory)
    volatile int* s_last = reinterpret_cast<volatile int*>(smem + 132);
    double* s_red = reinterpret_cast<double*>(smem + 256);          // [16 warps][2]
    if (threadIdx.x == 0) *s_last = (atomicAdd(p.ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    const int per_img = tiles_per_img * 16;
    constexpr int kRed = 512;                                       // reducing threads (16 warps)
    for (int n = 0; n < p.N; ++n) {
        double s = 0.0, q = 0.0;
        const float2* src = reinterpret_cast<const float2*>(p.ws) + (size_t)n * per_img;
        if (threadIdx.x < kRed) {
            for (int i = threadIdx.x; i < per_img; i += kRed) {
                const float2 v = __ldcg(src + i);
                s += (double)v.x;
                q += (double)v.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if (lane == 0) { s_red[warp * 2] = s; s_red[warp * 2 + 1] = q; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0;
            for (int k = 0; k < kRed / 32; ++k) { a += s_red[k * 2]; b += s_red[k * 2 + 1]; }
            p.logdet[n] = (p.accumulate ? p.logdet[n] : 0.f) + (float)a;
            if (p.sumsq) p.sumsq[n] = (float)b;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *p.ticket = 0;        // ready for the next launch / graph replay that reuses this buffer
}
}  // namespace

// Partial sums per sample that cwfa_coupling_tc writes: one per (16x16-pixel tile, epilogue warp).
extern "C" int cwfa_coupling_tc_tiles(int H, int W) { return ceil_div(W, kTW) * ceil_div(H, kTH) * 16; }

// Same contract as cwfa_conv_tc_coupling, for the CWFA sub-network shape: 3x3 conv from 64 hidden channels to
// Cout_p <= 96 (= one N block), ch <= 48.  workspace: 2 * N * cwfa_coupling_tc_tiles(H, W) floats.  With ticket != NULL
// (one int32, zero on entry, left zero on exit) the kernel itself reduces them into logdet / sumsq (last-CTA finalize, fixed
// order); with ticket == NULL reduce with cwfa_coupling_finalize(..., tiles = cwfa_coupling_tc_tiles(H, W), ...).
extern "C" int cwfa_coupling_tc(const void* b_c8, const void* w_packed, const float* bias, int N, int H, int W, int Cout,
                                int Cout_p, const float* cx, float* cy, const float* ct, float t_scale, const int32_t* perm,
                                int perm_axis, int ch, float clamp, float k_atan, int inverse, float* workspace, float* logdet,
                                float* sumsq, int accumulate, int32_t* ticket, int in_total_chunks, int in_chunk_off, int is_bf16,
                                void* stream) {
    if (N <= 0 || H <= 0 || W <= 0 || (int64_t)ch * H * W >= (1ll << 31) || !cy || !workspace || ch <= 0 || ch > 48 || Cout_p > kMaxBN || (Cout_p % 16) ||
        (ct ? Cout < ch : Cout < 2 * ch) || (perm && (perm_axis < 1 || perm_axis > 3)) || (!cx && !inverse) || (ticket && !logdet) ||
        in_chunk_off < 0 || in_chunk_off + kChunks > in_total_chunks) {
        set_error("coupling_tc: unsupported arguments (needs 64 -> Cout_p <= 96, ch <= 48)");
        return CWFA_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(b_c8) & 15) || (reinterpret_cast<uintptr_t>(w_packed) & 15)) {
        set_error("coupling_tc: pointers must be 16-byte aligned");
        return CWFA_EINVAL;
    }
    CpParams p{};
    p.N = N; p.H = H; p.W = W;
    p.tiles_x = ceil_div(W, kTW); p.tiles_y = ceil_div(H, kTH);
    const int64_t nt = (int64_t)p.tiles_x * p.tiles_y * N;
    if (nt > 0x7fffffff) { set_error("coupling_tc: too many tiles"); return CWFA_EINVAL; }
    p.num_tiles = (int)nt;
    p.BN = Cout_p; p.ch = ch; p.axis = perm_axis;
    p.w = (const uint8_t*)w_packed; p.bias = bias; p.x = cx; p.y = cy; p.t_ext = ct; p.perm = perm; p.ws = workspace;
    p.kk = clamp * k_atan; p.tscale = t_scale;
    p.logdet = logdet; p.sumsq = sumsq; p.ticket = ticket; p.accumulate = accumulate;
    CUtensorMap tmap;
    p.in_chunk_off = in_chunk_off;
    int rc = make_c8_tensor_map(&tmap, b_c8, N, in_total_chunks, H, W, kBW, kBH, kChunks, is_bf16);
    if (rc) return rc;
    typedef void (*KernT)(const CUtensorMap, const CpParams);
    static const KernT table[2][4] = {
        {coupling_tc_kernel<false, false, false>, coupling_tc_kernel<false, true, false>, coupling_tc_kernel<false, false, true>, coupling_tc_kernel<false, true, true>},
        {coupling_tc_kernel<true, false, false>, coupling_tc_kernel<true, true, false>, coupling_tc_kernel<true, false, true>, coupling_tc_kernel<true, true, true>}};
    const int mode = (ct ? 2 : 0) + (inverse ? 1 : 0);
    KernT kern = table[is_bf16 ? 1 : 0][mode];
    static bool attr_done[8] = {false, false, false, false, false, false, false, false};
    const int ki = (is_bf16 ? 4 : 0) + mode;
    if (!attr_done[ki]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_done[ki] = true;
    }
    const uint32_t wbytes = 9u * kChunks * Cout_p * 16u;
    p.off_a = (kOffW + wbytes + 127u) & ~127u;
    int stages = (int)((227u * 1024u - 1024u - p.off_a) / kA1Bytes);
    p.a_stages = stages > 3 ? 3 : stages;          // 2 at Cout_p = 96, 3 at Cout_p <= 64
    const size_t smem_bytes = 1024 + p.off_a + (size_t)p.a_stages * kA1Bytes;
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    kern<<<grid, kThreads, smem_bytes, (cudaStream_t)stream>>>(tmap, p);
    return check_launch("coupling_tc");
}
