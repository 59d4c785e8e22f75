// Implicit-GEMM conv2d / conv-transpose2d on tcgen05 (NCHW fp32 in/out, fp32-class accuracy).
//
// Replaces cuDNN in the encoders' heavy layers (reference: lib/net/FBNet.py:216-319 GlobalGenerator /
// ResnetBlock, lib/net/HGFilters.py + lib/net/net_util.py:258-280 ConvBlock) whenever Cin % 64 == 0.
//
//   D[128 pixels][NT channels] += A[128 x 64] * B[NT x 64]^T   per 64-wide K-chunk, K = taps * Cin,
//   k = tap * Cin + ci, so a chunk is one filter tap x 64 consecutive input channels.
//
// Same numerics as the occupancy MLP (mlp_tc.cu): activations and weights are split x = hi + lo in fp16
// and every k-step issues hi*Whi + hi*Wlo + lo*Whi with fp32 accumulation in tensor memory.
//   warp 0      weight producer: bulk copies of host-packed K-major SWIZZLE_128B tiles (hi | lo), 2-stage ring
//   warp 1      MMA issuer (tcgen05.mma.cta_group::1.kind::f16, M = 128, N = NT, A operand in TMEM)
//   warps 2-5   one output pixel (= TMEM lane) per thread: im2col gather of the chunk's 64 input values
//               (zero / reflection padding, stride, transposed-conv index maps), hi/lo split, tcgen05.st into
//               a double-buffered A region; afterwards the epilogue (TMEM -> +bias/residual/activation -> NCHW).
// Small spatial extents (the 32 x 32 ResnetBlocks) are filled across SMs with split-K: partial sums go to a
// workspace and k_splitk_finish reduces them deterministically.
#include <cuda_fp16.h>

#include "common.cuh"

namespace icon {

struct ConvTcParams {
    const float *x;        // [N][Cin][H][W]
    const uint8_t *wt;     // packed tiles: [n_tile][chunk] x (hi NT*128 B | lo NT*128 B)
    const float *bias;     // [Cout] or null   (applied here only when splits == 1)
    const float *res;      // residual or null (splits == 1)
    float *y;              // splits == 1: [N][Cout][OH][OW]; else partial [splits][N][Cout][OH][OW]
    int N, Cin, H, W, Cout, OH, OW, KH, KW, stride, pad, reflect, transposed, act;
    int chunks_total, splits;
};

constexpr int CT_THREADS = 192;

namespace tc {
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.b32 %0, 1, 0, q;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q;\n}" ::"r"(d),
                 "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void ld32(uint32_t addr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st16(uint32_t addr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t *>(&h);
    lo = *reinterpret_cast<uint32_t *>(&l);
}
}  // namespace tc

// grid: (pixel tiles, channel tiles, splits).  TMEM: [0,NT) accumulator, [256,384) two A buffers (hi 32 | lo 32).
template <int NT>
__global__ void __launch_bounds__(CT_THREADS, 1) k_conv_tc(ConvTcParams p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = s32(smem_raw), base = (raw + 1023u) & ~1023u;     // 2 stages x (hi NT*128 | lo NT*128)
    constexpr uint32_t STAGE = NT * 256;
    enum { B_BFULL0 = 0, B_BFULL1, B_BEMPTY0, B_BEMPTY1, B_AFULL0, B_AFULL1, B_AEMPTY0, B_AEMPTY1 };
    __shared__ uint64_t bar_acc;
    auto BAR = [&](int i) { return s32(&bars[i]); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int chunks_per = (p.chunks_total + p.splits - 1) / p.splits;
    const int c_begin = blockIdx.z * chunks_per, c_end = min(p.chunks_total, c_begin + chunks_per);
    const int nchunks = max(0, c_end - c_begin);
    const int n0 = blockIdx.y * NT;

    if (tid == 0) {
        mbar_init(BAR(B_BFULL0), 1); mbar_init(BAR(B_BFULL1), 1);
        mbar_init(BAR(B_BEMPTY0), 1); mbar_init(BAR(B_BEMPTY1), 1);
        mbar_init(BAR(B_AFULL0), 128); mbar_init(BAR(B_AFULL1), 128);
        mbar_init(BAR(B_AEMPTY0), 1); mbar_init(BAR(B_AEMPTY1), 1);
        mbar_init(s32(&bar_acc), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint8_t *src = p.wt + ((size_t)blockIdx.y * p.chunks_total + c_begin) * STAGE;
            for (int c = 0; c < nchunks; ++c) {
                const uint32_t s = c & 1, ph = (c >> 1) & 1;
                mbar_wait(BAR(B_BEMPTY0 + s), ph ^ 1);
                mbar_expect_tx(BAR(B_BFULL0 + s), STAGE);
                bulk_g2s(base + s * STAGE, src + (size_t)c * STAGE, STAGE, BAR(B_BFULL0 + s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int c = 0; c < nchunks; ++c) {
                const uint32_t s = c & 1, ph = (c >> 1) & 1;
                mbar_wait(BAR(B_AFULL0 + s), ph);
                mbar_wait(BAR(B_BFULL0 + s), ph);
                fence_after();
                const uint32_t a_hi = tmem + 256 + 64u * s, a_lo = a_hi + 32;
                const uint64_t bh = desc_sw128(base + s * STAGE), bl = desc_sw128(base + s * STAGE + NT * 128);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    mma_ts(tmem, a_hi + 8 * ks, bh + 2 * ks, IDESC, (c | ks) != 0);
                    mma_ts(tmem, a_hi + 8 * ks, bl + 2 * ks, IDESC, 1);
                    mma_ts(tmem, a_lo + 8 * ks, bh + 2 * ks, IDESC, 1);
                }
                commit(BAR(B_BEMPTY0 + s));
                commit(BAR(B_AEMPTY0 + s));
            }
            commit(s32(&bar_acc));
        }
    } else {
        // ---------------------------------------------------- gather + epilogue: one output pixel per thread
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16);
        const int64_t npix = (int64_t)p.N * p.OH * p.OW;
        const int64_t gp = (int64_t)blockIdx.x * 128 + r;
        const bool pv = gp < npix;
        int pn = 0, oy = 0, ox = 0;
        if (pv) {
            pn = (int)(gp / ((int64_t)p.OH * p.OW));
            const int rem = (int)(gp % ((int64_t)p.OH * p.OW));
            oy = rem / p.OW; ox = rem % p.OW;
        }
        const size_t plane = (size_t)p.H * p.W;
        const float *xn = p.x + (size_t)pn * p.Cin * plane;
        const int cpt = p.Cin / 64;                          // chunks per tap
        for (int c = 0; c < nchunks; ++c) {
            const uint32_t s = c & 1, ph = (c >> 1) & 1;
            const int ck = c_begin + c, tap = ck / cpt, ci0 = (ck % cpt) * 64;
            const int kh = tap / p.KW, kw = tap % p.KW;
            int iy, ix;
            bool ok = pv;
            if (!p.transposed) {
                iy = oy * p.stride - p.pad + kh;
                ix = ox * p.stride - p.pad + kw;
                if (p.reflect) {
                    iy = iy < 0 ? -iy : (iy >= p.H ? 2 * p.H - 2 - iy : iy);
                    ix = ix < 0 ? -ix : (ix >= p.W ? 2 * p.W - 2 - ix : ix);
                } else ok = ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
            } else {
                const int ty2 = oy + p.pad - kh, tx2 = ox + p.pad - kw;
                ok = ok && ty2 >= 0 && tx2 >= 0 && (ty2 % p.stride) == 0 && (tx2 % p.stride) == 0;
                iy = ty2 / p.stride; ix = tx2 / p.stride;
                ok = ok && iy < p.H && ix < p.W;
            }
            uint32_t hi[32], lo[32];
            if (ok) {
                const float *src = xn + (size_t)ci0 * plane + (size_t)iy * p.W + ix;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float a = __ldg(src + (size_t)(2 * j) * plane), b = __ldg(src + (size_t)(2 * j + 1) * plane);
                    split2(a, b, hi[j], lo[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { hi[j] = 0u; lo[j] = 0u; }
            }
            mbar_wait(BAR(B_AEMPTY0 + s), ph ^ 1);
            fence_after();
            const uint32_t a0 = tl + 256 + 64u * s;
            st16(a0, reinterpret_cast<const uint32_t(&)[16]>(hi[0]));
            st16(a0 + 16, reinterpret_cast<const uint32_t(&)[16]>(hi[16]));
            st16(a0 + 32, reinterpret_cast<const uint32_t(&)[16]>(lo[0]));
            st16(a0 + 48, reinterpret_cast<const uint32_t(&)[16]>(lo[16]));
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_before();
            mbar_arrive(BAR(B_AFULL0 + s));
        }
        // ---- epilogue
        mbar_wait(s32(&bar_acc), 0);
        fence_after();
        const size_t ohw = (size_t)p.OH * p.OW;
        float *yb = p.y + (p.splits > 1 ? (size_t)blockIdx.z * p.N * p.Cout * ohw : 0);
        const bool fin = p.splits == 1;
        for (int cb = 0; cb < NT; cb += 32) {
            uint32_t acc[32];
            ld32(tl + cb, acc);
            if (pv) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int co = n0 + cb + j;
                    if (co < p.Cout) {
                        const size_t o = ((size_t)pn * p.Cout + co) * ohw + (size_t)oy * p.OW + ox;
                        float v = nchunks > 0 ? __uint_as_float(acc[j]) : 0.f;
                        if (fin) {
                            if (p.bias) v += __ldg(p.bias + co);
                            if (p.res) v += p.res[o];
                            if (p.act == 1) v = fmaxf(v, 0.f);
                            else if (p.act == 2) v = tanhf(v);
                        }
                        yb[o] = v;
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// y = act(sum_s partial[s] + bias + res)
__global__ void k_splitk_finish(const float *__restrict__ part, int splits, const float *__restrict__ bias,
                                const float *__restrict__ res, float *__restrict__ y, int64_t total, int Cout,
                                int64_t ohw, int act) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += part[(size_t)s * total + i];
    if (bias) v += __ldg(bias + (int)((i / ohw) % Cout));
    if (res) v += res[i];
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = tanhf(v);
    y[i] = v;
}

template <int NT>
static int launch_conv_tc(const ConvTcParams &p, dim3 grid, cudaStream_t stream) {
    static bool attr_set[ICON_MAX_DEVICES] = {};
    const int smem = 2 * NT * 256 + 1024;
    if (device_needs_setup(attr_set))
        ICON_CUDA(cudaFuncSetAttribute(k_conv_tc<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_conv_tc<NT><<<grid, CT_THREADS, smem, stream>>>(p);
    ICON_LAUNCHED();
    return ICON_OK;
}

}  // namespace icon

using namespace icon;

extern "C" size_t icon_conv2d_tc_workspace_bytes(int N, int Cout, int OH, int OW, int splits) {
    return splits > 1 ? (size_t)splits * N * Cout * OH * OW * sizeof(float) : 0;
}

extern "C" int icon_conv2d_tc(const float *x, const void *wt_packed, const float *bias, const float *res, float *y,
                              int N, int Cin, int H, int W, int Cout, int KH, int KW, int stride, int pad, int out_pad,
                              int reflect, int transposed, int act, int n_tile, int splits, void *ws, size_t ws_bytes,
                              icon_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ICON_CHECK_ARG(x && wt_packed && y && N > 0 && Cin > 0 && Cout > 0, "icon_conv2d_tc: bad argument");
    ICON_CHECK_ARG(Cin % 64 == 0, "icon_conv2d_tc: Cin=%d must be a multiple of 64 (use icon_conv2d)", Cin);
    ICON_CHECK_ARG(n_tile == 64 || n_tile == 128 || n_tile == 256, "icon_conv2d_tc: n_tile must be 64, 128 or 256");
    ICON_CHECK_ARG(splits >= 1, "icon_conv2d_tc: splits >= 1");
    ICON_CHECK_ARG(((uintptr_t)wt_packed & 15) == 0, "icon_conv2d_tc: packed weights must be 16-byte aligned");
    ConvTcParams p{};
    p.x = x; p.wt = (const uint8_t *)wt_packed; p.bias = bias; p.res = res;
    p.N = N; p.Cin = Cin; p.H = H; p.W = W; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
    p.reflect = reflect; p.transposed = transposed; p.act = act;
    if (!transposed) { p.OH = (H + 2 * pad - KH) / stride + 1; p.OW = (W + 2 * pad - KW) / stride + 1; }
    else { p.OH = (H - 1) * stride - 2 * pad + KH + out_pad; p.OW = (W - 1) * stride - 2 * pad + KW + out_pad; }
    ICON_CHECK_ARG(p.OH > 0 && p.OW > 0, "icon_conv2d_tc: empty output");
    ICON_CHECK_ARG(!reflect || (pad < H && pad < W && !transposed), "icon_conv2d_tc: bad reflection padding");
    p.chunks_total = KH * KW * (Cin / 64);
    p.splits = splits;
    const size_t need = icon_conv2d_tc_workspace_bytes(N, Cout, p.OH, p.OW, splits);
    if (ws_bytes < need) { set_error("icon_conv2d_tc: workspace %zu < %zu", ws_bytes, need); return ICON_ENOSPC; }
    p.y = splits > 1 ? (float *)ws : y;
    const int64_t npix = (int64_t)N * p.OH * p.OW;
    dim3 grid((unsigned)((npix + 127) / 128), (unsigned)((Cout + n_tile - 1) / n_tile), (unsigned)splits);
    int rc;
    if (n_tile == 256) rc = launch_conv_tc<256>(p, grid, stream);
    else if (n_tile == 128) rc = launch_conv_tc<128>(p, grid, stream);
    else rc = launch_conv_tc<64>(p, grid, stream);
    if (rc) return rc;
    if (splits > 1) {
        const int64_t total = (int64_t)N * Cout * p.OH * p.OW;
        k_splitk_finish<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const float *)ws, splits, bias, res, y, total,
                                                                             Cout, (int64_t)p.OH * p.OW, act);
        ICON_LAUNCHED();
    }
    return ICON_OK;
}
