// Encoder convolutions, Blackwell-first: implicit GEMM on tcgen05 with BOTH operands staged by the TMA engine.
//
// Replaces cuDNN in the encoders (reference: lib/net/FBNet.py:216-319 GlobalGenerator / ResnetBlock,
// lib/net/HGFilters.py:49-197 + lib/net/net_util.py:258-280 HGFilter / HourGlass / ConvBlock).
//
// Activations live in HBM as NHWC, already split x = hi + lo into two fp16 tensors by the kernel that produced them
// (k_act_nhwc: normalisation + ReLU + split in one pass).  A K-chunk of the implicit GEMM is one filter tap x 64
// input channels, i.e. for a tile of BH x BW output pixels a [BH][BW][64] box of the input shifted by the tap:
// exactly one 4-D tiled TMA load (cp.async.bulk.tensor, SWIZZLE_128B) per operand half -- zero padding is the TMA's
// out-of-bounds fill, reflection padding is a halo the producer wrote, stride 2 is a space-to-depth layout the
// producer wrote (4 parity planes), a transposed convolution is 4 output phases with 1/2/2/4 taps each.  The box
// lands in shared memory as 128 rows (pixels) x 128 bytes (64 fp16), 16-byte chunks XOR-swizzled by row % 8 --
// the K-major SWIZZLE_128B operand layout of tcgen05.mma, so the MMA reads it through a descriptor with no thread
// ever touching the data.  Weights: host-packed tiles in the same layout (1-D bulk copies).
//
//   D[128 pixels][NT channels] += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi      (3 kind::f16 MMAs per k-step, fp32 in TMEM)
//
//   warp 0      producer: 2 tensor-map loads (A hi, A lo) + 1 bulk copy (B hi|lo) per chunk, STAGES-deep ring
//   warp 1      TMEM allocation (NT columns only) + MMA issue (one lane)
//   warps 2-5   epilogue: TMEM -> registers -> (+bias) -> fp32 NHWC store (optionally into a channel slice of a
//               wider tensor = torch.cat for free) and per-(image, channel) sum / sum-of-squares for the
//               Instance/GroupNorm that follows (fp64 atomics), so the norm needs no statistics pass.
// Small spatial extents (the 32 x 32 ResnetBlocks) fill the machine through split-K (partials + k_splitk_nhwc).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace icon {

struct TapDesc { int8_t dy, dx; uint8_t plane, wtap; };
constexpr int MAX_TAPS = 49;

struct ConvNhwcParams {
    const uint8_t *wt;       // [n tiles][wt_chunks] x (hi NT*128 B | lo NT*128 B)
    const float *bias;       // [Cout] or null
    float *out;              // fp32 NHWC [N][OHf][OWf][Cs], this conv writes channels [co_off, co_off + Cout)
    float *partial;          // splits > 1: [splits][N][Ht][Wt][Cout]
    double *stats;           // [N][Cout][2] (sum, sum of squares) or null
    int N, Ht, Wt;           // logical output grid of this launch (per image)
    int BW, BH, tiles_x, tiles_y;
    int OHf, OWf, osy, osx, ooy, oox;     // out (y, x) = (a * osy + ooy, b * osx + oox)
    int Cs, co_off, Cout;
    int cpt;                 // 64-channel chunks per tap
    int ntaps, nplanes, wt_chunks, splits;
    TapDesc taps[MAX_TAPS];
};

constexpr int CN_THREADS = 192;

template <int NT, int STAGES>
__global__ void __launch_bounds__(CN_THREADS, NT == 64 ? 2 : 1)
k_conv_nhwc(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
            const __grid_constant__ ConvNhwcParams p) {
    using namespace um;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES], bar_acc;
    __shared__ uint32_t tmem_slot;
    // NT <= 128: the hi | lo weight tiles of a chunk are contiguous in shared memory, i.e. ONE K-major operand of 2 NT
    // rows, so A_hi * [B_hi ; B_lo] is a single N = 2 NT instruction into two accumulators (summed in the epilogue).
    // Two MMAs per k-step instead of three, and the wide one runs at the full tensor rate (instructions with N <= 128
    // are bound by fetching the A operand from shared memory, not by the arithmetic).
    constexpr bool MERGE = NT <= 128;
    constexpr int TCOLS = MERGE ? 2 * NT : NT;
    constexpr uint32_t A_BYTES = 128 * 128;                   // one half (hi or lo) of the A tile
    constexpr uint32_t B_BYTES = NT * 256;                    // hi | lo
    constexpr uint32_t STAGE = 2 * A_BYTES + B_BYTES;
    const uint32_t base = (s32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();

    const int nchunks_all = p.ntaps * p.cpt;
    const int per = (nchunks_all + p.splits - 1) / p.splits;
    const int c_begin = blockIdx.z * per, c_end = min(nchunks_all, c_begin + per);
    const int nchunks = max(0, c_end - c_begin);

    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int n = t / p.tiles_y;
    const int x0 = tx * p.BW, y0 = ty * p.BH;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(s32(&bar_full[s]), 1); mbar_init(s32(&bar_empty[s]), 1); }
        mbar_init(s32(&bar_acc), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&map_hi);
        tma_prefetch_desc(&map_lo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_slot)), "n"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();                       // everything above is local set-up; the predecessor's outputs are read from here on

    if (warp == 0) {
        if (lane == 0) {
            const uint8_t *wsrc = p.wt + (size_t)blockIdx.y * p.wt_chunks * B_BYTES;
            for (int i = 0; i < nchunks; ++i) {
                const int c = c_begin + i, tap = c / p.cpt, cb = c - tap * p.cpt;
                const TapDesc td = p.taps[tap];
                const uint32_t s = i % STAGES, ph = (i / STAGES) & 1;
                mbar_wait(s32(&bar_empty[s]), ph ^ 1);
                const uint32_t full = s32(&bar_full[s]), dst = base + s * STAGE;
                mbar_expect_tx(full, STAGE);
                const int cx = x0 + td.dx, cy = y0 + td.dy, cn = n * p.nplanes + td.plane;
                tma_load_4d(dst, &map_hi, cb * 64, cx, cy, cn, full);
                tma_load_4d(dst + A_BYTES, &map_lo, cb * 64, cx, cy, cn, full);
                bulk_g2s(dst + 2 * A_BYTES, wsrc + (size_t)(td.wtap * p.cpt + cb) * B_BYTES, B_BYTES, full);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t IDESC = idesc_f16(128, NT);
            for (int i = 0; i < nchunks; ++i) {
                const uint32_t s = i % STAGES, ph = (i / STAGES) & 1;
                mbar_wait(s32(&bar_full[s]), ph);
                fence_after();
                const uint32_t sb = base + s * STAGE;
                const uint64_t ah = desc_sw128(sb), al = desc_sw128(sb + A_BYTES);
                const uint64_t bh = desc_sw128(sb + 2 * A_BYTES), bl = desc_sw128(sb + 2 * A_BYTES + NT * 128);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {               // 4 x (K = 16): +32 bytes inside the 128-byte row
                    if constexpr (MERGE) {
                        constexpr uint32_t IDESC2 = idesc_f16(128, 2 * NT);
                        mma_ss(tmem, ah + 2 * ks, bh + 2 * ks, IDESC2, (i | ks) != 0);     // [hi*Whi | hi*Wlo]
                        mma_ss(tmem, al + 2 * ks, bh + 2 * ks, IDESC, 1);                  // lo*Whi -> first accumulator
                    } else {
                        mma_ss(tmem, ah + 2 * ks, bh + 2 * ks, IDESC, (i | ks) != 0);
                        mma_ss(tmem, ah + 2 * ks, bl + 2 * ks, IDESC, 1);
                        mma_ss(tmem, al + 2 * ks, bh + 2 * ks, IDESC, 1);
                    }
                }
                (void)bl;
                commit(s32(&bar_empty[s]));
            }
            commit(s32(&bar_acc));
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        // TMEM hands every thread one PIXEL (32 channels at a time); global memory wants a warp on one pixel's
        // CHANNELS (128 contiguous bytes).  Each warp therefore transposes its 32 x 32 block through shared memory
        // (conflict-free, stride 33) and then walks its 32 pixels with lane = channel: one coalesced 128-byte store
        // per pixel, and the per-channel sums for the following norm fall out of the same walk.
        const int q4 = warp & 3, r = q4 * 32 + lane;                      // TMEM lane = pixel index inside the tile
        const int et = (warp - 2) * 32 + lane;                            // 0..127 among the epilogue threads
        const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16);
        mbar_wait(s32(&bar_acc), 0);
        fence_after();
        float *red = reinterpret_cast<float *>(smem_raw + (base - s32(smem_raw)));       // [128][33] + [4][32][2]
        float *part = red + 128 * 33;
        const int n0 = blockIdx.y * NT;
        const bool fin = p.splits == 1;
        const int bw_shift = 31 - __clz(p.BW), bw_mask = p.BW - 1;
        const size_t split_stride = (size_t)p.N * p.Ht * p.Wt * p.Cout;

        // (a) final tile + statistics: transposed walk (coalesced stores, sums from the same walk)
        // (b) final tile, no statistics / (c) split-K partial: straight 128-bit stores of the thread's own pixel
        const int a_me = y0 + (r >> bw_shift), b_me = x0 + (r & bw_mask);
        const bool pv = a_me < p.Ht && b_me < p.Wt;
        float *direct = nullptr;
        if (fin) direct = p.out + (((size_t)n * p.OHf + (size_t)(a_me * p.osy + p.ooy)) * p.OWf + (size_t)(b_me * p.osx + p.oox)) * p.Cs + p.co_off;
        else direct = p.partial + (size_t)blockIdx.z * split_stride + (((size_t)n * p.Ht + a_me) * p.Wt + b_me) * p.Cout;
        const bool walk = fin && p.stats != nullptr;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(direct) & 15) == 0);
        for (int cb = 0; cb < NT; cb += 32) {
            if (n0 + cb >= p.Cout) break;
            uint32_t acc[32];
            tmem_ld32(tl + cb, acc);
            if constexpr (MERGE) {
                uint32_t acc2[32];
                tmem_ld32(tl + NT + cb, acc2);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(acc2[j]));
            }
            if (nchunks == 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = 0u;
            }
            if (!walk) {
                if (fin && p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + cb + j < p.Cout) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __ldg(p.bias + n0 + cb + j));
                }
                if (pv) {
                    if (vec_ok && n0 + cb + 32 <= p.Cout) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<uint4 *>(direct + n0 + cb + j) = make_uint4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + cb + j < p.Cout) direct[n0 + cb + j] = __uint_as_float(acc[j]);
                    }
                }
                continue;
            }
            const int co = n0 + cb + lane;
            const bool cv = co < p.Cout;
#pragma unroll
            for (int j = 0; j < 32; ++j) red[r * 33 + j] = __uint_as_float(acc[j]);
            __syncwarp();
            const float bias = (p.bias && cv) ? __ldg(p.bias + co) : 0.f;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
                const int row = q4 * 32 + i;
                const int a = y0 + (row >> bw_shift), b = x0 + (row & bw_mask);
                if (cv && a < p.Ht && b < p.Wt) {
                    const float val = red[row * 33 + lane] + bias;
                    p.out[(((size_t)n * p.OHf + (size_t)(a * p.osy + p.ooy)) * p.OWf + (size_t)(b * p.osx + p.oox)) * p.Cs + p.co_off + co] = val;
                    s1 += val; s2 = fmaf(val, val, s2);
                }
            }
            __syncwarp();
            part[(q4 * 32 + lane) * 2] = s1; part[(q4 * 32 + lane) * 2 + 1] = s2;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et < 64) {
                const int j = et >> 1, w = et & 1, c2 = n0 + cb + j;
                if (c2 < p.Cout) {
                    const double tot = (double)part[(0 * 32 + j) * 2 + w] + (double)part[(1 * 32 + j) * 2 + w] +
                                       (double)part[(2 * 32 + j) * 2 + w] + (double)part[(3 * 32 + j) * 2 + w];
                    atomicAdd(p.stats + ((size_t)n * p.Cout + c2) * 2 + w, tot);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCOLS) : "memory");
    }
}

// split-K finish: out = bias + sum_s partial[s] in split order (deterministic) + per-(image, channel) statistics.
// Block = 32 pixels x 128 channels, 256 threads: lane = channel quad (128-bit loads, coalesced), warp = pixel slot,
// 4 pixels per thread, every load of a thread independent.  grid (ceil(HW / 32), ceil(Cout / 128), N).
__global__ void __launch_bounds__(256) k_splitk_nhwc(const __grid_constant__ ConvNhwcParams p) {
    __shared__ float4 sred[2][8][32];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.y * 128 + lane * 4;
    const int n = blockIdx.z;
    const int64_t hw = (int64_t)p.Ht * p.Wt;
    const size_t split_stride = (size_t)p.N * hw * p.Cout;
    const bool cv = c < p.Cout;                                 // Cout % 4 == 0 is checked by the host for this kernel
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cv && p.bias) bias = *reinterpret_cast<const float4 *>(p.bias + c);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t pix = (int64_t)blockIdx.x * 32 + k * 8 + w;
        if (!cv || pix >= hw) continue;
        const float *src = p.partial + ((size_t)n * hw + pix) * p.Cout + c;
        float4 v = bias;
        for (int s = 0; s < p.splits; ++s) {
            const float4 t = __ldcg(reinterpret_cast<const float4 *>(src + (size_t)s * split_stride));
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        const int a = (int)(pix / p.Wt), b = (int)(pix % p.Wt);
        *reinterpret_cast<float4 *>(p.out + (((size_t)n * p.OHf + (size_t)(a * p.osy + p.ooy)) * p.OWf + (size_t)(b * p.osx + p.oox)) * p.Cs +
                                    p.co_off + c) = v;
        s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
        s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
    }
    if (!p.stats) return;
    sred[0][w][lane] = s1; sred[1][w][lane] = s2;
    __syncthreads();
    const int t = threadIdx.x;                                  // 256 threads = 2 x 128 channels
    const int which = t >> 7, cc = t & 127, co = blockIdx.y * 128 + cc;
    if (co < p.Cout) {
        double tot = 0.0;
        for (int k = 0; k < 8; ++k) tot += (double)reinterpret_cast<const float *>(&sred[which][k][cc >> 2])[cc & 3];
        atomicAdd(p.stats + ((size_t)n * p.Cout + co) * 2 + which, tot);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp16 tensor, dims innermost first (elements), strides of dims 1..3 in elements; box [64][bw][bh][1], SWIZZLE_128B
static int make_map(CUtensorMap *m, const void *ptr, const int64_t dims[4], const int64_t strides[3], int bw, int bh) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return ICON_ECUDA; }
    cuuint64_t gd[4] = {(cuuint64_t)dims[0], (cuuint64_t)dims[1], (cuuint64_t)dims[2], (cuuint64_t)dims[3]};
    cuuint64_t gs[3] = {(cuuint64_t)strides[0] * 2, (cuuint64_t)strides[1] * 2, (cuuint64_t)strides[2] * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), gd, gs, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims %lld %lld %lld %lld strides %lld %lld %lld box %d %d", (int)r,
                  (long long)dims[0], (long long)dims[1], (long long)dims[2], (long long)dims[3], (long long)strides[0],
                  (long long)strides[1], (long long)strides[2], bw, bh);
        return ICON_ECUDA;
    }
    return ICON_OK;
}

template <int NT, int STAGES>
static int launch_conv_nhwc(const CUtensorMap &mh, const CUtensorMap &ml, const ConvNhwcParams &p, dim3 grid,
                            cudaStream_t stream) {
    constexpr int smem = STAGES * (2 * 128 * 128 + NT * 256) + 1024;
    static bool attr_set[ICON_MAX_DEVICES] = {};
    if (device_needs_setup(attr_set))
        ICON_CUDA(cudaFuncSetAttribute(k_conv_nhwc<NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ICON_CUDA(launch_pdl(k_conv_nhwc<NT, STAGES>, grid, dim3(CN_THREADS), (size_t)smem, stream, mh, ml, p));
    ICON_LAUNCHED();
    return ICON_OK;
}

}  // namespace icon

using namespace icon;

extern "C" size_t icon_conv_nhwc_workspace_bytes(int N, int Ht, int Wt, int Cout, int splits) {
    return splits > 1 ? (size_t)splits * N * Ht * Wt * Cout * sizeof(float) : 0;
}

extern "C" int icon_conv_nhwc(const void *a_hi, const void *a_lo, const int64_t *dims, const int64_t *strides,
                              const void *wt_packed, int wt_chunks, const float *bias, float *out, int OHf, int OWf, int Cs,
                              int co_off, int Cout, int N, int Ht, int Wt, int osy, int osx, int ooy, int oox, int nplanes,
                              int ntaps, const int *taps /* ntaps x (dy, dx, plane, wtap) */, int cpt, int n_tile, int splits,
                              double *stats, void *ws, size_t ws_bytes, icon_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ICON_CHECK_ARG(a_hi && a_lo && dims && strides && wt_packed && taps, "icon_conv_nhwc: null pointer");
    ICON_CHECK_ARG(out || splits > 1, "icon_conv_nhwc: out == NULL (park the split-K partials only) needs splits > 1");
    ICON_CHECK_ARG(N > 0 && Ht > 0 && Wt > 0 && Cout > 0 && cpt > 0 && nplanes > 0, "icon_conv_nhwc: bad size");
    ICON_CHECK_ARG(ntaps >= 1 && ntaps <= MAX_TAPS, "icon_conv_nhwc: 1..%d taps", MAX_TAPS);
    ICON_CHECK_ARG(n_tile == 64 || n_tile == 128 || n_tile == 256, "icon_conv_nhwc: n_tile must be 64, 128 or 256");
    ICON_CHECK_ARG(splits >= 1 && splits <= ntaps * cpt, "icon_conv_nhwc: 1 <= splits <= chunks");
    ICON_CHECK_ARG((((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)wt_packed) & 15) == 0, "icon_conv_nhwc: 16-byte alignment");
    ICON_CHECK_ARG(dims[0] % 8 == 0 && strides[0] % 8 == 0 && strides[1] % 8 == 0 && strides[2] % 8 == 0,
                   "icon_conv_nhwc: tensor-map strides must be multiples of 16 bytes");
    ICON_CHECK_ARG(co_off >= 0 && co_off + Cout <= Cs, "icon_conv_nhwc: channel slice outside the output tensor");
    ConvNhwcParams p{};
    p.wt = (const uint8_t *)wt_packed; p.bias = bias; p.out = out; p.stats = splits == 1 ? stats : nullptr;
    p.N = N; p.Ht = Ht; p.Wt = Wt;
    int bw = 1;
    while (bw < Wt && bw < 16) bw <<= 1;                       // 8 x 16 pixel tiles (taps of neighbouring tiles overlap
                                                               // in L2); narrower boxes for images under 16 wide
    p.BW = bw; p.BH = 128 / bw;
    p.tiles_x = (Wt + p.BW - 1) / p.BW; p.tiles_y = (Ht + p.BH - 1) / p.BH;
    p.OHf = OHf; p.OWf = OWf; p.osy = osy; p.osx = osx; p.ooy = ooy; p.oox = oox;
    p.Cs = Cs; p.co_off = co_off; p.Cout = Cout; p.cpt = cpt; p.ntaps = ntaps; p.nplanes = nplanes;
    p.wt_chunks = wt_chunks; p.splits = splits;
    for (int i = 0; i < ntaps; ++i) {
        p.taps[i].dy = (int8_t)taps[4 * i]; p.taps[i].dx = (int8_t)taps[4 * i + 1];
        p.taps[i].plane = (uint8_t)taps[4 * i + 2]; p.taps[i].wtap = (uint8_t)taps[4 * i + 3];
        ICON_CHECK_ARG(taps[4 * i + 2] >= 0 && taps[4 * i + 2] < nplanes && (taps[4 * i + 3] + 1) * cpt <= wt_chunks,
                       "icon_conv_nhwc: tap %d out of range", i);
    }
    ICON_CHECK_ARG(!out || ((Ht - 1) * osy + ooy < OHf && (Wt - 1) * osx + oox < OWf), "icon_conv_nhwc: output mapping outside the tensor");
    const size_t need = icon_conv_nhwc_workspace_bytes(N, Ht, Wt, Cout, splits);
    if (ws_bytes < need || (need && !ws)) { set_error("icon_conv_nhwc: workspace %zu < %zu", ws_bytes, need); return ICON_ENOSPC; }
    p.partial = splits > 1 ? (float *)ws : nullptr;
    ICON_CHECK_ARG(splits == 1 || !out || (Cout % 4 == 0 && Cs % 4 == 0 && co_off % 4 == 0 && ((uintptr_t)out & 15) == 0 &&
                                           (!bias || ((uintptr_t)bias & 15) == 0)),
                   "icon_conv_nhwc: split-K needs channel counts / offsets that are multiples of 4");
    CUtensorMap mh, ml;
    int rc = make_map(&mh, a_hi, dims, strides, p.BW, p.BH);
    if (rc) return rc;
    rc = make_map(&ml, a_lo, dims, strides, p.BW, p.BH);
    if (rc) return rc;
    dim3 grid((unsigned)(p.tiles_x * p.tiles_y * N), (unsigned)((Cout + n_tile - 1) / n_tile), (unsigned)splits);
    if (n_tile == 256) rc = launch_conv_nhwc<256, 2>(mh, ml, p, grid, stream);
    else if (n_tile == 128) rc = launch_conv_nhwc<128, 3>(mh, ml, p, grid, stream);
    else rc = launch_conv_nhwc<64, 2>(mh, ml, p, grid, stream);     // 97 KB: two CTAs per SM overlap prologue / epilogue
    if (rc) return rc;
    if (splits > 1 && out) {
        p.stats = stats;
        dim3 g2((unsigned)(((int64_t)Ht * Wt + 31) / 32), (unsigned)((Cout + 127) / 128), (unsigned)N);
        ICON_CUDA(launch_pdl(k_splitk_nhwc, g2, dim3(256), 0, stream, p));
        ICON_LAUNCHED();
    }
    return ICON_OK;
}
