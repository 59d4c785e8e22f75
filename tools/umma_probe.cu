// tcgen05 probe: settles, on the real B200, the operand-layout facts the fused MLP kernel relies
// on (shared-memory descriptor fields for K-major SWIZZLE_128B and no-swizzle tiles, A operand in
// tensor memory, tcgen05.ld/st lane mapping, fp16 instruction descriptor).  Not part of the product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_probe tools/umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int A_SMEM = 32768, B_SMEM = 65536;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params {
    const uint8_t *a_img; int a_bytes;      // SS: shared-memory image of A
    const uint8_t *b_img; int b_bytes;
    const uint32_t *a_tmem; int a_cols;     // TS: [128][a_cols] 32-bit words per row
    int mode;                               // 0 = SS, 1 = TS
    uint64_t a_desc, b_desc;                // descriptors without the start address
    int a_step, b_step;                     // start-address increment (16 B units) per k-step
    int a_tstep;                            // TS: TMEM column increment per k-step
    uint32_t idesc; int ksteps; int N;
    float *out;                             // [128][N]
};

__global__ void __launch_bounds__(128) probe(Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sa = smem, *sb = smem + A_SMEM;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = tid; i < p.a_bytes / 4; i += 128) ((uint32_t *)sa)[i] = ((const uint32_t *)p.a_img)[i];
    for (int i = tid; i < p.b_bytes / 4; i += 128) ((uint32_t *)sb)[i] = ((const uint32_t *)p.b_img)[i];
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    if (p.mode == 1) {
        // every thread writes its own row of A into TMEM columns [256, 256 + a_cols)
        for (int c = 0; c < p.a_cols; c += 8) {
            const uint32_t *src = p.a_tmem + (size_t)tid * p.a_cols + c;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(lane_addr + 256 + c),
                         "r"(src[0]), "r"(src[1]), "r"(src[2]), "r"(src[3]), "r"(src[4]), "r"(src[5]), "r"(src[6]), "r"(src[7]));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        uint64_t ad = p.a_desc | (uint64_t)((smem_u32(sa) >> 4) & 0x3FFF);
        uint64_t bd = p.b_desc | (uint64_t)((smem_u32(sb) >> 4) & 0x3FFF);
        for (int k = 0; k < p.ksteps; ++k) {
            uint32_t acc = k > 0;
            if (p.mode == 0) {
                asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n}" ::"r"(tmem),
                             "l"(ad + (uint64_t)(k * p.a_step)), "l"(bd + (uint64_t)(k * p.b_step)), "r"(p.idesc), "r"(acc));
            } else {
                asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q;\n}" ::"r"(tmem),
                             "r"(tmem + 256 + k * p.a_tstep), "l"(bd + (uint64_t)(k * p.b_step)), "r"(p.idesc), "r"(acc));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        long long spins = 0;
        while (!done && ++spins < 20000000ll) {
            asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\nselp.b32 %0, 1, 0, q;\n}"
                         : "=r"(done) : "r"(smem_u32(&bar)));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c = 0; c < p.N; c += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(lane_addr + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 8; ++j) p.out[(size_t)tid * p.N + c + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---------------------------------------------------------------- host
static uint64_t make_desc(uint32_t lbo_bytes, uint32_t sbo_bytes, int layout) {
    uint64_t d = 0;
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
static uint32_t make_idesc(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                           // D = F32
    d |= 0u << 7;                           // A = F16
    d |= 0u << 10;                          // B = F16
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

struct Mats {
    int K, N;
    std::vector<__half> A, B;               // A [128][K], B [N][K]
    std::vector<float> ref;                 // [128][N]
};
static Mats make_mats(int K, int N, int seed) {
    Mats m; m.K = K; m.N = N;
    m.A.resize(128 * K); m.B.resize((size_t)N * K); m.ref.assign(128 * (size_t)N, 0.f);
    srand(seed);
    for (auto &v : m.A) v = __float2half((float)(rand() % 17 - 8) / 8.f);
    for (auto &v : m.B) v = __float2half((float)(rand() % 13 - 6) / 4.f);
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < N; ++j) {
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += __half2float(m.A[i * K + k]) * __half2float(m.B[(size_t)j * K + k]);
            m.ref[(size_t)i * N + j] = s;
        }
    return m;
}
// K-major SWIZZLE_128B image: row pitch 128 B (64 halves), 16-B chunk index XOR (row % 8)
static std::vector<uint8_t> img_sw128(const std::vector<__half> &M, int rows, int K) {
    std::vector<uint8_t> im((size_t)rows * 128, 0);
    for (int r = 0; r < rows; ++r)
        for (int e = 0; e < K; ++e) {
            size_t off = (size_t)r * 128 + (size_t)(((e / 8) ^ (r % 8)) * 16) + (e % 8) * 2;
            memcpy(&im[off], &M[(size_t)r * K + e], 2);
        }
    return im;
}
// K-major no-swizzle image: core matrix = 8 rows x 16 B; kstride / gstride in bytes
static std::vector<uint8_t> img_nosw(const std::vector<__half> &M, int rows, int K, int kstride, int gstride) {
    size_t sz = (size_t)(K / 8) * kstride + (size_t)(rows / 8) * gstride + 128;
    std::vector<uint8_t> im(sz, 0);
    for (int r = 0; r < rows; ++r)
        for (int e = 0; e < K; ++e) {
            size_t off = (size_t)(e / 8) * kstride + (size_t)(r / 8) * gstride + (r % 8) * 16 + (e % 8) * 2;
            memcpy(&im[off], &M[(size_t)r * K + e], 2);
        }
    return im;
}

static double run(const char *name, const Mats &m, int mode, const std::vector<uint8_t> &aimg, const std::vector<uint8_t> &bimg,
                  const std::vector<uint32_t> &atm, int a_cols, uint64_t adesc, uint64_t bdesc, int a_step, int b_step,
                  int a_tstep, int ksteps) {
    Params p{};
    uint8_t *da = nullptr, *db = nullptr; uint32_t *dt = nullptr; float *dout = nullptr;
    size_t ab = (aimg.size() + 15) / 16 * 16, bb = (bimg.size() + 15) / 16 * 16;
    if (ab > A_SMEM || bb > B_SMEM) { printf("%-44s image too large\n", name); return -1; }
    CK(cudaMalloc(&da, ab ? ab : 16)); CK(cudaMalloc(&db, bb ? bb : 16));
    CK(cudaMemset(da, 0, ab ? ab : 16)); CK(cudaMemset(db, 0, bb ? bb : 16));
    if (aimg.size()) CK(cudaMemcpy(da, aimg.data(), aimg.size(), cudaMemcpyHostToDevice));
    if (bimg.size()) CK(cudaMemcpy(db, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    if (atm.size()) { CK(cudaMalloc(&dt, atm.size() * 4)); CK(cudaMemcpy(dt, atm.data(), atm.size() * 4, cudaMemcpyHostToDevice)); }
    CK(cudaMalloc(&dout, 128 * (size_t)m.N * 4)); CK(cudaMemset(dout, 0xff, 128 * (size_t)m.N * 4));
    p.a_img = da; p.a_bytes = (int)ab; p.b_img = db; p.b_bytes = (int)bb; p.a_tmem = dt; p.a_cols = a_cols; p.mode = mode;
    p.a_desc = adesc; p.b_desc = bdesc; p.a_step = a_step; p.b_step = b_step; p.a_tstep = a_tstep;
    p.idesc = make_idesc(128, m.N); p.ksteps = ksteps; p.N = m.N; p.out = dout;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM + B_SMEM));
    probe<<<1, 128, A_SMEM + B_SMEM>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s CUDA ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<float> out(128 * (size_t)m.N);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double mx = 0; int bad = 0;
    for (size_t i = 0; i < out.size(); ++i) { double d = fabs((double)out[i] - m.ref[i]); if (!(d <= 1e-3)) ++bad; if (d > mx || d != d) mx = d; }
    printf("%-44s max|err| = %-10.4g mismatches = %d / %zu  %s\n", name, mx, bad, out.size(), bad ? "FAIL" : "OK");
    cudaFree(da); cudaFree(db); if (dt) cudaFree(dt); cudaFree(dout);
    return mx;
}

int main() {
    // 1. SS, both operands K-major SWIZZLE_128B, K = 64 (4 k-steps, +32 B per step), N = 256
    {
        Mats m = make_mats(64, 256, 1);
        auto a = img_sw128(m.A, 128, 64), b = img_sw128(m.B, 256, 64);
        run("SS sw128 K=64 N=256 (SBO=1024, +2/step)", m, 0, a, b, {}, 0, make_desc(16, 1024, 2), make_desc(16, 1024, 2), 2, 2, 0, 4);
    }
    // 2. SS, no swizzle, K = 16: hypothesis H1: LBO = K-direction core stride, SBO = 8-row-group stride
    {
        Mats m = make_mats(16, 64, 2);
        int ka = 16 * 128, kb = 8 * 128;          // A: 16 groups, B: 8 groups; K cores far apart
        auto a = img_nosw(m.A, 128, 16, ka, 128), b = img_nosw(m.B, 64, 16, kb, 128);
        run("SS nosw K=16 N=64  H1 (LBO=K, SBO=rows)", m, 0, a, b, {}, 0, make_desc(ka, 128, 0), make_desc(kb, 128, 0), 0, 0, 0, 1);
        run("SS nosw K=16 N=64  H2 (LBO=rows, SBO=K)", m, 0, a, b, {}, 0, make_desc(128, ka, 0), make_desc(128, kb, 0), 0, 0, 0, 1);
    }
    // 3. SS no swizzle with a sub-tile of a taller B (rows 64..127 of a 512-row matrix): start offset + same LBO
    {
        Mats m = make_mats(16, 64, 3);
        std::vector<__half> Bfull(512 * 16, __float2half(0.f));
        for (int r = 0; r < 64; ++r) for (int e = 0; e < 16; ++e) Bfull[(size_t)(64 + r) * 16 + e] = m.B[(size_t)r * 16 + e];
        int kb = 64 * 128;
        auto a = img_nosw(m.A, 128, 16, 16 * 128, 128), bfull = img_nosw(Bfull, 512, 16, kb, 128);
        // descriptor start address += 8 groups * 128 B = 1024 B = 64 units
        uint64_t bd = make_desc(kb, 128, 0) + 64;
        run("SS nosw sub-tile of 512-row B (H1)", m, 0, a, bfull, {}, 0, make_desc(16 * 128, 128, 0), bd, 0, 0, 0, 1);
    }
    // 4. TS: A in tensor memory (row = lane, two halves per 32-bit column, even k in the low half), B sw128
    {
        Mats m = make_mats(64, 256, 4);
        auto b = img_sw128(m.B, 256, 64);
        std::vector<uint32_t> at(128 * 32);
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < 32; ++c) {
                uint16_t lo, hi;
                memcpy(&lo, &m.A[r * 64 + 2 * c], 2); memcpy(&hi, &m.A[r * 64 + 2 * c + 1], 2);
                at[r * 32 + c] = (uint32_t)lo | ((uint32_t)hi << 16);
            }
        run("TS A-in-TMEM K=64 N=256 (8 cols/step)", m, 1, {}, b, at, 32, 0, make_desc(16, 1024, 2), 0, 2, 8, 4);
    }
    // 5. TS with N = 128 and a longer K through two sw128 B tiles is just (4) repeated; check N=128, N=64
    {
        Mats m = make_mats(64, 128, 5);
        auto b = img_sw128(m.B, 128, 64);
        std::vector<uint32_t> at(128 * 32);
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < 32; ++c) {
                uint16_t lo, hi;
                memcpy(&lo, &m.A[r * 64 + 2 * c], 2); memcpy(&hi, &m.A[r * 64 + 2 * c + 1], 2);
                at[r * 32 + c] = (uint32_t)lo | ((uint32_t)hi << 16);
            }
        run("TS A-in-TMEM K=64 N=128", m, 1, {}, b, at, 32, 0, make_desc(16, 1024, 2), 0, 2, 8, 4);
    }
    return 0;
}
