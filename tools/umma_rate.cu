// tcgen05.mma issue-rate microbenchmark (one CTA per SM): cycles per MMA for M=128, kind::f16, with the A
// operand in TMEM (TS) or shared memory (SS), N = 256 / 128 / 64, optionally with concurrent bulk copies
// into another shared-memory region and concurrent tcgen05.ld/st traffic from other warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_rate tools/umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.b32 %0, 1, 0, q;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__global__ void __launch_bounds__(256) k_rate(int N, int ts, int copies, int ldst, int iters, const uint8_t *src, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_s;
    __shared__ volatile int stop;
    const uint32_t raw = s32(smem_raw), base = (raw + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
        stop = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0 && lane == 0) {
        const uint64_t bd = desc_sw128(base), ad = desc_sw128(base + 65536);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int ks = 0; ks < 4; ++ks) {
                if (ts) asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q;\n}" ::"r"(tmem), "r"(tmem + 256 + 8 * ks), "l"(bd + 2 * ks), "r"(idesc), "r"(1));
                else asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n}" ::"r"(tmem), "l"(ad + 2 * ks), "l"(bd + 2 * ks), "r"(idesc), "r"(1));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar[0])));
        mbar_wait(s32(&bar[0]), 0);
        out[blockIdx.x] = clock64() - t0;
        stop = 1;
    } else if (warp >= 1 && warp <= 2 && lane == 0 && copies) {
        // background bulk copies into a separate 64 KB region (two threads, 32 KB each, back to back)
        const int w = warp - 1;
        uint32_t ph = 0;
        while (!stop) {
            asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(s32(&bar[1 + w])), "r"(32768) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + 98304 + w * 32768), "l"(src + w * 32768), "r"(32768), "r"(s32(&bar[1 + w])) : "memory");
            mbar_wait(s32(&bar[1 + w]), ph); ph ^= 1;
        }
    } else if (warp >= 4 && ldst) {
        // background TMEM traffic: each warp reads 32 columns and writes 16 back, like the epilogue workers
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384;
        uint32_t r[32];
        while (!stop) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(tl));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tl + 64),
                "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
            asm volatile("tcgen05.wait::st.sync.aligned;");
            for (int k = 0; k < 200; ++k) asm volatile("" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
    uint8_t *src; CK(cudaMalloc(&src, 1 << 20)); CK(cudaMemset(src, 0, 1 << 20));
    long long *out; CK(cudaMalloc(&out, 148 * 8));
    const int smem = 65536 + 32768 + 65536 + 1024;
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 2000;
    struct { int N, ts, copies, ldst; } cfg[] = {{256, 1, 0, 0}, {256, 0, 0, 0}, {128, 1, 0, 0}, {64, 1, 0, 0}, {256, 1, 1, 0}, {256, 1, 0, 1}, {256, 1, 1, 1}, {128, 1, 1, 1}};
    for (auto c : cfg) for (int g : {1, 148}) {
        k_rate<<<g, 256, smem>>>(c.N, c.ts, c.copies, c.ldst, iters, src, out);
        CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, out, g * 8, cudaMemcpyDeviceToHost));
        double mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("N=%3d %s copies=%d tmem_ldst=%d grid=%3d: %6.1f cycles / MMA (ideal %d)\n", c.N, c.ts ? "TS" : "SS", c.copies, c.ldst, g, mx / (iters * 4.0), c.N / 2);
    }
    return 0;
}
