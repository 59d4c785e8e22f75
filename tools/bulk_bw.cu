// Microbenchmark: sustained global(L2-resident) -> shared bandwidth per SM of cp.async.bulk (1-D TMA
// bulk copy) vs grid size, copy size and copies in flight.  Decides whether the tcgen05 MLP's
// weight streaming is limited per SM or chip-wide.   nvcc -arch=sm_100a -O2 -o build/bulk_bw tools/bulk_bw.cu
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
// one thread per CTA streams `iters` copies of `bytes` each through `nst` smem stages
__global__ void k_bulk(const uint8_t *src, size_t src_bytes, int bytes, int nst, int iters, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
        const size_t nchunk = src_bytes / bytes;
        long long t0 = clock64();
        for (int it = 0; it < iters + nst; ++it) {
            const int s = it % nst;
            if (it >= nst) mbar_wait(s32(&bar[s]), ((it / nst) - 1) & 1);
            if (it < iters) {
                const uint8_t *p = src + ((size_t)(it + blockIdx.x * 7) % nchunk) * bytes;
                asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(s32(&bar[s])), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)s * bytes)), "l"(p), "r"(bytes), "r"(s32(&bar[s])) : "memory");
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}
// all threads stream with ld.global.v4 -> st.shared (no TMA)
__global__ void k_ldst(const uint8_t *src, size_t src_bytes, int bytes, int iters, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const size_t nchunk = src_bytes / bytes;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint4 *p = (const uint4 *)(src + ((size_t)(it + blockIdx.x * 7) % nchunk) * bytes);
        uint4 *d = (uint4 *)(sm + (size_t)(it & 1) * bytes);
        for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = __ldg(p + i);
        __syncthreads();
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
// `nthr` issuing threads (one per warp), each with its own stages and barriers
__global__ void k_bulk_mt(const uint8_t *src, size_t src_bytes, int bytes, int nst, int iters, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[32];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        uint64_t *b = bar + w * nst;
        uint8_t *buf = sm + (size_t)w * nst * bytes;
        for (int i = 0; i < nst; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&b[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
        const size_t nchunk = src_bytes / bytes;
        long long t0 = clock64();
        for (int it = 0; it < iters + nst; ++it) {
            const int s = it % nst;
            if (it >= nst) mbar_wait(s32(&b[s]), ((it / nst) - 1) & 1);
            if (it < iters) {
                const uint8_t *p = src + ((size_t)(it * 3 + w + blockIdx.x * 7) % nchunk) * bytes;
                asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(s32(&b[s])), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + (size_t)s * bytes)), "l"(p), "r"(bytes), "r"(s32(&b[s])) : "memory");
            }
        }
        if (w == 0) cycles[blockIdx.x] = clock64() - t0;
    }
}
int main() {
    const size_t src_bytes = 704512;   // ~ the MLP weight blob
    uint8_t *src; CK(cudaMalloc(&src, src_bytes)); CK(cudaMemset(src, 1, src_bytes));
    long long *cyc; CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_ldst, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int grids[] = {1, 8, 37, 74, 148};
    struct { int bytes, nst; } cfg[] = {{65536, 2}, {32768, 4}, {16384, 8}, {8192, 8}, {32768, 2}};
    printf("cp.async.bulk: bytes/clk/SM (chip TB/s at the measured clock)\n");
    for (auto c : cfg) for (int g : grids) {
        const int iters = 2000;
        k_bulk<<<g, 32, (size_t)c.bytes * c.nst>>>(src, src_bytes, c.bytes, c.nst, 50, cyc); CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_bulk<<<g, 32, (size_t)c.bytes * c.nst>>>(src, src_bytes, c.bytes, c.nst, iters, cyc);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[148]; CK(cudaMemcpy(h, cyc, g * 8, cudaMemcpyDeviceToHost));
        double mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("  copy %6d B x %d stages, grid %3d: %6.1f B/clk/SM   %6.2f TB/s chip  (%.2f GHz)\n", c.bytes, c.nst, g,
               (double)c.bytes * iters / mx, (double)c.bytes * iters * g / (ms * 1e-3) / 1e12, mx / (ms * 1e-3) / 1e9);
    }
    CK(cudaFuncSetAttribute(k_bulk_mt, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    printf("cp.async.bulk from several issuing threads (grid 148):\n");
    struct { int bytes, nst, nthr; } mt[] = {{32768, 2, 1}, {32768, 2, 2}, {16384, 2, 4}, {16384, 2, 2}, {8192, 2, 8}, {32768, 1, 4}, {16384, 1, 8}};
    for (auto c : mt) {
        const int iters = 2000, g = 148;
        k_bulk_mt<<<g, 32 * c.nthr, (size_t)c.bytes * c.nst * c.nthr>>>(src, src_bytes, c.bytes, c.nst, 50, cyc); CK(cudaDeviceSynchronize());
        k_bulk_mt<<<g, 32 * c.nthr, (size_t)c.bytes * c.nst * c.nthr>>>(src, src_bytes, c.bytes, c.nst, iters, cyc); CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, cyc, g * 8, cudaMemcpyDeviceToHost));
        double mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("  %d threads x copy %6d B x %d stages: %6.1f B/clk/SM total\n", c.nthr, c.bytes, c.nst, (double)c.bytes * iters * c.nthr / mx);
    }
    printf("ld.global.v4 -> st.shared, 256 threads:\n");
    for (int g : grids) {
        const int iters = 2000, bytes = 32768;
        k_ldst<<<g, 256, bytes * 2>>>(src, src_bytes, bytes, 50, cyc); CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_ldst<<<g, 256, bytes * 2>>>(src, src_bytes, bytes, iters, cyc);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[148]; CK(cudaMemcpy(h, cyc, g * 8, cudaMemcpyDeviceToHost));
        double mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("  grid %3d: %6.1f B/clk/SM   %6.2f TB/s chip\n", g, (double)bytes * iters / mx, (double)bytes * iters * g / (ms * 1e-3) / 1e12);
    }
    return 0;
}
