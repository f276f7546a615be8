// Latency / throughput of bulk global -> shared copies (the TMA engine) issued by one thread of one CTA per SM, from an
// L2-resident buffer: time from issue to mbarrier completion for one copy of N bytes, and for D copies of N bytes in
// flight (steady state), with `grid` CTAs running at once. Build: nvcc -arch=sm_100a -O3 -o tma_latency tma_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void wait(uint64_t* b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}
__global__ void k(const uint8_t* src, size_t src_bytes, int bytes, int depth, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint8_t* base = src + ((size_t)blockIdx.x * 7919 * 4096) % (src_bytes - (size_t)bytes * 64);
        long long t0 = clock64();
        int issued = 0, done = 0;
        for (; issued < depth; ++issued) { expect(&bar[issued], bytes); bulk(sm + issued * bytes, base + (size_t)(issued % 48) * bytes, bytes, &bar[issued]); }
        for (; done < iters; ++done) {
            const int s = done % depth;
            wait(&bar[s], (done / depth) & 1);
            if (issued < iters) { expect(&bar[s], bytes); bulk(sm + s * bytes, base + (size_t)(issued % 48) * bytes, bytes, &bar[s]); ++issued; }
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
}
int main() {
    size_t nb = 64u << 20;
    uint8_t* src; cudaMalloc(&src, nb); cudaMemset(src, 1, nb);
    long long* out; cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int sizes[] = {4096, 8192, 16384, 32768, 49152};
    for (int grid : {1, 104, 148}) for (int bytes : sizes) for (int depth : {1, 2, 3, 4, 6}) {
        if ((size_t)bytes * depth > 196608) continue;
        const int iters = 64;
        for (int rep = 0; rep < 2; ++rep) k<<<grid, 32, 200 * 1024>>>(src, nb, bytes, depth, iters, out);
        long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("grid %3d  %5d B x depth %d: %7.0f cycles per copy, %6.1f B/cycle/SM\n", grid, bytes, depth, (double)h / iters, (double)bytes * iters / h);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
