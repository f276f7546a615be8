// Micro-benchmark: cost of one dependent kernel node inside a CUDA graph on B200, with and without programmatic
// dependent launch (PDL), for the launch shapes the UNet plan uses. Build: nvcc -arch=sm_100a -O3 -o launch_latency ...
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_work(float* buf, int spin, int pdl) {
    extern __shared__ float sm[];
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // "prologue": touch smem
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    float v = buf[(blockIdx.x * blockDim.x + threadIdx.x) & 1023];
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    buf[(blockIdx.x * blockDim.x + threadIdx.x) & 1023] = v + sm[threadIdx.x] * 0.f;
}

static float run(int grid, int smem, int spin, int pdl, int n) {
    float* buf;
    CK(cudaMalloc(&buf, 4096));
    CK(cudaMemset(buf, 0, 4096));
    CK(cudaFuncSetAttribute(k_work, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < n; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(320);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        CK(cudaLaunchKernelEx(&cfg, k_work, buf, spin, pdl));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < 10; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
    cudaStreamDestroy(st);
    cudaFree(buf);
    return ms * 1000.f / (10.f * n);
}

int main() {
    const int n = 200;
    printf("us per dependent kernel node in a graph (320 threads/CTA)\n");
    printf("%6s %8s %8s | %8s %8s\n", "grid", "smem_KB", "spin_cyc", "plain", "PDL");
    int grids[] = {26, 104, 148, 416};
    int smems[] = {2 * 1024, 100 * 1024, 200 * 1024};
    int spins[] = {0, 4000, 14000};
    for (int gi = 0; gi < 4; ++gi)
        for (int si = 0; si < 3; ++si)
            for (int pi = 0; pi < 3; ++pi) {
                float a = run(grids[gi], smems[si], spins[pi], 0, n);
                float b = run(grids[gi], smems[si], spins[pi], 1, n);
                printf("%6d %8d %8d | %8.2f %8.2f\n", grids[gi], smems[si] / 1024, spins[pi], a, b);
            }
    return 0;
}
