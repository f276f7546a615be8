// GroupNorm (+SiLU) and LayerNorm kernels: fp32 channels-last in, bf16 out (the GEMM A operand).
// HBM-bound elementwise/reduction work: 128-bit loads, 64-bit bf16x4 stores, one pass for statistics and one
// pass for normalise+activate.
#include "ops.cuh"

namespace pd {
namespace {

constexpr int kGnThreads = 256;
constexpr int kGnIters = 8;  // independent 128-bit loads in flight per thread (HBM latency x bandwidth needs them)

// x [S][R][C] -> (sum, sumsq) per (sample, group), accumulated in double. Each block covers kGnIters * (256 / (C/4))
// rows; a thread keeps one channel quad, so its 8 row loads are independent and issued back to back.
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums,
                                                              int R, int C, int G) {
    grid_dep_launch();
    grid_dep_wait();
    __shared__ float red[kGnThreads][8];
    const int s = blockIdx.y;
    const int c4n = C >> 2;                     // float4 columns
    const int lanes_r = kGnThreads / c4n;       // rows processed per iteration
    const int tc = threadIdx.x % c4n;
    const int tr = threadIdx.x / c4n;
    const int r0 = blockIdx.x * (lanes_r * kGnIters) + tr;
    const float4* base = reinterpret_cast<const float4*>(x + ((size_t)s * R) * C) + tc;
    float4 v[kGnIters];
#pragma unroll
    for (int i = 0; i < kGnIters; ++i) {
        const int r = r0 + i * lanes_r;
        v[i] = r < R ? __ldg(base + (size_t)r * c4n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0, q0 = 0, q1 = 0, q2 = 0, q3 = 0;
#pragma unroll
    for (int i = 0; i < kGnIters; ++i) {
        a0 += v[i].x; a1 += v[i].y; a2 += v[i].z; a3 += v[i].w;
        q0 += v[i].x * v[i].x; q1 += v[i].y * v[i].y; q2 += v[i].z * v[i].z; q3 += v[i].w * v[i].w;
    }
    float* my = red[threadIdx.x];
    my[0] = a0; my[1] = a1; my[2] = a2; my[3] = a3; my[4] = q0; my[5] = q1; my[6] = q2; my[7] = q3;
    __syncthreads();
    const int cpg = C / G;
    for (int g = threadIdx.x; g < G; g += kGnThreads) {
        float sm = 0.f, sq = 0.f;
        for (int rr = 0; rr < lanes_r; ++rr)
            for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
                const float* e = red[rr * c4n + (c >> 2)];
                sm += e[c & 3];
                sq += e[4 + (c & 3)];
            }
        atomicAdd(&sums[((size_t)s * G + g) * 2 + 0], (double)sm);
        atomicAdd(&sums[((size_t)s * G + g) * 2 + 1], (double)sq);
    }
}

// y = act((x - mean) * rstd * gamma + beta) -> bf16. Same thread/row mapping as the stats kernel; a thread's channel
// quad is fixed, so mean/rstd/gamma/beta are computed once per thread (no smem, no block barrier) while its row
// loads are already in flight.
template <bool F32>   // F32: tf32-rounded fp32 operand instead of bf16 (common.cuh)
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const float* __restrict__ x,
                                                              const double* __restrict__ sums,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, void* __restrict__ y,
                                                              int R, int C, int G, float eps, int silu) {
    grid_dep_launch();
    grid_dep_wait();
    const int s = blockIdx.y;
    const int c4n = C >> 2;
    const int lanes_r = kGnThreads / c4n;
    const int tc = threadIdx.x % c4n;
    const int tr = threadIdx.x / c4n;
    const int r0 = blockIdx.x * (lanes_r * kGnIters) + tr;
    const float4* base = reinterpret_cast<const float4*>(x + ((size_t)s * R) * C) + tc;
    float4 v[kGnIters];
#pragma unroll
    for (int i = 0; i < kGnIters; ++i) {
        const int r = r0 + i * lanes_r;
        v[i] = r < R ? __ldg(base + (size_t)r * c4n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int cpg = C / G;
    const int c = tc * 4;
    float sc[4], sh[4];
    {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float gmv[4] = {gm.x, gm.y, gm.z, gm.w};
        const float btv[4] = {bt.x, bt.y, bt.z, bt.w};
        const double n = (double)R * cpg;
        int g_prev = -1;
        float mean = 0.f, rstd = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int g = (c + k) / cpg;
            if (g != g_prev) {
                const double m = sums[((size_t)s * G + g) * 2] / n;
                double var = sums[((size_t)s * G + g) * 2 + 1] / n - m * m;
                if (var < 0) var = 0;
                mean = (float)m;
                rstd = (float)(1.0 / sqrt(var + (double)eps));
                g_prev = g;
            }
            sc[k] = rstd * gmv[k];              // y = x * sc + sh
            sh[k] = btv[k] - mean * sc[k];
        }
    }
    const size_t ybase4 = ((size_t)s * R) * c4n + tc;   // float4 / bf16x4 index of (row 0, my channel quad)
#pragma unroll
    for (int i = 0; i < kGnIters; ++i) {
        const int r = r0 + i * lanes_r;
        if (r < R) {
            float o[4] = {fmaf(v[i].x, sc[0], sh[0]), fmaf(v[i].y, sc[1], sh[1]), fmaf(v[i].z, sc[2], sh[2]),
                          fmaf(v[i].w, sc[3], sh[3])};
            if (silu) {
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = silu_f(o[k]);
            }
            store_operand4<F32>(y, ybase4 + (size_t)r * c4n, o[0], o[1], o[2], o[3]);
        }
    }
}

// One warp per kLnRows consecutive output rows, NV float4 per lane per row: all row loads are issued before any
// reduction so that enough bytes are in flight. GATHER: PatchMerging3D 2x2 space-to-depth on the fly.
constexpr int kLnRows = 4;
template <int NV, bool GATHER, bool F32>
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, void* __restrict__ y, int P,
                                                         int C, float eps, int H, int W, int Cs) {
    grid_dep_launch();
    grid_dep_wait();
    constexpr int ROWS = NV <= 4 ? kLnRows : 1;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int row0 = warp * ROWS;
    if (row0 >= P) return;
    const int c4n = C >> 2;
    float4 v[ROWS][NV];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        const int row = row0 + rr;
        if constexpr (!GATHER) {
            const float4* rp = reinterpret_cast<const float4*>(x + (size_t)row * C);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c4 = lane + i * 32;
                v[rr][i] = (c4 < c4n && row < P) ? __ldg(rp + c4) : make_float4(0, 0, 0, 0);
            }
        } else {
            // output row = (f, h2, w2) over [F][H/2][W/2]; merged channel = (dh*2 + dw) * Cs + c
            const int W2 = W >> 1, H2 = H >> 1;
            const int w2 = row % W2;
            const int h2 = (row / W2) % H2;
            const int f = row / (W2 * H2);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c4 = lane + i * 32;
                if (c4 < c4n && row < P) {
                    const int c = c4 * 4;
                    const int seg = c / Cs;
                    const int cc = c - seg * Cs;
                    const int hh = 2 * h2 + (seg >> 1), ww = 2 * w2 + (seg & 1);
                    v[rr][i] = __ldg(reinterpret_cast<const float4*>(x + (((size_t)f * H + hh) * W + ww) * Cs + cc));
                } else {
                    v[rr][i] = make_float4(0, 0, 0, 0);
                }
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        const int row = row0 + rr;
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) sum += v[rr][i].x + v[rr][i].y + v[rr][i].z + v[rr][i].w;
        const float mean = warp_sum(sum) / (float)C;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c4 = lane + i * 32;
            if (c4 < c4n) {
                const float a = v[rr][i].x - mean, b = v[rr][i].y - mean, c = v[rr][i].z - mean, d = v[rr][i].w - mean;
                sq += a * a + b * b + c * c + d * d;
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
        if (row < P) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c4 = lane + i * 32;
                if (c4 < c4n) {
                    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
                    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c4);
                    store_operand4<F32>(y, (size_t)row * c4n + c4, (v[rr][i].x - mean) * rstd * gm.x + bt.x,
                                        (v[rr][i].y - mean) * rstd * gm.y + bt.y, (v[rr][i].z - mean) * rstd * gm.z + bt.z,
                                        (v[rr][i].w - mean) * rstd * gm.w + bt.w);
                }
            }
        }
    }
}

template <bool GATHER, bool F32>
int launch_ln(const float* x, const float* gamma, const float* beta, void* y, int P, int C, float eps, int H, int W,
              int Cs, cudaStream_t st) {
    PD_CHECK(C % 4 == 0 && C >= 4 && C <= 2048, PD_ERR_SHAPE, "layer_norm: unsupported C=%d", C);
    const int nv = ceil_div(C, 128);
    const int rows_per_warp = nv <= 4 ? kLnRows : 1;
    const int blocks = ceil_div(ceil_div(P, rows_per_warp), 8);
    if (nv <= 1) PD_LAUNCH((layer_norm_kernel<1, GATHER, F32>), blocks, 256, 0, st, x, gamma, beta, y, P, C, eps, H, W, Cs);
    else if (nv <= 2) PD_LAUNCH((layer_norm_kernel<2, GATHER, F32>), blocks, 256, 0, st, x, gamma, beta, y, P, C, eps, H, W, Cs);
    else if (nv <= 4) PD_LAUNCH((layer_norm_kernel<4, GATHER, F32>), blocks, 256, 0, st, x, gamma, beta, y, P, C, eps, H, W, Cs);
    else if (nv <= 8) PD_LAUNCH((layer_norm_kernel<8, GATHER, F32>), blocks, 256, 0, st, x, gamma, beta, y, P, C, eps, H, W, Cs);
    else PD_LAUNCH((layer_norm_kernel<16, GATHER, F32>), blocks, 256, 0, st, x, gamma, beta, y, P, C, eps, H, W, Cs);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace

int gn_stats(const float* x, double* sums, int S, int R, int C, int G, cudaStream_t st) {
    PD_CHECK(C % 4 == 0 && (kGnThreads % (C / 4) == 0) && C / 4 <= kGnThreads, PD_ERR_SHAPE, "gn_stats: unsupported C=%d",
             C);
    PD_CHECK(G > 0 && C % G == 0 && G <= 128, PD_ERR_SHAPE, "gn_stats: unsupported groups=%d for C=%d", G, C);
    dim3 grid(ceil_div(R, (kGnThreads / (C / 4)) * kGnIters), S);
    PD_LAUNCH(gn_stats_kernel, grid, kGnThreads, 0, st, x, sums, R, C, G);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int gn_apply(const float* x, const double* sums, const float* gamma, const float* beta, void* y, int S, int R, int C,
             int G, float eps, int silu, cudaStream_t st, int y_f32) {
    PD_CHECK(C % 4 == 0 && G > 0 && C % G == 0 && G <= 128, PD_ERR_SHAPE, "gn_apply: unsupported C=%d G=%d", C, G);
    PD_CHECK(kGnThreads % (C / 4) == 0 && C / 4 <= kGnThreads, PD_ERR_SHAPE, "gn_apply: unsupported C=%d", C);
    dim3 grid(ceil_div(R, (kGnThreads / (C / 4)) * kGnIters), S);
    if (y_f32) PD_LAUNCH(gn_apply_kernel<true>, grid, kGnThreads, 0, st, x, sums, gamma, beta, y, R, C, G, eps, silu);
    else PD_LAUNCH(gn_apply_kernel<false>, grid, kGnThreads, 0, st, x, sums, gamma, beta, y, R, C, G, eps, silu);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int layer_norm(const float* x, const float* gamma, const float* beta, void* y, int P, int C, float eps,
               cudaStream_t st, int y_f32) {
    return y_f32 ? launch_ln<false, true>(x, gamma, beta, y, P, C, eps, 0, 0, 0, st)
                 : launch_ln<false, false>(x, gamma, beta, y, P, C, eps, 0, 0, 0, st);
}

int patch_merge_ln(const float* x, const float* gamma, const float* beta, void* y, int BT, int H, int W, int C,
                   float eps, cudaStream_t st, int y_f32) {
    PD_CHECK(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, PD_ERR_SHAPE, "patch_merge_ln: H, W must be even (got %d, %d)", H, W);
    return y_f32 ? launch_ln<true, true>(x, gamma, beta, y, BT * (H / 2) * (W / 2), 4 * C, eps, H, W, C, st)
                 : launch_ln<true, false>(x, gamma, beta, y, BT * (H / 2) * (W / 2), 4 * C, eps, H, W, C, st);
}

}  // namespace pd
