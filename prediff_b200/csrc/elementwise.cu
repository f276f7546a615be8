// Elementwise / data-movement kernels and one-time weight repacking. All HBM-bound: vectorised, coalesced,
// grid-stride loops sized in multiples of the SM count.
#include "ops.cuh"

namespace pd {
namespace {

constexpr int kEwThreads = 256;
inline int ew_blocks(int64_t work_items) {
    int64_t b = ceil_div64(work_items, kEwThreads);
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <bool F32>
__global__ void __launch_bounds__(kEwThreads) unet_assemble_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ cond,
                                                                   float* __restrict__ of, void* __restrict__ ob,
                                                                   int B, int Tx, int Tc, int HW, int C, int Cpad) {
    grid_dep_launch();
    grid_dep_wait();
    const int T = Tx + Tc;
    const int c4n = Cpad >> 2;
    const int64_t total = (int64_t)B * T * HW * c4n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        const int64_t pos = i / c4n;  // (b, t, hw)
        const int hw = (int)(pos % HW);
        const int t = (int)((pos / HW) % T);
        const int b = (int)(pos / ((int64_t)HW * T));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C) {  // C % 4 == 0
            const float* src = t < Tc ? cond + (((size_t)b * Tc + t) * HW + hw) * C + c
                                      : x + (((size_t)b * Tx + (t - Tc)) * HW + hw) * C + c;
            v = __ldg(reinterpret_cast<const float4*>(src));
        } else if (c == C) {
            v.x = t < Tc ? 1.f : 0.f;  // observed-frame indicator channel
        }
        if (of) reinterpret_cast<float4*>(of)[i] = v;
        if (ob) store_operand4<F32>(ob, (size_t)i, v.x, v.y, v.z, v.w);
    }
}

__global__ void __launch_bounds__(kEwThreads) pos_embed_kernel(float* __restrict__ x, const float* __restrict__ Te,
                                                               const float* __restrict__ He,
                                                               const float* __restrict__ We, int B, int T, int H, int W,
                                                               int C) {
    grid_dep_launch();
    grid_dep_wait();
    const int c4n = C >> 2;
    const int64_t total = (int64_t)B * T * H * W * c4n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        int64_t pos = i / c4n;
        const int w = (int)(pos % W); pos /= W;
        const int h = (int)(pos % H); pos /= H;
        const int t = (int)(pos % T);
        float4 v = reinterpret_cast<float4*>(x)[i];
        const float4 a = __ldg(reinterpret_cast<const float4*>(Te + (size_t)t * C) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(He + (size_t)h * C) + c4);
        const float4 c = __ldg(reinterpret_cast<const float4*>(We + (size_t)w * C) + c4);
        // reference order: ((x + T) + H) + W
        v.x = ((v.x + a.x) + b.x) + c.x; v.y = ((v.y + a.y) + b.y) + c.y;
        v.z = ((v.z + a.z) + b.z) + c.z; v.w = ((v.w + a.w) + b.w) + c.w;
        reinterpret_cast<float4*>(x)[i] = v;
    }
}

template <bool F32>
__global__ void __launch_bounds__(kEwThreads) upsample2x_kernel(const float* __restrict__ x, void* __restrict__ y, int F,
                                                                int H, int W, int C) {
    grid_dep_launch();
    grid_dep_wait();
    const int c4n = C >> 2;
    const int H2 = 2 * H, W2 = 2 * W;
    const int64_t total = (int64_t)F * H2 * W2 * c4n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        int64_t pos = i / c4n;
        const int w = (int)(pos % W2); pos /= W2;
        const int h = (int)(pos % H2);
        const int f = (int)(pos / H2);
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)f * H + (h >> 1)) * W + (w >> 1)) * C) + c4);
        store_operand4<F32>(y, (size_t)i, v.x, v.y, v.z, v.w);
    }
}

template <bool F32>
__global__ void __launch_bounds__(kEwThreads) cast_bf16_kernel(const float* __restrict__ x, void* __restrict__ y, int S,
                                                               int64_t RC4, int64_t in_stride) {
    grid_dep_launch();
    grid_dep_wait();
    const int64_t total = (int64_t)S * RC4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / RC4, e = i - s * RC4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + s * in_stride) + e);
        store_operand4<F32>(y, (size_t)i, v.x, v.y, v.z, v.w);
    }
}

__global__ void __launch_bounds__(kEwThreads) parity_split_kernel(const float* __restrict__ x, bf16* __restrict__ y,
                                                                  int F, int H, int W, int C) {
    grid_dep_launch();
    grid_dep_wait();
    const int c4n = C >> 2;
    const int H2 = H >> 1, W2 = W >> 1;
    const int64_t total = (int64_t)F * H * W * c4n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        int64_t pos = i / c4n;
        const int w = (int)(pos % W); pos /= W;
        const int h = (int)(pos % H);
        const int f = (int)(pos / H);
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const int plane = (h & 1) * 2 + (w & 1);
        const size_t o = ((((size_t)f * 4 + plane) * H2 + (h >> 1)) * W2 + (w >> 1)) * c4n + c4;
        uint2 pk;
        pk.x = pack_bf16x2(v.x, v.y);
        pk.y = pack_bf16x2(v.z, v.w);
        reinterpret_cast<uint2*>(y)[o] = pk;
    }
}

__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, const int* __restrict__ step, int t_stride,
                                          float* __restrict__ out, int B, int dim) {
    grid_dep_launch();
    grid_dep_wait();
    if (step) t += (size_t)(*step) * t_stride;
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, k = i - b * half;
    // freqs = exp(-ln(10000) * k / half) in fp32, args = t * freqs (models/utils.py:78-81)
    const float freq = expf(-logf(10000.0f) * (float)k / (float)half);
    const float arg = (float)t[b] * freq;
    out[(size_t)b * dim + k] = cosf(arg);
    out[(size_t)b * dim + half + k] = sinf(arg);
}

// one warp per output feature n; loops over the (small) batch
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ in, const float* __restrict__ Wt,
                                                           const float* __restrict__ bias, float* __restrict__ out, int B,
                                                           int K, int N, int in_silu, int out_silu) {
    grid_dep_launch();
    grid_dep_wait();
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const float* w = Wt + (size_t)n * K;
    for (int b = 0; b < B; ++b) {
        const float* x = in + (size_t)b * K;
        float acc = 0.f;
        for (int k = lane; k < K; k += 32) {
            float xv = x[k];
            if (in_silu) xv = silu_f(xv);
            acc = fmaf(xv, __ldg(w + k), acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            float r = acc + (bias ? bias[n] : 0.f);
            out[(size_t)b * N + n] = out_silu ? silu_f(r) : r;
        }
    }
}

// VAE encoder conv_in: single input channel, 3x3, pad 1. One thread per (pixel, 4 output channels).
__global__ void __launch_bounds__(kEwThreads) conv3x3_c1_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, float* __restrict__ y,
                                                                   int F, int H, int W, int Cout) {
    grid_dep_launch();
    grid_dep_wait();
    const int c4n = Cout >> 2;
    const int64_t total = (int64_t)F * H * W * c4n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        int64_t pos = i / c4n;
        const int px = (int)(pos % W); pos /= W;
        const int py = (int)(pos % H);
        const int f = (int)(pos / H);
        float acc[4] = {bias[c], bias[c + 1], bias[c + 2], bias[c + 3]};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int yy = py + ky - 1, xx = px + kx - 1;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    const float v = __ldg(x + ((size_t)f * H + yy) * W + xx);
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[k] = fmaf(v, __ldg(w + (size_t)(c + k) * 9 + ky * 3 + kx), acc[k]);
                }
            }
        reinterpret_cast<float4*>(y)[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

// VAE decoder conv_out: Cin -> 1. One warp per output pixel; lanes split the channels.
__global__ void __launch_bounds__(256) conv3x3_c1_out_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                             float bias, float* __restrict__ y, int F, int H, int W,
                                                             int Cin) {
    grid_dep_launch();
    grid_dep_wait();
    const int64_t total = (int64_t)F * H * W;
    const int lane = threadIdx.x & 31;
    for (int64_t pix = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; pix < total;
         pix += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int px = (int)(pix % W);
        const int py = (int)((pix / W) % H);
        const int64_t f = pix / ((int64_t)W * H);
        float acc = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const bf16* src = x + (((size_t)f * H + yy) * W + xx) * Cin;
            const float* wt = w + (size_t)tap * Cin;
            for (int c = lane * 4; c < Cin; c += 128) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(src + c));
                const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
                const float4 ww = __ldg(reinterpret_cast<const float4*>(wt + c));
                acc = fmaf(a.x, ww.x, acc); acc = fmaf(a.y, ww.y, acc);
                acc = fmaf(b.x, ww.z, acc); acc = fmaf(b.y, ww.w, acc);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) y[pix] = acc + bias;
    }
}

template <bool F32>
__device__ __forceinline__ void store_operand1(void* y, int64_t i, float v) {
    if constexpr (F32) reinterpret_cast<float*>(y)[i] = tf32_rna(v);
    else reinterpret_cast<bf16*>(y)[i] = __float2bfloat16_rn(v);
}

template <bool F32>
__global__ void pack_linear_kernel(const float* __restrict__ w, void* __restrict__ out, int N, int K, int Kpad) {
    const int64_t total = (int64_t)N * Kpad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % Kpad);
        const int64_t n = i / Kpad;
        store_operand1<F32>(out, i, k < K ? w[n * K + k] : 0.f);
    }
}

template <bool F32>
__global__ void pack_conv_kernel(const float* __restrict__ w, void* __restrict__ out, int Co, int Ci, int taps,
                                 int Cipad) {
    const int64_t total = (int64_t)Co * taps * Cipad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cipad);
        const int tap = (int)((i / Cipad) % taps);
        const int64_t co = i / ((int64_t)Cipad * taps);
        store_operand1<F32>(out, i, ci < Ci ? w[(co * Ci + ci) * taps + tap] : 0.f);
    }
}

}  // namespace

int unet_assemble(const float* x, const float* cond, float* out_f32, void* out_bf16, int B, int Tx, int Tc, int HW,
                  int C, int Cpad, cudaStream_t st, int op_f32) {
    PD_CHECK(C % 4 == 0 && Cpad % 4 == 0 && Cpad > C, PD_ERR_SHAPE, "unet_assemble: C=%d Cpad=%d", C, Cpad);
    const int64_t total = (int64_t)B * (Tx + Tc) * HW * (Cpad / 4);
    if (op_f32) PD_LAUNCH(unet_assemble_kernel<true>, ew_blocks(total), kEwThreads, 0, st, x, cond, out_f32, out_bf16, B, Tx, Tc, HW, C, Cpad);
    else PD_LAUNCH(unet_assemble_kernel<false>, ew_blocks(total), kEwThreads, 0, st, x, cond, out_f32, out_bf16, B, Tx, Tc, HW, C, Cpad);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int pos_embed_add(float* x, const float* Te, const float* He, const float* We, int B, int T, int H, int W, int C,
                  cudaStream_t st) {
    PD_CHECK(C % 4 == 0, PD_ERR_SHAPE, "pos_embed_add: C=%d", C);
    const int64_t total = (int64_t)B * T * H * W * (C / 4);
    PD_LAUNCH(pos_embed_kernel, ew_blocks(total), kEwThreads, 0, st, x, Te, He, We, B, T, H, W, C);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int upsample2x_cast(const float* x, void* y, int F, int H, int W, int C, cudaStream_t st, int y_f32) {
    PD_CHECK(C % 4 == 0, PD_ERR_SHAPE, "upsample2x_cast: C=%d", C);
    const int64_t total = (int64_t)F * 4 * H * W * (C / 4);
    if (y_f32) PD_LAUNCH(upsample2x_kernel<true>, ew_blocks(total), kEwThreads, 0, st, x, y, F, H, W, C);
    else PD_LAUNCH(upsample2x_kernel<false>, ew_blocks(total), kEwThreads, 0, st, x, y, F, H, W, C);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int cast_bf16(const float* x, void* y, int S, int64_t RC, int64_t in_sample_stride, cudaStream_t st, int y_f32) {
    PD_CHECK(RC % 4 == 0 && in_sample_stride % 4 == 0, PD_ERR_SHAPE, "cast_bf16: sizes must be multiples of 4");
    const int64_t total = (int64_t)S * (RC / 4);
    if (y_f32) PD_LAUNCH(cast_bf16_kernel<true>, ew_blocks(total), kEwThreads, 0, st, x, y, S, RC / 4, in_sample_stride);
    else PD_LAUNCH(cast_bf16_kernel<false>, ew_blocks(total), kEwThreads, 0, st, x, y, S, RC / 4, in_sample_stride);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int parity_split_cast(const float* x, bf16* y, int F, int H, int W, int C, cudaStream_t st) {
    PD_CHECK(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, PD_ERR_SHAPE, "parity_split_cast: shape");
    const int64_t total = (int64_t)F * H * W * (C / 4);
    PD_LAUNCH(parity_split_kernel, ew_blocks(total), kEwThreads, 0, st, x, y, F, H, W, C);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int timestep_embedding(const int64_t* t, const int* step, int t_stride, float* out, int B, int dim,
                       cudaStream_t st) {
    PD_CHECK(dim % 2 == 0, PD_ERR_SHAPE, "timestep_embedding: dim must be even");
    const int total = B * (dim / 2);
    PD_LAUNCH(timestep_embedding_kernel, ceil_div(total, 128), 128, 0, st, t, step, t_stride ? t_stride : B, out, B, dim);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int small_linear(const float* in, const float* W, const float* bias, float* out, int B, int K, int N, int in_silu,
                 int out_silu, cudaStream_t st) {
    PD_LAUNCH(small_linear_kernel, ceil_div(N, 8), 256, 0, st, in, W, bias, out, B, K, N, in_silu, out_silu);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int conv3x3_c1_in(const float* x, const float* w, const float* bias, float* y, int F, int H, int W, int Cout,
                  cudaStream_t st) {
    PD_CHECK(Cout % 4 == 0, PD_ERR_SHAPE, "conv3x3_c1_in: Cout=%d", Cout);
    const int64_t total = (int64_t)F * H * W * (Cout / 4);
    PD_LAUNCH(conv3x3_c1_in_kernel, ew_blocks(total), kEwThreads, 0, st, x, w, bias, y, F, H, W, Cout);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int conv3x3_c1_out(const bf16* x, const float* w, float bias, float* y, int F, int H, int W, int Cin, cudaStream_t st) {
    PD_CHECK(Cin % 4 == 0, PD_ERR_SHAPE, "conv3x3_c1_out: Cin=%d", Cin);
    const int64_t total = (int64_t)F * H * W * 32;
    PD_LAUNCH(conv3x3_c1_out_kernel, ew_blocks(total), 256, 0, st, x, w, bias, y, F, H, W, Cin);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int pack_linear(const float* w, void* out, int N, int K, int Kpad, cudaStream_t st, int out_f32) {
    if (out_f32) pack_linear_kernel<true><<<ew_blocks((int64_t)N * Kpad), kEwThreads, 0, st>>>(w, out, N, K, Kpad);
    else pack_linear_kernel<false><<<ew_blocks((int64_t)N * Kpad), kEwThreads, 0, st>>>(w, out, N, K, Kpad);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int pack_conv(const float* w, void* out, int Co, int Ci, int taps, int Cipad, cudaStream_t st, int out_f32) {
    if (out_f32) pack_conv_kernel<true><<<ew_blocks((int64_t)Co * taps * Cipad), kEwThreads, 0, st>>>(w, out, Co, Ci, taps, Cipad);
    else pack_conv_kernel<false><<<ew_blocks((int64_t)Co * taps * Cipad), kEwThreads, 0, st>>>(w, out, Co, Ci, taps, Cipad);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
