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
    // one source quad per thread iteration -> its 2 x 2 copies (the first version walked the OUTPUT: four loads of every
    // source quad and one load in flight per thread; 83 us for a 120 MB VAE level = 1.4 TB/s)
    const int c4n = C >> 2;
    const int W2 = 2 * W;
    const int64_t total = (int64_t)F * H * W * c4n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        int64_t pos = i / c4n;
        const int w = (int)(pos % W); pos /= W;
        const int h = (int)(pos % H);
        const int f = (int)(pos / H);
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const size_t o = ((((size_t)f * 2 * H + 2 * h) * W2 + 2 * w) * c4n) + c4;   // output quad index of the top-left copy
        store_operand4<F32>(y, o, v.x, v.y, v.z, v.w);
        store_operand4<F32>(y, o + c4n, v.x, v.y, v.z, v.w);
        store_operand4<F32>(y, o + (size_t)W2 * c4n, v.x, v.y, v.z, v.w);
        store_operand4<F32>(y, o + (size_t)W2 * c4n + c4n, v.x, v.y, v.z, v.w);
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

// VAE encoder conv_in: single input channel, 3x3, pad 1 - write-bound (4 * Cout bytes per pixel out, 4 bytes in). A thread
// owns 4 output channels (fixed for its lifetime: the 36 weights + bias live in registers) and walks segments of 8 pixels
// along W: the segment's 3 x 10 input window is loaded once (broadcast across the 32 lanes that share the segment), every
// pixel is 36 FMAs and one 16-byte store, a warp writing 512 contiguous bytes per pixel. (Round 1 read its 36 weights per
// thread through __ldg with a 36-byte stride: 541 us for 28 frames = 5 % of the HBM peak; the first rewrite - one pixel per
// thread, weights from shared memory per tap - was bound by its 45 load wavefronts per 512 bytes stored: 193 us, 14 %.)
constexpr int kC1Seg = 8;
__global__ void __launch_bounds__(kEwThreads) conv3x3_c1_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, float* __restrict__ y,
                                                                   int F, int H, int W, int Cout) {
    const int c4n = Cout >> 2;                          // blockDim.x and the grid stride are multiples of c4n (host check)
    const int c = (threadIdx.x % c4n) * 4;
    float wr[9][4], br[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        br[k] = bias[c + k];
#pragma unroll
        for (int t = 0; t < 9; ++t) wr[t][k] = w[(size_t)(c + k) * 9 + t];
    }
    grid_dep_launch();
    grid_dep_wait();
    const int segs = (W + kC1Seg - 1) / kC1Seg;
    const int64_t total = (int64_t)F * H * segs * c4n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t pos = i / c4n;
        const int px0 = (int)(pos % segs) * kC1Seg; pos /= segs;
        const int py = (int)(pos % H);
        const int f = (int)(pos / H);
        const float* xf = x + (size_t)f * H * W;
        float win[3][kC1Seg + 2];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = py + ky - 1;
#pragma unroll
            for (int j = 0; j < kC1Seg + 2; ++j) {
                const int xx = px0 + j - 1;
                win[ky][j] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(xf + (size_t)yy * W + xx) : 0.f;
            }
        }
        float4* out = reinterpret_cast<float4*>(y + (((size_t)f * H + py) * W + px0) * Cout + c);
#pragma unroll
        for (int p = 0; p < kC1Seg; ++p) {
            if (px0 + p < W) {
                float a0 = br[0], a1 = br[1], a2 = br[2], a3 = br[3];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {   // same tap order as before: results are bit-identical
                        const float v = win[ky][p + kx];
                        a0 = fmaf(v, wr[ky * 3 + kx][0], a0); a1 = fmaf(v, wr[ky * 3 + kx][1], a1);
                        a2 = fmaf(v, wr[ky * 3 + kx][2], a2); a3 = fmaf(v, wr[ky * 3 + kx][3], a3);
                    }
                out[(size_t)p * c4n] = make_float4(a0, a1, a2, a3);
            }
        }
    }
}

// VAE decoder conv_out: Cin -> 1, 3x3, pad 1 - read-bound (2 * Cin bytes per pixel in, 4 bytes out). A block owns a
// 14 x 14 output tile = a 16 x 16 input tile with halo, one input pixel per thread: the thread reads its pixel's Cin
// channels once (contiguous) and forms the pixel's 9 per-tap dot products against the weights in shared memory (broadcast
// LDS.128); the 14 x 14 outputs are then sums of 9 shared-memory values. Every activation is read once (plus the 31 % halo)
// instead of nine times through a warp-per-pixel reduction (round 1: 461 us for 24 frames = 4 % of the HBM peak).
constexpr int kCoTile = 14;
__global__ void __launch_bounds__(256) conv3x3_c1_out_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                             float bias, float* __restrict__ y, int F, int H, int W,
                                                             int Cin) {
    extern __shared__ float s_co[];   // weights [9][Cin], then per-pixel taps [256][9]
    float* s_w = s_co;
    float* s_t = s_co + 9 * Cin;
    for (int i = threadIdx.x; i < 9 * Cin; i += blockDim.x) s_w[i] = w[i];
    grid_dep_launch();
    grid_dep_wait();
    __syncthreads();
    const int tiles_x = (W + kCoTile - 1) / kCoTile, tiles_y = (H + kCoTile - 1) / kCoTile;
    const int f = blockIdx.x / (tiles_x * tiles_y);
    const int tt = blockIdx.x - f * tiles_x * tiles_y;
    const int y0 = (tt / tiles_x) * kCoTile, x0 = (tt % tiles_x) * kCoTile;
    const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
    const int iy = y0 + ly - 1, ix = x0 + lx - 1;
    float t[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) t[k] = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const uint4* src = reinterpret_cast<const uint4*>(x + (((size_t)f * H + iy) * W + ix) * Cin);
        for (int c8 = 0; c8 < Cin / 8; ++c8) {
            const uint4 u = __ldg(src + c8);
            const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float4 w0 = *reinterpret_cast<const float4*>(s_w + k * Cin + c8 * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(s_w + k * Cin + c8 * 8 + 4);
                t[k] = fmaf(a0.x, w0.x, fmaf(a0.y, w0.y, fmaf(a1.x, w0.z, fmaf(a1.y, w0.w, t[k]))));
                t[k] = fmaf(a2.x, w1.x, fmaf(a2.y, w1.y, fmaf(a3.x, w1.z, fmaf(a3.y, w1.w, t[k]))));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) s_t[threadIdx.x * 9 + k] = t[k];
    __syncthreads();
    // output (oy, ox) of the tile = input-tile position (oy + 1, ox + 1); tap (ky, kx) reads input (oy + ky, ox + kx)
    if (ly < kCoTile && lx < kCoTile) {
        const int oy = y0 + ly, ox = x0 + lx;
        if (oy < H && ox < W) {
            float acc = bias;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) acc += s_t[((ly + ky) * 16 + (lx + kx)) * 9 + ky * 3 + kx];
            y[((size_t)f * H + oy) * W + ox] = acc;
        }
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
    const int64_t total = (int64_t)F * H * W * (C / 4);
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
    const int64_t total = (int64_t)F * H * ceil_div(W, kC1Seg) * (Cout / 4);
    PD_CHECK(Cout <= 1024 && kEwThreads % (Cout / 4) == 0, PD_ERR_SHAPE, "conv3x3_c1_in: Cout=%d (Cout / 4 must divide %d)", Cout,
             kEwThreads);
    PD_LAUNCH(conv3x3_c1_in_kernel, ew_blocks(total), kEwThreads, 0, st, x, w, bias, y, F, H, W, Cout);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int conv3x3_c1_out(const bf16* x, const float* w, float bias, float* y, int F, int H, int W, int Cin, cudaStream_t st) {
    PD_CHECK(Cin % 8 == 0 && Cin <= 512, PD_ERR_SHAPE, "conv3x3_c1_out: Cin=%d (multiple of 8, <= 512)", Cin);
    const int tiles = ceil_div(W, kCoTile) * ceil_div(H, kCoTile);
    PD_LAUNCH(conv3x3_c1_out_kernel, F * tiles, 256, (size_t)(9 * Cin + 256 * 9) * sizeof(float), st, x, w, bias, y, F, H, W, Cin);
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
