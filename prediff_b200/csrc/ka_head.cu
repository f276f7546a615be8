// Read-out head of the knowledge-alignment network and the alignment loss, forward and backward.
// Reference: NoisyCuboidTransformerEncoder.forward read-out (knowledge_alignment/models.py:500-528) =
// per frame GroupNorm(32)+SiLU -> AttentionPool3d (models.py:49-104: mean token + positional embedding, 1x1 qkv,
// QKVAttention models.py:19-46, token 0 -> c_proj), and SEVIRAvgIntensityAlignment.alignment_fn (sevir.py:76-83).
// Only query token 0 is ever read, so the pool is one softmax row per (frame, head).
#include "ops.cuh"

namespace pd {
namespace {

constexpr int kMaxTok = 96;

// tok[f][0][c] = mean_p o[f][p][c] + pos[0][c]; tok[f][1+p][c] = o[f][p][c] + pos[1+p][c], o = silu(GN(x)).
// One block per frame; thread = (token group, channel): kTokGroups groups walk the tokens interleaved (coalesced rows), their
// partial sums of the mean token are combined in group order (deterministic). (One group per frame - 24 blocks of 256 threads,
// each thread a serial walk over all tokens - took 41 us)
constexpr int kTokGroups = 4;
__global__ void __launch_bounds__(256 * kTokGroups) ka_tokens_kernel(const float* __restrict__ x, const double* __restrict__ sums,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     const float* __restrict__ pos /*[1+R][C]*/,
                                                                     bf16* __restrict__ tok, int R, int C, int G, float eps) {
    extern __shared__ float s_acc[];   // [kTokGroups][C]
    const int f = blockIdx.x;
    const int tg = threadIdx.x / 256, c0 = threadIdx.x % 256;
    const int cpg = C / G;
    const double n = (double)R * cpg;
    for (int c = c0; c < C; c += 256) {
        const int g = c / cpg;
        const double m = sums[((size_t)f * G + g) * 2] / n;
        double var = sums[((size_t)f * G + g) * 2 + 1] / n - m * m;
        if (var < 0) var = 0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float sc = rstd * gamma[c], sh = beta[c] - (float)m * sc;
        const float* xp = x + (size_t)f * R * C + c;
        bf16* tp = tok + (size_t)f * (R + 1) * C + c;
        float acc = 0.f;
        for (int p = tg; p < R; p += kTokGroups) {
            const float u = fmaf(xp[(size_t)p * C], sc, sh);
            const float o = u / (1.0f + __expf(-u));
            acc += o;
            tp[(size_t)(p + 1) * C] = __float2bfloat16_rn(o + pos[(size_t)(p + 1) * C + c]);
        }
        s_acc[tg * C + c] = acc;
    }
    __syncthreads();
    if (tg == 0) {
        for (int c = c0; c < C; c += 256) {
            float acc = s_acc[c];
#pragma unroll
            for (int k = 1; k < kTokGroups; ++k) acc += s_acc[k * C + c];
            tok[(size_t)f * (R + 1) * C + c] = __float2bfloat16_rn(acc / (float)R + pos[c]);
        }
    }
}

// One block per frame, kSub warps per head (the tokens of a head interleaved over them; partial results combined in warp
// order: deterministic). qkv fp32 [F][L][3C] (q | k | v, head-major channels).
// w = softmax_s(scale^2 q0 . k_s); a = sum_s w_s v_s; out[f] = cw . a + cb. Saves w for the backward.
constexpr int kSub = 4;
__global__ void __launch_bounds__(32 * 8 * kSub) ka_pool_kernel(const float* __restrict__ qkv, const float* __restrict__ cw,
                                                                float cb, float* __restrict__ wsave, float* __restrict__ out,
                                                                int L, int C, int heads) {
    __shared__ float s_w[8][kMaxTok];
    __shared__ float s_a[8][kSub][128];
    __shared__ float s_part[8];
    const int f = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int h = wid / kSub, sub = wid - h * kSub;   // blockDim = 32 * kSub * heads
    const int ch = C / heads;
    const float scale2 = rsqrtf((float)ch);   // (1/sqrt(sqrt(ch)))^2, models.py:37
    const float* base = qkv + (size_t)f * L * 3 * C;
    float q[4];
    for (int i = 0; i < 4; ++i) q[i] = (lane + 32 * i < ch) ? base[h * ch + lane + 32 * i] : 0.f;
    for (int s = sub; s < L; s += kSub) {
        const float* k = base + (size_t)s * 3 * C + C + h * ch;
        float d = 0.f;
        for (int i = 0; i < 4; ++i)
            if (lane + 32 * i < ch) d = fmaf(q[i], k[lane + 32 * i], d);
        d = warp_sum(d) * scale2;
        if (lane == 0) s_w[h][s] = d;
    }
    __syncthreads();
    if (sub == 0) {   // softmax over the head's L logits
        float mx = -INFINITY;
        for (int s = lane; s < L; s += 32) mx = fmaxf(mx, s_w[h][s]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int s = lane; s < L; s += 32) {
            const float e = __expf(s_w[h][s] - mx);
            s_w[h][s] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int s = lane; s < L; s += 32) {
            const float w = s_w[h][s] * inv;
            s_w[h][s] = w;
            wsave[((size_t)f * heads + h) * L + s] = w;
        }
    }
    __syncthreads();
    for (int i = 0; i < 4; ++i) {
        const int d = lane + 32 * i;
        float a = 0.f;
        if (d < ch)
            for (int s = sub; s < L; s += kSub) a = fmaf(s_w[h][s], base[(size_t)s * 3 * C + 2 * C + h * ch + d], a);
        if (d < 128) s_a[h][sub][d] = a;
    }
    __syncthreads();
    if (sub == 0) {
        float part = 0.f;
        for (int i = 0; i < 4; ++i) {
            const int d = lane + 32 * i;
            if (d < ch) {
                float a = s_a[h][0][d];
#pragma unroll
                for (int k = 1; k < kSub; ++k) a += s_a[h][k][d];
                part = fmaf(cw[h * ch + d], a, part);
            }
        }
        part = warp_sum(part);
        if (lane == 0) s_part[h] = part;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = cb;
        for (int i = 0; i < heads; ++i) r += s_part[i];
        out[f] = r;
    }
}

// dout[f] = guide_scale * d || mean_T out - target ||_2 / d out[f]  (norm over the whole batch, sevir.py:82)
__global__ void ka_loss_grad_kernel(const float* __restrict__ out, const float* __restrict__ target,
                                    float* __restrict__ dout, float* __restrict__ loss, int B, int T, float guide_scale) {
    __shared__ float s_diff[1024];
    __shared__ float s_norm;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float m = 0.f;
        for (int t = 0; t < T; ++t) m += out[b * T + t];
        s_diff[b] = m / (float)T - target[b];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ss = 0.f;
        for (int b = 0; b < B; ++b) ss = fmaf(s_diff[b], s_diff[b], ss);
        s_norm = sqrtf(ss);
        if (loss) *loss = s_norm;
    }
    __syncthreads();
    const float inv = s_norm > 0.f ? guide_scale / (s_norm * (float)T) : 0.f;
    for (int i = threadIdx.x; i < B * T; i += blockDim.x) dout[i] = s_diff[i / T] * inv;
}

// Backward of ka_pool: dqkv bf16 [F][L][3C] (dq is non-zero for token 0 only). Same block shape as ka_pool_kernel: kSub warps per
// head share the tokens; the two reductions over tokens (dot, dq) are combined in warp order.
__global__ void __launch_bounds__(32 * 8 * kSub) ka_pool_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ cw,
                                                                    const float* __restrict__ wsave, const float* __restrict__ dout,
                                                                    bf16* __restrict__ dqkv, int L, int C, int heads) {
    __shared__ float s_dl[8][kMaxTok];
    __shared__ float s_dq[8][kSub][128];
    const int f = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int h = wid / kSub, sub = wid - h * kSub;
    const int ch = C / heads;
    const float scale2 = rsqrtf((float)ch);
    const float* base = qkv + (size_t)f * L * 3 * C;
    bf16* obase = dqkv + (size_t)f * L * 3 * C;
    const float* w = wsave + ((size_t)f * heads + h) * L;
    const float go = dout[f];
    float da[4], q[4];
    for (int i = 0; i < 4; ++i) {
        const int d = lane + 32 * i;
        da[i] = d < ch ? go * cw[h * ch + d] : 0.f;
        q[i] = d < ch ? base[h * ch + d] : 0.f;
    }
    // dw_s = da . v_s
    for (int s = sub; s < L; s += kSub) {
        const float* v = base + (size_t)s * 3 * C + 2 * C + h * ch;
        float d = 0.f;
        for (int i = 0; i < 4; ++i)
            if (lane + 32 * i < ch) d = fmaf(da[i], v[lane + 32 * i], d);
        d = warp_sum(d);
        if (lane == 0) s_dl[h][s] = d;
    }
    __syncthreads();
    // dot = sum_s w_s dw_s, in token order on every warp of the head (identical values)
    float dot = 0.f;
    for (int s = 0; s < L; ++s) dot = fmaf(w[s], s_dl[h][s], dot);
    __syncthreads();
    if (sub == 0)
        for (int s = lane; s < L; s += 32) s_dl[h][s] = w[s] * (s_dl[h][s] - dot) * scale2;   // d logit_s * scale^2
    __syncthreads();
    for (int i = 0; i < 4; ++i) {
        const int d = lane + 32 * i;
        float dq = 0.f;
        if (d < ch) {
            for (int s = sub; s < L; s += kSub) {
                const float dl = s_dl[h][s];
                dq = fmaf(dl, base[(size_t)s * 3 * C + C + h * ch + d], dq);
                bf16* o = obase + (size_t)s * 3 * C + h * ch + d;
                if (s > 0) o[0] = __float2bfloat16_rn(0.f);
                o[C] = __float2bfloat16_rn(dl * q[i]);
                o[2 * C] = __float2bfloat16_rn(w[s] * da[i]);
            }
        }
        if (d < 128) s_dq[h][sub][d] = dq;
    }
    __syncthreads();
    if (sub == 0) {
        for (int i = 0; i < 4; ++i) {
            const int d = lane + 32 * i;
            if (d >= ch) continue;
            float dq = s_dq[h][0][d];
#pragma unroll
            for (int k = 1; k < kSub; ++k) dq += s_dq[h][k][d];
            obase[h * ch + d] = __float2bfloat16_rn(dq);
        }
    }
}

// do[f][p][c] = dtok[f][1+p][c] + dtok[f][0][c] / R
__global__ void __launch_bounds__(256) ka_tokens_bwd_kernel(const float* __restrict__ dtok, float* __restrict__ dout, int F,
                                                            int R, int C) {
    const int c4n = C >> 2;
    const int64_t total = (int64_t)F * R * c4n;
    const float inv = 1.0f / (float)R;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        const int64_t fp = i / c4n;
        const int p = (int)(fp % R);
        const int64_t f = fp / R;
        const float4 a = __ldg(reinterpret_cast<const float4*>(dtok + ((size_t)f * (R + 1) + p + 1) * C) + c4);
        const float4 m = __ldg(reinterpret_cast<const float4*>(dtok + ((size_t)f * (R + 1)) * C) + c4);
        reinterpret_cast<float4*>(dout)[i] = make_float4(fmaf(m.x, inv, a.x), fmaf(m.y, inv, a.y), fmaf(m.z, inv, a.z),
                                                         fmaf(m.w, inv, a.w));
    }
}

__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
    const int64_t total = (int64_t)R * C;   // in [R][C] -> out [C][R]
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i % R);
        const int64_t c = i / R;
        out[i] = in[(int64_t)r * C + c];
    }
}

}  // namespace

int ka_tokens(const float* x, const double* sums, const float* gamma, const float* beta, const float* pos_tc, bf16* tok,
              int F, int R, int C, int G, float eps, cudaStream_t st) {
    PD_CHECK(G > 0 && C % G == 0, PD_ERR_SHAPE, "ka_tokens: C=%d G=%d", C, G);
    ka_tokens_kernel<<<F, 256 * kTokGroups, (size_t)kTokGroups * C * sizeof(float), st>>>(x, sums, gamma, beta, pos_tc, tok, R, C, G, eps);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int ka_pool(const float* qkv, const float* cw, float cb, float* wsave, float* out, int F, int L, int C, int heads,
            cudaStream_t st) {
    PD_CHECK(heads >= 1 && heads <= 8 && C % heads == 0 && C / heads <= 128 && L <= kMaxTok, PD_ERR_SHAPE,
             "ka_pool: C=%d heads=%d L=%d", C, heads, L);
    ka_pool_kernel<<<F, 32 * kSub * heads, 0, st>>>(qkv, cw, cb, wsave, out, L, C, heads);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int ka_loss_grad(const float* out, const float* target, float* dout, float* loss, int B, int T, float guide_scale,
                 cudaStream_t st) {
    PD_CHECK(B >= 1 && B <= 1024, PD_ERR_SHAPE, "ka_loss_grad: batch %d", B);
    ka_loss_grad_kernel<<<1, 256, 0, st>>>(out, target, dout, loss, B, T, guide_scale);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int ka_pool_bwd(const float* qkv, const float* cw, const float* wsave, const float* dout, bf16* dqkv, int F, int L, int C,
                int heads, cudaStream_t st) {
    PD_CHECK(heads >= 1 && heads <= 8 && C % heads == 0 && C / heads <= 128 && L <= kMaxTok, PD_ERR_SHAPE,
             "ka_pool_bwd: C=%d heads=%d L=%d", C, heads, L);
    ka_pool_bwd_kernel<<<F, 32 * kSub * heads, 0, st>>>(qkv, cw, wsave, dout, dqkv, L, C, heads);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int ka_tokens_bwd(const float* dtok, float* dout, int F, int R, int C, cudaStream_t st) {
    PD_CHECK(C % 4 == 0, PD_ERR_SHAPE, "ka_tokens_bwd: C=%d", C);
    const int64_t total = (int64_t)F * R * (C / 4);
    int blocks = (int)ceil_div64(total, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    ka_tokens_bwd_kernel<<<blocks, 256, 0, st>>>(dtok, dout, F, R, C);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int transpose_f32(const float* in, float* out, int R, int C, cudaStream_t st) {
    int blocks = (int)ceil_div64((int64_t)R * C, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    transpose_f32_kernel<<<blocks, 256, 0, st>>>(in, out, R, C);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
