// Global vectors of the cuboid attention (reference: cuboid_transformer.py:864-945 CuboidSelfAttentionLayer with
// use_global_vector / separate_global_qkv=False; :1130-1145 the stack block's residual + global FFN;
// cuboid_transformer_unet.py:124-126, 432-434, 449-450, 489-490 the UNet's init_global_vectors and level projections).
// K vectors per sample (K <= 32) ride beside the token grid: B * K rows of width C, three orders of magnitude fewer rows
// than the grid has tokens, so their linears are fp32 CUDA-core kernels on the fp32 weights as loaded (no repack, no
// operand rounding); the one kernel with real work is the global queries' attention over all num_cuboids * volume slots.
#include "ops.cuh"
#include <algorithm>
#include <climits>

namespace pd {
namespace {

constexpr int kGvRows = 16;      // rows per block of gv_linear
constexpr int kGvKChunk = 512;   // K elements of those rows staged in shared memory at a time
constexpr int kGvWarps = 8;      // one output column per warp

// out[r][n] = (res ? res[r][n] : 0) + act( sum_k f(in[r][k]) W[n][k] + bias[n] ),  f = LayerNorm(ln_gamma, ln_beta) or id.
// One block: kGvRows rows x 8 columns. The rows' inputs are staged in shared memory (normalised on the way in); a warp owns
// one output column: its lanes walk the weight row with 16-byte loads (coalesced, read once per row group) against 16-byte
// shared-memory reads of the 16 rows, and the 16 row sums are reduced by shuffles. (The first version gave a warp four
// columns and scalar loads: 32 blocks for a 32 x 2048 x 512 problem, 59 us; grid = N / 8 x M / 16 now.)
__global__ void __launch_bounds__(kGvWarps * 32) gv_linear_kernel(const float* __restrict__ in, const float* __restrict__ ln_gamma,
                                                                  const float* __restrict__ ln_beta, const float* __restrict__ W,
                                                                  const float* __restrict__ bias, const float* res,
                                                                  float* out_f32, bf16* __restrict__ out_bf16, int M, int K, int N,
                                                                  int act, float ln_eps) {
    grid_dep_launch();
    grid_dep_wait();
    __shared__ __align__(16) float s_in[kGvRows][kGvKChunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.y * kGvRows;
    const int n = blockIdx.x * kGvWarps + warp;
    float acc[kGvRows];
#pragma unroll
    for (int r = 0; r < kGvRows; ++r) acc[r] = 0.f;
    for (int k0 = 0; k0 < K; k0 += kGvKChunk) {
        const int kc = min(kGvKChunk, K - k0);   // K % 4 == 0 (host check): kc is a multiple of 4
        // the warp's slice of its weight row: all loads issued before the staging below (the fp32 weights of the global path
        // are HBM-cold every step; one 16-byte load at a time per lane was latency-bound)
        float4 wreg[kGvKChunk / 128];
        if (n < N) {
            const float* w = W + (size_t)n * K + k0;
#pragma unroll
            for (int i = 0; i < kGvKChunk / 128; ++i) {
                const int k = 4 * lane + 128 * i;
                wreg[i] = k < kc ? __ldg(reinterpret_cast<const float4*>(w + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();
        // stage (and normalise: the host guarantees K <= kGvKChunk then, so the chunk is the whole row)
        for (int r = warp; r < kGvRows; r += kGvWarps) {
            const int row = r0 + r;
            if (row >= M) {
                for (int k = lane; k < kc; k += 32) s_in[r][k] = 0.f;
                continue;
            }
            const float* x = in + (size_t)row * K + k0;
            if (ln_gamma) {   // torch.nn.LayerNorm: biased variance around the mean, eps inside the square root
                float xv[kGvKChunk / 32];   // the row, once, in registers
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < kGvKChunk / 32; ++i) {
                    const int k = lane + 32 * i;
                    xv[i] = k < kc ? x[k] : 0.f;
                    s += xv[i];
                }
                const float mean = warp_sum(s) / (float)kc;
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < kGvKChunk / 32; ++i) {
                    const float d = lane + 32 * i < kc ? xv[i] - mean : 0.f;
                    v = fmaf(d, d, v);
                }
                const float rstd = rsqrtf(warp_sum(v) / (float)kc + ln_eps);
#pragma unroll
                for (int i = 0; i < kGvKChunk / 32; ++i) {
                    const int k = lane + 32 * i;
                    if (k < kc) s_in[r][k] = (xv[i] - mean) * rstd * ln_gamma[k] + ln_beta[k];
                }
            } else {
                for (int k = 4 * lane; k < kc; k += 128)
                    *reinterpret_cast<float4*>(&s_in[r][k]) = *reinterpret_cast<const float4*>(x + k);
            }
        }
        __syncthreads();
        if (n < N) {
#pragma unroll
            for (int i = 0; i < kGvKChunk / 128; ++i) {
                const int k = 4 * lane + 128 * i;
                if (k < kc) {
                    const float4 wv = wreg[i];
#pragma unroll
                    for (int r = 0; r < kGvRows; ++r) {
                        const float4 xv = *reinterpret_cast<const float4*>(&s_in[r][k]);
                        acc[r] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[r]))));
                    }
                }
            }
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int r = 0; r < kGvRows; ++r) {
        const float t = warp_sum(acc[r]);
        if (lane == r) mine = t;
    }
    const int row = r0 + lane;
    if (n < N && lane < kGvRows && row < M) {
        float v = mine + (bias ? bias[n] : 0.f);
        if (act == 1) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));   // nn.GELU() (erf form)
        if (res) v += res[(size_t)row * N + n];
        if (out_f32) out_f32[(size_t)row * N + n] = v;
        if (out_bf16) out_bf16[(size_t)row * N + n] = __float2bfloat16(v);
    }
}

// g[b][k][:] = init[k][:]   (cuboid_transformer_unet.py:432-434: init_global_vectors.expand(batch, K, C))
__global__ void gv_broadcast_kernel(const float* __restrict__ init, float* __restrict__ g, int B, int KC) {
    grid_dep_launch();
    grid_dep_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * KC) g[i] = init[i % KC];
}

constexpr int kGaKeys = 64;      // keys per pass of a block
constexpr int kGaThreads = 128;
constexpr int kMaxGlobal = 32;   // K <= 32

// Global-to-local(+global) attention (cuboid_transformer.py:928-945): the K global queries of (sample b, head h) over a
// range of the key list = the layer's num_cuboids * volume slots in cuboid order (slot -> token row through `tok`; -1 = a
// zero-padded slot whose k = v = 0 still takes part unless `gmask` hides it) followed, with self-attention, by the K global
// keys. grid (splits, heads, B); every block writes an un-normalised partial {m[K], l[K], o[K][HD]} that
// global_attention_combine_kernel merges in split order.
// KMAX: K rounded up to 8 / 16 / 32 - the bound of every per-query loop, so that 8 vectors do not pay the instruction
// stream of 32 (predicated-off FMAs still issue: the first version ran 30 k warp instructions per block, 41 us).
template <int HD, int KMAX>
__global__ void __launch_bounds__(kGaThreads) global_attention_kernel(const GvQuery a, const int* __restrict__ tok,
                                                                      const int* __restrict__ gmask, int n_slots, int n_keys,
                                                                      int keys_per_split, int K, int N, int C, int heads,
                                                                      float* __restrict__ part) {
    grid_dep_launch();
    grid_dep_wait();
    constexpr int QG = kGaThreads / HD;            // query groups of the PV phase (threads = (group, channel))
    constexpr int NACC = KMAX / QG > 0 ? KMAX / QG : 1;          // accumulators per thread
    __shared__ __align__(16) float s_q[KMAX][HD];
    __shared__ float s_p[KMAX][kGaKeys];
    __shared__ __align__(16) bf16 s_v[kGaKeys][HD];   // the pass's value rows (the PV loop must not chase global pointers)
    __shared__ long long s_row[kGaKeys];           // element offset of the key's q|k|v row (-1: zero row), or LLONG_MIN = masked
    __shared__ float s_m[KMAX], s_l[KMAX], s_alpha[KMAX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int C3 = 3 * C;
    const float scale = rsqrtf((float)HD);
    for (int i = tid; i < KMAX * HD; i += kGaThreads) {
        const int g = i / HD, d = i - g * HD;
        s_q[g][d] = g < K ? a.q[((size_t)b * K + g) * a.q_ld + h * HD + d] * scale : 0.f;   // q_global * scale (:899)
    }
    if (tid < KMAX) {
        s_m[tid] = -INFINITY;
        s_l[tid] = 0.f;
    }
    const bf16* base = a.tok_kv + (size_t)b * N * C3 + h * HD;
    const bf16* skb = a.sk ? a.sk + (size_t)b * K * a.s_ld + h * HD : nullptr;   // self-attention: the global keys / values
    const bf16* svb = a.sv ? a.sv + (size_t)b * K * a.s_ld + h * HD : nullptr;
    // separate_global_qkv: the global keys meet another query set (g2g_global_q) than the token keys do; it is read from
    // global memory by the <= 32 threads that hold a global key (one pass of one block per (sample, head))
    const float* sqb = (a.sq && a.sq != a.q) ? a.sq + (size_t)b * K * a.sq_ld + h * HD : nullptr;
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    const int qg = tid / HD, dch = tid - qg * HD;
    const int k_begin = split * keys_per_split, k_end = min(n_keys, k_begin + keys_per_split);
    for (int kb = k_begin; kb < k_end; kb += kGaKeys) {
        __syncthreads();
        // ---- threads 0-63: scores, one thread per key; threads 64-127: the same keys' value rows -> shared memory ----
        {
            const int kt = tid & (kGaKeys - 1);
            const int j = kb + kt;
            long long row = LLONG_MIN;
            int isg = 0;
            if (j < k_end) {
                if (j < n_slots) {
                    if (!gmask || gmask[j]) {
                        const int t = tok[j];
                        row = t >= 0 ? (long long)t * C3 : -1;
                    }
                } else {
                    row = (long long)(j - n_slots) * a.s_ld;
                    isg = 1;
                }
            }
            const bf16* kp = isg ? skb + row : base + (row >= 0 ? row : 0) + C;
            const bf16* vp = isg ? svb + row : base + (row >= 0 ? row : 0) + 2 * C;
            if (tid < kGaKeys) {
                s_row[kt] = row;
                uint4 u[HD / 8];
                if (row >= 0) {
#pragma unroll
                    for (int i = 0; i < HD / 8; ++i) u[i] = *reinterpret_cast<const uint4*>(kp + 8 * i);   // all loads in flight
                }
                float sc[KMAX];
#pragma unroll
                for (int g = 0; g < KMAX; ++g) sc[g] = 0.f;
                if (row >= 0 && isg && sqb) {   // a global key under separate_global_qkv: queries from global memory
#pragma unroll 1
                    for (int g = 0; g < K; ++g) {
                        const float* qg2 = sqb + (size_t)g * a.sq_ld;
                        float acc2 = 0.f;
#pragma unroll
                        for (int i = 0; i < HD / 8; ++i) {
                            const float2 f0 = unpack_bf16x2(u[i].x), f1 = unpack_bf16x2(u[i].y), f2 = unpack_bf16x2(u[i].z),
                                         f3 = unpack_bf16x2(u[i].w);
                            const float4 qa = __ldg(reinterpret_cast<const float4*>(qg2 + 8 * i));
                            const float4 qb = __ldg(reinterpret_cast<const float4*>(qg2 + 8 * i + 4));
                            acc2 = fmaf(qa.x, f0.x, acc2); acc2 = fmaf(qa.y, f0.y, acc2); acc2 = fmaf(qa.z, f1.x, acc2);
                            acc2 = fmaf(qa.w, f1.y, acc2); acc2 = fmaf(qb.x, f2.x, acc2); acc2 = fmaf(qb.y, f2.y, acc2);
                            acc2 = fmaf(qb.z, f3.x, acc2); acc2 = fmaf(qb.w, f3.y, acc2);
                        }
                        s_p[g][kt] = acc2 * scale;
                    }
                } else if (row >= 0) {
#pragma unroll
                    for (int i = 0; i < HD / 8; ++i) {
                        const float2 f0 = unpack_bf16x2(u[i].x), f1 = unpack_bf16x2(u[i].y), f2 = unpack_bf16x2(u[i].z),
                                     f3 = unpack_bf16x2(u[i].w);
#pragma unroll
                        for (int g = 0; g < KMAX; ++g) {   // rows K .. KMAX - 1 of s_q are zero
                            const float4 qa = *reinterpret_cast<const float4*>(&s_q[g][8 * i]);
                            const float4 qb = *reinterpret_cast<const float4*>(&s_q[g][8 * i + 4]);
                            float a = sc[g];
                            a = fmaf(qa.x, f0.x, a); a = fmaf(qa.y, f0.y, a); a = fmaf(qa.z, f1.x, a); a = fmaf(qa.w, f1.y, a);
                            a = fmaf(qb.x, f2.x, a); a = fmaf(qb.y, f2.y, a); a = fmaf(qb.z, f3.x, a); a = fmaf(qb.w, f3.y, a);
                            sc[g] = a;
                        }
                    }
                }
                if (!(row >= 0 && isg && sqb)) {
#pragma unroll
                    for (int g = 0; g < KMAX; ++g) s_p[g][kt] = row == LLONG_MIN ? -INFINITY : sc[g];
                }
            } else {
                uint4* dst = reinterpret_cast<uint4*>(&s_v[kt][0]);
                if (row >= 0) {
#pragma unroll
                    for (int i = 0; i < HD / 8; ++i) dst[i] = *reinterpret_cast<const uint4*>(vp + 8 * i);
                } else {   // masked or a zero-padded slot: v = 0
#pragma unroll
                    for (int i = 0; i < HD / 8; ++i) dst[i] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
        __syncthreads();
        // ---- online softmax: one warp per query (two keys per lane) ----
        for (int g = warp; g < K; g += kGaThreads / 32) {
            const float v0 = s_p[g][lane], v1 = s_p[g][lane + 32];
            const float m_old = s_m[g];
            const float m_new = fmaxf(m_old, warp_max(fmaxf(v0, v1)));
            const float p0 = v0 == -INFINITY ? 0.f : __expf(v0 - m_new), p1 = v1 == -INFINITY ? 0.f : __expf(v1 - m_new);
            const float alpha = m_new == -INFINITY ? 1.f : __expf(m_old - m_new);
            const float sum = warp_sum(p0 + p1);
            s_p[g][lane] = p0;
            s_p[g][lane + 32] = p1;
            __syncwarp();   // every lane has read s_m[g] before lane 0 replaces it
            if (lane == 0) {
                s_m[g] = m_new;
                s_l[g] = s_l[g] * alpha + sum;
                s_alpha[g] = alpha;
            }
        }
        __syncthreads();
        // ---- O += P V: thread = (query group, channel) ----
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            const int g = qg + QG * i;
            if (g < K) acc[i] *= s_alpha[g];
        }
#pragma unroll 4
        for (int j = 0; j < kGaKeys; ++j) {   // masked keys carry p = 0, zero-padded slots v = 0
            const float v = __bfloat162float(s_v[j][dch]);
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                const int g = qg + QG * i;
                if (g < K) acc[i] = fmaf(s_p[g][j], v, acc[i]);
            }
        }
    }
    __syncthreads();
    float* dst = part + (((size_t)b * heads + h) * gridDim.x + split) * (size_t)K * (HD + 2);
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        const int g = qg + QG * i;
        if (g < K) dst[(size_t)g * (HD + 2) + 2 + dch] = acc[i];
    }
    if (tid < K) {
        dst[(size_t)tid * (HD + 2)] = s_m[tid];
        dst[(size_t)tid * (HD + 2) + 1] = s_l[tid];
    }
}

// out[b][g][h * HD + d] = sum_s o_s e^{m_s - M} / sum_s l_s e^{m_s - M}, splits taken in order (deterministic).
// One block per (query g, head, sample), thread = channel: the partials of a split are read coalesced and the loads of
// different splits are independent (the first version looped over splits AND outputs in 16 blocks: 26 x 2 dependent L2 round
// trips per output, 50 us - more than the attention itself).
__global__ void global_attention_combine_kernel(const float* __restrict__ part, float* __restrict__ out, int splits, int K, int HD,
                                                int C, int heads) {
    grid_dep_launch();
    grid_dep_wait();
    const int g = blockIdx.x, h = blockIdx.y, b = blockIdx.z, d = threadIdx.x;
    const float* p = part + (((size_t)b * heads + h) * splits * K + g) * (size_t)(HD + 2);
    const size_t step = (size_t)K * (HD + 2);   // between the splits of one query
    float M = -INFINITY;
#pragma unroll 8
    for (int s = 0; s < splits; ++s) M = fmaxf(M, __ldg(p + s * step));
    float L = 0.f, O = 0.f;
#pragma unroll 8
    for (int s = 0; s < splits; ++s) {
        const float* q = p + s * step;
        const float m = __ldg(q);
        const float w = m == -INFINITY ? 0.f : __expf(m - M);
        L = fmaf(__ldg(q + 1), w, L);
        O = fmaf(__ldg(q + 2 + d), w, O);
    }
    out[((size_t)b * K + g) * C + h * HD + d] = L > 0.f ? O / L : 0.f;
}

}  // namespace

int gv_linear(const float* in, const float* ln_gamma, const float* ln_beta, const float* W, const float* bias, const float* res,
              float* out_f32, bf16* out_bf16, int M, int K, int N, int act, cudaStream_t st) {
    PD_CHECK(in && W && (out_f32 || out_bf16), PD_ERR_ARG, "gv_linear: null pointer");
    PD_CHECK(M >= 1 && K >= 4 && K % 4 == 0 && N >= 1 && (act == 0 || act == 1), PD_ERR_ARG,
             "gv_linear: M=%d K=%d (a multiple of 4) N=%d act=%d", M, K, N, act);
    PD_CHECK(!ln_gamma || (ln_beta && K <= kGvKChunk), PD_ERR_SHAPE, "gv_linear: fused LayerNorm needs K <= %d (got %d)",
             kGvKChunk, K);
    dim3 grid(ceil_div(N, kGvWarps), ceil_div(M, kGvRows));
    PD_LAUNCH(gv_linear_kernel, grid, kGvWarps * 32, 0, st, in, ln_gamma, ln_beta, W, bias, res, out_f32, out_bf16, M, K, N, act,
              1e-5f);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int gv_broadcast(const float* init, float* g, int B, int K, int C, cudaStream_t st) {
    const int n = B * K * C;
    PD_LAUNCH(gv_broadcast_kernel, ceil_div(n, 256), 256, 0, st, init, g, B, K * C);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

// keys per block: a multiple of the 64-key pass, at most 32 splits (3328 slots -> 26 x 128, 832 -> 13 x 64)
int global_attention_keys_per_split(int n_keys) { return std::max(kGaKeys, ceil_div(ceil_div(n_keys, 32), kGaKeys) * kGaKeys); }
int global_attention_splits(int n_keys) { return std::max(1, ceil_div(n_keys, global_attention_keys_per_split(n_keys))); }

size_t global_attention_workspace_floats(int B, int heads, int K, int hd, int n_keys) {
    return (size_t)B * heads * global_attention_splits(n_keys) * K * (hd + 2);
}

int global_attention(const GvQuery& a, float* out, float* workspace, int B, int N, int C, int heads, int K, const CuboidDev& g,
                     cudaStream_t st) {
    PD_CHECK(a.q && a.tok_kv && out && workspace, PD_ERR_ARG, "global_attention: null pointer");
    const int self_attn = a.sq != nullptr;
    PD_CHECK(!self_attn || (a.sk && a.sv && a.s_ld > 0), PD_ERR_ARG, "global_attention: self-attention needs keys and values");
    PD_CHECK(a.q_ld % 4 == 0 && (!self_attn || (a.sq_ld % 4 == 0 && a.s_ld % 8 == 0)), PD_ERR_ARG,
             "global_attention: row strides must keep 16-byte alignment");
    PD_CHECK(K >= 1 && K <= kMaxGlobal, PD_ERR_SHAPE, "global vectors: %d (1..%d are built)", K, kMaxGlobal);
    PD_CHECK(C % heads == 0, PD_ERR_SHAPE, "global_attention: C=%d heads=%d", C, heads);
    const int hd = C / heads;
    const int n_slots = g.num_cuboids * g.volume;
    const int n_keys = n_slots + (self_attn ? K : 0);
    const int splits = global_attention_splits(n_keys);
    const int per = global_attention_keys_per_split(n_keys);
    dim3 grid(splits, heads, B);
#define PD_GA(HDV, KM)                                                                                                         \
    PD_LAUNCH((global_attention_kernel<HDV, KM>), grid, kGaThreads, 0, st, a, g.tok, g.gmask, n_slots, n_keys, per, K, N, C, \
              heads, workspace)
#define PD_GA_K(HDV)                          \
    case HDV:                                 \
        if (K <= 8) PD_GA(HDV, 8);            \
        else if (K <= 16) PD_GA(HDV, 16);     \
        else PD_GA(HDV, 32);                  \
        break
    switch (hd) {
        PD_GA_K(16);
        PD_GA_K(32);
        PD_GA_K(64);
        PD_GA_K(128);
        default: set_error("global_attention: unsupported head dim %d", hd); return PD_ERR_SHAPE;
    }
#undef PD_GA_K
#undef PD_GA
    PD_LAUNCH_CHECK();
    PD_LAUNCH(global_attention_combine_kernel, dim3(K, heads, B), hd, 0, st, (const float*)workspace, out, splits, K, hd, C, heads);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
