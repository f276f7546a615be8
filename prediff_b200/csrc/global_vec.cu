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
constexpr int kGvColsPerWarp = 4;
constexpr int kGvWarps = 8;

// out[r][n] = (res ? res[r][n] : 0) + act( sum_k f(in[r][k]) W[n][k] + bias[n] ),  f = LayerNorm(ln_gamma, ln_beta) or id.
// One block: kGvRows rows x 32 columns; a warp owns 4 columns, its lanes split K (coalesced weight rows), the rows'
// inputs come from shared memory (lane-consecutive k: conflict free), 16 row sums per column reduced by shuffles.
__global__ void __launch_bounds__(kGvWarps * 32) gv_linear_kernel(const float* __restrict__ in, const float* __restrict__ ln_gamma,
                                                                  const float* __restrict__ ln_beta, const float* __restrict__ W,
                                                                  const float* __restrict__ bias, const float* res,
                                                                  float* out_f32, bf16* __restrict__ out_bf16, int M, int K, int N,
                                                                  int act, float ln_eps) {
    grid_dep_launch();
    grid_dep_wait();
    __shared__ float s_in[kGvRows][kGvKChunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.y * kGvRows;
    const int n0 = blockIdx.x * (kGvWarps * kGvColsPerWarp) + warp * kGvColsPerWarp;
    float acc[kGvColsPerWarp][kGvRows];
#pragma unroll
    for (int c = 0; c < kGvColsPerWarp; ++c)
#pragma unroll
        for (int r = 0; r < kGvRows; ++r) acc[c][r] = 0.f;
    for (int k0 = 0; k0 < K; k0 += kGvKChunk) {
        const int kc = min(kGvKChunk, K - k0);
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
                float s = 0.f;
                for (int k = lane; k < kc; k += 32) s += x[k];
                const float mean = warp_sum(s) / (float)kc;
                float v = 0.f;
                for (int k = lane; k < kc; k += 32) {
                    const float d = x[k] - mean;
                    v = fmaf(d, d, v);
                }
                const float rstd = rsqrtf(warp_sum(v) / (float)kc + ln_eps);
                for (int k = lane; k < kc; k += 32) s_in[r][k] = (x[k] - mean) * rstd * ln_gamma[k] + ln_beta[k];
            } else {
                for (int k = lane; k < kc; k += 32) s_in[r][k] = x[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < kGvColsPerWarp; ++c) {
            const int n = n0 + c;
            if (n >= N) continue;
            const float* w = W + (size_t)n * K + k0;
            for (int k = lane; k < kc; k += 32) {
                const float wv = __ldg(w + k);
#pragma unroll
                for (int r = 0; r < kGvRows; ++r) acc[c][r] = fmaf(s_in[r][k], wv, acc[c][r]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < kGvColsPerWarp; ++c) {
        const int n = n0 + c;
        float mine = 0.f;
#pragma unroll
        for (int r = 0; r < kGvRows; ++r) {
            const float t = warp_sum(acc[c][r]);
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
template <int HD>
__global__ void __launch_bounds__(kGaThreads) global_attention_kernel(const float* __restrict__ gqkv, const bf16* __restrict__ qkv,
                                                                      const bf16* __restrict__ gkv, const int* __restrict__ tok,
                                                                      const int* __restrict__ gmask, int n_slots, int n_keys,
                                                                      int keys_per_split, int K, int N, int C, int heads,
                                                                      float* __restrict__ part) {
    grid_dep_launch();
    grid_dep_wait();
    constexpr int QG = kGaThreads / HD;            // query groups of the PV phase (threads = (group, channel))
    constexpr int NACC = kMaxGlobal / QG;          // accumulators per thread
    __shared__ float s_q[kMaxGlobal][HD];
    __shared__ float s_p[kMaxGlobal][kGaKeys];
    __shared__ long long s_row[kGaKeys];           // element offset of the key's q|k|v row (< 0: zero row), or LLONG_MIN = masked
    __shared__ int s_isg[kGaKeys];
    __shared__ float s_m[kMaxGlobal], s_l[kMaxGlobal], s_alpha[kMaxGlobal];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int C3 = 3 * C;
    const float scale = rsqrtf((float)HD);
    for (int i = tid; i < K * HD; i += kGaThreads) {
        const int g = i / HD, d = i - g * HD;
        s_q[g][d] = gqkv[((size_t)b * K + g) * C3 + h * HD + d] * scale;   // q_global * scale (:899)
    }
    if (tid < kMaxGlobal) {
        s_m[tid] = -INFINITY;
        s_l[tid] = 0.f;
    }
    const bf16* base = qkv + (size_t)b * N * C3 + h * HD;
    const bf16* gbase = gkv + (size_t)b * K * C3 + h * HD;
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    const int qg = tid / HD, dch = tid - qg * HD;
    const int k_begin = split * keys_per_split, k_end = min(n_keys, k_begin + keys_per_split);
    for (int kb = k_begin; kb < k_end; kb += kGaKeys) {
        __syncthreads();
        // ---- scores: one thread per key ----
        if (tid < kGaKeys) {
            const int j = kb + tid;
            long long row = LLONG_MIN;
            int isg = 0;
            if (j < k_end) {
                if (j < n_slots) {
                    if (!gmask || gmask[j]) {
                        const int t = tok[j];
                        row = t >= 0 ? (long long)t * C3 : -1;
                    }
                } else {
                    row = (long long)(j - n_slots) * C3;
                    isg = 1;
                }
            }
            s_row[tid] = row;
            s_isg[tid] = isg;
            float sc[kMaxGlobal];
#pragma unroll
            for (int g = 0; g < kMaxGlobal; ++g) sc[g] = 0.f;
            if (row >= 0) {
                const bf16* kr = (isg ? gbase : base) + row + C;
#pragma unroll 2
                for (int d0 = 0; d0 < HD; d0 += 8) {
                    const uint4 u = *reinterpret_cast<const uint4*>(kr + d0);
                    const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                    const float kv[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
                    for (int g = 0; g < kMaxGlobal; ++g) {
                        if (g < K) {
                            float a = sc[g];
#pragma unroll
                            for (int e = 0; e < 8; ++e) a = fmaf(s_q[g][d0 + e], kv[e], a);
                            sc[g] = a;
                        }
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < kMaxGlobal; ++g)
                if (g < K) s_p[g][tid] = row == LLONG_MIN ? -INFINITY : sc[g];
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
        for (int j = 0; j < kGaKeys; ++j) {
            const long long row = s_row[j];
            if (row < 0) continue;   // masked, or a zero row (v = 0)
            const float v = __bfloat162float(((s_isg[j] ? gbase : base) + row + 2 * C)[dch]);
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

// out[b][g][h * HD + d] = sum_s o_s e^{m_s - M} / sum_s l_s e^{m_s - M}, splits taken in order (deterministic)
__global__ void global_attention_combine_kernel(const float* __restrict__ part, float* __restrict__ out, int splits, int K, int HD,
                                                int C, int heads) {
    grid_dep_launch();
    grid_dep_wait();
    const int h = blockIdx.x, b = blockIdx.y;
    const float* p = part + ((size_t)b * heads + h) * splits * (size_t)K * (HD + 2);
    for (int i = threadIdx.x; i < K * HD; i += blockDim.x) {
        const int g = i / HD, d = i - g * HD;
        float M = -INFINITY;
        for (int s = 0; s < splits; ++s) M = fmaxf(M, p[((size_t)s * K + g) * (HD + 2)]);
        float L = 0.f, O = 0.f;
        for (int s = 0; s < splits; ++s) {
            const float* q = p + ((size_t)s * K + g) * (HD + 2);
            const float w = q[0] == -INFINITY ? 0.f : __expf(q[0] - M);
            L = fmaf(q[1], w, L);
            O = fmaf(q[2 + d], w, O);
        }
        out[((size_t)b * K + g) * C + h * HD + d] = L > 0.f ? O / L : 0.f;
    }
}

}  // namespace

int gv_linear(const float* in, const float* ln_gamma, const float* ln_beta, const float* W, const float* bias, const float* res,
              float* out_f32, bf16* out_bf16, int M, int K, int N, int act, cudaStream_t st) {
    PD_CHECK(in && W && (out_f32 || out_bf16), PD_ERR_ARG, "gv_linear: null pointer");
    PD_CHECK(M >= 1 && K >= 1 && N >= 1 && (act == 0 || act == 1), PD_ERR_ARG, "gv_linear: M=%d K=%d N=%d act=%d", M, K, N, act);
    PD_CHECK(!ln_gamma || (ln_beta && K <= kGvKChunk), PD_ERR_SHAPE, "gv_linear: fused LayerNorm needs K <= %d (got %d)",
             kGvKChunk, K);
    dim3 grid(ceil_div(N, kGvWarps * kGvColsPerWarp), ceil_div(M, kGvRows));
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

int global_attention_splits(int n_keys) { return std::max(1, std::min(32, ceil_div(n_keys, 4 * kGaKeys))); }

size_t global_attention_workspace_floats(int B, int heads, int K, int hd, int n_keys) {
    return (size_t)B * heads * global_attention_splits(n_keys) * K * (hd + 2);
}

int global_attention(const float* gqkv, const bf16* qkv, const bf16* gkv, float* out, float* workspace, int B, int N, int C,
                     int heads, int K, int self_attn, const CuboidDev& g, cudaStream_t st) {
    PD_CHECK(gqkv && qkv && gkv && out && workspace, PD_ERR_ARG, "global_attention: null pointer");
    PD_CHECK(K >= 1 && K <= kMaxGlobal, PD_ERR_SHAPE, "global vectors: %d (1..%d are built)", K, kMaxGlobal);
    PD_CHECK(C % heads == 0, PD_ERR_SHAPE, "global_attention: C=%d heads=%d", C, heads);
    const int hd = C / heads;
    const int n_slots = g.num_cuboids * g.volume;
    const int n_keys = n_slots + (self_attn ? K : 0);
    const int splits = global_attention_splits(n_keys);
    const int per = ceil_div(ceil_div(n_keys, splits), kGaKeys) * kGaKeys;
    dim3 grid(splits, heads, B);
    switch (hd) {
#define PD_GA(HDV)                                                                                                              \
    case HDV:                                                                                                                   \
        PD_LAUNCH((global_attention_kernel<HDV>), grid, kGaThreads, 0, st, gqkv, qkv, gkv, g.tok, g.gmask, n_slots, n_keys, per, K, \
                  N, C, heads, workspace);                                                                                      \
        break
        PD_GA(16);
        PD_GA(32);
        PD_GA(64);
        PD_GA(128);
#undef PD_GA
        default: set_error("global_attention: unsupported head dim %d", hd); return PD_ERR_SHAPE;
    }
    PD_LAUNCH_CHECK();
    PD_LAUNCH(global_attention_combine_kernel, dim3(heads, B), 128, 0, st, (const float*)workspace, out, splits, K, hd, C, heads);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
