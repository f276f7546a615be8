// Input-gradient (dgrad-only) kernels for the knowledge-alignment guidance g = guide_scale * d||mean_T U(z_t,t) - y||/dz_t
// (reference: diffusion/knowledge_alignment/sevir.py:76-104, alignment_pl.py:441-445 - torch.autograd there).
// No weight gradients exist on this path, so a backward pass never needs the GEMM input activations: it needs the
// inputs of the non-linearities only (GroupNorm / LayerNorm inputs, GELU pre-activations, attention q|k|v).
// Every kernel writes the bf16 copy the next dgrad GEMM consumes next to the fp32 gradient it finishes.
#include "ops.cuh"

namespace pd {
namespace {

constexpr int kGnThreads = 256;
constexpr int kGnIters = 4;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// Shared prologue of the GroupNorm(+SiLU) backward kernels: for this thread's channel quad, the affine map
// xhat = x * rs[k] + sh[k], u = xhat * gm[k] + bt[k] and the factors of d(silu(u))/du.
struct GnQuad {
    float rs[4], sh[4], gm[4], bt[4];
    int grp[4];
};
__device__ __forceinline__ void gn_quad_load(GnQuad& q, const double* __restrict__ sums, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, int s, int c, int R, int C, int G, float eps) {
    const int cpg = C / G;
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
    q.gm[0] = gm.x; q.gm[1] = gm.y; q.gm[2] = gm.z; q.gm[3] = gm.w;
    q.bt[0] = bt.x; q.bt[1] = bt.y; q.bt[2] = bt.z; q.bt[3] = bt.w;
    const double n = (double)R * cpg;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int g = (c + k) / cpg;
        const double m = sums[((size_t)s * G + g) * 2] / n;
        double var = sums[((size_t)s * G + g) * 2 + 1] / n - m * m;
        if (var < 0) var = 0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        q.rs[k] = rstd;
        q.sh[k] = -(float)m * rstd;
        q.grp[k] = g;
    }
}
// d(act(u))/du * dy with act = SiLU (silu != 0) or identity
__device__ __forceinline__ float act_bwd(float u, float dy, int silu) {
    if (!silu) return dy;
    const float sg = sigmoid_f(u);
    return dy * sg * (1.0f + u * (1.0f - sg));
}

// pass 1: bsums[s][g] += (sum dxhat, sum dxhat * xhat) with dxhat = dy * act'(u) * gamma
__global__ void __launch_bounds__(kGnThreads) gn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  const double* __restrict__ sums,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  double* __restrict__ bsums, int R, int C, int G, float eps,
                                                                  int silu) {
    __shared__ float red[kGnThreads][8];
    const int s = blockIdx.y;
    const int c4n = C >> 2;
    const int lanes_r = kGnThreads / c4n;
    const int tc = threadIdx.x % c4n;
    const int tr = threadIdx.x / c4n;
    const int r0 = blockIdx.x * (lanes_r * kGnIters) + tr;
    const float4* xb = reinterpret_cast<const float4*>(x + ((size_t)s * R) * C) + tc;
    const float4* db = reinterpret_cast<const float4*>(dy + ((size_t)s * R) * C) + tc;
    float4 xv[kGnIters], dv[kGnIters];
#pragma unroll
    for (int i = 0; i < kGnIters; ++i) {
        const int r = r0 + i * lanes_r;
        xv[i] = r < R ? __ldg(xb + (size_t)r * c4n) : make_float4(0.f, 0.f, 0.f, 0.f);
        dv[i] = r < R ? __ldg(db + (size_t)r * c4n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    GnQuad q;
    gn_quad_load(q, sums, gamma, beta, s, tc * 4, R, C, G, eps);
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < kGnIters; ++i) {
        const float xs[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
        const float ds[4] = {dv[i].x, dv[i].y, dv[i].z, dv[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float xh = fmaf(xs[k], q.rs[k], q.sh[k]);
            const float u = fmaf(xh, q.gm[k], q.bt[k]);
            const float dxh = act_bwd(u, ds[k], silu) * q.gm[k];   // rows beyond R contribute dy = 0
            a[k] += dxh;
            b[k] = fmaf(dxh, xh, b[k]);
        }
    }
    float* my = red[threadIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) { my[k] = a[k]; my[4 + k] = b[k]; }
    __syncthreads();
    const int cpg = C / G;
    for (int g = threadIdx.x; g < G; g += kGnThreads) {
        float s1 = 0.f, s2 = 0.f;
        for (int rr = 0; rr < lanes_r; ++rr)
            for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
                const float* e = red[rr * c4n + (c >> 2)];
                s1 += e[c & 3];
                s2 += e[4 + (c & 3)];
            }
        atomicAdd(&bsums[((size_t)s * G + g) * 2 + 0], (double)s1);
        atomicAdd(&bsums[((size_t)s * G + g) * 2 + 1], (double)s2);
    }
}

// pass 2: dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)); dx_io (+)= dx, dxb = bf16(dx_io)
__global__ void __launch_bounds__(kGnThreads) gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  const double* __restrict__ sums,
                                                                  const double* __restrict__ bsums,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float* __restrict__ dx_io,
                                                                  bf16* __restrict__ dxb, int R, int C, int G, float eps,
                                                                  int silu, int accumulate) {
    const int s = blockIdx.y;
    const int c4n = C >> 2;
    const int lanes_r = kGnThreads / c4n;
    const int tc = threadIdx.x % c4n;
    const int tr = threadIdx.x / c4n;
    const int r0 = blockIdx.x * (lanes_r * kGnIters) + tr;
    const size_t sbase = ((size_t)s * R) * C;
    const float4* xb = reinterpret_cast<const float4*>(x + sbase) + tc;
    const float4* db = reinterpret_cast<const float4*>(dy + sbase) + tc;
    float4 xv[kGnIters], dv[kGnIters], ov[kGnIters];
#pragma unroll
    for (int i = 0; i < kGnIters; ++i) {
        const int r = r0 + i * lanes_r;
        const bool ok = r < R;
        xv[i] = ok ? __ldg(xb + (size_t)r * c4n) : make_float4(0.f, 0.f, 0.f, 0.f);
        dv[i] = ok ? __ldg(db + (size_t)r * c4n) : make_float4(0.f, 0.f, 0.f, 0.f);
        ov[i] = (ok && accumulate && dx_io) ? reinterpret_cast<const float4*>(dx_io + sbase)[(size_t)r * c4n + tc]
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    GnQuad q;
    gn_quad_load(q, sums, gamma, beta, s, tc * 4, R, C, G, eps);
    const int cpg = C / G;
    const float inv_n = 1.0f / ((float)R * (float)cpg);
    float m1[4], m2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        m1[k] = (float)bsums[((size_t)s * G + q.grp[k]) * 2] * inv_n;
        m2[k] = (float)bsums[((size_t)s * G + q.grp[k]) * 2 + 1] * inv_n;
    }
#pragma unroll
    for (int i = 0; i < kGnIters; ++i) {
        const int r = r0 + i * lanes_r;
        if (r >= R) continue;
        const float xs[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
        const float ds[4] = {dv[i].x, dv[i].y, dv[i].z, dv[i].w};
        float o[4] = {ov[i].x, ov[i].y, ov[i].z, ov[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float xh = fmaf(xs[k], q.rs[k], q.sh[k]);
            const float u = fmaf(xh, q.gm[k], q.bt[k]);
            const float dxh = act_bwd(u, ds[k], silu) * q.gm[k];
            o[k] += q.rs[k] * (dxh - m1[k] - xh * m2[k]);
        }
        if (dx_io) reinterpret_cast<float4*>(dx_io + sbase)[(size_t)r * c4n + tc] = make_float4(o[0], o[1], o[2], o[3]);
        if (dxb) {
            uint2 pk;
            pk.x = pack_bf16x2(o[0], o[1]);
            pk.y = pack_bf16x2(o[2], o[3]);
            reinterpret_cast<uint2*>(dxb + sbase)[(size_t)r * c4n + tc] = pk;
        }
    }
}

// LayerNorm backward, one warp per row (NV float4 per lane). y = xhat * gamma + beta, dy given:
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma;  dx_io += dx (or = dx), dxb = bf16(dx_io).
// SCATTER: the row is a PatchMerging3D merged row (2x2 space-to-depth gather of x, cuboid_transformer.py:286-294);
// its gradient is scattered back to the four source positions (plain store: x has no other consumer there).
template <int NV, bool SCATTER>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ dy, float* __restrict__ dx_io,
                                                     bf16* __restrict__ dxb, int P, int C, float eps, int accumulate, int H,
                                                     int W, int Cs) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= P) return;
    const int c4n = C >> 2;
    size_t off[NV];   // element offset of each float4 of this row inside x / dx
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c4 = lane + i * 32;
        const int c = c4 * 4;
        if constexpr (!SCATTER) {
            off[i] = (size_t)row * C + c;
        } else {
            const int W2 = W >> 1, H2 = H >> 1;
            const int w2 = row % W2;
            const int h2 = (row / W2) % H2;
            const int f = row / (W2 * H2);
            const int seg = c / Cs;
            const int cc = c - seg * Cs;
            off[i] = (((size_t)f * H + 2 * h2 + (seg >> 1)) * W + 2 * w2 + (seg & 1)) * Cs + cc;
        }
    }
    float4 xv[NV], gv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c4 = lane + i * 32;
        if (c4 < c4n) {
            xv[i] = __ldg(reinterpret_cast<const float4*>(x + off[i]));
            const float4 d = __ldg(reinterpret_cast<const float4*>(dy + (size_t)row * C) + c4);
            const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
            gv[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        } else {
            xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) sum += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < c4n) {
            const float a = xv[i].x - mean, b = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
            sq += a * a + b * b + c * c + d * d;
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < c4n) {
            xv[i].x = (xv[i].x - mean) * rstd; xv[i].y = (xv[i].y - mean) * rstd;
            xv[i].z = (xv[i].z - mean) * rstd; xv[i].w = (xv[i].w - mean) * rstd;
            s1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
            s2 += gv[i].x * xv[i].x + gv[i].y * xv[i].y + gv[i].z * xv[i].z + gv[i].w * xv[i].w;
        }
    }
    const float m1 = warp_sum(s1) / (float)C, m2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < c4n) {
            float4 o = make_float4(rstd * (gv[i].x - m1 - xv[i].x * m2), rstd * (gv[i].y - m1 - xv[i].y * m2),
                                   rstd * (gv[i].z - m1 - xv[i].z * m2), rstd * (gv[i].w - m1 - xv[i].w * m2));
            if (accumulate) {
                const float4 p = *reinterpret_cast<const float4*>(dx_io + off[i]);
                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            *reinterpret_cast<float4*>(dx_io + off[i]) = o;
            if (dxb) {
                uint2 pk;
                pk.x = pack_bf16x2(o.x, o.y);
                pk.y = pack_bf16x2(o.z, o.w);
                *reinterpret_cast<uint2*>(dxb + off[i]) = pk;
            }
        }
    }
}

template <bool SCATTER>
int launch_ln_bwd(const float* x, const float* gamma, const float* dy, float* dx_io, bf16* dxb, int P, int C, float eps,
                  int accumulate, int H, int W, int Cs, cudaStream_t st) {
    PD_CHECK(C % 128 == 0 && C <= 1024, PD_ERR_SHAPE, "layer_norm_bwd: unsupported C=%d", C);
    const int nv = C / 128;
    const int blocks = ceil_div(P, 8);
#define PD_LNB(NVV) ln_bwd_kernel<NVV, SCATTER><<<blocks, 256, 0, st>>>(x, gamma, dy, dx_io, dxb, P, C, eps, accumulate, H, W, Cs)
    if (nv == 1) PD_LNB(1);
    else if (nv == 2) PD_LNB(2);
    else if (nv <= 4) PD_LNB(4);
    else PD_LNB(8);
#undef PD_LNB
    PD_LAUNCH_CHECK();
    return PD_OK;
}

// exact-erf GELU on the saved fp32 pre-activation -> bf16 (the forward the backward differentiates)
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ pre, bf16* __restrict__ y, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(pre) + i);
        const float a[4] = {v.x, v.y, v.z, v.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = 0.5f * a[k] * (1.0f + erff(a[k] * 0.70710678118654752f));
        uint2 pk;
        pk.x = pack_bf16x2(o[0], o[1]);
        pk.y = pack_bf16x2(o[2], o[3]);
        reinterpret_cast<uint2*>(y)[i] = pk;
    }
}
// dpre = dmid * (Phi(x) + x phi(x))
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ pre, const bf16* __restrict__ dmid,
                                                       bf16* __restrict__ dpre, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(pre) + i);
        const uint2 du = __ldg(reinterpret_cast<const uint2*>(dmid) + i);
        const float2 d01 = unpack_bf16x2(du.x), d23 = unpack_bf16x2(du.y);
        const float a[4] = {v.x, v.y, v.z, v.w};
        const float d[4] = {d01.x, d01.y, d23.x, d23.y};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float cdf = 0.5f * (1.0f + erff(a[k] * 0.70710678118654752f));
            const float pdf = 0.3989422804014327f * __expf(-0.5f * a[k] * a[k]);
            o[k] = d[k] * (cdf + a[k] * pdf);
        }
        uint2 pk;
        pk.x = pack_bf16x2(o[0], o[1]);
        pk.y = pack_bf16x2(o[2], o[3]);
        reinterpret_cast<uint2*>(dpre)[i] = pk;
    }
}

// Backward of the axial attention core. One block per line (<= 16 tokens), one warp per head:
//   S = scale q k^T + bias, P = softmax(S); dV = P^T dO; dP = dO V^T; dS = P o (dP - rowsum(P o dP));
//   dQ = scale dS K; dK = scale dS^T Q.   q|k|v and dO staged in smem (bf16), all math fp32 on CUDA cores
// (sequence <= 16: the whole op is a fraction of a percent of the guidance FLOPs).
constexpr int kMaxLine = 16;
template <int HD>
__global__ void __launch_bounds__(128) axial_attention_bwd_kernel(const bf16* __restrict__ qkv,
                                                                  const float* __restrict__ bias_table,
                                                                  const bf16* __restrict__ dout, bf16* __restrict__ dqkv,
                                                                  int T, int H, int W, int C, int heads, int axis) {
    extern __shared__ __align__(16) uint8_t smem_attb[];
    const int C3 = 3 * C;
    const int ld = C3 + 8, ldo = C + 8;
    bf16* s_qkv = reinterpret_cast<bf16*>(smem_attb);                 // [16][3C + 8]
    bf16* s_do = s_qkv + (size_t)kMaxLine * ld;                        // [16][C + 8]
    float* s_mat = reinterpret_cast<float*>(s_do + (size_t)kMaxLine * ldo);   // per warp: P[16][17], dS[16][17]
    int L, stride, base;
    {
        const int line = blockIdx.x;
        if (axis == 0) {
            L = T; stride = H * W;
            const int hw = line % (H * W), b = line / (H * W);
            base = b * T * H * W + hw;
        } else if (axis == 1) {
            L = H; stride = W;
            const int w = line % W, bt = line / W;
            base = bt * H * W + w;
        } else {
            L = W; stride = 1;
            base = line * W;
        }
    }
    {
        const int vq = C3 / 8, vo = C / 8;
        for (int i = threadIdx.x; i < kMaxLine * vq; i += blockDim.x) {
            const int r = i / vq, v = i - r * vq;
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            if (r < L) val = __ldg(reinterpret_cast<const uint4*>(qkv + (size_t)(base + r * stride) * C3) + v);
            *reinterpret_cast<uint4*>(s_qkv + (size_t)r * ld + v * 8) = val;
        }
        for (int i = threadIdx.x; i < kMaxLine * vo; i += blockDim.x) {
            const int r = i / vo, v = i - r * vo;
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            if (r < L) val = __ldg(reinterpret_cast<const uint4*>(dout + (size_t)(base + r * stride) * C) + v);
            *reinterpret_cast<uint4*>(s_do + (size_t)r * ldo + v * 8) = val;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    float* sP = s_mat + (size_t)wid * (2 * 16 * 17);
    float* sD = sP + 16 * 17;
    const float scale = rsqrtf((float)HD);
    for (int h = wid; h < heads; h += blockDim.x >> 5) {
        const bf16* sq = s_qkv + h * HD;
        const bf16* sk = s_qkv + C + h * HD;
        const bf16* sv = s_qkv + 2 * C + h * HD;
        const bf16* so = s_do + h * HD;
        // ---- S and dP: lane -> row i = lane / 2, columns j0 .. j0+7 ----
        {
            const int i = lane >> 1, j0 = (lane & 1) * 8;
            float sacc[8], pacc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { sacc[j] = 0.f; pacc[j] = 0.f; }
            for (int d = 0; d < HD; d += 2) {
                const float2 qv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sq + (size_t)i * ld + d));
                const float2 ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(so + (size_t)i * ldo + d));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sk + (size_t)(j0 + j) * ld + d));
                    const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sv + (size_t)(j0 + j) * ld + d));
                    sacc[j] = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, sacc[j]));
                    pacc[j] = fmaf(ov.x, vv.x, fmaf(ov.y, vv.y, pacc[j]));
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int jj = j0 + j;
                float s = -INFINITY;
                if (i < L && jj < L) s = sacc[j] * scale + __ldg(bias_table + (i - jj + L - 1) * heads + h);
                sP[i * 17 + jj] = s;
                sD[i * 17 + jj] = pacc[j];
            }
        }
        __syncwarp();
        // ---- softmax + dS per row (lanes 0..15) ----
        if (lane < 16) {
            const int i = lane;
            float mx = -INFINITY;
            for (int j = 0; j < 16; ++j) mx = fmaxf(mx, sP[i * 17 + j]);
            if (mx == -INFINITY) mx = 0.f;
            float p[16], sum = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                p[j] = __expf(sP[i * 17 + j] - mx);
                sum += p[j];
            }
            const float inv = sum > 0.f ? 1.f / sum : 0.f;
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                p[j] *= inv;
                dot = fmaf(p[j], sD[i * 17 + j], dot);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                sP[i * 17 + j] = p[j];
                sD[i * 17 + j] = p[j] * (sD[i * 17 + j] - dot) * scale;   // scale folded in: dQ = dS K, dK = dS^T Q
            }
        }
        __syncwarp();
        // ---- dQ, dK, dV: lane -> columns d = lane, lane + 32, ... ----
        for (int d = lane; d < HD; d += 32) {
            float qc[16], kc[16], vc[16], oc[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                qc[r] = __bfloat162float(sq[(size_t)r * ld + d]);
                kc[r] = __bfloat162float(sk[(size_t)r * ld + d]);
                oc[r] = __bfloat162float(so[(size_t)r * ldo + d]);
            }
            (void)vc;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (r >= L) break;
                float dq = 0.f, dk = 0.f, dv = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    dq = fmaf(sD[r * 17 + j], kc[j], dq);     // dQ[r] = sum_j dS[r][j] K[j]
                    dk = fmaf(sD[j * 17 + r], qc[j], dk);     // dK[r] = sum_i dS[i][r] Q[i]
                    dv = fmaf(sP[j * 17 + r], oc[j], dv);     // dV[r] = sum_i P[i][r] dO[i]
                }
                bf16* orow = dqkv + (size_t)(base + r * stride) * C3 + h * HD + d;
                orow[0] = __float2bfloat16_rn(dq);
                orow[C] = __float2bfloat16_rn(dk);
                orow[2 * C] = __float2bfloat16_rn(dv);
            }
        }
        __syncwarp();
    }
}

__global__ void pack_linear_t_kernel(const float* __restrict__ w, bf16* __restrict__ out, int N, int K) {
    const int64_t total = (int64_t)N * K;   // out [K][N]
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const int64_t k = i / N;
        out[i] = __float2bfloat16_rn(w[(int64_t)n * K + k]);
    }
}
// w fp32 [Co][Ci][taps] -> bf16 [Ci][taps][Co] with the taps reversed (dgrad of a stride-1 "same" correlation)
__global__ void pack_conv_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Co, int Ci, int taps) {
    const int64_t total = (int64_t)Ci * taps * Co;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % Co);
        const int tp = (int)((i / Co) % taps);
        const int64_t ci = i / ((int64_t)Co * taps);
        out[i] = __float2bfloat16_rn(w[((int64_t)co * Ci + ci) * taps + (taps - 1 - tp)]);
    }
}

inline int ew_blocks(int64_t items) {
    int64_t b = ceil_div64(items, 256);
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int gn_bwd(const float* x, const float* dy, const double* sums, double* bsums, const float* gamma, const float* beta,
           float* dx_io, bf16* dxb, int S, int R, int C, int G, float eps, int silu, int accumulate, cudaStream_t st) {
    PD_CHECK(C % 4 == 0 && (kGnThreads % (C / 4) == 0) && C / 4 <= kGnThreads, PD_ERR_SHAPE, "gn_bwd: unsupported C=%d", C);
    PD_CHECK(G > 0 && C % G == 0 && G <= 128, PD_ERR_SHAPE, "gn_bwd: unsupported groups=%d for C=%d", G, C);
    PD_CHECK(dx_io || dxb, PD_ERR_ARG, "gn_bwd: no output");
    PD_CHECK(!accumulate || dx_io, PD_ERR_ARG, "gn_bwd: accumulate needs dx_io");
    dim3 grid(ceil_div(R, (kGnThreads / (C / 4)) * kGnIters), S);
    gn_bwd_stats_kernel<<<grid, kGnThreads, 0, st>>>(x, dy, sums, gamma, beta, bsums, R, C, G, eps, silu);
    PD_LAUNCH_CHECK();
    gn_bwd_apply_kernel<<<grid, kGnThreads, 0, st>>>(x, dy, sums, bsums, gamma, beta, dx_io, dxb, R, C, G, eps, silu,
                                                     accumulate);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int layer_norm_bwd(const float* x, const float* gamma, const float* dy, float* dx_io, bf16* dxb, int P, int C, float eps,
                   int accumulate, cudaStream_t st) {
    PD_CHECK(dx_io, PD_ERR_ARG, "layer_norm_bwd: dx_io required");
    return launch_ln_bwd<false>(x, gamma, dy, dx_io, dxb, P, C, eps, accumulate, 0, 0, 0, st);
}

int patch_merge_ln_bwd(const float* x, const float* gamma, const float* dy, float* dx, bf16* dxb, int BT, int H, int W,
                       int C, float eps, cudaStream_t st) {
    PD_CHECK(H % 2 == 0 && W % 2 == 0 && C % 4 == 0 && dx, PD_ERR_SHAPE, "patch_merge_ln_bwd: shape");
    return launch_ln_bwd<true>(x, gamma, dy, dx, dxb, BT * (H / 2) * (W / 2), 4 * C, eps, 0, H, W, C, st);
}

int gelu_fwd(const float* pre, bf16* y, int64_t n, cudaStream_t st) {
    PD_CHECK(n % 4 == 0, PD_ERR_SHAPE, "gelu_fwd: n %% 4");
    gelu_fwd_kernel<<<ew_blocks(n / 4), 256, 0, st>>>(pre, y, n / 4);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int gelu_bwd(const float* pre, const bf16* dmid, bf16* dpre, int64_t n, cudaStream_t st) {
    PD_CHECK(n % 4 == 0, PD_ERR_SHAPE, "gelu_bwd: n %% 4");
    gelu_bwd_kernel<<<ew_blocks(n / 4), 256, 0, st>>>(pre, dmid, dpre, n / 4);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int axial_attention_bwd(const bf16* qkv, const float* bias_table, const bf16* dout, bf16* dqkv, int B, int T, int H,
                        int W, int C, int heads, int axis, cudaStream_t st) {
    PD_CHECK(axis >= 0 && axis <= 2, PD_ERR_ARG, "axial_attention_bwd: axis %d", axis);
    const int L = axis == 0 ? T : (axis == 1 ? H : W);
    PD_CHECK(L >= 1 && L <= kMaxLine, PD_ERR_SHAPE, "axial_attention_bwd: line length %d > %d", L, kMaxLine);
    PD_CHECK(C % heads == 0 && C % 8 == 0, PD_ERR_SHAPE, "axial_attention_bwd: C=%d heads=%d", C, heads);
    const int hd = C / heads;
    const int lines = B * T * H * W / L;
    const int threads = heads * 32 > 128 ? 128 : heads * 32;
    const size_t smem = (size_t)kMaxLine * (3 * C + 8 + C + 8) * sizeof(bf16) + (size_t)(threads / 32) * 2 * 16 * 17 * 4;
    PD_CHECK(smem <= 160 * 1024, PD_ERR_SHAPE, "axial_attention_bwd: line of %zu bytes does not fit in smem", smem);
#define PD_LAUNCH_AXB(HDV)                                                                                             \
    do {                                                                                                               \
        static bool attr_set = false;                                                                                  \
        if (!attr_set) {                                                                                               \
            PD_CUDA(cudaFuncSetAttribute(axial_attention_bwd_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         160 * 1024));                                                                 \
            attr_set = true;                                                                                           \
        }                                                                                                              \
        axial_attention_bwd_kernel<HDV><<<lines, threads, smem, st>>>(qkv, bias_table, dout, dqkv, T, H, W, C, heads,  \
                                                                      axis);                                           \
    } while (0)
    switch (hd) {
        case 16: PD_LAUNCH_AXB(16); break;
        case 32: PD_LAUNCH_AXB(32); break;
        case 64: PD_LAUNCH_AXB(64); break;
        case 128: PD_LAUNCH_AXB(128); break;
        default: set_error("axial_attention_bwd: unsupported head dim %d", hd); return PD_ERR_SHAPE;
    }
#undef PD_LAUNCH_AXB
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int pack_linear_t(const float* w, bf16* out, int N, int K, cudaStream_t st) {
    pack_linear_t_kernel<<<ew_blocks((int64_t)N * K), 256, 0, st>>>(w, out, N, K);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int pack_conv_dgrad(const float* w, bf16* out, int Co, int Ci, int taps, cudaStream_t st) {
    pack_conv_dgrad_kernel<<<ew_blocks((int64_t)Co * Ci * taps), 256, 0, st>>>(w, out, Co, Ci, taps);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
