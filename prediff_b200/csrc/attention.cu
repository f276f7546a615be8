// Attention cores. The cuboid attention of the shipped SEVIR-LR config is axial (sequence length 8..16 per
// softmax, 0.25 % of the step's FLOPs), so it is a CUDA-core kernel that reads q/k/v straight out of the QKV
// GEMM output with strided (axial) addressing - the reference's cuboid_reorder / reverse copies never exist.
#include "ops.cuh"
#include <cstdlib>

namespace pd {
namespace {

constexpr int kMaxLine = 16;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D(16x8, f32) += A(16x8, tf32, row) * B(8x8, tf32, col)
__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + lo with hi, lo representable in tf32 (hi = rna(x), lo = rna(x - hi)): the 3xTF32 operand split
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gsrc) : "memory");
}

// One block per line (<= 16 tokens along the attended axis), one warp per head. The line's q|k|v rows are staged
// in smem with cp.async (rows padded by 16 B so ldmatrix is bank-conflict free); per head the 16x16 score tile and
// the 16xHD output come from warp-level mma.sync (the sequence is far too short to fill a 128-row tcgen05 tile:
// this op is 0.25 % of the step's FLOPs), softmax stays in the accumulator registers.
// GV: the sample's global vectors (cuboid_transformer.py:902-913) ride along as up to 16 more keys per line - their k / v rows
// (GvKeys, ops.cuh) staged beside the line, scores without position bias, never masked - one more 16-key tile through the
// same fragments; with separate_global_qkv the line's own rows of a second query tensor (gk.q2) meet those keys. (The
// shipped config has no global vectors and runs the GV = false instantiation, unchanged.)
template <int HD, bool GV>
__global__ void __launch_bounds__(128) axial_attention_kernel(const bf16* __restrict__ qkv,
                                                              const float* __restrict__ bias_table,
                                                              bf16* __restrict__ out, int T, int H, int W, int C,
                                                              int heads, int axis, const GvKeys gk) {
    grid_dep_launch();
    grid_dep_wait();
    extern __shared__ __align__(16) uint8_t smem_att[];
    bf16* s_qkv = reinterpret_cast<bf16*>(smem_att);  // [16][3C + 8]
    const int C3 = 3 * C;
    const int ld = C3 + 8;
    bf16* s_g = s_qkv + (size_t)kMaxLine * ld;        // GV: [16][2C + 8] = k | v rows of the global vectors
    const int ldg = 2 * C + 8;
    bf16* s_q2 = s_g + (size_t)kMaxLine * ldg;        // GV, gk.q2: [16][C + 8] = the line's queries for the global keys
    const int ldq2 = C + 8;
    const int n_global = gk.n;
    int L, stride, base;
    {
        const int line = blockIdx.x;
        if (axis == 0) {  // along T: lines enumerate (b, h, w)
            L = T; stride = H * W;
            const int hw = line % (H * W), b = line / (H * W);
            base = b * T * H * W + hw;
        } else if (axis == 1) {  // along H: lines enumerate (b, t, w)
            L = H; stride = W;
            const int w = line % W, bt = line / W;
            base = bt * H * W + w;
        } else {  // along W: lines enumerate (b, t, h)
            L = W; stride = 1;
            base = line * W;
        }
    }
    {
        const int vec_per_row = C3 / 8;  // 16-byte vectors
        const int total = kMaxLine * vec_per_row;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int r = i / vec_per_row, v = i - r * vec_per_row;
            bf16* dst = s_qkv + (size_t)r * ld + v * 8;
            if (r < L) cp_async16(dst, qkv + (size_t)(base + r * stride) * C3 + v * 8);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);  // padded tokens: finite (zero) rows
        }
        if constexpr (GV) {
            const size_t grow = (size_t)(base / (T * H * W)) * n_global * gk.ld;   // the sample's global rows
            const int vpr = C / 8;
            for (int i = threadIdx.x; i < kMaxLine * 2 * vpr; i += blockDim.x) {
                const int r = i / (2 * vpr), v = i - r * 2 * vpr;   // v < vpr: key part, else value part
                bf16* dst = s_g + (size_t)r * ldg + v * 8;
                if (r < n_global) cp_async16(dst, (v < vpr ? gk.k + v * 8 : gk.v + (v - vpr) * 8) + grow + (size_t)r * gk.ld);
                else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
            if (gk.q2) {
                for (int i = threadIdx.x; i < kMaxLine * vpr; i += blockDim.x) {
                    const int r = i / vpr, v = i - r * vpr;
                    bf16* dst = s_q2 + (size_t)r * ldq2 + v * 8;
                    if (r < L) cp_async16(dst, gk.q2 + (size_t)(base + r * stride) * gk.q2_ld + v * 8);
                    else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const float scale = rsqrtf((float)HD);
    constexpr int NK = GV ? 8 : 4;   // scores per thread and row half: 4 local (+ 4 global) keys
    for (int h = threadIdx.x >> 5; h < heads; h += blockDim.x >> 5) {
        const bf16* sq = s_qkv + h * HD;
        const bf16* sk = s_qkv + C + h * HD;
        const bf16* sv = s_qkv + 2 * C + h * HD;
        const bf16* sgk = s_g + h * HD;
        const bf16* sgv = s_g + C + h * HD;
        // ---- S = Q K^T : 16 x 16 (+ 16 global keys), 8-wide key tiles ----
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
        float s2[4] = {0.f, 0.f, 0.f, 0.f}, s3[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
            uint32_t a[4], b[4];
            ldmatrix_x4(a, sq + (size_t)((lane & 7) + 8 * ((lane >> 3) & 1)) * ld + kk * 16 + 8 * (lane >> 4));
            ldmatrix_x4(b, sk + (size_t)((lane & 7) + 8 * (lane >> 4)) * ld + kk * 16 + 8 * ((lane >> 3) & 1));
            mma_bf16_16816(s0, a, b[0], b[1]);
            mma_bf16_16816(s1, a, b[2], b[3]);
            if constexpr (GV) {
                uint32_t bg[4];
                ldmatrix_x4(bg, sgk + (size_t)((lane & 7) + 8 * (lane >> 4)) * ldg + kk * 16 + 8 * ((lane >> 3) & 1));
                if (gk.q2) {   // separate_global_qkv: the tokens' l2g_q rows instead of q
                    uint32_t a2[4];
                    ldmatrix_x4(a2, s_q2 + h * HD + (size_t)((lane & 7) + 8 * ((lane >> 3) & 1)) * ldq2 + kk * 16 + 8 * (lane >> 4));
                    mma_bf16_16816(s2, a2, bg[0], bg[1]);
                    mma_bf16_16816(s3, a2, bg[2], bg[3]);
                } else {
                    mma_bf16_16816(s2, a, bg[0], bg[1]);
                    mma_bf16_16816(s3, a, bg[2], bg[3]);
                }
            }
        }
        // thread holds rows g, g+8; keys {2tq, 2tq+1} (s0), {8+2tq, 9+2tq} (s1) and the same slots of the global tile (s2, s3)
        float p[2][NK];  // [row half][key slot]
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            const int i = g + 8 * rh;
            float v[NK];
            v[0] = s0[2 * rh]; v[1] = s0[2 * rh + 1]; v[2] = s1[2 * rh]; v[3] = s1[2 * rh + 1];
            if constexpr (GV) { v[4] = s2[2 * rh]; v[5] = s2[2 * rh + 1]; v[6] = s3[2 * rh]; v[7] = s3[2 * rh + 1]; }
            const int j[4] = {2 * tq, 2 * tq + 1, 8 + 2 * tq, 9 + 2 * tq};
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (j[k] < L && i < L) v[k] = v[k] * scale + __ldg(bias_table + (i - j[k] + L - 1) * heads + h);
                else v[k] = -INFINITY;
                mx = fmaxf(mx, v[k]);
            }
            if constexpr (GV) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[4 + k] = (j[k] < n_global && i < L) ? v[4 + k] * scale : -INFINITY;
                    mx = fmaxf(mx, v[4 + k]);
                }
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            if (mx == -INFINITY) mx = 0.f;  // padded query row
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                v[k] = __expf(v[k] - mx);
                sum += v[k];
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
            for (int k = 0; k < NK; ++k) p[rh][k] = v[k] * inv;
        }
        // P (C-fragment layout) is already the A fragment of the next mma: a0=(g, k lo), a1=(g+8, k lo), a2/a3 = k hi
        uint32_t pa[4], pg[4] = {0u, 0u, 0u, 0u};
        pa[0] = pack_bf16x2(p[0][0], p[0][1]);
        pa[1] = pack_bf16x2(p[1][0], p[1][1]);
        pa[2] = pack_bf16x2(p[0][2], p[0][3]);
        pa[3] = pack_bf16x2(p[1][2], p[1][3]);
        if constexpr (GV) {
            pg[0] = pack_bf16x2(p[0][NK - 4], p[0][NK - 3]);
            pg[1] = pack_bf16x2(p[1][NK - 4], p[1][NK - 3]);
            pg[2] = pack_bf16x2(p[0][NK - 2], p[0][NK - 1]);
            pg[3] = pack_bf16x2(p[1][NK - 2], p[1][NK - 1]);
        }
        // ---- O = P V : 16 x HD ----
        bf16* o_lo = out + (size_t)(base + g * stride) * C + h * HD + 2 * tq;
        bf16* o_hi = out + (size_t)(base + (g + 8) * stride) * C + h * HD + 2 * tq;
#pragma unroll
        for (int jn = 0; jn < HD / 8; jn += 2) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, sv + (size_t)((lane & 7) + 8 * ((lane >> 3) & 1)) * ld + 8 * (jn + (lane >> 4)));
            float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
            mma_bf16_16816(o0, pa, b[0], b[1]);
            mma_bf16_16816(o1, pa, b[2], b[3]);
            if constexpr (GV) {
                uint32_t bg[4];
                ldmatrix_x4_trans(bg, sgv + (size_t)((lane & 7) + 8 * ((lane >> 3) & 1)) * ldg + 8 * (jn + (lane >> 4)));
                mma_bf16_16816(o0, pg, bg[0], bg[1]);
                mma_bf16_16816(o1, pg, bg[2], bg[3]);
            }
            if (g < L) {
                *reinterpret_cast<uint32_t*>(o_lo + 8 * jn) = pack_bf16x2(o0[0], o0[1]);
                *reinterpret_cast<uint32_t*>(o_lo + 8 * jn + 8) = pack_bf16x2(o1[0], o1[1]);
            }
            if (g + 8 < L) {
                *reinterpret_cast<uint32_t*>(o_hi + 8 * jn) = pack_bf16x2(o0[2], o0[3]);
                *reinterpret_cast<uint32_t*>(o_hi + 8 * jn + 8) = pack_bf16x2(o1[2], o1[3]);
            }
        }
    }
}

// fp32 variant for the TF32-class precision mode (PD_PRECISION_TF32): q|k|v arrive as fp32 from the QKV GEMM and the
// whole core (scores, bias, softmax, PV) runs in fp32 on the CUDA cores - 1.62 of the step's 653 GFLOP, so there is
// nothing to win with tensor cores, and nothing of the reference's fp32 attention arithmetic is rounded away. The output
// is the projection GEMM's A operand: fp32 rounded to tf32. One block per line, one warp per head. Round 2: the two products
// run on mma.sync m16n8k8 with the 3xTF32 operand split (x = hi + lo, hi*hi + hi*lo + lo*hi, fp32 accumulate): fp32-accurate
// (the test bars did not move) at ~650 instead of ~1 450 instructions per (line, head) - the FMA version was issue-bound at 20 %
// occupancy (ncu: 35 % issue slots busy, 28 us per level-0 launch).
template <int HD, bool PAIR>   // PAIR: two lines of <= 8 tokens per block
__global__ void __launch_bounds__(128) axial_attention_f32_kernel(const float* __restrict__ qkv,
                                                                  const float* __restrict__ bias_table,
                                                                  float* __restrict__ out, int T, int H, int W, int C,
                                                                  int heads, int axis, int total_lines) {
    grid_dep_launch();
    grid_dep_wait();
    extern __shared__ __align__(16) uint8_t smem_att[];
    float* s_qkv = reinterpret_cast<float*>(smem_att);  // [16][3C + 4]
    const int C3 = 3 * C;
    const int ld = C3 + 4;
    // Lines of <= 8 tokens travel in pairs (rows 0-7 / 8-15 of the 16-row MMA tile, scores across the two masked out): a block
    // is `lpb` lines; slot row r = (line r / rows_per, position r % rows_per).
    const int L = axis == 0 ? T : (axis == 1 ? H : W);
    const int stride = axis == 0 ? H * W : (axis == 1 ? W : 1);
    constexpr int lpb = PAIR ? 2 : 1, rows_per = kMaxLine / lpb;
    int base_of[lpb];
#pragma unroll
    for (int u = 0; u < lpb; ++u) {
        const int line = blockIdx.x * lpb + u;
        int b0;
        if (axis == 0) {
            const int hw = line % (H * W), b = line / (H * W);
            b0 = b * T * H * W + hw;
        } else if (axis == 1) {
            const int w = line % W, bt = line / W;
            b0 = bt * H * W + w;
        } else {
            b0 = line * W;
        }
        base_of[u] = line < total_lines ? b0 : -1;
    }
    {
        const int vec_per_row = C3 / 4;
        const int total = kMaxLine * vec_per_row;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int r = i / vec_per_row, v = i - r * vec_per_row;
            const int u = r / rows_per, pos = r - u * rows_per;
            const int bu = PAIR ? (u ? base_of[lpb - 1] : base_of[0]) : base_of[0];
            float* dst = s_qkv + (size_t)r * ld + v * 4;
            if (pos < L && bu >= 0) cp_async16(dst, qkv + (size_t)(bu + pos * stride) * C3 + v * 4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const float scale = rsqrtf((float)HD);
    // per-warp scratch for the probability tile (C-fragment layout -> A-fragment layout of the second product)
    float* s_p = s_qkv + (size_t)kMaxLine * ld + warp * (16 * 20);
    for (int h = warp; h < heads; h += blockDim.x >> 5) {
        const float* sq = s_qkv + h * HD;
        const float* sk = s_qkv + C + h * HD;
        const float* sv = s_qkv + 2 * C + h * HD;
        // ---- S = Q K^T (16 x 16, K = HD) on mma.sync m16n8k8 with split operands x = hi + lo (both tf32): hi*hi + hi*lo +
        // lo*hi carries ~21 mantissa bits per product, fp32 accumulate - the result is fp32-accurate ----
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int kk = 0; kk < HD / 8; ++kk) {
            uint32_t ah[4], al[4];
            split_tf32(sq[(size_t)g * ld + 8 * kk + tq], ah[0], al[0]);
            split_tf32(sq[(size_t)(g + 8) * ld + 8 * kk + tq], ah[1], al[1]);
            split_tf32(sq[(size_t)g * ld + 8 * kk + tq + 4], ah[2], al[2]);
            split_tf32(sq[(size_t)(g + 8) * ld + 8 * kk + tq + 4], ah[3], al[3]);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                uint32_t bh[2], bl[2];
                split_tf32(sk[(size_t)(g + 8 * nt) * ld + 8 * kk + tq], bh[0], bl[0]);
                split_tf32(sk[(size_t)(g + 8 * nt) * ld + 8 * kk + tq + 4], bh[1], bl[1]);
                float (&acc)[4] = nt ? s1 : s0;
                mma_tf32_1688(acc, al, bh[0], bh[1]);
                mma_tf32_1688(acc, ah, bl[0], bl[1]);
                mma_tf32_1688(acc, ah, bh[0], bh[1]);
            }
        }
        // thread holds rows g, g+8; keys {2tq, 2tq+1} (s0) and {8+2tq, 9+2tq} (s1)
        // reference order (cuboid_transformer.py:849-861): q * scale, then q k^T, + bias, softmax
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            const int i = g + 8 * rh;
            const int iu = i / rows_per, ip = i - iu * rows_per;   // (line of the pair, position) of the query slot
            float v[4] = {s0[2 * rh], s0[2 * rh + 1], s1[2 * rh], s1[2 * rh + 1]};
            const int j[4] = {2 * tq, 2 * tq + 1, 8 + 2 * tq, 9 + 2 * tq};
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int ju = j[k] / rows_per, jp = j[k] - ju * rows_per;
                if (ju == iu && jp < L && ip < L) v[k] = v[k] * scale + __ldg(bias_table + (ip - jp + L - 1) * heads + h);
                else v[k] = -INFINITY;
                mx = fmaxf(mx, v[k]);
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            if (mx == -INFINITY) mx = 0.f;
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = expf(v[k] - mx);
                sum += v[k];
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = sum > 0.f ? 1.f / sum : 0.f;
            float* pr = s_p + i * 20;
            *reinterpret_cast<float2*>(pr + 2 * tq) = make_float2(v[0] * inv, v[1] * inv);
            *reinterpret_cast<float2*>(pr + 8 + 2 * tq) = make_float2(v[2] * inv, v[3] * inv);
        }
        __syncwarp();
        // ---- O = P V (16 x HD, K = 16 keys), same split ----
        uint32_t ph[2][4], pl[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            split_tf32(s_p[g * 20 + 8 * ks + tq], ph[ks][0], pl[ks][0]);
            split_tf32(s_p[(g + 8) * 20 + 8 * ks + tq], ph[ks][1], pl[ks][1]);
            split_tf32(s_p[g * 20 + 8 * ks + tq + 4], ph[ks][2], pl[ks][2]);
            split_tf32(s_p[(g + 8) * 20 + 8 * ks + tq + 4], ph[ks][3], pl[ks][3]);
        }
        __syncwarp();   // the scratch tile is free for the next head
        // slot rows g and g + 8 -> (line, position) -> token
        const int u_lo = g / rows_per, p_lo = g - u_lo * rows_per, u_hi = (g + 8) / rows_per, p_hi = g + 8 - u_hi * rows_per;
        const int b_lo = base_of[0], b_hi = base_of[lpb - 1];   // rows 0-7 / 8-15 (the same line unless PAIR)
        const bool ok_lo = p_lo < L && b_lo >= 0, ok_hi = p_hi < L && b_hi >= 0;
        float* o_lo = out + (size_t)((ok_lo ? b_lo : 0) + p_lo * stride) * C + h * HD + 2 * tq;
        float* o_hi = out + (size_t)((ok_hi ? b_hi : 0) + p_hi * stride) * C + h * HD + 2 * tq;
#pragma unroll 2
        for (int nt = 0; nt < HD / 8; ++nt) {
            float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t bh[2], bl[2];
                split_tf32(sv[(size_t)(8 * ks + tq) * ld + 8 * nt + g], bh[0], bl[0]);
                split_tf32(sv[(size_t)(8 * ks + tq + 4) * ld + 8 * nt + g], bh[1], bl[1]);
                mma_tf32_1688(o, pl[ks], bh[0], bh[1]);
                mma_tf32_1688(o, ph[ks], bl[0], bl[1]);
                mma_tf32_1688(o, ph[ks], bh[0], bh[1]);
            }
            if (ok_lo) *reinterpret_cast<float2*>(o_lo + 8 * nt) = make_float2(tf32_rna(o[0]), tf32_rna(o[1]));
            if (ok_hi) *reinterpret_cast<float2*>(o_hi + 8 * nt) = make_float2(tf32_rna(o[2]), tf32_rna(o[3]));
        }
    }
}

// ---- general cuboid attention ---------------------------------------------------------------------------------
// Any cuboid size / strategy ('l' local, 'd' dilated) / shifted window / end padding of CuboidSelfAttentionLayer
// (cuboid_transformer.py:812-966). The reference pads, rolls and reorders the activations into
// (num_cuboids, volume, C) and builds a (num_cuboids, volume, volume) mask; here none of those tensors exist: a
// per-layer table gives, for every (cuboid, slot), the token row it holds (-1 = padding) and its shifted-window region
// label (-1 = masked out), and the kernel gathers q|k|v rows straight from the QKV GEMM output.
//   score(i, j) = q_i k_j / sqrt(hd) + table[rel[i] - rel[j] + rel_off]   if lab[i] == lab[j] >= 0, else masked
// Flash-style: one block per (64-query tile, cuboid, sample x head), 4 warps x 16 query rows, keys streamed through
// shared memory in chunks of 64 with an online softmax; QK^T and PV on warp-level mma.sync m16n8k16 (bf16, fp32
// accumulate). Padding slots contribute zero q/k/v rows (the reference pads after its LayerNorm and qkv has no bias).
constexpr int kQTile = 64;
constexpr int kGlobalRow = 1 << 30;   // key metadata of the global-vector chunk: row = kGlobalRow + index, label = kGlobalLab
constexpr int kGlobalLab = -2;
constexpr int kMaxRelSmem = 4096;    // bias-table rows of one head staged in shared memory (<= 16 KB per block)

// Two schedules, chosen by the launcher (see cuboid_attention): kv_stages == 2 double-buffers the K/V chunks (the gather
// of chunk c+1 is in flight under the math of chunk c); kv_stages == 1 fetches a chunk after the previous one is
// consumed and keeps the footprint at three tiles. n_rel_smem > 0: the head's column of the relative-position table
// sits in shared memory instead of being gathered from L2 per score.
template <int HD>
__global__ void __launch_bounds__(128) cuboid_attention_kernel(const bf16* __restrict__ qkv,
                                                               const float* __restrict__ bias_table,
                                                               bf16* __restrict__ out, const int* __restrict__ tok,
                                                               const int* __restrict__ lab, const int* __restrict__ rel,
                                                               const int* __restrict__ dstp, int N, int C, int heads, int vol,
                                                               int rel_off, int n_rel_smem, int kv_stages,
                                                               const GvKeys gk) {
    grid_dep_launch();
    grid_dep_wait();
    constexpr int LD = HD + 8;        // row pitch: +16 B keeps ldmatrix bank-conflict free
    constexpr int VPR = HD / 8;       // 16-byte vectors per row
    extern __shared__ __align__(16) uint8_t smem_cub[];
    bf16* sQ = reinterpret_cast<bf16*>(smem_cub);
    bf16* sQ2 = sQ + kQTile * LD;                                  // gk.q2 only: the tile's queries for the global keys
    bf16* sKV = sQ + (gk.q2 ? 2 : 1) * kQTile * LD;                // [kv_stages][K | V][64][LD]
    const int n_global = gk.n;
    int* s_qtok = reinterpret_cast<int*>(sKV + kv_stages * 2 * kQTile * LD);
    int* s_qlab = s_qtok + kQTile;
    int* s_qrel = s_qlab + kQTile;
    int* s_kmeta = s_qrel + kQTile;                                // [2 stages][tok | lab | rel][64]
    float* s_bias = reinterpret_cast<float*>(s_kmeta + 6 * kQTile);   // [n_rel_smem] (this head's table column)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int c = blockIdx.y, b = blockIdx.z / heads, h = blockIdx.z - b * heads;
    const int q0 = blockIdx.x * kQTile;
    const int C3 = 3 * C;
    const int* ctok = tok + (size_t)c * vol;
    const int* clab = lab + (size_t)c * vol;
    const bf16* base = qkv + (size_t)b * N * C3 + h * HD;
    const bool bias_in_smem = n_rel_smem > 0;
    // global vectors (cuboid_transformer.py:902-913): after the cuboid's own slots every query sees the n_global <= 64 global
    // keys (GvKeys: k / v rows [B][n_global][ld]) as one more key chunk - never masked, no position bias
    const bf16* gkb = n_global > 0 ? gk.k + (size_t)b * n_global * gk.ld + h * HD : nullptr;
    const bf16* gvb = n_global > 0 ? gk.v + (size_t)b * n_global * gk.ld + h * HD : nullptr;

    if (tid < kQTile) {
        const int i = q0 + tid;
        const bool in = i < vol;
        s_qtok[tid] = in ? ctok[i] : -1;
        s_qlab[tid] = in ? clab[i] : -1;
        s_qrel[tid] = in ? rel[i] : 0;
    }
    for (int i = tid; i < n_rel_smem; i += 128) s_bias[i] = __ldg(bias_table + (size_t)i * heads + h);

    // chunk loader: slot metadata by the first 64 threads, then (after a block barrier) the row gathers by everyone
    const int n_chunks = (vol + kQTile - 1) / kQTile;
    const int n_total = n_chunks + (n_global > 0 ? 1 : 0);
    auto load_meta = [&](int ch, int stage) {
        if (tid < kQTile) {
            int* m = s_kmeta + stage * 3 * kQTile;
            if (ch < n_chunks) {
                const int j = ch * kQTile + tid;
                const bool in = j < vol;
                m[tid] = in ? ctok[j] : -1;
                m[kQTile + tid] = in ? clab[j] : -1;   // slots past the cuboid's end are always masked
                m[2 * kQTile + tid] = in ? rel[j] : 0;
            } else {   // the global keys
                const bool in = tid < n_global;
                m[tid] = in ? kGlobalRow + tid : -1;
                m[kQTile + tid] = in ? kGlobalLab : -1;
                m[2 * kQTile + tid] = 0;
            }
        }
    };
    auto load_rows = [&](int stage) {
        const int* m = s_kmeta + stage * 3 * kQTile;
        bf16* sK = sKV + stage * 2 * kQTile * LD;
        bf16* sV = sK + kQTile * LD;
        for (int i = tid; i < kQTile * VPR; i += 128) {
            const int r = i / VPR, v = i - r * VPR;
            const int t = m[r];
            bf16* dk = sK + r * LD + v * 8;
            bf16* dv = sV + r * LD + v * 8;
            if (t >= kGlobalRow) {
                cp_async16(dk, gkb + (size_t)(t - kGlobalRow) * gk.ld + v * 8);
                cp_async16(dv, gvb + (size_t)(t - kGlobalRow) * gk.ld + v * 8);
            } else if (t >= 0) {
                const bf16* src = base + (size_t)t * C3 + v * 8;
                cp_async16(dk, src + C);
                cp_async16(dv, src + 2 * C);
            } else {
                *reinterpret_cast<uint4*>(dk) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(dv) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    load_meta(0, 0);
    __syncthreads();
    for (int i = tid; i < kQTile * VPR; i += 128) {   // Q tile: same commit group as chunk 0
        const int r = i / VPR, v = i - r * VPR;
        bf16* dst = sQ + r * LD + v * 8;
        const int t = s_qtok[r];
        if (t >= 0) cp_async16(dst, base + (size_t)t * C3 + v * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        if (gk.q2) {   // separate_global_qkv: l2g_q rows of the same tokens
            bf16* dst2 = sQ2 + r * LD + v * 8;
            if (t >= 0) cp_async16(dst2, gk.q2 + ((size_t)b * N + t) * gk.q2_ld + h * HD + v * 8);
            else *reinterpret_cast<uint4*>(dst2) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    load_rows(0);

    const float scale = rsqrtf((float)HD);
    const int r_lo = warp * 16 + g;
    const int qlab[2] = {s_qlab[r_lo], s_qlab[r_lo + 8]};
    const int qrel[2] = {s_qrel[r_lo] + rel_off, s_qrel[r_lo + 8] + rel_off};
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[HD / 8][4];
#pragma unroll
    for (int jn = 0; jn < HD / 8; ++jn) o[jn][0] = o[jn][1] = o[jn][2] = o[jn][3] = 0.f;

    const bool two_stage = kv_stages == 2;
    for (int ch = 0; ch < n_total; ++ch) {
        const int stage = two_stage ? (ch & 1) : 0;
        if (two_stage && ch + 1 < n_total) {   // prefetch the next chunk into the other stage (its last readers
            load_meta(ch + 1, stage ^ 1);          // passed the barrier that closes iteration ch - 1)
            __syncthreads();
            load_rows(stage ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            if (!two_stage && ch > 0) {   // single stage: the chunk is fetched after the previous one is consumed
                load_meta(ch, 0);
                __syncthreads();
                load_rows(0);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const bf16* sK = sKV + stage * 2 * kQTile * LD;
        const bf16* sV = sK + kQTile * LD;
        const int* s_klab = s_kmeta + stage * 3 * kQTile + kQTile;
        const int* s_krel = s_klab + kQTile;

        // ---- S = Q K^T : 16 query rows x 64 keys per warp (8 key tiles of 8) ----
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        const bf16* sQa = (gk.q2 && ch >= n_chunks) ? sQ2 : sQ;   // the global chunk may have its own queries
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
            uint32_t a[4];
            ldmatrix_x4(a, sQa + (size_t)(warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + kk * 16 + 8 * (lane >> 4));
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
                uint32_t bb[4];
                ldmatrix_x4(bb, sK + (size_t)(kb * 16 + (lane & 7) + 8 * (lane >> 4)) * LD + kk * 16 + 8 * ((lane >> 3) & 1));
                mma_bf16_16816(s[2 * kb], a, bb[0], bb[1]);
                mma_bf16_16816(s[2 * kb + 1], a, bb[2], bb[3]);
            }
        }
        // ---- bias, mask (thread: rows g / g+8, keys nt*8 + 2tq + {0,1}) ----
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int jl = nt * 8 + 2 * tq + e;
                const int kl = s_klab[jl], kr = s_krel[jl];
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    float v = -INFINITY;
                    if (kl == kGlobalLab) {
                        v = s[nt][2 * rh + e] * scale;
                    } else if (qlab[rh] >= 0 && kl == qlab[rh]) {
                        const int idx = qrel[rh] - kr;
                        const float bias = bias_in_smem ? s_bias[idx] : __ldg(bias_table + (size_t)idx * heads + h);
                        v = s[nt][2 * rh + e] * scale + bias;
                    }
                    s[nt][2 * rh + e] = v;
                    mx[rh] = fmaxf(mx[rh], v);
                }
            }
        // ---- online softmax ----
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            float m = mx[rh];
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            const float m_new = fmaxf(m_run[rh], m);
            const float alpha = (m_new == -INFINITY) ? 1.f : __expf(m_run[rh] - m_new);
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float v = s[nt][2 * rh + e];
                    const float p = (v == -INFINITY) ? 0.f : __expf(v - m_new);
                    s[nt][2 * rh + e] = p;
                    sum += p;
                }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            l_run[rh] = l_run[rh] * alpha + sum;
            m_run[rh] = m_new;
#pragma unroll
            for (int jn = 0; jn < HD / 8; ++jn) {
                o[jn][2 * rh] *= alpha;
                o[jn][2 * rh + 1] *= alpha;
            }
        }
        // ---- O += P V : the score fragments are already the A fragments of the next mma ----
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
            uint32_t pa[4];
            pa[0] = pack_bf16x2(s[2 * kb][0], s[2 * kb][1]);
            pa[1] = pack_bf16x2(s[2 * kb][2], s[2 * kb][3]);
            pa[2] = pack_bf16x2(s[2 * kb + 1][0], s[2 * kb + 1][1]);
            pa[3] = pack_bf16x2(s[2 * kb + 1][2], s[2 * kb + 1][3]);
#pragma unroll
            for (int jn = 0; jn < HD / 8; jn += 2) {
                uint32_t bb[4];
                ldmatrix_x4_trans(bb, sV + (size_t)(kb * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + 8 * (jn + (lane >> 4)));
                mma_bf16_16816(o[jn], pa, bb[0], bb[1]);
                mma_bf16_16816(o[jn + 1], pa, bb[2], bb[3]);
            }
        }
        __syncthreads();   // this stage (rows + metadata) is free for the next fetch into it
    }
    // ---- normalise and scatter the rows of real tokens (padding slots are dropped = the reference's unpadding) ----
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
        const int r = r_lo + 8 * rh;
        // 'nearest' padding: the slot's destination differs from the token its rows were copied from
        const int t = dstp ? (q0 + r < vol ? dstp[(size_t)c * vol + q0 + r] : -1) : s_qtok[r];
        if (t < 0) continue;
        const float inv = l_run[rh] > 0.f ? 1.f / l_run[rh] : 0.f;
        bf16* dst = out + ((size_t)b * N + t) * C + h * HD + 2 * tq;
#pragma unroll
        for (int jn = 0; jn < HD / 8; ++jn)
            *reinterpret_cast<uint32_t*>(dst + 8 * jn) = pack_bf16x2(o[jn][2 * rh] * inv, o[jn][2 * rh + 1] * inv);
    }
}

// p = softmax(scale * s) per row; one warp per row, L <= 1024, L % 32 == 0.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, bf16* __restrict__ p, int rows,
                                                           int L, float scale) {
    grid_dep_launch();
    grid_dep_wait();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* in = s + (size_t)row * L;
    float v[32];
    const int n = L / 32;
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) {
            v[i] = in[lane + i * 32] * scale;
            mx = fmaxf(mx, v[i]);
        }
    mx = warp_max(mx);
    float den = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) {
            v[i] = __expf(v[i] - mx);
            den += v[i];
        }
    den = warp_sum(den);
    const float inv = 1.f / den;
    bf16* o = p + (size_t)row * L;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) o[lane + i * 32] = __float2bfloat16_rn(v[i] * inv);
}

// in [S][R][ld_in] -> out [S][C][R], 32x32 tiles through padded smem.
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int R,
                                                             int C, int ld_in) {
    grid_dep_launch();
    grid_dep_wait();
    __shared__ bf16 tile[32][34];
    const int s = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int r = r0 + k, c = c0 + tx;
        tile[k][tx] = (r < R && c < C) ? in[((size_t)s * R + r) * ld_in + c] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, r = r0 + tx;
        if (r < R && c < C) out[((size_t)s * C + c) * R + r] = tile[tx][k];
    }
}

}  // namespace

int axial_attention(const void* qkv_v, const float* bias_table, void* out_v, int B, int T, int H, int W, int C, int heads,
                    int axis, cudaStream_t st, int f32, const GvKeys* gkp) {
    const GvKeys gk = gkp ? *gkp : GvKeys();
    const int n_global = gk.n;
    PD_CHECK(n_global >= 0 && n_global <= kMaxLine && (n_global == 0 || (gk.k && gk.v && !f32)), PD_ERR_ARG,
             "axial_attention: %d global vectors (at most %d, bf16 operands)", n_global, kMaxLine);
    PD_CHECK(axis >= 0 && axis <= 2, PD_ERR_ARG, "axial_attention: axis %d", axis);
    const int L = axis == 0 ? T : (axis == 1 ? H : W);
    PD_CHECK(L >= 1 && L <= kMaxLine, PD_ERR_SHAPE,
             "axial_attention: line length %d > %d (only the axial pattern of the shipped config is built)", L, kMaxLine);
    PD_CHECK(C % heads == 0 && C % 8 == 0, PD_ERR_SHAPE, "axial_attention: C=%d heads=%d", C, heads);
    const int hd = C / heads;
    const int lines = B * T * H * W / L;
    const int threads = heads * 32 > 128 ? 128 : heads * 32;
    if (f32) {
        const float* qkv = static_cast<const float*>(qkv_v);
        float* out = static_cast<float*>(out_v);
        const size_t smem = (size_t)kMaxLine * (3 * C + 4) * sizeof(float) + (size_t)(threads / 32) * 16 * 20 * sizeof(float);
        PD_CHECK(smem <= 200 * 1024, PD_ERR_SHAPE, "axial_attention (fp32): line of %zu bytes does not fit in smem", smem);
#define PD_LAUNCH_AXF(HDV)                                                                                              \
    do {                                                                                                                \
        static bool attr_set = false;                                                                                   \
        if (!attr_set) {                                                                                                \
            PD_CUDA(cudaFuncSetAttribute(axial_attention_f32_kernel<HDV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         200 * 1024));                                                                  \
            PD_CUDA(cudaFuncSetAttribute(axial_attention_f32_kernel<HDV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         200 * 1024));                                                                  \
            attr_set = true;                                                                                            \
        }                                                                                                               \
        if (L <= 8)                                                                                                     \
            PD_LAUNCH((axial_attention_f32_kernel<HDV, true>), (lines + 1) / 2, threads, smem, st, qkv, bias_table, out, T, H, W, \
                      C, heads, axis, lines);                                                                           \
        else                                                                                                            \
            PD_LAUNCH((axial_attention_f32_kernel<HDV, false>), lines, threads, smem, st, qkv, bias_table, out, T, H, W, C,   \
                      heads, axis, lines);                                                                              \
    } while (0)
        switch (hd) {
            case 16: PD_LAUNCH_AXF(16); break;
            case 32: PD_LAUNCH_AXF(32); break;
            case 64: PD_LAUNCH_AXF(64); break;
            case 128: PD_LAUNCH_AXF(128); break;
            default: set_error("axial_attention: unsupported head dim %d", hd); return PD_ERR_SHAPE;
        }
#undef PD_LAUNCH_AXF
        PD_LAUNCH_CHECK();
        return PD_OK;
    }
    const bf16* qkv = static_cast<const bf16*>(qkv_v);
    bf16* out = static_cast<bf16*>(out_v);
    const size_t smem = (size_t)kMaxLine * (3 * C + 8) * sizeof(bf16) + (n_global ? (size_t)kMaxLine * (2 * C + 8) * sizeof(bf16) : 0) +
                        (n_global && gk.q2 ? (size_t)kMaxLine * (C + 8) * sizeof(bf16) : 0);
#define PD_LAUNCH_AX(HDV)                                                                                            \
    do {                                                                                                             \
        static bool attr_set = false;                                                                                \
        if (!attr_set) {                                                                                             \
            PD_CUDA(cudaFuncSetAttribute((axial_attention_kernel<HDV, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         160 * 1024));                                                               \
            PD_CUDA(cudaFuncSetAttribute((axial_attention_kernel<HDV, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         160 * 1024));                                                               \
            attr_set = true;                                                                                         \
        }                                                                                                            \
        if (n_global)                                                                                                \
            PD_LAUNCH((axial_attention_kernel<HDV, true>), lines, threads, smem, st, qkv, bias_table, out, T, H, W, C, heads, \
                      axis, gk);                                                                                     \
        else                                                                                                         \
            PD_LAUNCH((axial_attention_kernel<HDV, false>), lines, threads, smem, st, qkv, bias_table, out, T, H, W, C, heads, \
                      axis, gk);                                                                                     \
    } while (0)
    PD_CHECK(smem <= 160 * 1024, PD_ERR_SHAPE, "axial_attention: line of %zu bytes does not fit in smem", smem);
    switch (hd) {
        case 16: PD_LAUNCH_AX(16); break;
        case 32: PD_LAUNCH_AX(32); break;
        case 64: PD_LAUNCH_AX(64); break;
        case 128: PD_LAUNCH_AX(128); break;
        default: set_error("axial_attention: unsupported head dim %d", hd); return PD_ERR_SHAPE;
    }
#undef PD_LAUNCH_AX
    PD_LAUNCH_CHECK();
    return PD_OK;
}

// Host-side geometry of one CuboidSelfAttentionLayer (same tables as prediff_b200/patterns.py::layer_geometry; both are
// tested against the unmodified reference's cuboid_reorder / compute_cuboid_self_attention_mask).
int build_cuboid_tables(int T, int H, int W, const CuboidLayerSpec& spec, int padding_type, CuboidTables* g) {
    PD_CHECK(padding_type >= 0 && padding_type <= 2, PD_ERR_ARG,
             "cuboid attention: padding_type %d (0 = zeros, 1 = ignore, 2 = nearest)", padding_type);
    const int dims[3] = {T, H, W};
    int n[3], padded[3];
    for (int a = 0; a < 3; ++a) {
        PD_CHECK(spec.size[a] >= 1 && spec.shift[a] >= 0 && (spec.strategy[a] == 0 || spec.strategy[a] == 1) && dims[a] >= 1,
                 PD_ERR_ARG, "cuboid attention: bad layer spec on axis %d", a);
        // update_cuboid_size_shift_size (cuboid_transformer.py:563-592)
        g->size[a] = spec.size[a];
        g->shift[a] = spec.strategy[a] == 1 ? 0 : spec.shift[a];
        if (dims[a] <= spec.size[a]) {
            g->size[a] = dims[a];
            g->shift[a] = 0;
        }
        PD_CHECK(g->shift[a] < g->size[a], PD_ERR_ARG, "cuboid attention: shift %d >= cuboid size %d", g->shift[a], g->size[a]);
        g->pad[a] = (g->size[a] - dims[a] % g->size[a]) % g->size[a];
        padded[a] = dims[a] + g->pad[a];
        n[a] = padded[a] / g->size[a];
    }
    const int nc = n[0] * n[1] * n[2], vol = g->size[0] * g->size[1] * g->size[2];
    g->num_cuboids = nc;
    g->volume = vol;
    g->tok.assign((size_t)nc * vol, -1);
    g->lab.assign((size_t)nc * vol, -1);
    g->dst.clear();
    g->gmask.clear();
    // 'nearest' (models/utils.py:228-270): the padded grid is F.interpolate(x, size = padded) - position o copies token
    // floor(o * dims / padded) - and the result is F.interpolate(y, size = dims): token t takes position
    // floor(t * padded / dims). Both with torch's float32 index arithmetic (scale = in / out, src = min(floorf(dst * scale),
    // in - 1)). Without padding on an axis both maps are the identity.
    const bool nearest = padding_type == 2 && (g->pad[0] > 0 || g->pad[1] > 0 || g->pad[2] > 0);
    std::vector<int> near_src[3], near_dst[3];   // per axis: padded position -> source token / destination token (-1: none)
    if (nearest) {
        g->dst.assign((size_t)nc * vol, -1);
        for (int a = 0; a < 3; ++a) {
            near_src[a].resize(padded[a]);
            near_dst[a].assign(padded[a], -1);
            const float up = (float)dims[a] / (float)padded[a], down = (float)padded[a] / (float)dims[a];
            for (int o = 0; o < padded[a]; ++o) {
                const int sidx = (int)floorf((float)o * up);
                near_src[a][o] = sidx < dims[a] - 1 ? sidx : dims[a] - 1;
            }
            for (int t = 0; t < dims[a]; ++t) {
                const int o = (int)floorf((float)t * down);
                near_dst[a][o < padded[a] - 1 ? o : padded[a] - 1] = t;
            }
        }
    }
    const bool any_shift = g->shift[0] > 0 || g->shift[1] > 0 || g->shift[2] > 0;
    for (int c = 0; c < nc; ++c) {
        const int cub[3] = {c / (n[1] * n[2]), (c / n[2]) % n[1], c % n[2]};
        for (int i = 0; i < vol; ++i) {
            const int inn[3] = {i / (g->size[1] * g->size[2]), (i / g->size[2]) % g->size[1], i % g->size[2]};
            bool valid = true;
            int label = 0, src[3];
            for (int a = 0; a < 3; ++a) {
                // cuboid_reorder (:388-429): 'l' splits the axis as (block, in-block), 'd' as (in-block, block)
                const int p = spec.strategy[a] == 0 ? cub[a] * g->size[a] + inn[a] : inn[a] * n[a] + cub[a];
                // shifted-window region of the rolled frame (:513-522): three slices per axis, one if the shift is 0
                const int la = g->shift[a] == 0 ? 2 : (p < padded[a] - g->size[a] ? 0 : (p < padded[a] - g->shift[a] ? 1 : 2));
                label = label * 3 + la;
                src[a] = any_shift ? (p + g->shift[a]) % padded[a] : p;   // torch.roll(x, -shift)
                valid = valid && src[a] < dims[a];
            }
            const size_t k = (size_t)c * vol + i;
            if (nearest) {   // src[] is the position in the padded (un-rolled) grid
                const int st[3] = {near_src[0][src[0]], near_src[1][src[1]], near_src[2][src[2]]};
                const int dt[3] = {near_dst[0][src[0]], near_dst[1][src[1]], near_dst[2][src[2]]};
                g->tok[k] = (st[0] * H + st[1]) * W + st[2];
                g->dst[k] = (dt[0] >= 0 && dt[1] >= 0 && dt[2] >= 0) ? (dt[0] * H + dt[1]) * W + dt[2] : -1;
                g->lab[k] = label;
                continue;
            }
            g->tok[k] = valid ? (src[0] * H + src[1]) * W + src[2] : -1;
            g->lab[k] = (!valid && padding_type == 1) ? -1 : label;
        }
    }
    // global vectors, 'ignore' padding (:915-924): the mask of the global queries over the nc * vol slots is the validity grid
    // of the padded, rolled frame flattened in RASTER order - the reference applies it to the cuboid-ordered keys unchanged
    if (padding_type == 1) {
        g->gmask.resize((size_t)nc * vol);
        for (int s = 0; s < nc * vol; ++s) {
            const int p[3] = {s / (padded[1] * padded[2]), (s / padded[2]) % padded[1], s % padded[2]};
            bool valid = true;
            for (int a = 0; a < 3; ++a) valid = valid && (any_shift ? (p[a] + g->shift[a]) % padded[a] : p[a]) < dims[a];
            g->gmask[s] = valid ? 1 : 0;
        }
    }
    // relative_position_index[:vol, :vol] is built from the CONSTRUCTOR's cuboid size (:714-734, 855-857)
    const int b1 = spec.size[1], b2 = spec.size[2];
    const int s1 = (2 * b1 - 1) * (2 * b2 - 1), s2 = 2 * b2 - 1;
    g->rel.resize(vol);
    for (int i = 0; i < vol; ++i) g->rel[i] = (i / (b1 * b2)) * s1 + ((i / b2) % b1) * s2 + i % b2;
    g->rel_off = (spec.size[0] - 1) * s1 + (b1 - 1) * s2 + (b2 - 1);
    g->n_rel = (2 * spec.size[0] - 1) * s1;   // rows of relative_position_bias_table
    // the axial fast path: one non-unit axis spanning the whole dimension, no shift / padding / dilation effects
    g->axial_axis = -1;
    int non_unit = 0, ax = 0;
    for (int a = 0; a < 3; ++a)
        if (g->size[a] > 1) { ++non_unit; ax = a; }
    bool same = true;
    for (int a = 0; a < 3; ++a) same = same && g->size[a] == spec.size[a];
    if (non_unit == 1 && same && g->size[ax] == dims[ax] && dims[ax] <= kMaxLine && !any_shift) g->axial_axis = ax;
    return PD_OK;
}

int cuboid_attention(const bf16* qkv, const float* bias_table, bf16* out, int B, int N, int C, int heads,
                     const CuboidDev& g, cudaStream_t st, int impl, const GvKeys* gkp) {
    PD_CHECK(C % heads == 0, PD_ERR_SHAPE, "cuboid_attention: C=%d heads=%d", C, heads);
    const int hd = C / heads;
    const GvKeys gk = gkp ? *gkp : GvKeys();
    const int n_global = gk.n;
    PD_CHECK(n_global >= 0 && n_global <= kQTile && (n_global == 0 || (gk.k && gk.v && impl != 2)), PD_ERR_ARG,
             "cuboid_attention: %d global vectors (at most %d, on the mma.sync kernel)", n_global, kQTile);
    if (n_global == 0 && (impl == 2 || (impl == 0 && cuboid_attention_tc_eligible(hd, g.volume))))
        return cuboid_attention_tc(qkv, bias_table, out, B, N, C, heads, g, st);
    PD_CHECK(g.num_cuboids >= 1 && g.num_cuboids <= 65535 && B * heads <= 65535, PD_ERR_SHAPE,
             "cuboid_attention: %d cuboids, %d sample-heads exceed the grid limits", g.num_cuboids, B * heads);
    dim3 grid(ceil_div(g.volume, kQTile), g.num_cuboids, B * heads);
    // the head's table column goes to shared memory when it fits (PD_CUBOID_GLOBAL_BIAS=1 keeps the global gather: A/B)
    // Schedule (measured on B200, batch 4, profiles/trace_r01f_* = single stage + global gather vs trace_r01g_* = two
    // stages + table column in shared memory): volumes of 128-256 with small tables gain 25-30 % from the pipelined
    // variant; a single-chunk cuboid has nothing to prefetch and only pays the second stage's shared memory (fewer
    // resident blocks); the 24 025-row table of full attention costs 96 KB per block = one block per SM and was 1.7x
    // slower than gathering from L2 with 6+ resident blocks. So: two stages only when there is more than one chunk, the
    // table column in shared memory only up to kMaxRelSmem rows. PD_CUBOID_GLOBAL_BIAS / PD_CUBOID_ONE_STAGE force
    // the other choice (A/B).
    static const bool global_bias = getenv("PD_CUBOID_GLOBAL_BIAS") != nullptr;
    static const bool one_stage = getenv("PD_CUBOID_ONE_STAGE") != nullptr;
    const int n_rel_smem = (!global_bias && g.n_rel <= kMaxRelSmem) ? g.n_rel : 0;
    const bool big_table = g.n_rel > kMaxRelSmem;
    const int kv_stages = (g.volume > kQTile && !one_stage && !big_table) ? 2 : 1;
#define PD_LAUNCH_CUB(HDV)                                                                                          \
    do {                                                                                                            \
        const size_t smem = (size_t)(1 + (gk.q2 ? 1 : 0) + 2 * kv_stages) * kQTile * (HDV + 8) * sizeof(bf16) +     \
                            9 * kQTile * sizeof(int) + (size_t)n_rel_smem * sizeof(float);                          \
        static size_t attr_bytes = 0;                                                                               \
        if (smem > attr_bytes) {                                                                                    \
            PD_CUDA(cudaFuncSetAttribute(cuboid_attention_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                               \
            attr_bytes = smem;                                                                                      \
        }                                                                                                           \
        PD_LAUNCH((cuboid_attention_kernel<HDV>), grid, 128, smem, st, qkv, bias_table, out, g.tok, g.lab, g.rel, g.dst, N, \
                  C, heads, g.volume, g.rel_off, n_rel_smem, kv_stages, gk);                                        \
    } while (0)
    switch (hd) {
        case 16: PD_LAUNCH_CUB(16); break;
        case 32: PD_LAUNCH_CUB(32); break;
        case 64: PD_LAUNCH_CUB(64); break;
        case 128: PD_LAUNCH_CUB(128); break;
        default: set_error("cuboid_attention: unsupported head dim %d", hd); return PD_ERR_SHAPE;
    }
#undef PD_LAUNCH_CUB
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int CuboidTablesDev::upload(const CuboidTables& t) {
    const size_t nt = t.tok.size(), nr = t.rel.size(), nd = t.dst.size(), ng = t.gmask.size();
    if (cudaMalloc(&mem, (2 * nt + nr + nd + ng) * sizeof(int)) != cudaSuccess) {
        mem = nullptr;
        set_error("cuboid tables: cudaMalloc of %zu ints failed", 2 * nt + nr + nd + ng);
        return PD_ERR_CUDA;
    }
    int* p = static_cast<int*>(mem);
    PD_CUDA(cudaMemcpy(p, t.tok.data(), nt * sizeof(int), cudaMemcpyHostToDevice));
    PD_CUDA(cudaMemcpy(p + nt, t.lab.data(), nt * sizeof(int), cudaMemcpyHostToDevice));
    PD_CUDA(cudaMemcpy(p + 2 * nt, t.rel.data(), nr * sizeof(int), cudaMemcpyHostToDevice));
    dev.tok = p;
    dev.lab = p + nt;
    dev.rel = p + 2 * nt;
    dev.dst = nullptr;
    if (nd) {
        PD_CUDA(cudaMemcpy(p + 2 * nt + nr, t.dst.data(), nd * sizeof(int), cudaMemcpyHostToDevice));
        dev.dst = p + 2 * nt + nr;
    }
    dev.gmask = nullptr;
    if (ng) {
        PD_CUDA(cudaMemcpy(p + 2 * nt + nr + nd, t.gmask.data(), ng * sizeof(int), cudaMemcpyHostToDevice));
        dev.gmask = p + 2 * nt + nr + nd;
    }
    dev.num_cuboids = t.num_cuboids;
    dev.volume = t.volume;
    dev.rel_off = t.rel_off;
    dev.n_rel = t.n_rel;
    return PD_OK;
}
CuboidTablesDev::~CuboidTablesDev() {
    if (mem) cudaFree(mem);
}

int softmax_rows(const float* s, bf16* p, int rows, int L, float scale, cudaStream_t st) {
    PD_CHECK(L % 32 == 0 && L <= 1024, PD_ERR_SHAPE, "softmax_rows: L=%d", L);
    PD_LAUNCH(softmax_rows_kernel, ceil_div(rows, 8), 256, 0, st, s, p, rows, L, scale);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int transpose_bf16(const bf16* in, bf16* out, int S, int R, int C, int ld_in, cudaStream_t st) {
    dim3 grid(ceil_div(C, 32), ceil_div(R, 32), S);
    PD_LAUNCH(transpose_bf16_kernel, grid, 256, 0, st, in, out, R, C, ld_in);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
