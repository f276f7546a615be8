// Attention cores. The cuboid attention of the shipped SEVIR-LR config is axial (sequence length 8..16 per
// softmax, 0.25 % of the step's FLOPs), so it is a CUDA-core kernel that reads q/k/v straight out of the QKV
// GEMM output with strided (axial) addressing - the reference's cuboid_reorder / reverse copies never exist.
#include "ops.cuh"

namespace pd {
namespace {

constexpr int kMaxLine = 16;

// One block per line (all heads): stage the line's L rows of [3C] bf16 in smem, then warp h handles head h.
// Two lanes per query row, each owning half of the head dim.
template <int HD>
__global__ void __launch_bounds__(256) axial_attention_kernel(const bf16* __restrict__ qkv,
                                                              const float* __restrict__ bias_table,
                                                              bf16* __restrict__ out, int T, int H, int W, int C,
                                                              int heads, int axis) {
    extern __shared__ __align__(16) uint8_t smem_att[];
    bf16* s_qkv = reinterpret_cast<bf16*>(smem_att);  // [L][3C]
    const int C3 = 3 * C;
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
    // cooperative, coalesced copy of the L token rows
    {
        const int vec_per_row = C3 / 8;  // 16-byte vectors
        const int total = L * vec_per_row;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int r = i / vec_per_row, v = i - r * vec_per_row;
            const uint4* src = reinterpret_cast<const uint4*>(qkv + (size_t)(base + r * stride) * C3) + v;
            reinterpret_cast<uint4*>(s_qkv + (size_t)r * C3)[v] = __ldg(src);
        }
    }
    __syncthreads();

    constexpr int DPL = HD / 2;  // dims per lane
    const int lane = threadIdx.x & 31;
    const int i = lane >> 1;     // query slot
    const int part = lane & 1;
    const float scale = rsqrtf((float)HD);
    for (int h = threadIdx.x >> 5; h < heads; h += blockDim.x >> 5) {
        const bool active = i < L;
        const int qi = active ? i : 0;
        float q[DPL];
        {
            const bf16* qp = s_qkv + (size_t)qi * C3 + h * HD + part * DPL;
#pragma unroll
            for (int d = 0; d < DPL; d += 2) {
                const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(qp + d));
                q[d] = f.x * scale;
                q[d + 1] = f.y * scale;
            }
        }
        float sc[kMaxLine];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kMaxLine; ++j) {
            float acc = 0.f;
            if (j < L) {
                const bf16* kp = s_qkv + (size_t)j * C3 + C + h * HD + part * DPL;
#pragma unroll
                for (int d = 0; d < DPL; d += 2) {
                    const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(kp + d));
                    acc = fmaf(q[d], f.x, acc);
                    acc = fmaf(q[d + 1], f.y, acc);
                }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (j < L) {
                acc += __ldg(bias_table + (qi - j + L - 1) * heads + h);
                mx = fmaxf(mx, acc);
            }
            sc[j] = acc;
        }
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxLine; ++j) {
            const float e = (j < L) ? __expf(sc[j] - mx) : 0.f;
            sc[j] = e;
            den += e;
        }
        const float inv = 1.f / den;
        float o[DPL];
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxLine; ++j) {
            if (j < L) {
                const float pj = sc[j] * inv;
                const bf16* vp = s_qkv + (size_t)j * C3 + 2 * C + h * HD + part * DPL;
#pragma unroll
                for (int d = 0; d < DPL; d += 2) {
                    const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vp + d));
                    o[d] = fmaf(pj, f.x, o[d]);
                    o[d + 1] = fmaf(pj, f.y, o[d + 1]);
                }
            }
        }
        if (active) {
            bf16* op = out + (size_t)(base + i * stride) * C + h * HD + part * DPL;
            if constexpr (DPL % 8 == 0) {
#pragma unroll
                for (int d = 0; d < DPL; d += 8) {
                    uint4 pk;
                    pk.x = pack_bf16x2(o[d], o[d + 1]);
                    pk.y = pack_bf16x2(o[d + 2], o[d + 3]);
                    pk.z = pack_bf16x2(o[d + 4], o[d + 5]);
                    pk.w = pack_bf16x2(o[d + 6], o[d + 7]);
                    *reinterpret_cast<uint4*>(op + d) = pk;
                }
            } else {
#pragma unroll
                for (int d = 0; d < DPL; d += 2) *reinterpret_cast<uint32_t*>(op + d) = pack_bf16x2(o[d], o[d + 1]);
            }
        }
    }
}

// p = softmax(scale * s) per row; one warp per row, L <= 1024, L % 32 == 0.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, bf16* __restrict__ p, int rows,
                                                           int L, float scale) {
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

int axial_attention(const bf16* qkv, const float* bias_table, bf16* out, int B, int T, int H, int W, int C, int heads,
                    int axis, cudaStream_t st) {
    PD_CHECK(axis >= 0 && axis <= 2, PD_ERR_ARG, "axial_attention: axis %d", axis);
    const int L = axis == 0 ? T : (axis == 1 ? H : W);
    PD_CHECK(L >= 1 && L <= kMaxLine, PD_ERR_SHAPE,
             "axial_attention: line length %d > %d (only the axial pattern of the shipped config is built)", L, kMaxLine);
    PD_CHECK(C % heads == 0 && C % 8 == 0, PD_ERR_SHAPE, "axial_attention: C=%d heads=%d", C, heads);
    const int hd = C / heads;
    const int lines = B * T * H * W / L;
    const size_t smem = (size_t)L * 3 * C * sizeof(bf16);
    const int threads = heads * 32 > 256 ? 256 : heads * 32;
#define PD_LAUNCH_AX(HDV)                                                                                            \
    do {                                                                                                             \
        static bool attr_set = false;                                                                                \
        if (!attr_set) {                                                                                             \
            PD_CUDA(cudaFuncSetAttribute(axial_attention_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                         160 * 1024));                                                               \
            attr_set = true;                                                                                         \
        }                                                                                                            \
        axial_attention_kernel<HDV><<<lines, threads, smem, st>>>(qkv, bias_table, out, T, H, W, C, heads, axis);    \
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

int softmax_rows(const float* s, bf16* p, int rows, int L, float scale, cudaStream_t st) {
    PD_CHECK(L % 32 == 0 && L <= 1024, PD_ERR_SHAPE, "softmax_rows: L=%d", L);
    softmax_rows_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(s, p, rows, L, scale);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int transpose_bf16(const bf16* in, bf16* out, int S, int R, int C, int ld_in, cudaStream_t st) {
    dim3 grid(ceil_div(C, 32), ceil_div(R, 32), S);
    transpose_bf16_kernel<<<grid, 256, 0, st>>>(in, out, R, C, ld_in);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
