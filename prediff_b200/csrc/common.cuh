// Common host/device helpers for the prediff_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/prediff_b200.h"

namespace pd {

// ---- error plumbing -------------------------------------------------------------------------
// Every extern "C" entry returns int (0 = ok, negative = failure) and never throws; the message of
// the last failure on this thread is kept for pd_last_error().
// error codes: PD_OK / PD_ERR_* from the public header

void set_error(const char* fmt, ...);
const char* last_error();

#define PD_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::pd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PD_ERR_CUDA;                                                              \
        }                                                                                          \
    } while (0)

#define PD_CHECK(cond, code, ...)                  \
    do {                                           \
        if (!(cond)) {                             \
            ::pd::set_error(__VA_ARGS__);          \
            return (code);                         \
        }                                          \
    } while (0)

#define PD_TRY(expr)                  \
    do {                              \
        int _rc = (expr);             \
        if (_rc != 0) return _rc;     \
    } while (0)

#define PD_LAUNCH_CHECK() PD_CUDA(cudaGetLastError())

typedef __nv_bfloat16 bf16;

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch -----------------------------------------------------------
// The denoise step is a chain of ~300 dependent launches of single-wave kernels, so what the chain costs is the sum of
// (launch latency + prologue + first-operand latency) per kernel. Every kernel on the sampling path therefore
//   * calls grid_dep_launch() first: the next kernel of the stream may be dispatched as soon as all CTAs of this one
//     are resident, and runs its own prologue (barrier init, TMEM allocation, tensor-map fetch, weight-tile TMA loads -
//     nothing that depends on this kernel) under this kernel's execution;
//   * calls grid_dep_wait() before its first access to memory another kernel writes or reads: it returns once the
//     preceding kernel has completed and flushed (which, by induction, means every earlier kernel of the stream has).
// Both are no-ops when the launch carries no programmatic attribute. PD_NO_PDL=1 disables the attribute.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// <<<grid, block, smem, st>>> with the programmatic-stream-serialization attribute (and an optional cluster shape).
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, dim3 cluster,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    unsigned n = 0;
    if (pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster.x * cluster.y * cluster.z > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster.x;
        at[n].val.clusterDim.y = cluster.y;
        at[n].val.clusterDim.z = cluster.z;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define PD_LAUNCH(kernel, grid, block, smem, st, ...) \
    PD_CUDA(::pd::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, dim3(1, 1, 1), __VA_ARGS__))
#endif

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// x * sigmoid(x): 1 FMUL + MUFU.EX2 + FADD + MUFU.RCP + FMUL
__device__ __forceinline__ float silu_f(float x) {
    return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
}
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) = 0.5 x + 0.5 |x| (1 - erfc(|x| / sqrt 2)) with
// erfc(|x| / sqrt 2) = 2^P(|x|), P = degree-6 least-squares fit of log2 erfc on [0, 4.2 sqrt 2] (weighted by the
// sensitivity of GELU to it; |x| clamped at the end of the range, where erfc < 3e-9). fp32 Horner + ex2.approx:
// max abs error 3.6e-7, max rel error 9e-6 where |GELU| > 1e-2 (checked on 1.8M points against scipy erf).
// 11 instructions with ONE MUFU op: the bf16 GEMM epilogues are MUFU-throughput-bound (16 lanes/clk/SM) on this
// function - the former Abramowitz-Stegun 7.1.26 form needed MUFU.RCP + MUFU.EX2 and cost 7000 cycles per 128 x 256
// tile; erff() costs ~85 instructions.
__device__ __forceinline__ float gelu_fast(float x) {
    const float ax = fminf(fabsf(x), 5.939697f);
    float p = fmaf(3.406753603e-05f, ax, -7.722947048e-04f);
    p = fmaf(p, ax, 8.069103584e-03f);
    p = fmaf(p, ax, -5.335454270e-02f);
    p = fmaf(p, ax, -4.588471353e-01f);
    p = fmaf(p, ax, -1.151165724f);
    p = fmaf(p, ax, 2.320643716e-06f);
    const float e = ex2_approx(p);       // erfc(|x| / sqrt 2)
    const float h = 0.5f * fabsf(x);
    return fmaf(0.5f, x, fmaf(-h, e, h));
}

// Two GELUs whose results are stored as bf16 (the A operand of the next GEMM), returned as a packed bf16x2: the same
// max(x, 0) - 0.5 |x| 2^P(|x|) identity, with the exponent polynomial evaluated for both values at once in half2
// (degree 4: in fp16 arithmetic the error floor is ~1e-3 in P whatever the degree) and everything that touches the
// magnitude of the result - the exponential, |x|, the final multiply-add - in fp32. 9.5 instructions per element instead
// of 14.75 (the fused-FFN GELU epilogues are issue-bound: two warps per scheduler, 128 elements per thread and chunk).
// Error vs the exact GELU after the bf16 rounding: rel-RMS 1.530e-3 against 1.529e-3 for exact -> bf16 (2.5 M points,
// N(0, 1.5) and uniform [-8, 8]; tools/gelu_half2_check.py); 12 % of the results differ by one bf16 ulp.
__device__ __forceinline__ uint32_t gelu_pair_bf16(float x0, float x1) {
    const __half2 xh = __floats2half2_rn(x0, x1);
    const __half2 ax = __hmin2(__habs2(xh), __float2half2_rn(5.939697f));
    __half2 p = __hfma2(__float2half2_rn(3.7584315550e-03f), ax, __float2half2_rn(-4.3446286322e-02f));
    p = __hfma2(p, ax, __float2half2_rn(-4.6924023712e-01f));
    p = __hfma2(p, ax, __float2half2_rn(-1.1464713489f));
    p = __hfma2(p, ax, __float2half2_rn(-6.8215604968e-04f));
    const float2 pf = __half22float2(p);
    const float e0 = ex2_approx(pf.x), e1 = ex2_approx(pf.y);   // erfc(|x| / sqrt 2)
    const float r0 = fmaf(-0.5f * fabsf(x0), e0, fmaxf(x0, 0.f));
    const float r1 = fmaf(-0.5f * fabsf(x1), e1, fmaxf(x1, 0.f));
    __nv_bfloat162 t = __floats2bfloat162_rn(r0, r1);
    return *reinterpret_cast<uint32_t*>(&t);
}

// ---- operand precision ------------------------------------------------------------------------------------------
// The tensor-core GEMMs take bf16 operands (default) or tf32 operands (fp32 containers; `precision` = PD_PRECISION_TF32:
// the reference's own GPU arithmetic, cfg.yaml:32 float32_matmul_precision "high"). Every producer of a GEMM operand
// (norm / activation / attention / cast kernels, operand-producing GEMM epilogues, the weight repack) stores either
// bf16, or fp32 rounded to the nearest tf32 - so the tensor core's truncation of the low 13 mantissa bits is exact
// instead of a one-sided error.
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// Stores 4 consecutive operand elements at element index 4 * i4 of y: bf16 (8 bytes) or tf32-rounded fp32 (16 bytes).
template <bool F32>
__device__ __forceinline__ void store_operand4(void* y, size_t i4, float a, float b, float c, float d) {
    if constexpr (F32) {
        reinterpret_cast<float4*>(y)[i4] = make_float4(tf32_rna(a), tf32_rna(b), tf32_rna(c), tf32_rna(d));
    } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(y)[i4] = pk;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- L2 weight prefetch ----------------------------------------------------------------------------------------
// The UNet's bf16 weights (274 MB) do not fit the 126 MB L2, so every GEMM of a denoise step finds its weights in HBM:
// measured with cold weights a GEMM CTA runs 3-10 % longer (tools/gemm_phases.py, PD_PHASE_COLD=1), all of it on the
// step's critical path. Each GEMM-family kernel therefore starts by asking the L2 for the weights of the NEXT GEMM of
// the plan (up to three ranges; every CTA requests its 1/gridDim share with cp.async.bulk.prefetch.L2), which then
// stream in from HBM underneath this kernel's own work. Plan::link_prefetch() chains the ranges.
struct WRange {
    const uint8_t* p[3];
    uint32_t n[3];   // bytes, multiples of 16
};
__device__ __forceinline__ void prefetch_l2_share(const WRange& r, unsigned cta, unsigned nctas) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const uint32_t total = r.n[i];
        if (total == 0) continue;
        const uint32_t share = ((total + nctas - 1) / nctas + 255u) & ~255u;
        uint32_t off = cta * share;
        if (off >= total) continue;
        uint32_t len = total - off < share ? total - off : share;
        const uint8_t* src = r.p[i] + off;
        while (len > 0) {
            const uint32_t piece = len < 16384u ? len : 16384u;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(piece) : "memory");
            src += piece;
            len -= piece;
        }
    }
}

// GroupNorm statistics accumulated by the epilogue that PRODUCES the tensor (the GroupNorm that follows then needs no
// statistics pass of its own): the lane holds 32 consecutive channels of one output row as eight float4 cells whose sums
// and sums of squares are s[8], q[8]; cpg = channels per group (8, 16 or 32). The 32 rows of the warp belong to one
// sample; dst = &sums[(sample * G + first group of this chunk) * 2] in the [S][G][2] double table of gn_stats().
__device__ __forceinline__ void gn_chunk_accumulate(const float (&s)[8], const float (&q)[8], bool valid, int cpg, double* dst,
                                                    int lane) {
    // v[2 g + {0, 1}] = (sum, sum of squares) of group g of this chunk for the lane's row; unused entries are zero
    float v[8];
    int nv;
    if (cpg == 8) {
#pragma unroll
        for (int g = 0; g < 4; ++g) { v[2 * g] = s[2 * g] + s[2 * g + 1]; v[2 * g + 1] = q[2 * g] + q[2 * g + 1]; }
        nv = 8;
    } else if (cpg == 16) {
        v[0] = (s[0] + s[1]) + (s[2] + s[3]); v[1] = (q[0] + q[1]) + (q[2] + q[3]);
        v[2] = (s[4] + s[5]) + (s[6] + s[7]); v[3] = (q[4] + q[5]) + (q[6] + q[7]);
        v[4] = v[5] = v[6] = v[7] = 0.f;
        nv = 4;
    } else {
        v[0] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
        v[1] = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
        v[2] = v[3] = v[4] = v[5] = v[6] = v[7] = 0.f;
        nv = 2;
    }
    if (!valid) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
    // Transposing reduction over the 32 rows: 4 + 2 + 1 + 1 + 1 = 9 shuffles in 5 dependent steps. (The first version ran
    // eight 5-step butterflies one after the other - 40 dependent shuffle / add pairs per chunk with two warps per
    // scheduler to hide them, seen as a serial chain in the ncu source view - and cost the epilogue 6-7 us.)
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
    float a0 = b4 ? v[4] : v[0], a1 = b4 ? v[5] : v[1], a2 = b4 ? v[6] : v[2], a3 = b4 ? v[7] : v[3];
    const float s0 = b4 ? v[0] : v[4], s1 = b4 ? v[1] : v[5], s2 = b4 ? v[2] : v[6], s3 = b4 ? v[3] : v[7];
    a0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    a1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    a2 += __shfl_xor_sync(0xffffffffu, s2, 16);
    a3 += __shfl_xor_sync(0xffffffffu, s3, 16);
    float c0 = b3 ? a2 : a0, c1 = b3 ? a3 : a1;
    const float t0 = b3 ? a0 : a2, t1 = b3 ? a1 : a3;
    c0 += __shfl_xor_sync(0xffffffffu, t0, 8);
    c1 += __shfl_xor_sync(0xffffffffu, t1, 8);
    float e = b2 ? c1 : c0;
    const float u = b2 ? c0 : c1;
    e += __shfl_xor_sync(0xffffffffu, u, 4);
    e += __shfl_xor_sync(0xffffffffu, e, 2);
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    const int k = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);   // the entry this lane quad ended up with
    if ((lane & 3) == 0 && k < nv) atomicAdd(dst + k, (double)e);
}

// For the GEMM epilogues: re-reads the lane's 128-byte row of the swizzled fp32 slab it has just written (cell i sits at
// ((i ^ sw) << 4)) and accumulates its statistics. Only instantiated in the GN = true kernel variants: the shuffle tree
// inside the chunk loop slowed every epilogue by ~5 % (measured) when it was compiled into the common kernels, used or not.
__device__ __forceinline__ void gn_chunk_from_slab(const uint8_t* my_row, uint32_t sw, bool valid, double* sums, int cpg,
                                                   int groups, int gsample, int col0, int lane) {
    float gs[8], gq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4));
        gs[i] = (a.x + a.y) + (a.z + a.w);
        gq[i] = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
    }
    // gsample: GroupNorm sample of the warp's rows (computed once per warp by the caller); col0: first channel of the chunk;
    // cpg is 8, 16 or 32, so the group index is a shift
    const int shift = cpg == 8 ? 3 : (cpg == 16 ? 4 : 5);
    gn_chunk_accumulate(gs, gq, valid, cpg, sums + ((size_t)gsample * groups + (col0 >> shift)) * 2, lane);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(t);
}

}  // namespace pd
