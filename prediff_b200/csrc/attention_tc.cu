// tcgen05 cuboid self-attention tile: softmax(Q K^T / sqrt(hd) + rel-pos bias, cuboid / shifted-window mask) V for cuboid
// volumes >= 128 (reference: CuboidSelfAttentionLayer.forward, src/prediff/models/cuboid_transformer/
// cuboid_transformer.py:849-861 scores + bias, :531-560 masked softmax, :949 P V; patterns with such volumes:
// cuboid_transformer_patterns.py:40-118 - video_swin_PxM, divided_st, full).
//
// One CTA = one 128-query tile of one (cuboid, head, sample); the keys of the cuboid stream through in chunks of 128.
//   warps 0-3  softmax: thread = query row = TMEM lane. tcgen05.ld the row of S, + bias, mask, online softmax in fp32,
//              P (bf16) -> 128-byte-swizzled K-major smem tile (the A operand of P V); rescales O in TMEM
//              (tcgen05.ld / st) only when a row maximum of the warp moved; final 1/l normalise + scatter to token order.
//   warps 4-6  loaders: cp.async row gathers through the layer's slot table (token row or padding) into swizzled tiles:
//              Q once, K / V chunks in a 2-deep ring, plus the chunk's mask labels / rel-pos codes and the WINDOW of the
//              head's bias-table column that this (query tile, key chunk) pair can touch (a few hundred contiguous rows
//              even for full attention's 24 025-row table - staged per pair, not per column and not gathered per score).
//   warp 7     one lane issues the MMAs: S_c = Q K_c^T (hd/16 x 128x128x16, both operands K-major) into one of two
//              128-column TMEM buffers - issued one chunk ahead so it runs under the softmax of the previous chunk - and
//              O += P_c V_c (8 x 128 x hd x 16, V consumed in its natural [key][channel] layout as an MN-major operand).
// TMEM: 2 x 128 columns of S + hd columns of O (<= 384 of 512). Single-chunk cuboids (volume 128) take a light
// instantiation (one S buffer, no ring, 256 TMEM columns, <= 100 KB smem) so two CTAs share an SM.
#include "ops.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace pd {
namespace {

constexpr int kTile = 128;                 // queries per CTA = keys per chunk
constexpr int kSlab = kTile * 128;         // 128 rows x 128 bytes (64 bf16): one swizzled slab, 16 KB
constexpr int kBiasWin = 2048;             // staged bias-window rows per (tile, chunk); wider windows gather from L2
constexpr int kSoftmaxThreads = 128, kLoaderThreads = 96;   // 8 warps in all: up to 255 registers for the softmax rows
constexpr int kThreadsTc = kSoftmaxThreads + kLoaderThreads + 32;

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// K-major SW128 descriptor (rows of 128 bytes, 8-row groups 1024 B apart) - the GEMM kernels' descriptor.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) { return ptx::make_smem_desc_sw128(addr); }
// MN-major SW128 descriptor: 128-byte rows hold 64 consecutive N (channel) elements of one K (key) index; 8 consecutive
// keys form a 1024-byte swizzle atom (SBO = stride between 8-key groups); LBO = stride between 64-channel blocks.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// kind::f16 instruction descriptor, bf16 x bf16 -> fp32, A K-major, B K-major or MN-major (bit 16).
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

template <int HD, bool MULTI>
struct TcCfg {
    static constexpr int kSlabs = HD / 64;                 // 64-channel slabs per operand tile
    static constexpr int kOpBytes = kSlabs * kSlab;        // one Q / K / V tile
    static constexpr int kStages = MULTI ? 2 : 1;
    static constexpr int kSBufs = MULTI ? 2 : 1;
    static constexpr int kPBytes = 2 * kSlab;              // P: 128 x 128 bf16 = two key slabs
    static constexpr int kOffQ = 0;
    static constexpr int kOffK = kOffQ + kOpBytes;
    static constexpr int kOffV = kOffK + kStages * kOpBytes;
    static constexpr int kOffP = kOffV + kStages * kOpBytes;
    static constexpr int kOffMeta = kOffP + kPBytes;
    // meta: q tok/lab/rel [3][128] ints, per stage k lab/rel [2][128] ints + window (lo, width) + bias window floats
    static constexpr int kMetaBytes = 3 * kTile * 4 + kStages * (2 * kTile * 4 + 16 + kBiasWin * 4) + 64 /*reduction slots*/ + 128 /*barriers*/;
    static constexpr int kSmem = kOffMeta + kMetaBytes + 1024 /*alignment slack*/;
    static constexpr int kTmemCols = MULTI ? 512 : 256;
    static constexpr int kColO = kSBufs * kTile;
};

template <int HD, bool MULTI>
__global__ void __launch_bounds__(kThreadsTc, 1)
cuboid_attention_tc_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias_table, bf16* __restrict__ out,
                           const int* __restrict__ tok, const int* __restrict__ lab, const int* __restrict__ rel, int N, int C,
                           int heads, int vol, int rel_off, int n_rel) {
    using Cfg = TcCfg<HD, MULTI>;
    constexpr int ST = Cfg::kStages, SB = Cfg::kSBufs;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    int* s_qtok = reinterpret_cast<int*>(smem + Cfg::kOffMeta);
    int* s_qlab = s_qtok + kTile;
    int* s_qrel = s_qlab + kTile;
    int* s_kmeta = s_qrel + kTile;                                   // [ST][lab | rel][128]
    int* s_win = s_kmeta + ST * 2 * kTile;                            // [ST][4]: lo, width (0 = gather from global), kmin, kmax
    float* s_bias = reinterpret_cast<float*>(s_win + ST * 4);         // [ST][kBiasWin]
    int* s_red = reinterpret_cast<int*>(s_bias + ST * kBiasWin);      // [0] qmin, [1] qmax
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 16);
    uint64_t* kv_full = bars;            // [2]
    uint64_t* kv_empty = bars + 2;       // [2]
    uint64_t* s_full = bars + 4;         // [2]
    uint64_t* p_ready = bars + 6;
    uint64_t* o_done = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cub = blockIdx.y, b = blockIdx.z / heads, h = blockIdx.z - b * heads;
    const int q0 = blockIdx.x * kTile;
    const int C3 = 3 * C;
    const int* ctok = tok + (size_t)cub * vol;
    const int* clab = lab + (size_t)cub * vol;
    const bf16* base = qkv + (size_t)b * N * C3 + h * HD;
    const int n_chunks = (vol + kTile - 1) / kTile;
    const uint32_t smem_base = ptx::smem_u32(smem);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&kv_full[i], kLoaderThreads);
            ptx::mbar_init(&kv_empty[i], 1);
            ptx::mbar_init(&s_full[i], 1);
        }
        ptx::mbar_init(p_ready, kSoftmaxThreads);
        ptx::mbar_init(o_done, 1);
        ptx::fence_barrier_init();
        s_red[0] = 0x7fffffff;
        s_red[1] = -0x7fffffff;
    }
    if (warp == 7) {
        ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    grid_dep_launch();
    grid_dep_wait();

    // query-tile metadata (first 128 threads) + the range of its rel-pos codes (for the bias windows)
    if (tid < kTile) {
        const int i = q0 + tid;
        const bool in = i < vol;
        const int r = in ? rel[i] : 0;
        s_qtok[tid] = in ? ctok[i] : -1;
        s_qlab[tid] = in ? clab[i] : -1;
        s_qrel[tid] = r;
        int mn = in ? r : 0x7fffffff, mx = in ? r : -0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) {
            atomicMin(&s_red[0], mn);
            atomicMax(&s_red[1], mx);
        }
    }
    __syncthreads();

    if (warp >= 4 && warp < 7) {
        // ================================ loaders ================================
        const int lt = tid - kSoftmaxThreads;   // 0..95
        const int qmin = s_red[0], qmax = s_red[1];
        // row gather into a swizzled tile: row r, 16-byte chunk c of slab s -> s * 16 KB + r * 128 + ((c ^ (r & 7)) << 4)
        auto gather = [&](uint32_t tile_addr, const int* toks_smem_or_null, int slot0, int which) {
            constexpr int CH = HD / 8;   // 16-byte chunks per row
            for (int i = lt; i < kTile * CH; i += kLoaderThreads) {
                const int r = i / CH, c = i - r * CH;
                const int slot = slot0 + r;
                const int t = toks_smem_or_null ? toks_smem_or_null[r] : (slot < vol ? ctok[slot] : -1);
                const uint32_t dst = tile_addr + (uint32_t)((c >> 3) * kSlab + r * 128 + (((c & 7) ^ (r & 7)) << 4));
                if (t >= 0) cp_async16(dst, base + (size_t)t * C3 + which * C + c * 8);
                else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
            }
        };
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % ST;
            if (c >= ST) ptx::mbar_wait(&kv_empty[s], ((c / ST) - 1) & 1);
            const int k0 = c * kTile;
            int* km = s_kmeta + s * 2 * kTile;
            int* win = s_win + s * 4;
            {
                int mn = 0x7fffffff, mx = -0x7fffffff;
                for (int jl = lt; jl < kTile; jl += kLoaderThreads) {
                    const int j = k0 + jl;
                    const bool in = j < vol;
                    const int r = in ? rel[j] : 0;
                    km[jl] = in ? clab[j] : -1;
                    km[kTile + jl] = r;
                    if (in) { mn = min(mn, r); mx = max(mx, r); }
                }
                if (lt == 0) { win[2] = 0x7fffffff; win[3] = -0x7fffffff; }
                named_bar_sync(2, kLoaderThreads);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                if (lane == 0) {
                    atomicMin(&win[2], mn);
                    atomicMax(&win[3], mx);
                }
            }
            if (c == 0) gather(smem_base + Cfg::kOffQ, s_qtok, q0, 0);
            gather(smem_base + Cfg::kOffK + s * Cfg::kOpBytes, nullptr, k0, 1);
            gather(smem_base + Cfg::kOffV + s * Cfg::kOpBytes, nullptr, k0, 2);
            named_bar_sync(2, kLoaderThreads);
            {   // bias window of this (query tile, key chunk): table rows [lo, lo + width)
                const int lo = qmin - win[3] + rel_off;
                const int width = (qmax - win[2] + rel_off) - lo + 1;
                const bool staged = width > 0 && width <= kBiasWin && lo >= 0 && lo + width <= n_rel;
                float* sb = s_bias + s * kBiasWin;
                if (staged)
                    for (int w = lt; w < width; w += kLoaderThreads) sb[w] = __ldg(bias_table + (size_t)(lo + w) * heads + h);
                named_bar_sync(2, kLoaderThreads);   // everyone has read win[2..3] before they are rewritten below
                if (lt == 0) { win[0] = lo; win[1] = staged ? width : 0; }
            }
            cp_async_wait_all();
            ptx::fence_proxy_async();     // generic-proxy writes (cp.async / st.shared) -> visible to the tensor core
            ptx::mbar_arrive(&kv_full[s]);
        }
    } else if (warp == 7) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(kTile, kTile, false);
            constexpr uint32_t idesc_o = idesc_bf16(kTile, HD, true);
            const uint32_t q_addr = smem_base + Cfg::kOffQ, p_addr = smem_base + Cfg::kOffP;
            auto issue_s = [&](int c) {
                const int s = c % ST;
                ptx::mbar_wait(&kv_full[s], (c / ST) & 1);
                ptx::tc_fence_after();
                const uint32_t k_addr = smem_base + Cfg::kOffK + s * Cfg::kOpBytes;
                const uint32_t d = tmem_base + (uint32_t)((c % SB) * kTile);
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) {
                    const uint32_t off = (uint32_t)((kk >> 2) * kSlab + (kk & 3) * 32);
                    ptx::umma_f16(d, desc_kmajor(q_addr + off), desc_kmajor(k_addr + off), idesc_s, kk ? 1u : 0u);
                }
                ptx::umma_commit(&s_full[c % SB]);
            };
            issue_s(0);
            for (int c = 0; c < n_chunks; ++c) {
                if (MULTI && c + 1 < n_chunks) issue_s(c + 1);   // runs under the softmax of chunk c
                ptx::mbar_wait(p_ready, c & 1);
                ptx::tc_fence_after();
                const uint32_t v_addr = smem_base + Cfg::kOffV + (c % ST) * Cfg::kOpBytes;
                const uint32_t d = tmem_base + Cfg::kColO;
#pragma unroll
                for (int kk = 0; kk < kTile / 16; ++kk) {   // 16 keys per MMA: A advances 32 B inside its slab, V 16 rows
                    const uint32_t a_off = (uint32_t)((kk >> 2) * kSlab + (kk & 3) * 32);
                    ptx::umma_f16(d, desc_kmajor(p_addr + a_off), desc_mnmajor(v_addr + kk * 2048, kSlab), idesc_o,
                                  (c | kk) ? 1u : 0u);
                }
                ptx::umma_commit(&kv_empty[c % ST]);
                ptx::umma_commit(o_done);
            }
        }
    } else {
        // ================================ softmax (thread = query row) ================================
        const int r = tid;                        // row of the tile = TMEM lane
        const int qlab = s_qlab[r];
        const int qrel = s_qrel[r] + rel_off;
        const float scale = rsqrtf((float)HD) * 1.4426950408889634f;   // scores kept in log2 units: exp2 below
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        uint8_t* p_row = smem + Cfg::kOffP + r * 128;
        const uint32_t sw = static_cast<uint32_t>(r & 7);
        float m_run = -INFINITY, l_run = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % ST;
            ptx::mbar_wait(&s_full[c % SB], (c / SB) & 1);
            ptx::tc_fence_after();
            const int* km = s_kmeta + s * 2 * kTile;
            const int win_lo = s_win[s * 4], win_w = s_win[s * 4 + 1];
            const float* sb = s_bias + s * kBiasWin;
            float v[kTile];
            float mx = -INFINITY;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(t_lane + (uint32_t)((c % SB) * kTile + g * 32), raw);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int jj = g * 32 + j;
                    const int kl = km[jj];
                    float x = -INFINITY;
                    if (qlab >= 0 && kl == qlab) {
                        const int idx = qrel - km[kTile + jj];
                        const float bias = win_w ? sb[idx - win_lo] : __ldg(bias_table + (size_t)idx * heads + h);
                        x = fmaf(__uint_as_float(raw[j]), scale, bias * 1.4426950408889634f);
                    }
                    v[jj] = x;
                    mx = fmaxf(mx, x);
                }
            }
            const float m_new = fmaxf(m_run, mx);
            const float alpha = (m_new == -INFINITY) ? 1.f : ex2_approx(m_run - m_new);   // m_run = -inf -> 0
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < kTile; ++j) {
                const float p = (v[j] == -INFINITY) ? 0.f : ex2_approx(v[j] - m_new);
                v[j] = p;
                sum += p;
            }
            l_run = l_run * alpha + sum;
            m_run = m_new;
            if (c > 0) {
                ptx::mbar_wait(o_done, (c - 1) & 1);   // P V of the previous chunk is complete: P tile free, O stable
                ptx::tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {   // a row maximum of this warp moved: rescale its O rows
#pragma unroll
                    for (int g = 0; g < HD / 32; ++g) {
                        uint32_t o[32];
                        ptx::tmem_ld_32x32(t_lane + (uint32_t)(Cfg::kColO + g * 32), o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                        ptx::tmem_st_32x32(t_lane + (uint32_t)(Cfg::kColO + g * 32), o);
                    }
                    ptx::tmem_st_wait();
                }
            }
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) {   // 8 probabilities -> one 16-byte cell of the swizzled P tile
                const uint4 pk = make_uint4(pack_bf16x2(v[8 * ch], v[8 * ch + 1]), pack_bf16x2(v[8 * ch + 2], v[8 * ch + 3]),
                                            pack_bf16x2(v[8 * ch + 4], v[8 * ch + 5]), pack_bf16x2(v[8 * ch + 6], v[8 * ch + 7]));
                *reinterpret_cast<uint4*>(p_row + (ch >> 3) * kSlab + (((static_cast<uint32_t>(ch) & 7) ^ sw) << 4)) = pk;
            }
            ptx::fence_proxy_async();
            ptx::tc_fence_before();
            ptx::mbar_arrive(p_ready);
        }
        ptx::mbar_wait(o_done, (n_chunks - 1) & 1);
        ptx::tc_fence_after();
        const int t = s_qtok[r];
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        bf16* dst = out + ((size_t)b * N + (t >= 0 ? t : 0)) * C + h * HD;
#pragma unroll
        for (int g = 0; g < HD / 32; ++g) {
            uint32_t o[32];
            ptx::tmem_ld_32x32(t_lane + (uint32_t)(Cfg::kColO + g * 32), o);
            ptx::tmem_ld_wait();
            if (t >= 0) {   // padding slots are dropped (= the reference's un-padding / reverse reorder)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(dst + g * 32 + 8 * j) =
                        make_uint4(pack_bf16x2(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                                   pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                                   pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                                   pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 7) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int HD, bool MULTI>
int launch_tc(const bf16* qkv, const float* bias_table, bf16* out, int B, int N, int C, int heads, const CuboidDev& g,
              cudaStream_t st) {
    using Cfg = TcCfg<HD, MULTI>;
    static bool attr_set = false;
    if (!attr_set) {
        PD_CUDA(cudaFuncSetAttribute(cuboid_attention_tc_kernel<HD, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::kSmem));
        attr_set = true;
    }
    dim3 grid(ceil_div(g.volume, kTile), g.num_cuboids, B * heads);
    PD_LAUNCH((cuboid_attention_tc_kernel<HD, MULTI>), grid, kThreadsTc, Cfg::kSmem, st, qkv, bias_table, out, g.tok, g.lab,
              g.rel, N, C, heads, g.volume, g.rel_off, g.n_rel);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace

bool cuboid_attention_tc_eligible(int hd, int volume) {
    static const bool off = getenv("PD_CUBOID_NO_TC") != nullptr;
    return !off && volume >= kTile && (hd == 64 || hd == 128);
}

int cuboid_attention_tc(const bf16* qkv, const float* bias_table, bf16* out, int B, int N, int C, int heads,
                        const CuboidDev& g, cudaStream_t st) {
    PD_CHECK(C % heads == 0, PD_ERR_SHAPE, "cuboid_attention_tc: C=%d heads=%d", C, heads);
    const int hd = C / heads;
    PD_CHECK(hd == 64 || hd == 128, PD_ERR_SHAPE, "cuboid_attention_tc: head dim %d (64 or 128)", hd);
    PD_CHECK(g.num_cuboids >= 1 && g.num_cuboids <= 65535 && B * heads <= 65535, PD_ERR_SHAPE,
             "cuboid_attention_tc: %d cuboids, %d sample-heads exceed the grid limits", g.num_cuboids, B * heads);
    const bool multi = g.volume > kTile;
    if (hd == 64) return multi ? launch_tc<64, true>(qkv, bias_table, out, B, N, C, heads, g, st)
                               : launch_tc<64, false>(qkv, bias_table, out, B, N, C, heads, g, st);
    return multi ? launch_tc<128, true>(qkv, bias_table, out, B, N, C, heads, g, st)
                 : launch_tc<128, false>(qkv, bias_table, out, B, N, C, heads, g, st);
}

}  // namespace pd
