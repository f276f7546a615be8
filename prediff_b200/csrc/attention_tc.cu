// tcgen05 cuboid self-attention tile: softmax(Q K^T / sqrt(hd) + rel-pos bias, cuboid / shifted-window mask) V for cuboid
// volumes >= 128 (reference: CuboidSelfAttentionLayer.forward, src/prediff/models/cuboid_transformer/
// cuboid_transformer.py:849-861 scores + bias, :531-560 masked softmax, :949 P V; patterns with such volumes:
// cuboid_transformer_patterns.py:40-118 - video_swin_PxM, divided_st, full).
//
// One CTA = one 128-query tile of one (cuboid, head, sample); the keys of the cuboid stream through in chunks of 128.
//   warps 0-7  softmax: two threads per query row (= TMEM lane; warps w and w + 4 share a lane quarter and split the 128
//              keys of a chunk / the hd columns of O). tcgen05.ld the half row of S, + bias, mask, online softmax in fp32
//              (the two halves exchange their row maxima through shared memory, one 64-thread named barrier per chunk),
//              P (bf16) -> 128-byte-swizzled K-major smem tile (the A operand of P V); rescales O in TMEM
//              (tcgen05.ld / st) only when a row maximum of the warp moved; final 1/l normalise + scatter to token order.
//   warps 8-10 loaders: cp.async row gathers through the layer's slot table (token row or padding) into swizzled tiles:
//              Q once, K / V chunks in a 2-deep ring, plus the chunk's mask labels / rel-pos codes and the WINDOW of the
//              head's bias-table column that this (query tile, key chunk) pair can touch (a few hundred contiguous rows
//              even for full attention's 24 025-row table - staged per pair, not per column and not gathered per score).
//   warp 11    one lane issues the MMAs: S_c = Q K_c^T (hd/16 x 128x128x16, both operands K-major) into one of two
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
constexpr int kSoftmaxThreads = 256, kLoaderThreads = 96;   // 8 + 3 + 1 warps
constexpr int kThreadsTc = kSoftmaxThreads + kLoaderThreads + 32;
constexpr int kLoaderWarp0 = kSoftmaxThreads / 32, kMmaWarp = kLoaderWarp0 + kLoaderThreads / 32;

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// explicit shared-space accesses (pointers derived from the aligned dynamic-smem base compile to generic LD / ST otherwise)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float ldsf(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

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
    static constexpr int kMetaStages = kStages + 1;        // chunk metadata / bias windows are prepared one chunk ahead of K / V
    static constexpr int kPBytes = 2 * kSlab;              // P: 128 x 128 bf16 = two key slabs
    static constexpr int kOffQ = 0;
    static constexpr int kOffK = kOffQ + kOpBytes;
    static constexpr int kOffV = kOffK + kStages * kOpBytes;
    static constexpr int kOffP = kOffV + kStages * kOpBytes;
    static constexpr int kOffMeta = kOffP + kPBytes;
    // meta: q tok/lab/rel [3][128] ints, per stage k lab/rel/tok [3][128] ints + window (lo, width) + bias window floats,
    // row-maximum / row-sum exchange between the two threads of a row [chunk parity][2][128] floats
    static constexpr int kMetaBytes = 3 * kTile * 4 + kMetaStages * (3 * kTile * 4 + 32 + kBiasWin * 4) + 4 * kTile * 4 +
                                      64 /*reduction slots*/ + 128 /*barriers*/;
    static constexpr int kSmem = kOffMeta + kMetaBytes + 1024 /*alignment slack*/;
    static constexpr int kTmemCols = MULTI ? 512 : 256;
    static constexpr int kColO = kSBufs * kTile;
};

template <int HD, bool MULTI>
__global__ void __launch_bounds__(kThreadsTc, 1)
cuboid_attention_tc_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias_table, bf16* __restrict__ out,
                           const int* __restrict__ tok, const int* __restrict__ lab, const int* __restrict__ rel,
                           const int* __restrict__ dstp, int N, int C, int heads, int vol, int rel_off, int n_rel) {
    using Cfg = TcCfg<HD, MULTI>;
    constexpr int ST = Cfg::kStages, SB = Cfg::kSBufs, MST = Cfg::kMetaStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    int* s_qtok = reinterpret_cast<int*>(smem + Cfg::kOffMeta);
    int* s_qlab = s_qtok + kTile;
    int* s_qrel = s_qlab + kTile;
    int* s_kmeta = s_qrel + kTile;                                   // [MST][packed key word | rel | tok][128]
    int* s_win = s_kmeta + MST * 3 * kTile;                           // [MST][8]: lo, width (0 = gather from global), kmin, kmax,
                                                                      //          label min, label max, uniform label or -2
    float* s_bias = reinterpret_cast<float*>(s_win + MST * 8);        // [MST][kBiasWin]
    float* s_xchg = s_bias + MST * kBiasWin;                          // [chunk parity][2][128]: row max / sum of the other half
    int* s_red = reinterpret_cast<int*>(s_xchg + 4 * kTile);          // [0] qmin, [1] qmax
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
    if (warp == kMmaWarp) {
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

    if (warp >= kLoaderWarp0 && warp < kMmaWarp) {
        // ================================ loaders ================================
        const int lt = tid - kSoftmaxThreads;   // 0..95
        const int qmin = s_red[0], qmax = s_red[1];
        // Row gathers into swizzled tiles: row r, 16-byte chunk c of slab s -> s * 16 KB + r * 128 + ((c ^ (r & 7)) << 4).
        // A thread keeps one chunk column c and walks the rows; the token rows of a pass are read from shared memory in one
        // batch before any cp.async is issued (with the slot table read from global inside the loop every iteration was a
        // dependent L2 round trip in front of its cp.async: ~11 k cycles per chunk, measured), and K and V share them.
        constexpr int CH = HD / 8;                        // 16-byte chunks per row
        constexpr int RP = kLoaderThreads / CH;           // rows per pass
        constexpr int NP = (kTile + RP - 1) / RP;         // passes
        const int gc = lt % CH, gr0 = lt / CH;
        const uint32_t g_dst = (uint32_t)((gc >> 3) * kSlab);
        auto gather = [&](const int* toks, uint32_t tile0, int which0, uint32_t tile1, int which1) {
            int t[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int r = gr0 + k * RP;
                t[k] = r < kTile ? toks[r] : -2;
            }
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int r = gr0 + k * RP;
                if (t[k] == -2) continue;
                const uint32_t off = g_dst + (uint32_t)(r * 128 + (((gc & 7) ^ (r & 7)) << 4));
                if (t[k] >= 0) {
                    const bf16* src = base + (size_t)t[k] * C3 + gc * 8;
                    cp_async16(tile0 + off, src + which0 * C);
                    if (which1 >= 0) cp_async16(tile1 + off, src + which1 * C);
                } else {
                    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(tile0 + off), "r"(0u) : "memory");
                    if (which1 >= 0) asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(tile1 + off), "r"(0u) : "memory");
                }
            }
        };
        // Chunk metadata + bias window of chunk c into meta stage c % MST. Nothing here depends on a K / V stage being
        // free, so it is done one chunk ahead, under the previous chunk's copies.
        auto prepare_meta = [&](int c) {
            const int ms = c % MST;
            const int k0 = c * kTile;
            int* km = s_kmeta + ms * 3 * kTile;
            int* win = s_win + ms * 8;
            int mn = 0x7fffffff, mx = -0x7fffffff, lmn = 0x7fffffff, lmx = -0x7fffffff;
            for (int jl = lt; jl < kTile; jl += kLoaderThreads) {
                const int j = k0 + jl;
                const bool in = j < vol;
                const int r = in ? rel[j] : 0;
                const int lb = in ? clab[j] : -1;
                km[jl] = lb;
                km[kTile + jl] = r;
                km[2 * kTile + jl] = in ? ctok[j] : -1;
                if (in) { mn = min(mn, r); mx = max(mx, r); }
                lmn = min(lmn, lb);
                lmx = max(lmx, lb);
            }
            if (lt == 0) { win[2] = 0x7fffffff; win[3] = -0x7fffffff; win[4] = 0x7fffffff; win[5] = -0x7fffffff; }
            named_bar_sync(2, kLoaderThreads);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                lmn = min(lmn, __shfl_xor_sync(0xffffffffu, lmn, o));
                lmx = max(lmx, __shfl_xor_sync(0xffffffffu, lmx, o));
            }
            if (lane == 0) {
                atomicMin(&win[2], mn);
                atomicMax(&win[3], mx);
                atomicMin(&win[4], lmn);
                atomicMax(&win[5], lmx);
            }
            named_bar_sync(2, kLoaderThreads);
            // bias window of this (query tile, key chunk): table rows [lo, lo + width), pre-multiplied by log2(e);
            // packed key word for the softmax loop: (label + 1) << 16 | (kmax - rel_j): window index = (rel_i - qmin) + that
            const int kmin = win[2], kmax = win[3], labmin = win[4], labmax = win[5];
            const int lo = qmin - kmax + rel_off;
            const int width = (qmax - kmin + rel_off) - lo + 1;
            const bool staged = width > 0 && width <= kBiasWin && lo >= 0 && lo + width <= n_rel;
            float* sb = s_bias + ms * kBiasWin;
            if (staged)
                for (int w = lt; w < width; w += kLoaderThreads)
                    sb[w] = 1.4426950408889634f * __ldg(bias_table + (size_t)(lo + w) * heads + h);
            for (int jl = lt; jl < kTile; jl += kLoaderThreads) {
                const int lb = km[jl];
                reinterpret_cast<uint32_t*>(km)[jl] = ((uint32_t)(lb + 1) << 16) | (uint32_t)(lb >= 0 ? kmax - km[kTile + jl] : 0);
            }
            if (lt == 0) {   // win[6]: the label every key of the chunk carries (fast path of the softmax loop), or -2
                win[0] = lo; win[1] = staged ? width : 0; win[6] = (labmin == labmax && labmin >= 0) ? labmin : -2;
            }
            named_bar_sync(2, kLoaderThreads);   // tok rows of this stage are complete before anyone gathers through them
        };
        prepare_meta(0);
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % ST;
            if (c >= ST) ptx::mbar_wait(&kv_empty[s], ((c / ST) - 1) & 1);
            const int* toks = s_kmeta + (c % MST) * 3 * kTile + 2 * kTile;
            if (c == 0) gather(s_qtok, smem_base + Cfg::kOffQ, 0, 0u, -1);
            gather(toks, smem_base + Cfg::kOffK + s * Cfg::kOpBytes, 1, smem_base + Cfg::kOffV + s * Cfg::kOpBytes, 2);
            asm volatile("cp.async.commit_group;" ::: "memory");
            // (meta stage (c + 1) % MST was last read for chunk c - ST, whose P V completed before the kv_empty wait above)
            if (c + 1 < n_chunks) prepare_meta(c + 1);   // under this chunk's copies
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            ptx::fence_proxy_async();     // generic-proxy writes (cp.async / st.shared) -> visible to the tensor core
            ptx::mbar_arrive(&kv_full[s]);
        }
    } else if (warp == kMmaWarp) {
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
        // ================================ softmax (two threads per query row) ================================
        const int q = warp & 3, hf = warp >> 2;   // TMEM lane quarter; column half (keys of a chunk / channels of O)
        const int r = q * 32 + lane;              // row of the tile = TMEM lane
        constexpr int KH = kTile / 2;             // keys per thread and chunk
        constexpr int OH = HD / 2;                // O columns per thread
        const int qlab = s_qlab[r];
        const int qd = qlab >= 0 ? s_qrel[r] - s_red[0] : 0;   // my rel-pos code relative to the tile's minimum (window row offset)
        const float scale = rsqrtf((float)HD) * 1.4426950408889634f;   // scores kept in log2 units: exp2 below
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* p_row = smem + Cfg::kOffP + hf * kSlab + r * 128;     // my 64 keys = key slab hf of the P tile
        const uint32_t sw = static_cast<uint32_t>(r & 7);
        float m_run = -INFINITY, l_run = 0.f;     // l_run: partial row sum over my column half
        for (int c = 0; c < n_chunks; ++c) {
            ptx::mbar_wait(&s_full[c % SB], (c / SB) & 1);
            ptx::tc_fence_after();
            const int ms = c % MST;
            const uint32_t km_a = smem_base + (uint32_t)(Cfg::kOffMeta + (3 * kTile + ms * 3 * kTile + hf * KH) * 4);
            const uint32_t sb_a = smem_base + (uint32_t)(Cfg::kOffMeta + (3 * kTile + MST * 3 * kTile + MST * 8 + ms * kBiasWin) * 4);
            const int win_lo = s_win[ms * 8], win_w = s_win[ms * 8 + 1], uni = s_win[ms * 8 + 6];
            uint32_t raw[KH];
            {
                uint32_t(&lo32)[32] = *reinterpret_cast<uint32_t(*)[32]>(raw);
                uint32_t(&hi32)[32] = *reinterpret_cast<uint32_t(*)[32]>(raw + 32);
                ptx::tmem_ld_32x32(t_lane + (uint32_t)((c % SB) * kTile + hf * KH), lo32);
                ptx::tmem_ld_32x32(t_lane + (uint32_t)((c % SB) * kTile + hf * KH + 32), hi32);
                ptx::tmem_ld_wait();
            }
            float v[KH];
            float mx = -INFINITY;
            const uint32_t qlab1 = (uint32_t)(qlab + 1);        // 0 = masked-out query row
            const uint32_t sbq = sb_a + (uint32_t)qd * 4;       // my row's base inside the staged window
            if (win_w && uni >= 0) {
                // fast path (unshifted windows, no padding in this chunk): every key carries the same label
                const bool on = uni == qlab;
#pragma unroll
                for (int j4 = 0; j4 < KH / 4; ++j4) {
                    const uint4 kp = lds128(km_a + j4 * 16);
                    const uint32_t kw[4] = {kp.x, kp.y, kp.z, kp.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float bias = ldsf(sbq + ((kw[e] & 0xFFFFu) << 2));
                        const float x = on ? fmaf(__uint_as_float(raw[4 * j4 + e]), scale, bias) : -INFINITY;
                        v[4 * j4 + e] = x;
                        mx = fmaxf(mx, x);
                    }
                }
            } else {
#pragma unroll
                for (int j4 = 0; j4 < KH / 4; ++j4) {
                    const uint4 kp = lds128(km_a + j4 * 16);
                    const uint32_t kw[4] = {kp.x, kp.y, kp.z, kp.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = -INFINITY;
                        if (qlab1 != 0u && (kw[e] >> 16) == qlab1) {
                            const uint32_t widx = (uint32_t)qd + (kw[e] & 0xFFFFu);
                            const float bias = win_w ? ldsf(sb_a + (widx << 2))
                                                     : 1.4426950408889634f * __ldg(bias_table + (size_t)(win_lo + (int)widx) * heads + h);
                            x = fmaf(__uint_as_float(raw[4 * j4 + e]), scale, bias);
                        }
                        v[4 * j4 + e] = x;
                        mx = fmaxf(mx, x);
                    }
                }
            }
            // row maximum over both column halves (the partner thread sits in warp w ^ 4)
            float* xc = s_xchg + (c & 1) * 2 * kTile;   // double-buffered by chunk parity: the partner may still be reading
            xc[hf * kTile + r] = mx;                    // the previous chunk's value when this thread gets here
            named_bar_sync(3 + q, 64);
            mx = fmaxf(mx, xc[(hf ^ 1) * kTile + r]);
            const float m_new = fmaxf(m_run, mx);
            const float alpha = (m_new == -INFINITY) ? 1.f : ex2_approx(m_run - m_new);   // m_run = -inf -> 0
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < KH; ++j) {
                const float p = (v[j] == -INFINITY) ? 0.f : ex2_approx(v[j] - m_new);
                v[j] = p;
                sum += p;
            }
            l_run = l_run * alpha + sum;
            m_run = m_new;
            if (c > 0) {
                ptx::mbar_wait(o_done, (c - 1) & 1);   // P V of the previous chunk is complete: P tile free, O stable
                ptx::tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {   // a row maximum of this warp moved: rescale my O columns
#pragma unroll
                    for (int g = 0; g < OH / 32; ++g) {
                        uint32_t o[32];
                        ptx::tmem_ld_32x32(t_lane + (uint32_t)(Cfg::kColO + hf * OH + g * 32), o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                        ptx::tmem_st_32x32(t_lane + (uint32_t)(Cfg::kColO + hf * OH + g * 32), o);
                    }
                    ptx::tmem_st_wait();
                }
            }
#pragma unroll
            for (int ch = 0; ch < KH / 8; ++ch) {   // 8 probabilities -> one 16-byte cell of the swizzled P tile
                const uint4 pk = make_uint4(pack_bf16x2(v[8 * ch], v[8 * ch + 1]), pack_bf16x2(v[8 * ch + 2], v[8 * ch + 3]),
                                            pack_bf16x2(v[8 * ch + 4], v[8 * ch + 5]), pack_bf16x2(v[8 * ch + 6], v[8 * ch + 7]));
                *reinterpret_cast<uint4*>(p_row + ((static_cast<uint32_t>(ch) ^ sw) << 4)) = pk;
            }
            ptx::fence_proxy_async();
            ptx::tc_fence_before();
            ptx::mbar_arrive(p_ready);
        }
        // total row sum = my half + the partner's (same m_run on both sides)
        float* xl = s_xchg + (n_chunks & 1) * 2 * kTile;   // the buffer the last chunk did not use
        xl[hf * kTile + r] = l_run;
        named_bar_sync(3 + q, 64);
        const float l_tot = l_run + xl[(hf ^ 1) * kTile + r];
        ptx::mbar_wait(o_done, (n_chunks - 1) & 1);
        ptx::tc_fence_after();
        // 'nearest' padding: the slot's destination differs from the token its rows were copied from
        const int t = dstp ? (q0 + r < vol ? dstp[(size_t)cub * vol + q0 + r] : -1) : s_qtok[r];
        const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
        bf16* dst = out + ((size_t)b * N + (t >= 0 ? t : 0)) * C + h * HD + hf * OH;
#pragma unroll
        for (int g = 0; g < OH / 32; ++g) {
            uint32_t o[32];
            ptx::tmem_ld_32x32(t_lane + (uint32_t)(Cfg::kColO + hf * OH + g * 32), o);
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
    if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
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
              g.rel, g.dst, N, C, heads, g.volume, g.rel_off, g.n_rel);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace

// Where the tile kernel is chosen (measured at batch 4, graph replay, profiles/cuboid_attention_r02.txt): it wins 1.3-2.8x
// for volumes of 256 and up and for head dim 128; a volume-128, head-dim-64 cuboid (video_swin_2x8 at level 0: one chunk,
// 448 one-tile CTAs) is a latency chain per CTA with nothing to pipeline, where the small mma.sync blocks - six to a SM -
// still win (24 vs 30 us), so that case stays on the warp-level kernel.
bool cuboid_attention_tc_eligible(int hd, int volume) {
    static const bool off = getenv("PD_CUBOID_NO_TC") != nullptr;
    if (off || volume < kTile) return false;
    return hd == 128 || (hd == 64 && volume > kTile);
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
