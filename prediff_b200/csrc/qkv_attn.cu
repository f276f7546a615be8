// Fused QKV projection + axial attention core: att = softmax(q k^T / sqrt(hd) + rel-pos bias) v with q|k|v = ln Wqkv^T
// computed inside the kernel - the bf16 q|k|v tensor (20 MB per layer at level 0, batch 4) never exists and two launches
// of the dependent chain (QKV GEMM 10.3 / 10.9 us + axial_attention_kernel 5.5 - 6.9 us per layer) become one.
// Reference: CuboidSelfAttentionLayer.forward (src/prediff/models/cuboid_transformer/cuboid_transformer.py:812-861, 949)
// for the axial cuboids (T,1,1) / (1,H,1) / (1,1,W) of the shipped pattern (cuboid_transformer_patterns.py:21-37):
// `qkv` Linear without bias (:735), q scaled by hd^-0.5 (:849), relative position bias (:855-859), softmax, attn @ v.
//
// The projection is per token, so a CTA is free to pick WHICH 128 tokens form its row tile: it takes a box of the
// (W, H, T) token grid that holds whole attention lines - (W, 8, 1) / (8, H, 1) / (8, 1, T) tokens at level 0 - through
// one rank-5 TMA map over the normalised activations (the box order is the row order of the A tile). Per head the CTA
// then owns every q, k and v row its lines need:
//   warp 0    TMA producer: the A tile (C / 64 k-blocks of 16 KB, resident for all heads of the CTA) and, per (head,
//             k-block), a ring stage [q_h | k_h | v_h rows of Wqkv][64] = three boxes of hd weight rows stacked into one
//             K-major B tile of 3 hd rows;
//   warp 1    one lane issues tcgen05.mma 128 x 3hd x 16 (hd = 128: two of 128 x 192) into a TMEM buffer per head (double
//             buffered when 3 hd <= 256, so the GEMM of head h + 1 runs under the attention of head h);
//   warps 2-9 thread = row: tcgen05.ld -> bf16 (the rounding the separate QKV GEMM applied) -> a staging area in shared
//             memory laid out LINE-major (rows of one attention line adjacent, odd 16-byte strides: the row-per-thread
//             stores and the ldmatrix reads below are both bank-conflict free); then one warp per line: S = Q K^T and
//             O = P V on warp-level mma.sync m16n8k16 with the softmax in the accumulator fragments - a line has 8 - 16
//             tokens, far below a tcgen05 tile, exactly as in axial_attention_kernel (attention.cu) whose arithmetic this
//             repeats instruction for instruction; O goes back through the line's dead q slot to coalesced 16-byte stores.
// Grid: (row tiles, heads / heads per CTA). Level 0 (C 256, hd 64): 104 - 128 tiles x 1 (four heads per CTA, A read once);
// level 1 (C 512, hd 128): 28 - 32 tiles x 4 heads.
#include "gemm.cuh"
#include "ops.cuh"
#include "ptx.cuh"

namespace pd {
namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kTileA = 128 * 64 * 2;   // one k-block of the A tile
constexpr int kMaxKb = 8;              // C <= 512
constexpr int kMaxStages = 6;
constexpr int kBarBytes = 1280;     // mbarriers + TMEM slot (256 B) + the CTA's slice of the bias table (512 B) + line table (512 B)
constexpr int kBiasMaxBytes = 512;

struct QkvAttnParams {
    const float* bias_table;   // [2 L - 1][heads]
    bf16* out;                 // [B][T][H][W][C]
    int T, H, W, C, heads, axis, L;
    int bw, bh, bt;            // token box of a row tile (full extent along `axis`)
    int nw, nh, nt;            // boxes per sample along W, H, T
    int kblocks;               // C / 64
    int hpc;                   // heads per CTA
    int stages;                // weight ring depth
    int stg_off;               // byte offset of the staging area (0 = aliases the dead A tile: hpc == 1)
    int row_stride, line_stride;   // staging: bytes between the rows of a line / between lines (odd multiples of 16)
    int ring_off, bar_off;
    WRange pf;
    unsigned long long* dbg;
    int dbg_x, dbg_y;          // the CTA that writes the stamps
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD>
struct QkvCfg {
    static constexpr int NQ = 3 * HD;                                   // accumulator columns per head
    static constexpr int BUFCOLS = NQ <= 64 ? 64 : NQ <= 128 ? 128 : NQ <= 256 ? 256 : 512;
    static constexpr int NBUF = BUFCOLS <= 256 ? 2 : 1;
    static constexpr int ALLOC = BUFCOLS * NBUF;
    static constexpr int NMMA = NQ <= 256 ? 1 : 2;                      // instructions per k-step
    static constexpr int MMA_N = NQ / NMMA;
    static constexpr int STAGE = NQ * 128;                              // bytes of one weight stage
    static constexpr int NCH = (NQ + 31) / 32;                          // 32-column chunks of the accumulator
};

template <int HD>
__global__ void __launch_bounds__(kThreads, 1)
qkv_attn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ QkvAttnParams p) {
    using Cfg = QkvCfg<HD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem + p.ring_off;
    uint8_t* stg = smem + p.stg_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
    uint64_t* a_full = bars;                       // [kMaxKb]
    uint64_t* b_full = a_full + kMaxKb;            // [kMaxStages]
    uint64_t* b_empty = b_full + kMaxStages;       // [kMaxStages]
    uint64_t* acc_full = b_empty + kMaxStages;     // [2]
    uint64_t* acc_empty = acc_full + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* s_bias = reinterpret_cast<float*>(smem + p.bar_off + 256);   // [2 L - 1][hpc]: the CTA's heads of the table
    int* s_tok = reinterpret_cast<int*>(smem + p.bar_off + 256 + kBiasMaxBytes);   // [lines]: global token of position 0, -1 = outside

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int KB = p.kblocks;
    const int n_it = p.hpc * KB;                   // weight stages this CTA consumes
    const int head0 = blockIdx.y * p.hpc;
    // row tile -> box origin
    int tile = blockIdx.x;
    const int iw = tile % p.nw; tile /= p.nw;
    const int ih = tile % p.nh; tile /= p.nh;
    const int it_ = tile % p.nt;
    const int b = tile / p.nt;
    const int w0 = iw * p.bw, h0 = ih * p.bh, t0 = it_ * p.bt;
    const int rows_box = p.bw * p.bh * p.bt;
    const int L = p.L;
    const int n_lines = rows_box / L;
    const int gstride = p.axis == 2 ? 1 : (p.axis == 1 ? p.W : p.H * p.W);   // tokens between positions of a line
    unsigned long long* dbg = (p.dbg && (int)blockIdx.x == p.dbg_x && (int)blockIdx.y == p.dbg_y) ? p.dbg : nullptr;
#define PD_QSTAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
    const bool stamper = threadIdx.x == 64;
    if (stamper) PD_QSTAMP(0);
    if (stamper && p.dbg && p.dbg_x < 0) {   // every CTA: entry time + SM id (buffer of 32 + 3 * CTAs words)
        unsigned long long gt;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long* d = p.dbg + 32 + 3 * (blockIdx.y * gridDim.x + blockIdx.x);
        d[0] = gt;
        d[2] = smid;
    }
    if (stamper && dbg) {
        unsigned long long gt;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        dbg[8] = gt;
        dbg[11] = smid;
    }
    if (threadIdx.x == 32) prefetch_l2_share(p.pf, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxKb; ++s) ptx::mbar_init(&a_full[s], 1);
        for (int s = 0; s < kMaxStages; ++s) {
            ptx::mbar_init(&b_full[s], 1);
            ptx::mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&acc_full[s], 1);
            ptx::mbar_init(&acc_empty[s], 1);
        }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_w);
    }
    // weight stage = (head hh of the CTA, k-block kb) -> ring slot s: q, k and v rows of the head
    auto load_stage = [&](int hh, int kb, int s) {
        const int h = head0 + hh;
        uint8_t* dst = ring + s * Cfg::STAGE;
        ptx::mbar_arrive_expect_tx(&b_full[s], Cfg::STAGE);
#pragma unroll
        for (int part = 0; part < 3; ++part)
            ptx::tma_load_2d(dst + part * HD * 128, &tmap_w, &b_full[s], kb * 64, part * p.C + h * HD);
    };
    const int n_pre = n_it < p.stages ? n_it : p.stages;
    // the first weight stages do not depend on the preceding kernel: requested before the dependency wait
    if (threadIdx.x == 0) {
        int hh = 0, kb = 0;
        for (int it = 0; it < n_pre; ++it) {   // n_pre <= stages: slot = it
            load_stage(hh, kb, it);
            if (++kb == KB) { kb = 0; ++hh; }
        }
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, Cfg::ALLOC);
        ptx::tmem_relinquish();
    }
    if (warp >= 2) {   // a parameter, not an activation: no dependency on the preceding kernel
        const int n_b = (2 * p.L - 1) * p.hpc;
        for (int i = threadIdx.x - 64; i < n_b; i += 32 * kEpiWarps) {
            const int rel = i / p.hpc, hh = i - rel * p.hpc;
            s_bias[i] = __ldg(p.bias_table + rel * p.heads + head0 + hh);
        }
        for (int line = threadIdx.x - 64; line < n_lines; line += 32 * kEpiWarps) {
            bool ok;
            long tok0;
            if (p.axis == 2) {
                const int hl = line % p.bh, tl = line / p.bh;
                ok = h0 + hl < p.H && t0 + tl < p.T;
                tok0 = ((long)(b * p.T + t0 + tl) * p.H + h0 + hl) * p.W;
            } else if (p.axis == 1) {
                const int wl = line % p.bw, tl = line / p.bw;
                ok = w0 + wl < p.W && t0 + tl < p.T;
                tok0 = ((long)(b * p.T + t0 + tl) * p.H) * p.W + w0 + wl;
            } else {
                const int wl = line % p.bw, hl = line / p.bw;
                ok = w0 + wl < p.W && h0 + hl < p.H;
                tok0 = ((long)b * p.T * p.H + h0 + hl) * p.W + w0 + wl;
            }
            s_tok[line] = ok ? (int)tok0 : -1;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    grid_dep_launch();
    grid_dep_wait();
    if (stamper) PD_QSTAMP(1);

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                ptx::mbar_arrive_expect_tx(&a_full[kb], (uint32_t)rows_box * 128u);
                ptx::tma_load_5d(smem + kb * kTileA, &tmap_a, &a_full[kb], kb * 64, w0, h0, t0, b);
            }
            // stages n_pre ... : n_pre == stages here (or nothing is left), so the slot walk restarts at 0
            int s = 0, par = 0, hh = n_pre / KB, kb = n_pre - hh * KB;
#pragma unroll 1
            for (int it = n_pre; it < n_it; ++it) {
                ptx::mbar_wait(&b_empty[s], par);
                load_stage(hh, kb, s);
                if (++kb == KB) { kb = 0; ++hh; }
                if (++s == p.stages) { s = 0; par ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(128, Cfg::MMA_N);
            int s = 0, par = 0;
#pragma unroll 1
            for (int hh = 0; hh < p.hpc; ++hh) {
                const int buf = hh % Cfg::NBUF;
                if (hh >= Cfg::NBUF) {
                    ptx::mbar_wait(&acc_empty[buf], ((hh / Cfg::NBUF) - 1) & 1);
                    ptx::tc_fence_after();
                }
                const uint32_t d_addr = tmem_base + buf * Cfg::BUFCOLS;
#pragma unroll 1
                for (int kb = 0; kb < KB; ++kb) {
                    if (hh == 0) ptx::mbar_wait(&a_full[kb], 0);
                    ptx::mbar_wait(&b_full[s], par);
                    ptx::tc_fence_after();
                    if (hh == 0 && kb == 0) PD_QSTAMP(16);
                    const uint32_t a_addr = ptx::smem_u32(smem + kb * kTileA);
                    const uint32_t b_addr = ptx::smem_u32(ring + s * Cfg::STAGE);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int m = 0; m < Cfg::NMMA; ++m)
                            ptx::umma_f16(d_addr + m * Cfg::MMA_N, ptx::make_smem_desc_sw128(a_addr + k * 32),
                                          ptx::make_smem_desc_sw128(b_addr + m * Cfg::MMA_N * 128 + k * 32), idesc,
                                          (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&b_empty[s]);
                    if (++s == p.stages) { s = 0; par ^= 1; }
                }
                ptx::umma_commit(&acc_full[buf]);
                if (hh == 0) PD_QSTAMP(17);
            }
        }
    } else {
        const int e = warp - 2, q = warp & 3, half = e >> 2;
        const int r = q * 32 + lane;                       // row of the tile = TMEM lane
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // staging slot of this thread's row: (line, position along the attended axis)
        int my_line, my_pos;
        if (p.axis == 2) { my_line = r / L; my_pos = r - my_line * L; }
        else if (p.axis == 1) { const int wl = r % p.bw, rest = r / p.bw; my_pos = rest % L; my_line = (rest / L) * p.bw + wl; }
        else { const int per = p.bw * p.bh; my_pos = r / per; my_line = r - my_pos * per; }
        const bool row_used = r < rows_box;
        const uint32_t stg_u32 = ptx::smem_u32(stg);
        const uint32_t my_row = stg_u32 + my_line * p.line_stride + my_pos * p.row_stride;
        // A 16-row MMA group holds G = 16 / L whole lines (two for the 8-token lines of level 1): slot x of a group =
        // (line x / L of the group, position x % L); scores between different lines of a group are masked out.
        const int G = 16 / L;
        const int n_groups = (n_lines + G - 1) / G;
        const int ls = p.line_stride, rs = p.row_stride;
        const int g = lane >> 2, tq = lane & 3;
        const int rq = (lane & 7) + 8 * ((lane >> 3) & 1), rk = (lane & 7) + 8 * (lane >> 4);   // ldmatrix rows of this lane
        const int rq_sub = rq / L, rk_sub = rk / L;
        const int rq_off = rq_sub * ls + (rq - rq_sub * L) * rs, rk_off = rk_sub * ls + (rk - rk_sub * L) * rs;
        int i_sub[2], i_off[2], rel[2][4];   // rel: (pos_i - pos_j + L - 1) * hpc, or -1 for a pair of different lines
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            const int x = g + 8 * rh;
            i_sub[rh] = x / L;
            const int i_pos = x - i_sub[rh] * L;
            i_off[rh] = i_sub[rh] * ls + i_pos * rs;
            const int jj[4] = {2 * tq, 2 * tq + 1, 8 + 2 * tq, 9 + 2 * tq};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j_sub = jj[k] / L, j_pos = jj[k] - j_sub * L;
                rel[rh][k] = (i_sub[rh] < G && j_sub == i_sub[rh]) ? (i_pos - j_pos + L - 1) * p.hpc : -1;
            }
        }
        const float scale = rsqrtf((float)HD);
#pragma unroll 1
        for (int hh = 0; hh < p.hpc; ++hh) {
            const int buf = hh % Cfg::NBUF;
            const int h = head0 + hh;
            ptx::mbar_wait(&acc_full[buf], (hh / Cfg::NBUF) & 1);
            ptx::tc_fence_after();
            if (stamper && hh == 0) PD_QSTAMP(2);
            // ---- accumulator -> bf16 -> staging (chunk c of 32 columns belongs to column half c & 1), two chunks in flight ----
            constexpr int MYMAX = (Cfg::NCH + 1) / 2;
#pragma unroll
            for (int k0 = 0; k0 < MYMAX; k0 += 2) {
                uint32_t v[2][32];
                const int c0 = half + 2 * k0, c1 = c0 + 2;
                const bool has0 = c0 < Cfg::NCH, has1 = k0 + 1 < MYMAX && c1 < Cfg::NCH;   // warp-uniform
                if (has0) ptx::tmem_ld_32x32(t_lane + buf * Cfg::BUFCOLS + c0 * 32, v[0]);
                if (has1) ptx::tmem_ld_32x32(t_lane + buf * Cfg::BUFCOLS + c1 * 32, v[1]);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c = u ? c1 : c0;
                    if (row_used && (u ? has1 : has0)) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (c * 32 + i * 8 < Cfg::NQ)
                                ptx::st_shared_v4(my_row + c * 64 + i * 16,
                                                  pack_bf16x2(__uint_as_float(v[u][8 * i]), __uint_as_float(v[u][8 * i + 1])),
                                                  pack_bf16x2(__uint_as_float(v[u][8 * i + 2]), __uint_as_float(v[u][8 * i + 3])),
                                                  pack_bf16x2(__uint_as_float(v[u][8 * i + 4]), __uint_as_float(v[u][8 * i + 5])),
                                                  pack_bf16x2(__uint_as_float(v[u][8 * i + 6]), __uint_as_float(v[u][8 * i + 7])));
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            if (threadIdx.x == 64) ptx::mbar_arrive(&acc_empty[buf]);   // every warp has read its columns out
            if (stamper && hh == 0) PD_QSTAMP(3);
            // ---- one warp per group of lines ----
#pragma unroll 1
            for (int gi = e; gi < n_groups; gi += kEpiWarps) {
                const int line0 = gi * G;
                const int nl = n_lines - line0 < G ? n_lines - line0 : G;   // lines of this group that exist
                const uint32_t gbase = stg_u32 + line0 * ls;
                // rows of slots without a line read the group's first row (finite values; their scores are masked)
                const uint32_t aq = gbase + (rq_sub < nl ? rq_off : 0), ak = gbase + (rk_sub < nl ? rk_off : 0);
                // ---- S = Q K^T : 16 x 16, two 8-wide key tiles ----
                float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t qa = aq + 16 * (lane >> 4);
                const uint32_t ka = ak + HD * 2 + 16 * ((lane >> 3) & 1);
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) {
                    uint32_t a[4], bb[4];
                    ldsm_x4(a, qa + kk * 32);
                    ldsm_x4(bb, ka + kk * 32);
                    mma_16816(s0, a, bb[0], bb[1]);
                    mma_16816(s1, a, bb[2], bb[3]);
                }
                // thread holds slots g, g+8 (queries) x {2tq, 2tq+1} (s0) and {8+2tq, 9+2tq} (s1) (keys)
                float pr[2][4];
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const bool row_ok = i_sub[rh] < nl;
                    float v[4] = {s0[2 * rh], s0[2 * rh + 1], s1[2 * rh], s1[2 * rh + 1]};
                    float mx = -INFINITY;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (row_ok && rel[rh][k] >= 0) v[k] = v[k] * scale + s_bias[rel[rh][k] + hh];
                        else v[k] = -INFINITY;
                        mx = fmaxf(mx, v[k]);
                    }
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                    if (mx == -INFINITY) mx = 0.f;   // slot without a line
                    float sum = 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[k] = __expf(v[k] - mx);
                        sum += v[k];
                    }
                    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                    const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) pr[rh][k] = v[k] * inv;
                }
                uint32_t pa[4];
                pa[0] = pack_bf16x2(pr[0][0], pr[0][1]);
                pa[1] = pack_bf16x2(pr[1][0], pr[1][1]);
                pa[2] = pack_bf16x2(pr[0][2], pr[0][3]);
                pa[3] = pack_bf16x2(pr[1][2], pr[1][3]);
                // ---- O = P V : 16 x HD, written over the (dead) q columns of the group's rows ----
                const uint32_t va = aq + HD * 4 + 16 * (lane >> 4);
                uint32_t ov[HD / 16][4];
#pragma unroll
                for (int jn = 0; jn < HD / 8; jn += 2) {
                    uint32_t bb[4];
                    ldsm_x4_t(bb, va + jn * 16);
                    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_16816(o0, pa, bb[0], bb[1]);
                    mma_16816(o1, pa, bb[2], bb[3]);
                    ov[jn / 2][0] = pack_bf16x2(o0[0], o0[1]);
                    ov[jn / 2][1] = pack_bf16x2(o1[0], o1[1]);
                    ov[jn / 2][2] = pack_bf16x2(o0[2], o0[3]);
                    ov[jn / 2][3] = pack_bf16x2(o1[2], o1[3]);
                }
                __syncwarp();   // every lane's ldmatrix of q is complete before the slots are overwritten
                {
                    const uint32_t lo = gbase + i_off[0] + tq * 4, hi = gbase + i_off[1] + tq * 4;
                    const bool lo_ok = i_sub[0] < nl, hi_ok = i_sub[1] < nl;
#pragma unroll
                    for (int jn = 0; jn < HD / 8; jn += 2) {
                        if (lo_ok) {
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(lo + jn * 16), "r"(ov[jn / 2][0]) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(lo + jn * 16 + 16), "r"(ov[jn / 2][1]) : "memory");
                        }
                        if (hi_ok) {
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(hi + jn * 16), "r"(ov[jn / 2][2]) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(hi + jn * 16 + 16), "r"(ov[jn / 2][3]) : "memory");
                        }
                    }
                }
                __syncwarp();
                constexpr int VPR = HD / 8;   // 16-byte vectors per output row
#pragma unroll 1
                for (int sub = 0; sub < nl; ++sub) {
                    const int tok0 = s_tok[line0 + sub];
                    if (tok0 < 0) continue;   // a line of the box that lies outside the token grid
                    bf16* obase = p.out + (size_t)tok0 * p.C + h * HD;
                    const uint32_t lbase = gbase + sub * ls;
                    for (int idx = lane; idx < L * VPR; idx += 32) {
                        const int i = idx / VPR, cv = idx - i * VPR;
                        uint4 val;
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                                     : "r"(lbase + i * rs + cv * 16));
                        *reinterpret_cast<uint4*>(obase + (size_t)i * gstride * p.C + cv * 8) = val;
                    }
                }
            }
            if (stamper && hh == 0) PD_QSTAMP(4);
            if (hh + 1 < p.hpc) asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");   // staging free again
        }
        if (stamper) PD_QSTAMP(5);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::ALLOC);
    if (stamper && p.dbg && p.dbg_x < 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.dbg[32 + 3 * (blockIdx.y * gridDim.x + blockIdx.x) + 1] = gt;
    }
    if (stamper && dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg[9] = gt;
        dbg[10] = clock64();
    }
#undef PD_QSTAMP
}

struct QkvAttnOpImpl {
    CUtensorMap tmap_a, tmap_w;
    QkvAttnParams p;
    int hd;
    unsigned grid_x, grid_y;
    size_t smem;
    WRange own_w;
};
static_assert(sizeof(QkvAttnOpImpl) <= sizeof(QkvAttnOp), "QkvAttnOp storage too small");

// smallest box extent <= cap that covers n in the fewest boxes (the last box may hang over the edge: TMA zero-fills the
// rows, the kernel skips the lines)
int fewest_boxes(int n, int cap) {
    if (cap < 1) cap = 1;
    if (cap > n) cap = n;
    const int boxes = (n + cap - 1) / cap;
    return (n + boxes - 1) / boxes;
}

template <int HD>
int launch_hd(const QkvAttnOpImpl& op, cudaStream_t st) {
    static size_t attr_smem = 0;
    if (op.smem > attr_smem) {
        PD_CUDA(cudaFuncSetAttribute(qkv_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.smem));
        attr_smem = op.smem;
    }
    PD_CUDA(launch_pdl(qkv_attn_kernel<HD>, dim3(op.grid_x, op.grid_y), dim3(kThreads), op.smem, st, dim3(1, 1, 1), op.tmap_a,
                       op.tmap_w, op.p));
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace

bool qkv_attn_supported(int T, int H, int W, int C, int heads, int axis) {
    if (axis < 0 || axis > 2 || heads <= 0 || C % heads != 0 || C % 64 != 0 || C > 64 * kMaxKb) return false;
    const int hd = C / heads;
    if (hd != 16 && hd != 32 && hd != 64 && hd != 128) return false;
    const int L = axis == 0 ? T : (axis == 1 ? H : W);
    return L >= 1 && L <= 16 && T <= 256 && H <= 256 && W <= 256;
}

int qkv_attn_make(QkvAttnOp* op_, const bf16* ln, const bf16* wqkv, const float* bias_table, bf16* out, int B, int T, int H,
                  int W, int C, int heads, int axis) {
    PD_TRY(gemm_init());
    QkvAttnOpImpl* op = reinterpret_cast<QkvAttnOpImpl*>(op_);
    PD_CHECK(ln && wqkv && bias_table && out && B >= 1, PD_ERR_ARG, "qkv_attn: null argument");
    PD_CHECK(qkv_attn_supported(T, H, W, C, heads, axis), PD_ERR_SHAPE,
             "qkv_attn: unsupported shape T=%d H=%d W=%d C=%d heads=%d axis=%d", T, H, W, C, heads, axis);
    const int hd = C / heads;
    QkvAttnParams& p = op->p;
    p = QkvAttnParams{};
    p.bias_table = bias_table; p.out = out;
    p.T = T; p.H = H; p.W = W; p.C = C; p.heads = heads; p.axis = axis;
    p.L = axis == 0 ? T : (axis == 1 ? H : W);
    // token box: full extent along the attended axis, then grow W, H, T (in that order) up to 128 rows
    int box[3] = {1, 1, 1};            // W, H, T
    const int dim[3] = {W, H, T};
    const int ax = 2 - axis;           // index into box[] / dim[]
    box[ax] = dim[ax];
    int rows = box[ax];
    for (int d = 0; d < 3; ++d) {
        if (d == ax) continue;
        box[d] = fewest_boxes(dim[d], 128 / rows);
        rows *= box[d];
    }
    p.bw = box[0]; p.bh = box[1]; p.bt = box[2];
    p.nw = ceil_div(W, p.bw); p.nh = ceil_div(H, p.bh); p.nt = ceil_div(T, p.bt);
    p.kblocks = C / 64;
    // heads per CTA: the A tile is read once per CTA, so few CTAs with all heads while that still fills the machine
    const long tiles = (long)B * p.nw * p.nh * p.nt;
    p.hpc = heads;
    while (p.hpc > 1 && p.hpc % 2 == 0 && tiles * (heads / p.hpc) * 2 <= kNumSMs) p.hpc /= 2;
    if (3 * hd > 256) p.hpc = 1;       // a single TMEM buffer: nothing to pipeline between heads
    p.row_stride = 6 * hd + 16;
    p.line_stride = p.L * p.row_stride;
    if ((p.line_stride / 16) % 2 == 0) p.line_stride += 16;
    const int n_lines = 128 / p.L;
    const int stg_bytes = (n_lines * p.line_stride + 1023) / 1024 * 1024;
    const int a_bytes = p.kblocks * kTileA;
    const int stage = 3 * hd * 128;
    p.ring_off = a_bytes;
    const int budget = 227 * 1024 - 1024 - kBarBytes;
    if (p.hpc == 1) {   // staging aliases the A tile + ring once the only accumulator is complete
        p.stg_off = 0;
        int stages = (budget - a_bytes) / stage;
        const int need = p.kblocks;
        p.stages = stages > need ? need : stages;
        if (p.stages > kMaxStages) p.stages = kMaxStages;
        PD_CHECK(p.stages >= 2 || p.stages == need, PD_ERR_SHAPE, "qkv_attn: shared memory too small for C=%d hd=%d", C, hd);
        const int pipe = a_bytes + p.stages * stage;
        p.bar_off = pipe > stg_bytes ? pipe : stg_bytes;
    } else {
        int stages = (budget - a_bytes - stg_bytes) / stage;
        const int need = p.hpc * p.kblocks;
        p.stages = stages > need ? need : stages;
        if (p.stages > kMaxStages) p.stages = kMaxStages;
        PD_CHECK(p.stages >= 2, PD_ERR_SHAPE, "qkv_attn: shared memory too small for C=%d hd=%d", C, hd);
        p.stg_off = a_bytes + p.stages * stage;
        p.bar_off = p.stg_off + stg_bytes;
    }
    PD_CHECK((2 * p.L - 1) * p.hpc * 4 <= kBiasMaxBytes && (long)B * T * H * W < (1l << 30), PD_ERR_SHAPE,
             "qkv_attn: bias slice / token count out of range");
    op->smem = (size_t)p.bar_off + kBarBytes + 1024;
    PD_CHECK(op->smem <= 227 * 1024, PD_ERR_SHAPE, "qkv_attn: %zu bytes of shared memory (C=%d hd=%d)", op->smem, C, hd);
    op->hd = hd;
    op->grid_x = (unsigned)tiles;
    op->grid_y = (unsigned)(heads / p.hpc);
    // A: the normalised activations as (C, W, H, T, B); a row tile = one token box, one 64-channel k-block per load
    const uint64_t dims_a[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)B};
    const uint64_t st_a[4] = {(uint64_t)C * 2, (uint64_t)C * 2 * W, (uint64_t)C * 2 * W * H, (uint64_t)C * 2 * W * H * T};
    const uint32_t box_a[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bt, 1};
    PD_TRY(tmap_encode_sw128(&op->tmap_a, true, 5, ln, dims_a, st_a, box_a));
    const uint64_t dims_w[2] = {(uint64_t)C, (uint64_t)3 * C}, st_w[1] = {(uint64_t)C * 2};
    const uint32_t box_w[2] = {64, (uint32_t)hd};
    PD_TRY(tmap_encode_sw128(&op->tmap_w, true, 2, wqkv, dims_w, st_w, box_w));
    p.pf = WRange{};
    p.dbg = nullptr;
    op->own_w = WRange{};
    op->own_w.p[0] = reinterpret_cast<const uint8_t*>(wqkv);
    op->own_w.n[0] = (uint32_t)((size_t)3 * C * C * 2);
    return PD_OK;
}

void qkv_attn_set_prefetch(QkvAttnOp* op_, const WRange& next) { reinterpret_cast<QkvAttnOpImpl*>(op_)->p.pf = next; }
void qkv_attn_set_dbg(QkvAttnOp* op_, unsigned long long* stamps, int cta_x, int cta_y) {
    QkvAttnOpImpl* op = reinterpret_cast<QkvAttnOpImpl*>(op_);
    op->p.dbg = stamps;
    op->p.dbg_x = cta_x;
    op->p.dbg_y = cta_y;
}
WRange qkv_attn_weights(const QkvAttnOp& op_) { return reinterpret_cast<const QkvAttnOpImpl&>(op_).own_w; }
double qkv_attn_flops(const QkvAttnOp& op_) {
    const QkvAttnParams& p = reinterpret_cast<const QkvAttnOpImpl&>(op_).p;
    const QkvAttnOpImpl& op = reinterpret_cast<const QkvAttnOpImpl&>(op_);
    const double tokens = (double)(op.grid_x / (unsigned)(p.nw * p.nh * p.nt)) * p.T * p.H * p.W;
    return 2.0 * tokens * p.C * 3.0 * p.C + 4.0 * tokens * p.L * p.C;
}

int qkv_attn_launch(const QkvAttnOp& op_, cudaStream_t st) {
    const QkvAttnOpImpl& op = reinterpret_cast<const QkvAttnOpImpl&>(op_);
    switch (op.hd) {
        case 16: return launch_hd<16>(op, st);
        case 32: return launch_hd<32>(op, st);
        case 64: return launch_hd<64>(op, st);
        case 128: return launch_hd<128>(op, st);
        default: set_error("qkv_attn: unsupported head dim %d", op.hd); return PD_ERR_SHAPE;
    }
}

}  // namespace pd
