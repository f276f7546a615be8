// Fused PositionwiseFFN for the width-512 level: x <- x + W2 GELU(W1 ln + b1) + b2 and the LayerNorm that follows, in ONE
// kernel, with the hidden dimension (2048) split over a 4-CTA thread-block cluster (VERDICT r01 "next" 2).
// Reference: PositionwiseFFN.forward (src/prediff/models/cuboid_transformer/cuboid_transformer.py:182-208), pre-norm input
// `ln` produced by the preceding projection epilogue - or, PROJ variant, by the kernel itself: the attention output projection
// + residual + pre-norm (cuboid_transformer.py:952, 1151) run in front (G0 / E0 below). Replaces the level-1 projection GEMM,
// FFN-1 GEMM (persistent kernel, 208 tiles on 148 SMs, bf16 `mid` round trip) and FFN-2 GEMM (K = 2048 on 104 CTAs): 26 row
// tiles x 4 = 104 CTAs, one launch.
//
// CTA j of a cluster owns the 128 rows of its row tile and hidden columns [512 j, 512 j + 512):
//   G0   (PROJ) TMEM columns [0, 128) = att . Wp[128 j ...]^T: eight k-blocks of (att 16 KB + Wp tile 16 KB) through four
//        stages laid over the ring and hidden-slice tiles 4-5;
//   E0   (PROJ) x1 = (G0 + bp) + x for the CTA's 128 columns -> x (TMA); row sums of x1 exchanged with st.async onto the
//        peers' armed barriers; LayerNorm(x1) slice -> bf16 slab -> TMA store into `ln`, one releasing arrival per CTA on
//        every CTA's ln_ready barrier; all four CTAs then load the complete normalised tile as the A operand of G1;
//   G1   acc[c] (TMEM columns [256 c, +256)) = ln . W1[512 j + 256 c ...]^T, c = 0, 1: K = 512 in 8 k-blocks of (A 16 KB + W1
//        tile 32 KB). The TMA ring borrows the still-unused hidden-slice region: FOUR stages during G1(0), three during
//        G1(1) (a round trip of a 48 KB stage is ~2 000 cycles under load; two stages ran at 820-925 cycles per k-block
//        against the tensor core's 512);
//   E1   acc[c] + b1 -> GELU -> bf16 -> 128-byte-swizzled K-major tiles: the CTA's 128 x 512 hidden slice (128 KB of
//        shared memory) = the A operand of G2; E1(0) runs under G1(1);
//   G2   partial[128 x 512] (all 512 TMEM columns, re-used) = hidden slice . W2[:, 512 j ...]^T: 8 k-blocks x 2 output halves
//        of 256 through three 32 KB stages; the first four (k-block, half 0) steps only need E1(0) and run under E1(1);
//   R    reduce-scatter of the four partials THROUGH L2: each CTA writes the three 128 x 128 slices of its partial that
//        belong to the other CTAs to a global workspace (thread = row, consecutive lanes -> consecutive 16 bytes), a
//        releasing arrival on each destination's barrier, then CTA j adds the four partials of output columns [128 j, +128) in RANK
//        ORDER (deterministic, batch-invariant) + b2 + residual -> x. (The first version exchanged the slices through
//        distributed shared memory: 192 KB out and 192 KB in per SM at the SM-to-SM network's ~17 B/clk took 22 k cycles,
//        and the row-per-thread global read-modify-write of x another 21 k - measured with the phase stamps below; the
//        kernel was slower than the two GEMMs it replaced. Now the residual tile arrives by TMA into swizzled slabs, is
//        combined in place and bulk-stored.)
//   LN   row statistics over the 512 columns: per-CTA partial sums exchanged through DSMEM (st.async, 8 bytes per row, onto
//        the receiver's armed barrier), totals in rank order -> LayerNorm of the CTA's 128 columns -> bf16 slab -> TMA store (the next layer's
//        pre-norm); optional GroupNorm statistics of the new x rows for the resblock that follows the stack.
//   warp 0: TMA producer; warp 1: TMEM allocation + one lane issuing tcgen05.mma 128 x 256 x 16; warps 2-9: E0 / E1 / R / LN
//   (thread = row; warps w and w + 4 share a TMEM lane quarter and split the columns). The kernel has no cluster-wide
//   barrier besides the split-phase one that publishes the mbarrier initialisation: every remote access is one the
//   receiving CTA waits for.
#include "gemm.cuh"
#include "ops.cuh"
#include "ptx.cuh"

namespace pd {
namespace {

constexpr int kC = 512, kHid = 2048, kCl = 4;
constexpr int kHs = kHid / kCl;              // hidden columns per CTA (512)
constexpr int kOs = kC / kCl;                // output columns finalised per CTA (128)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kTileA = 128 * 64 * 2;         // 16 KB
constexpr int kTileW = 256 * 64 * 2;         // 32 KB
constexpr int kStage = kTileA + kTileW;      // 48 KB
constexpr int kRingBytes = 2 * kStage;       // 96 KB: G1 slots 0 and 1, later the three W2 stages of G2
constexpr int kMidBytes = 8 * kTileA;        // 128 KB: eight k-blocks of the GELU'd hidden slice
constexpr int kPipeBytes = kRingBytes + kMidBytes;   // 224 KB
// G1 stage slots: 0, 1 = the ring; 2 ("X") = hidden-slice tiles 4-6 (written by E1(1), i.e. after all of G1); 3 ("Y") =
// hidden-slice tiles 0-2 (written by E1(0): usable during G1(0) only)
__host__ __device__ constexpr int slot_off(int s) { return s == 0 ? 0 : s == 1 ? kStage : s == 2 ? kRingBytes + 4 * kTileA : kRingBytes; }
// k-block `it` (0-15: 8 of G1(0), 8 of G1(1)) -> slot and how often that slot has been used before
__host__ __device__ constexpr int g1_slot(int it) { return it < 8 ? (it & 3) : (it - 8) % 3; }
__host__ __device__ constexpr int g1_use(int it) { return it < 8 ? (it >> 2) : 2 + (it - 8) / 3; }
constexpr int kG1Pre = 4;                    // W1 tiles requested before the dependency wait
constexpr int kG2Stages = 3;
// reduce phase (aliases the dead ring + hidden slice): per epilogue warp two fp32 slabs (32 rows x 32 columns, 128-byte
// swizzled rows: residual in by TMA, result out by TMA) and one bf16 slab (32 rows x 64 columns) for the LayerNorm output
constexpr int kOffSlabF = 0;                              // 8 warps x 2 x 4 KB
constexpr int kOffSlabB = kOffSlabF + kEpiWarps * 2 * 4096;   // 8 warps x 4 KB
constexpr int kOffLnX = kOffSlabB + kEpiWarps * 4096;     // [2][128] float2: the two column halves of a row inside the CTA
constexpr int kOffLnPeer = kOffLnX + 2 * 128 * 8;         // [4][128] float2: per-CTA row sums, by sender rank
static_assert(kOffLnPeer + kCl * 128 * 8 <= kPipeBytes, "reduce buffers must fit in the dead ring + mid region");
constexpr int kBarBytes = 512;
constexpr int kSmem = kPipeBytes + 1024 + kBarBytes;
// PROJ variant, phase G0 (x1 = x + att Wp^T + bp for the CTA's 128 output columns): four stages of (att tile 16 KB + Wp tile
// 16 KB) = the 96 KB ring + hidden-slice tiles 4-5; E0's residual / LayerNorm slabs sit in hidden-slice tiles 0-3 (8 KB per
// warp) and its row-statistics exchange in tile 7 - all of it dead again before E1 writes the hidden slice.
constexpr int kTileP = 128 * 64 * 2;         // Wp tile: 128 output rows x 64
constexpr int kStage0 = kTileA + kTileP;     // 32 KB
__host__ __device__ constexpr int g0_off(int s) { return s < 3 ? s * kStage0 : kRingBytes + 4 * kTileA; }
constexpr int kOffSlab0 = kRingBytes;                      // 8 warps x 2 x 4 KB (fp32; the first one is re-used for bf16)
constexpr int kOffLn0X = kRingBytes + 7 * kTileA;          // [2][128] float2 (tile 7: no G1 slot covers it)
constexpr int kOffLn0Peer = kOffLn0X + 2 * 128 * 8;        // [4][128] float2
constexpr int kG1PreProj = 3;                // W1 tiles requested once G0 is complete (slot 3 still holds E0's slabs)
constexpr size_t kWsFloatsPerTile = (size_t)kCl * (kCl - 1) * 128 * kOs;   // 4 destinations x 3 senders x 128 x 128

struct FfnClParams {
    const float* b1;
    const float* b2;
    const float* ln_gamma;   // null: no fused LayerNorm output
    const float* ln_beta;
    float* ws;               // [tiles][4 dst][3 src][32 cells][128 rows] float4: partial slices in flight between CTAs
    float ln_eps;
    int M;
    double* gn_sums;
    int gn_cpg, gn_groups, gn_rows;
    unsigned long long* dbg;   // optional clock64() stamps of CTA 0 (tools/ffn_cluster_phases.py): see PD_CSTAMP sites
    // PROJ variant: attention output projection + residual + the FFN's pre-norm in front (x1 = x + att Wp^T + bp)
    const float* bp;
    const float* ln1_gamma;
    const float* ln1_beta;
};

// receive slot of sender s at destination d (s != d): senders in rank order, skipping d itself
__device__ __forceinline__ int recv_slot(int s, int d) { return s < d ? s : s - 1; }

template <bool PROJ, bool GN>
__global__ void __launch_bounds__(kThreads, 1)
ffn_cluster_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w1,
                   const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_x,
                   const __grid_constant__ CUtensorMap tmap_ln, const __grid_constant__ CUtensorMap tmap_att,
                   const __grid_constant__ CUtensorMap tmap_wp, const __grid_constant__ CUtensorMap tmap_ln1,
                   const __grid_constant__ FfnClParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sMid = smem + kRingBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPipeBytes);
    uint64_t* w_full = bars;                 // [4]: G1 slots
    uint64_t* w_empty = w_full + 4;          // [4]
    uint64_t* v_full = w_empty + 4;          // [3]: G2 stages
    uint64_t* v_empty = v_full + kG2Stages;  // [3]
    uint64_t* acc_full = v_empty + kG2Stages;   // [2]: G1(c) complete
    uint64_t* mid_full = acc_full + 2;       // [2]: E1(c) has written its four k-blocks (and read acc[c] out)
    uint64_t* acc2_full = mid_full + 2;      // [2]: output columns [0, 256) of the partial complete / all of it
    uint64_t* res_bar = acc2_full + 2;       // [8 warps][2]: residual slabs landed
    uint64_t* g0_full = res_bar + 2 * kEpiWarps;   // [4]: PROJ, G0 stages
    uint64_t* g0_empty = g0_full + 4;           // [4]
    uint64_t* acc0_full = g0_empty + 4;         // [1]: x1 columns complete in TMEM columns [0, 128)
    uint64_t* stat_bar = acc0_full + 1;         // [1]: the three peers' row sums of x1 have landed (st.async complete_tx)
    uint64_t* ln_ready = stat_bar + 1;          // [1]: every CTA of the cluster has published its LayerNorm(x1) slice
    uint64_t* stat2_bar = ln_ready + 1;         // [1]: the three peers' row sums of the final x have landed
    uint64_t* slices_ready = stat2_bar + 1;     // [1]: the three peers have written this CTA's slices of their partials
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slices_ready + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int j = (int)ptx::cluster_ctarank();            // hidden slice / output-column slice of this CTA
    const int tile = blockIdx.x / kCl;
    const int row_tile = tile * 128;
    unsigned long long* dbg = blockIdx.x == 0 ? p.dbg : nullptr;
#define PD_CSTAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
    const bool stamper = threadIdx.x == 64;   // lane 0 of the first epilogue warp
    if (stamper) PD_CSTAMP(0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) {
            ptx::mbar_init(&w_full[s], 1);
            ptx::mbar_init(&w_empty[s], 1);
        }
        for (int s = 0; s < kG2Stages; ++s) {
            ptx::mbar_init(&v_full[s], 1);
            ptx::mbar_init(&v_empty[s], 1);
        }
        for (int c = 0; c < 2; ++c) {
            ptx::mbar_init(&acc_full[c], 1);
            ptx::mbar_init(&mid_full[c], kEpiWarps);
        }
        ptx::mbar_init(&acc2_full[0], 1);
        ptx::mbar_init(&acc2_full[1], 1);
        for (int s = 0; s < 2 * kEpiWarps; ++s) ptx::mbar_init(&res_bar[s], 1);
        ptx::mbar_init(stat2_bar, 1);
        ptx::mbar_arrive_expect_tx(stat2_bar, (kCl - 1) * 128 * 8);   // armed long before any peer reaches its final LayerNorm
        ptx::mbar_init(slices_ready, kCl - 1);
        if (PROJ) {
            for (int s = 0; s < 4; ++s) {
                ptx::mbar_init(&g0_full[s], 1);
                ptx::mbar_init(&g0_empty[s], 1);
            }
            ptx::mbar_init(acc0_full, 1);
            ptx::mbar_init(stat_bar, 1);
            ptx::mbar_init(ln_ready, kCl);
            ptx::mbar_arrive_expect_tx(stat_bar, (kCl - 1) * 128 * 8);   // armed before any peer can send
        }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_w1);
        ptx::prefetch_tmap(&tmap_w2);
        ptx::prefetch_tmap(&tmap_x);
        ptx::prefetch_tmap(&tmap_ln);
        if (PROJ) {
            ptx::prefetch_tmap(&tmap_att);
            ptx::prefetch_tmap(&tmap_wp);
            ptx::prefetch_tmap(&tmap_ln1);
        }
    }
    // weight tiles of the first stages do not depend on the preceding kernel: requested before the dependency wait
    if (threadIdx.x == 0) {
        if (PROJ) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                ptx::mbar_arrive_expect_tx(&g0_full[it], kStage0);
                ptx::tma_load_2d(smem + g0_off(it) + kTileA, &tmap_wp, &g0_full[it], it * 64, j * kOs);
            }
        } else {
#pragma unroll
            for (int it = 0; it < kG1Pre; ++it) {
                ptx::mbar_arrive_expect_tx(&w_full[g1_slot(it)], kStage);
                ptx::tma_load_2d(smem + slot_off(g1_slot(it)) + kTileA, &tmap_w1, &w_full[g1_slot(it)], it * 64, j * kHs);
            }
        }
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Peers exchange row sums / arrivals through this CTA's barriers and the kernel has no full cluster barrier at all:
    // split-phase - arrive now (the barrier initialisation above is published by fence.mbarrier_init), wait before the
    // first remote access (E0 in the PROJ variant, the hand-over of the partial slices otherwise)
    ptx::cluster_arrive();
    if (PROJ && warp == 2 && lane < 12) {   // the CTA's 128 columns of bp / ln1 gamma / ln1 beta -> L1 (4 lines each)
        const float* v = lane < 4 ? p.bp : (lane < 8 ? p.ln1_gamma : p.ln1_beta);
        ptx::prefetch_l1(v + j * kOs + (lane & 3) * 32);
    }
    grid_dep_launch();
    grid_dep_wait();
    if (stamper) PD_CSTAMP(1);

    // G2 step order: (k-block, output half); the first four steps need E1(0) only
    auto g2_step = [](int i, int* kb, int* nh) {
        if (i < 4) { *kb = i; *nh = 0; }
        else if (i < 8) { *kb = i - 4; *nh = 1; }
        else if (i < 12) { *kb = i - 4; *nh = 0; }
        else { *kb = i - 8; *nh = 1; }
    };

    if (warp == 0) {
        if (lane == 0) {
            constexpr int n_pre = PROJ ? kG1PreProj : kG1Pre;
            if (PROJ) {
#pragma unroll
                for (int it = 0; it < 8; ++it) {   // G0: att k-block + Wp tile per stage
                    const int s = it & 3;
                    uint8_t* st = smem + g0_off(s);
                    if (it >= 4) {
                        ptx::mbar_wait(&g0_empty[s], 0);
                        ptx::mbar_arrive_expect_tx(&g0_full[s], kStage0);
                        ptx::tma_load_2d(st + kTileA, &tmap_wp, &g0_full[s], it * 64, j * kOs);
                    }
                    ptx::tma_load_3d(st, &tmap_att, &g0_full[s], it * 64, row_tile, 0);
                }
                ptx::mbar_wait(acc0_full, 0);   // every G0 MMA has completed: the G0 stages are dead
#pragma unroll
                for (int it = 0; it < n_pre; ++it) {
                    ptx::mbar_arrive_expect_tx(&w_full[g1_slot(it)], kStage);
                    ptx::tma_load_2d(smem + slot_off(g1_slot(it)) + kTileA, &tmap_w1, &w_full[g1_slot(it)], it * 64, j * kHs);
                }
                // LayerNorm(x1) of the whole row tile = the four CTAs' slices, published through L2
                ptx::mbar_wait_cluster(ln_ready, 0);
                ptx::fence_proxy_async_all();
            }
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int c = it >> 3, kb = it & 7, s = g1_slot(it), u = g1_use(it);
                uint8_t* st = smem + slot_off(s);
                if (it >= n_pre) {   // the first W1 tiles were requested in the prologue (PROJ: right after G0)
                    if (u > 0) ptx::mbar_wait(&w_empty[s], (u - 1) & 1);
                    ptx::mbar_arrive_expect_tx(&w_full[s], kStage);
                    ptx::tma_load_2d(st + kTileA, &tmap_w1, &w_full[s], kb * 64, j * kHs + c * 256);
                }
                ptx::tma_load_3d(st, &tmap_a, &w_full[s], kb * 64, row_tile, 0);
            }
            // G2 stages alias ring slots 0 / 1: their last G1 uses (k-blocks 14 and 15) must have been consumed
            ptx::mbar_wait(&w_empty[0], g1_use(14) & 1);
#pragma unroll 1
            for (int i = 0; i < 16; ++i) {
                int kb, nh;
                g2_step(i, &kb, &nh);
                const int s = i % kG2Stages, u = i / kG2Stages;
                if (i == 1) ptx::mbar_wait(&w_empty[1], g1_use(15) & 1);
                if (u > 0) ptx::mbar_wait(&v_empty[s], (u - 1) & 1);
                ptx::mbar_arrive_expect_tx(&v_full[s], kTileW);
                ptx::tma_load_2d(smem + s * kTileW, &tmap_w2, &v_full[s], j * kHs + kb * 64, nh * 256);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 256);
            if (PROJ) {   // G0: TMEM columns [0, 128) = att . Wp[128 j ...]^T
                constexpr uint32_t idesc0 = ptx::make_idesc_bf16(128, kOs);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int s = it & 3;
                    ptx::mbar_wait(&g0_full[s], (it >> 2) & 1);
                    ptx::tc_fence_after();
                    if (it == 0) PD_CSTAMP(22);
                    const uint32_t a_addr = ptx::smem_u32(smem + g0_off(s));
                    const uint32_t b_addr = a_addr + kTileA;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16(tmem_base, ptx::make_smem_desc_sw128(a_addr + k * 32),
                                      ptx::make_smem_desc_sw128(b_addr + k * 32), idesc0, (it | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&g0_empty[s]);
                }
                ptx::umma_commit(acc0_full);
                PD_CSTAMP(23);
            }
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int c = it >> 3, kb = it & 7, s = g1_slot(it), u = g1_use(it);
                ptx::mbar_wait(&w_full[s], u & 1);
                ptx::tc_fence_after();
                if (it == 0) PD_CSTAMP(16);
                const uint32_t a_addr = ptx::smem_u32(smem + slot_off(s));
                const uint32_t b_addr = a_addr + kTileA;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::umma_f16(tmem_base + c * 256, ptx::make_smem_desc_sw128(a_addr + k * 32),
                                  ptx::make_smem_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                ptx::umma_commit(&w_empty[s]);
                if (kb == 7) {
                    ptx::umma_commit(&acc_full[c]);
                    PD_CSTAMP(17 + c);
                }
            }
#pragma unroll 1
            for (int i = 0; i < 16; ++i) {
                int kb, nh;
                g2_step(i, &kb, &nh);
                if (i == 0) { ptx::mbar_wait(&mid_full[0], 0); ptx::tc_fence_after(); }   // hidden k-blocks 0-3, acc[0] read out
                if (i == 4) { PD_CSTAMP(19); ptx::mbar_wait(&mid_full[1], 0); ptx::tc_fence_after(); PD_CSTAMP(20); }   // 4-7
                const int s = i % kG2Stages;
                ptx::mbar_wait(&v_full[s], (i / kG2Stages) & 1);
                ptx::tc_fence_after();
                const uint32_t a_addr = ptx::smem_u32(sMid + kb * kTileA);
                const uint32_t b_addr = ptx::smem_u32(smem + s * kTileW);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::umma_f16(tmem_base + nh * 256, ptx::make_smem_desc_sw128(a_addr + k * 32),
                                  ptx::make_smem_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                ptx::umma_commit(&v_empty[s]);
                if (i == 11) ptx::umma_commit(&acc2_full[0]);   // output half 0 has seen all eight k-blocks
            }
            ptx::umma_commit(&acc2_full[1]);
            PD_CSTAMP(21);
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        if (PROJ) {
            // ---- E0: x1 = (att Wp^T + bp) + x for the CTA's 128 columns -> x; LayerNorm(x1) -> the pre-norm tensor ----
            const int e0 = warp - 2, r0 = q * 32 + lane;
            const int c0 = j * kOs + half * 64, rw0 = row_tile + q * 32;
            uint8_t* slab = smem + kOffSlab0 + e0 * 8192;          // two fp32 slabs (32 rows x 32 columns, swizzled rows)
            uint64_t* rb = res_bar + 2 * e0;
            if (lane == 0) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    ptx::mbar_arrive_expect_tx(&rb[g], 4096);
                    ptx::tma_load_3d(slab + g * 4096, &tmap_x, &rb[g], c0 + g * 32, rw0, 0);
                }
            }
            ptx::mbar_wait(acc0_full, 0);
            ptx::tc_fence_after();
            if (stamper) PD_CSTAMP(24);
            float xr[64];
            {
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_32x32(t_lane + half * 64, v0);
                ptx::tmem_ld_32x32(t_lane + half * 64 + 32, v1);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
#pragma unroll
                for (int i = 0; i < 32; ++i) { xr[i] = __uint_as_float(v0[i]); xr[32 + i] = __uint_as_float(v1[i]); }
            }
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                ptx::mbar_wait(&rb[g], 0);
                uint8_t* my_row = slab + g * 4096 + lane * 128;
                // all eight residual cells first: a generic-address store between two loads would serialise them (the
                // compiler must assume they alias)
                float4 resv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    resv[i] = *reinterpret_cast<const float4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bp + c0 + g * 32 + 4 * i));
                    float4* cell = reinterpret_cast<float4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4));
                    const float4 res = resv[i];
                    float4 a;
                    a.x = (xr[g * 32 + 4 * i] + bb.x) + res.x; a.y = (xr[g * 32 + 4 * i + 1] + bb.y) + res.y;
                    a.z = (xr[g * 32 + 4 * i + 2] + bb.z) + res.z; a.w = (xr[g * 32 + 4 * i + 3] + bb.w) + res.w;
                    *cell = a;
                    xr[g * 32 + 4 * i] = a.x; xr[g * 32 + 4 * i + 1] = a.y; xr[g * 32 + 4 * i + 2] = a.z; xr[g * 32 + 4 * i + 3] = a.w;
                    s1 += (a.x + a.y) + (a.z + a.w);
                    s2 += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
                }
            }
            if (stamper) PD_CSTAMP(26);
            // row statistics over the 512 columns: two threads per row inside the CTA, then the four CTAs (st.async onto the
            // peers' armed barriers: no release fence on the sending side)
            float2* ln_x = reinterpret_cast<float2*>(smem + kOffLn0X);
            float2* ln_peer = reinterpret_cast<float2*>(smem + kOffLn0Peer);
            ln_x[half * 128 + r0] = make_float2(s1, s2);
            asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
            const float2 o = ln_x[(half ^ 1) * 128 + r0];
            const float t1 = half == 0 ? s1 + o.x : o.x + s1, t2 = half == 0 ? s2 + o.y : o.y + s2;
            ptx::cluster_wait();   // peers have initialised their barriers (arrive at kernel start)
            if (half == 0) {
#pragma unroll
                for (int dd = 1; dd < kCl; ++dd) {
                    const uint32_t d = (uint32_t)((j + dd) & (kCl - 1));
                    ptx::st_async_cluster_f32x2(ptx::mapa(ptx::smem_u32(&ln_peer[j * 128 + r0]), d), t1, t2,
                                                ptx::mapa(ptx::smem_u32(stat_bar), d));
                }
            }
            ptx::mbar_wait_cluster(stat_bar, 0);
            float tot1 = 0.f, tot2 = 0.f;
#pragma unroll
            for (int sdr = 0; sdr < kCl; ++sdr) {   // rank order, the CTA's own total in its place
                float2 t = ln_peer[sdr * 128 + r0];
                if (sdr == j) t = make_float2(t1, t2);
                tot1 = sdr == 0 ? t.x : tot1 + t.x;
                tot2 = sdr == 0 ? t.y : tot2 + t.y;
            }
            const float mean = tot1 * (1.0f / kC);
            const float var = fmaxf(tot2 * (1.0f / kC) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + p.ln_eps);
            if (stamper) PD_CSTAMP(27);
            // bf16 slab: the (still empty) A part of G1's ring slots 0 / 1 - refilled only after ln_ready
            uint8_t* bslab = smem + (e0 < 4 ? 0 : kStage) + (e0 & 3) * 4096;
            uint8_t* brow = bslab + lane * 128;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln1_gamma + c0 + 8 * i));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.ln1_gamma + c0 + 8 * i + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln1_beta + c0 + 8 * i));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.ln1_beta + c0 + 8 * i + 4));
                uint4 pk;
                pk.x = pack_bf16x2(fmaf((xr[8 * i] - mean) * rstd, g0.x, b0.x), fmaf((xr[8 * i + 1] - mean) * rstd, g0.y, b0.y));
                pk.y = pack_bf16x2(fmaf((xr[8 * i + 2] - mean) * rstd, g0.z, b0.z), fmaf((xr[8 * i + 3] - mean) * rstd, g0.w, b0.w));
                pk.z = pack_bf16x2(fmaf((xr[8 * i + 4] - mean) * rstd, g1.x, b1.x), fmaf((xr[8 * i + 5] - mean) * rstd, g1.y, b1.y));
                pk.w = pack_bf16x2(fmaf((xr[8 * i + 6] - mean) * rstd, g1.z, b1.z), fmaf((xr[8 * i + 7] - mean) * rstd, g1.w, b1.w));
                *reinterpret_cast<uint4*>(brow + ((static_cast<uint32_t>(i) ^ sw) << 4)) = pk;
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                ptx::tma_store_3d(&tmap_ln1, bslab, c0, rw0, 0);
                ptx::bulk_commit();
                // x1 -> x behind it (only the reduce phase at the end of the kernel reads it back): a younger bulk group
#pragma unroll
                for (int g = 0; g < 2; ++g) ptx::tma_store_3d(&tmap_x, slab + g * 4096, c0 + g * 32, rw0, 0);   // rows past M are clipped
                ptx::bulk_commit();
                ptx::bulk_wait_all<1>();            // LayerNorm(x1) of this warp's rows is written
                if (stamper) PD_CSTAMP(28);
            }
            // one release per CTA: the eight warps' slabs are complete at the named barrier, one thread then arrives on the
            // ln_ready barrier of every CTA of the cluster (release at cluster scope; the readers acquire, then fence the
            // async proxy before their TMA loads)
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            if (threadIdx.x == 64) {
#pragma unroll
                for (int d = 0; d < kCl; ++d) ptx::mbar_arrive_remote(ptx::mapa(ptx::smem_u32(ln_ready), (uint32_t)d));
            }
            if (lane == 0) ptx::bulk_wait_read<0>();   // the x stores have read the slabs: E1 may overwrite them
            __syncwarp();
            if (stamper) PD_CSTAMP(25);
        }
        // ---- E1: GELU'd hidden slice -> swizzled A tiles of G2 ----
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            ptx::mbar_wait(&acc_full[c], 0);
            ptx::tc_fence_after();
            if (stamper) PD_CSTAMP(2 + 2 * c);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {          // phase h: hidden columns [128 h, 128 h + 128) of the 256-wide sub-chunk
                const int col = h * 128 + half * 64;   // this warp's 64 columns = k-block (4 c + 2 h + half) of the slice
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_32x32(t_lane + c * 256 + col, v0);
                ptx::tmem_ld_32x32(t_lane + c * 256 + col + 32, v1);
                ptx::tmem_ld_wait();
                const uint32_t my_row = ptx::smem_u32(sMid + (c * 4 + h * 2 + half) * kTileA + (q * 32 + lane) * 128);
                const float* bias = p.b1 + j * kHs + c * 256 + col;
#pragma unroll
                for (int cell = 0; cell < 8; ++cell) {
                    const float4 bv0 = __ldg(reinterpret_cast<const float4*>(bias + cell * 8));
                    const float4 bv1 = __ldg(reinterpret_cast<const float4*>(bias + cell * 8 + 4));
                    const float bb[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
                    uint32_t pk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t r0 = cell < 4 ? v0[cell * 8 + 2 * k] : v1[(cell - 4) * 8 + 2 * k];
                        const uint32_t r1 = cell < 4 ? v0[cell * 8 + 2 * k + 1] : v1[(cell - 4) * 8 + 2 * k + 1];
                        pk[k] = gelu_pair_bf16(__uint_as_float(r0) + bb[2 * k], __uint_as_float(r1) + bb[2 * k + 1]);
                    }
                    ptx::st_shared_v4(my_row + ((static_cast<uint32_t>(cell) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
                }
            }
            ptx::fence_proxy_async();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&mid_full[c]);
            if (stamper) PD_CSTAMP(3 + 2 * c);
        }
    }

    // ---- R: reduce-scatter of the four partials through the L2-resident workspace ----
    const int e = warp - 2, q = warp & 3, half = e >> 2;
    const int r = q * 32 + lane;            // row inside the tile (epilogue threads)
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int col0 = j * kOs + half * 64;   // the thread's 64 finished columns
    const int row0 = row_tile + q * 32;     // the warp's 32 rows
    uint8_t* slabF = smem + kOffSlabF + (warp >= 2 ? e : 0) * (2 * 4096);
    uint8_t* slabB = smem + kOffSlabB + (warp >= 2 ? e : 0) * 4096;
    uint64_t* my_bar = res_bar + 2 * (warp >= 2 ? e : 0);
    float4* ws4 = reinterpret_cast<float4*>(p.ws) + (size_t)tile * (kWsFloatsPerTile / 4);
    if (warp >= 2) {
        if (!PROJ) ptx::cluster_wait();   // peers have initialised their barriers (PROJ: waited for in E0)
        // slices for the owners of output columns [0, 256) leave while G2 still works on columns [256, 512)
#pragma unroll 1
        for (int d = 0; d < kCl; ++d) {                               // destination CTA: owner of output columns [128 d, +128)
            if (d == 0) { ptx::mbar_wait(&acc2_full[0], 0); ptx::tc_fence_after(); }
            if (d == 2) {
                ptx::mbar_wait(&acc2_full[1], 0);   // this CTA's partial is complete; ring and hidden slice are dead
                ptx::tc_fence_after();
                if (stamper) PD_CSTAMP(6);
                if (lane == 0) {   // residual rows of the warp's two 32-column groups (rows past M arrive as zeros)
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        ptx::mbar_arrive_expect_tx(&my_bar[g], 4096);
                        ptx::tma_load_3d(slabF + g * 4096, &tmap_x, &my_bar[g], col0 + g * 32, row0, 0);
                    }
                }
            }
            if (d == j) continue;
            float4* dst = ws4 + (size_t)((d * 3 + recv_slot(j, d)) * 32 + half * 16) * 128 + r;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(t_lane + (uint32_t)(d * kOs + half * 64 + g * 32), v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[(g * 8 + i) * 128] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
        }
        if (stamper) PD_CSTAMP(8);
        // hand-over: the CTA barrier orders the eight warps' stores before one thread, whose fence + releasing arrivals on the
        // three peers' barriers publish them (cumulativity); each CTA then waits for ITS three senders only - no cluster-wide
        // barrier (the full barrier.cluster here cost 4.4 k cycles and held everybody for the slowest CTA's last slice)
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        if (threadIdx.x == 64) {
            __threadfence();
#pragma unroll
            for (int dd = 1; dd < kCl; ++dd)
                ptx::mbar_arrive_remote(ptx::mapa(ptx::smem_u32(slices_ready), (uint32_t)((j + dd) & (kCl - 1))));
        }
        ptx::mbar_wait_cluster(slices_ready, 0);
    }
    if (warp < 2) ptx::cluster_wait();      // second half of the split-phase barrier of the prologue
    if (stamper) PD_CSTAMP(9);
    float s1 = 0.f, s2 = 0.f;
    const bool row_ok = row_tile + r < p.M;
    if (warp >= 2) {
        const float4* src = ws4 + (size_t)(j * 3 * 32 + half * 16) * 128 + r;   // + (slot * 32 + cell) * 128
        // four batches of four 16-byte cells (batch b: column group g = b >> 1, cells 4 (b & 1) + [0, 4)); the loads of
        // batch b + 1 are in flight while batch b is combined (24 x 16 bytes per thread against ~1 000 cycles of L2 latency)
        float4 part[2][3][4];
        auto fetch = [&](int b, float4 (&dstp)[3][4]) {
#pragma unroll
            for (int sl = 0; sl < 3; ++sl)
#pragma unroll
                for (int i = 0; i < 4; ++i) dstp[sl][i] = __ldcg(src + (size_t)(sl * 32 + b * 4 + i) * 128);
        };
        fetch(0, part[0]);
        uint32_t v[32];
        float gs[8], gq[8];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int g = b >> 1, hb = b & 1;
            if (b + 1 < 4) fetch(b + 1, part[(b + 1) & 1]);
            if (hb == 0) {
                ptx::tmem_ld_32x32(t_lane + (uint32_t)(col0 + g * 32), v);
                ptx::tmem_ld_wait();
                ptx::mbar_wait(&my_bar[g], PROJ ? 1 : 0);
            }
            uint8_t* my_row = slabF + g * 4096 + lane * 128;
            float4 resv[4];   // loads before the stores of the batch (generic addresses: a store in between serialises them)
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
                resv[ii] = *reinterpret_cast<const float4*>(my_row + ((static_cast<uint32_t>(hb * 4 + ii) ^ sw) << 4));
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = hb * 4 + ii;
                const float4 own = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                               __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                const float4 (&pt)[3][4] = part[b & 1];
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int s = 0; s < kCl; ++s) {   // rank order, whoever computes: deterministic and batch-invariant
                    // slot of sender s at this destination: s < j -> s, s > j -> s - 1 (j is uniform per CTA)
                    float4 t;
                    if (s == 0) t = j == 0 ? own : pt[0][ii];
                    else if (s == 1) t = j == 1 ? own : (j < 1 ? pt[0][ii] : pt[1][ii]);
                    else if (s == 2) t = j == 2 ? own : (j < 2 ? pt[1][ii] : pt[2][ii]);
                    else t = j == 3 ? own : pt[2][ii];
                    if (s == 0) acc = t;
                    else { acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
                }
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + g * 32 + 4 * i));
                float4* cell = reinterpret_cast<float4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4));
                const float4 res = resv[ii];
                acc.x = (acc.x + bb.x) + res.x; acc.y = (acc.y + bb.y) + res.y;
                acc.z = (acc.z + bb.z) + res.z; acc.w = (acc.w + bb.w) + res.w;
                *cell = acc;
                gs[i] = (acc.x + acc.y) + (acc.z + acc.w);
                gq[i] = (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
                s1 += gs[i];
                s2 += gq[i];
            }
            if (hb == 1) {
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_3d(&tmap_x, slabF + g * 4096, col0 + g * 32, row0, 0);   // rows past M are clipped
                    ptx::bulk_commit();
                }
                if constexpr (GN) {   // statistics for the GroupNorm that reads this output (the warp's 32 rows: one sample)
                    const int shift = p.gn_cpg == 8 ? 3 : (p.gn_cpg == 16 ? 4 : 5);
                    const int gsample = row0 / p.gn_rows;
                    if (row0 < p.M)
                        gn_chunk_accumulate(gs, gq, row_ok, p.gn_cpg,
                                            p.gn_sums + ((size_t)gsample * p.gn_groups + ((col0 + g * 32) >> shift)) * 2, lane);
                }
            }
        }
    }
    if (stamper) PD_CSTAMP(10);
    if (p.ln_gamma != nullptr) {
        // ---- LN: row statistics over all 512 columns = 2 threads x 4 CTAs ----
        float2* ln_x = reinterpret_cast<float2*>(smem + kOffLnX);
        float2* ln_peer = reinterpret_cast<float2*>(smem + kOffLnPeer);
        if (warp >= 2) {
            ln_x[half * 128 + r] = make_float2(s1, s2);
            asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
            const float2 o = ln_x[(half ^ 1) * 128 + r];
            // both halves form the same CTA total (half 0's value first, fixed order)
            const float t1 = half == 0 ? s1 + o.x : o.x + s1, t2 = half == 0 ? s2 + o.y : o.y + s2;
            // st.async onto the peers' armed barrier instead of DSMEM stores + a full cluster barrier (2.4 k cycles): every CTA
            // waits for its three peers' sums before it goes on, so nobody exits while a peer still writes into its memory
            if (half == 0) {
#pragma unroll
                for (int dd = 1; dd < kCl; ++dd) {
                    const uint32_t d = (uint32_t)((j + dd) & (kCl - 1));
                    ptx::st_async_cluster_f32x2(ptx::mapa(ptx::smem_u32(&ln_peer[j * 128 + r]), d), t1, t2,
                                                ptx::mapa(ptx::smem_u32(stat2_bar), d));
                }
            }
            ptx::mbar_wait_cluster(stat2_bar, 0);
            if (stamper) PD_CSTAMP(11);
            float tot1 = 0.f, tot2 = 0.f;
#pragma unroll
            for (int sdr = 0; sdr < kCl; ++sdr) {   // rank order, the CTA's own total in its place
                float2 t = ln_peer[sdr * 128 + r];
                if (sdr == j) t = make_float2(t1, t2);
                tot1 = sdr == 0 ? t.x : tot1 + t.x;
                tot2 = sdr == 0 ? t.y : tot2 + t.y;
            }
            const float mean = tot1 * (1.0f / kC);
            const float var = fmaxf(tot2 * (1.0f / kC) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + p.ln_eps);
            uint8_t* brow = slabB + lane * 128;
            float4 av[16];   // loads before the stores (generic addresses)
#pragma unroll
            for (int i = 0; i < 16; ++i)
                av[i] = *reinterpret_cast<const float4*>(slabF + (i >> 3) * 4096 + lane * 128 + ((static_cast<uint32_t>(i & 7) ^ sw) << 4));
#pragma unroll
            for (int i = 0; i < 8; ++i) {   // 8 values (two cells of the fp32 slabs) -> one 16-byte cell of the bf16 slab
                const float4 a0 = av[2 * i];
                const float4 a1 = av[2 * i + 1];
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col0 + 8 * i));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col0 + 8 * i + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col0 + 8 * i));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col0 + 8 * i + 4));
                uint4 pk;
                pk.x = pack_bf16x2(fmaf((a0.x - mean) * rstd, g0.x, b0.x), fmaf((a0.y - mean) * rstd, g0.y, b0.y));
                pk.y = pack_bf16x2(fmaf((a0.z - mean) * rstd, g0.z, b0.z), fmaf((a0.w - mean) * rstd, g0.w, b0.w));
                pk.z = pack_bf16x2(fmaf((a1.x - mean) * rstd, g1.x, b1.x), fmaf((a1.y - mean) * rstd, g1.y, b1.y));
                pk.w = pack_bf16x2(fmaf((a1.z - mean) * rstd, g1.z, b1.z), fmaf((a1.w - mean) * rstd, g1.w, b1.w));
                *reinterpret_cast<uint4*>(brow + ((static_cast<uint32_t>(i) ^ sw) << 4)) = pk;
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                ptx::tma_store_3d(&tmap_ln, slabB, col0, row0, 0);
                ptx::bulk_commit();
            }
        }
    }
    if (warp >= 2 && lane == 0) ptx::bulk_wait_read<0>();   // the slabs must outlive the bulk stores' reads
    // nobody exits while a peer can still write into its shared memory: every remote access (row sums, barrier arrivals) is
    // one the receiving CTA waits for before it goes on
    if (stamper) PD_CSTAMP(12);
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
#undef PD_CSTAMP
}

struct FfnClusterOpImpl {
    CUtensorMap tmap_a, tmap_w1, tmap_w2, tmap_x, tmap_ln, tmap_att, tmap_wp, tmap_ln1;
    FfnClParams p;
    int tiles;
    int proj;
    WRange own_w;
};
static_assert(sizeof(FfnClusterOpImpl) <= sizeof(FfnClusterOp), "FfnClusterOp storage too small");

}  // namespace

size_t ffn_cluster_workspace_bytes(int M) { return (size_t)ceil_div(M, 128) * kWsFloatsPerTile * sizeof(float); }

int ffn_cluster_make(FfnClusterOp* op_, const bf16* ln_in, int M, const bf16* w1, const float* b1, const bf16* w2,
                     const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta, bf16* ln_out, float ln_eps,
                     float* workspace, const FfnProjArgs* proj) {
    PD_TRY(gemm_init());
    FfnClusterOpImpl* op = reinterpret_cast<FfnClusterOpImpl*>(op_);
    PD_CHECK(ln_in && w1 && b1 && w2 && b2 && x_inout && workspace && M >= 1, PD_ERR_ARG, "ffn_cluster: null argument");
    PD_CHECK((ln_gamma != nullptr) == (ln_out != nullptr) && (ln_gamma != nullptr) == (ln_beta != nullptr), PD_ERR_ARG,
             "ffn_cluster: LayerNorm output needs gamma, beta and the output tensor");
    PD_CHECK(!proj || (proj->att && proj->wp && proj->bp && proj->ln1_gamma && proj->ln1_beta), PD_ERR_ARG,
             "ffn_cluster: incomplete projection arguments");
    static bool attr_set = false;
    if (!attr_set) {
        PD_CUDA(cudaFuncSetAttribute(ffn_cluster_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        PD_CUDA(cudaFuncSetAttribute(ffn_cluster_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        PD_CUDA(cudaFuncSetAttribute(ffn_cluster_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        PD_CUDA(cudaFuncSetAttribute(ffn_cluster_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set = true;
    }
    const uint64_t dims_a[3] = {kC, (uint64_t)M, 1}, st_a[2] = {kC * 2, (uint64_t)kC * 2 * M};
    const uint32_t box_a[3] = {64, 128, 1};
    PD_TRY(tmap_encode_sw128(&op->tmap_a, true, 3, ln_in, dims_a, st_a, box_a));
    const uint64_t d1[2] = {kC, kHid}, s1[1] = {kC * 2};       // W1 [2048][512] (K-major)
    const uint64_t d2[2] = {kHid, kC}, s2[1] = {kHid * 2};     // W2 [512][2048]
    const uint32_t box[2] = {64, 256};
    PD_TRY(tmap_encode_sw128(&op->tmap_w1, true, 2, w1, d1, s1, box));
    PD_TRY(tmap_encode_sw128(&op->tmap_w2, true, 2, w2, d2, s2, box));
    // epilogue slabs: 32 rows x 128 bytes (32 fp32 columns of x / 64 bf16 columns of the LayerNorm output)
    const uint64_t dims_x[3] = {kC, (uint64_t)M, 1}, st_x[2] = {kC * 4, (uint64_t)kC * 4 * M};
    const uint32_t box_x[3] = {32, 32, 1};
    PD_TRY(tmap_encode_sw128(&op->tmap_x, false, 3, x_inout, dims_x, st_x, box_x));
    const uint32_t box_ln[3] = {64, 32, 1};
    PD_TRY(tmap_encode_sw128(&op->tmap_ln, true, 3, ln_out ? ln_out : ln_in, dims_a, st_a, box_ln));
    op->proj = proj != nullptr;
    op->tmap_att = op->tmap_a; op->tmap_wp = op->tmap_w1; op->tmap_ln1 = op->tmap_ln;
    op->p.bp = op->p.ln1_gamma = op->p.ln1_beta = nullptr;
    if (proj) {   // x1 = x + att Wp^T + bp; LayerNorm(x1; ln1) -> ln_in (the tensor G1 then loads as its A operand)
        PD_TRY(tmap_encode_sw128(&op->tmap_att, true, 3, proj->att, dims_a, st_a, box_a));
        const uint64_t dp[2] = {kC, kC}, sp[1] = {kC * 2};     // Wp [512][512]
        const uint32_t boxp[2] = {64, (uint32_t)kOs};
        PD_TRY(tmap_encode_sw128(&op->tmap_wp, true, 2, proj->wp, dp, sp, boxp));
        PD_TRY(tmap_encode_sw128(&op->tmap_ln1, true, 3, ln_in, dims_a, st_a, box_ln));
        op->p.bp = proj->bp; op->p.ln1_gamma = proj->ln1_gamma; op->p.ln1_beta = proj->ln1_beta;
    }
    op->p.b1 = b1; op->p.b2 = b2; op->p.ln_gamma = ln_gamma; op->p.ln_beta = ln_beta;
    op->p.ws = workspace; op->p.ln_eps = ln_eps; op->p.M = M;
    op->p.gn_sums = nullptr; op->p.gn_cpg = op->p.gn_groups = op->p.gn_rows = 0;
    op->p.dbg = nullptr;
    op->tiles = ceil_div(M, 128);
    op->own_w = WRange{};
    op->own_w.p[0] = reinterpret_cast<const uint8_t*>(w1); op->own_w.n[0] = (uint32_t)((size_t)kHid * kC * 2);
    op->own_w.p[1] = reinterpret_cast<const uint8_t*>(w2); op->own_w.n[1] = (uint32_t)((size_t)kHid * kC * 2);
    if (proj) { op->own_w.p[2] = reinterpret_cast<const uint8_t*>(proj->wp); op->own_w.n[2] = (uint32_t)((size_t)kC * kC * 2); }
    return PD_OK;
}

int ffn_cluster_set_gn(FfnClusterOp* op_, double* gn_sums, int groups, int rows) {
    FfnClusterOpImpl* op = reinterpret_cast<FfnClusterOpImpl*>(op_);
    PD_CHECK(gn_sums && gemm_gn_fusable(kC, groups, rows), PD_ERR_SHAPE, "ffn_cluster: GroupNorm statistics not fusable");
    op->p.gn_sums = gn_sums; op->p.gn_groups = groups; op->p.gn_rows = rows; op->p.gn_cpg = kC / groups;
    return PD_OK;
}

void ffn_cluster_set_dbg(FfnClusterOp* op_, unsigned long long* stamps) {
    reinterpret_cast<FfnClusterOpImpl*>(op_)->p.dbg = stamps;
}

WRange ffn_cluster_weights(const FfnClusterOp& op_) { return reinterpret_cast<const FfnClusterOpImpl&>(op_).own_w; }

int ffn_cluster_launch(const FfnClusterOp& op_, cudaStream_t st) {
    const FfnClusterOpImpl& op = reinterpret_cast<const FfnClusterOpImpl&>(op_);
#define PD_FCL_LAUNCH(PROJ, GN)                                                                                           \
    PD_CUDA(launch_pdl(ffn_cluster_kernel<PROJ, GN>, dim3(op.tiles * kCl), dim3(kThreads), (size_t)kSmem, st, dim3(kCl, 1, 1), \
                       op.tmap_a, op.tmap_w1, op.tmap_w2, op.tmap_x, op.tmap_ln, op.tmap_att, op.tmap_wp, op.tmap_ln1, op.p))
    if (op.proj) {
        if (op.p.gn_sums) PD_FCL_LAUNCH(true, true);
        else PD_FCL_LAUNCH(true, false);
    } else {
        if (op.p.gn_sums) PD_FCL_LAUNCH(false, true);
        else PD_FCL_LAUNCH(false, false);
    }
#undef PD_FCL_LAUNCH
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
