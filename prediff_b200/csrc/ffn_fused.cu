// Fused PositionwiseFFN for width-256 levels: x <- x + W2 GELU(W1 ln + b1) + b2 (and, optionally, the LayerNorm that
// follows) in ONE kernel per 128-row tile - the 128 x 1024 hidden activation never leaves the SM.
// Reference: PositionwiseFFN.forward (src/prediff/models/cuboid_transformer/cuboid_transformer.py:182-208), pre-norm
// input `ln` produced by the preceding epilogue. Replaces the FFN-1 GEMM (bf16 `mid` round trip through HBM/L2:
// 27 MB written + read per call at batch 4) and the FFN-2 GEMM.
//
//   warp 0    : TMA producer - the A operand of GEMM-1 (128 x 256 bf16 = four 16 KB k-block tiles) is RESIDENT in shared
//               memory for all four hidden chunks; a 3-stage ring of 32 KB stages carries nothing but weight tiles (W1 tile
//               = 256 hidden rows x 64, W2 tile = 256 output rows x 64, PROJ: Wp tiles first), so every ring load is
//               independent of the preceding kernel and of this kernel's own epilogues. (Until round 2 the A k-blocks
//               travelled with the W1 tiles in 48 KB stages and were re-read for every chunk; one trip of a stage round
//               the ring - commit, producer wake-up, two TMA issues at ~270 cycles each, ~800 cycles of transfer, consumer
//               wake-up - took ~1 500 cycles, so three stages sustained one k-block per ~730 cycles against the tensor
//               core's 512, with ONE CTA on the chip as with 148: tools/micro/tma_latency.cu, tools/ffn_phases.py.)
//   warp 1    : one lane issues tcgen05.mma 128 x 256 x 16 (an N = 128 MMA costs the same ~128 cycles, measured):
//                 G1(c):  acc1  = ln . W1[c]^T                       c = 0..3, 256 hidden columns per chunk
//                 G2h(c): acc2 += gelu_c[:, 128h : 128h+128] . W2[:, ...]^T   (two K halves, each as soon as it is ready)
//               order G1(0) | G1(c+1) G2h0(c) G2h1(c) | ... : GEMM-1 of the next chunk runs under the first half of this
//               chunk's GELU epilogue (acc1 is copied to registers at once), the first half of GEMM-2 under the second
//   warps 2-9 : E1(c), two phases of 128 hidden columns: acc1 (TMEM) -> + b1 -> GELU -> bf16 -> 128B-swizzled K-major
//               smem tiles = the A operand of G2(c); acc1 is handed back right after the second phase's TMEM loads;
//               final: acc2 -> + b2 + residual (TMA-loaded) -> x (TMA store) -> fused LayerNorm -> bf16 (TMA store)
// TMEM: acc1 (256 columns) + acc2 (256 columns) = 512 columns.
#include "gemm.cuh"
#include "ops.cuh"
#include "ptx.cuh"

namespace pd {
namespace {

constexpr int kC = 256;            // model width (K of GEMM-1, N of GEMM-2)
constexpr int kHid = 1024;         // hidden width
constexpr int kChunk = 256;        // hidden columns per chunk
constexpr int kNumChunks = kHid / kChunk;
#ifndef PD_FFN_EPI_WARPS
#define PD_FFN_EPI_WARPS 16
#endif
constexpr int kEpiWarps = PD_FFN_EPI_WARPS;   // 8 or 16: kParts warps share a TMEM lane quarter and split its columns
constexpr int kParts = kEpiWarps / 4;
constexpr int kCh = 8 / kParts;               // 32-column chunks of the 256-wide row per thread (prologue / final epilogue)
constexpr int kIs = 16 / kEpiWarps;           // 4 KB slabs per warp that fit into the 64 KB mid region
constexpr int kThreads = 64 + 32 * kEpiWarps;
static_assert(kEpiWarps == 8 || kEpiWarps == 16, "epilogue warps: two or four per TMEM lane quarter");
constexpr int kTileA = 128 * 64 * 2;         // 16 KB: 128 rows x 64 bf16
constexpr int kTileW = 256 * 64 * 2;         // 32 KB: 256 rows x 64 bf16
constexpr int kABytes = 4 * kTileA;          // 64 KB: the resident A operand (four k-blocks)
constexpr int kStages = 3;
constexpr int kRingBytes = kStages * kTileW; // 96 KB
constexpr int kMidBytes = 4 * kTileA;        // 64 KB: four k-blocks of the GELU'd chunk
constexpr int kPipeBytes = kABytes + kRingBytes + kMidBytes;   // 224 KB
constexpr int kBarBytes = 512;
// the dynamic shared memory window is declared 1024-byte aligned (checked at kernel start): no alignment slack
constexpr int kSmem = kPipeBytes + kBarBytes;
static_assert(kSmem <= 232448, "over the 227 KB per-CTA limit");
static_assert(kABytes + kRingBytes >= kEpiWarps * kCh * 4096 + 4 * kParts * 32 * 8,
              "fp32 epilogue slabs + the row-statistics exchange alias A + the ring");
static_assert(kMidBytes >= kEpiWarps * (kCh / 2) * 4096 && kMidBytes >= kEpiWarps * kIs * 4096, "bf16 / residual slabs alias mid");

struct FfnParams {
    const float* b1;
    const float* b2;
    const float* ln_gamma;   // null: no fused LayerNorm output
    const float* ln_beta;
    float ln_eps;
    int M;
    unsigned long long* dbg;   // optional clock64() stamps of CTA 0 (tools/ffn_phases.py): see PD_FSTAMP sites
    // PROJ variant (attention output projection fused in front): x1 = x + att Wp^T + bp never leaves the SM
    const float* bp;         // proj bias
    const float* ln1_gamma;  // the FFN's own pre-norm, applied to x1 inside the kernel
    const float* ln1_beta;
    // fused GroupNorm statistics of the new x rows (gemm.cuh GemmEpilogue::gn_sums): the resblock that follows a stack
    double* gn_sums;
    int gn_cpg, gn_groups, gn_rows;
    WRange pf;               // weights of the next GEMM of the plan, requested into L2 at kernel start (common.cuh)
};

// PROJ = false: the A operand of GEMM-1 is the (already normalised) tensor behind tmap_a, loaded once; acc2 starts at zero
//               and the residual is added in the final epilogue.
// PROJ = true : the kernel first builds x1 = (x + bp) + att . Wp^T in the acc2 columns of TMEM (GEMM-0 reads `att` from the
//               resident A tiles; meanwhile the epilogue warps park x + bp in the idle acc1 columns through tcgen05.st and
//               add the two afterwards), normalises it (the FFN's pre-norm) straight into the resident A tiles as swizzled
//               bf16 - no round trip through global memory - and GEMM-2 keeps accumulating onto x1, so no residual is ever
//               re-loaded.
template <bool PROJ, bool GN>   // GN: also accumulate GroupNorm statistics of the new x rows (FfnParams::gn_sums)
__global__ void __launch_bounds__(kThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_x,
                 const __grid_constant__ CUtensorMap tmap_ln, const __grid_constant__ CUtensorMap tmap_att,
                 const __grid_constant__ CUtensorMap tmap_wp,
                 const __grid_constant__ FfnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sRing = smem + kABytes;
    uint8_t* sMid = smem + kABytes + kRingBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPipeBytes);
    uint64_t* w_full = bars;                 // [kStages]
    uint64_t* w_empty = w_full + kStages;    // [kStages]
    uint64_t* acc1_full = w_empty + kStages; // [1]
    uint64_t* acc1_empty = acc1_full + 1;    // [1]
    uint64_t* mid_full = acc1_empty + 1;     // [2] (K halves)
    uint64_t* mid_empty = mid_full + 2;      // [2]
    uint64_t* acc2_full = mid_empty + 2;     // [1] (PROJ: phase 0 = x1 complete, phase 1 = FFN complete)
    uint64_t* acc2_init = acc2_full + 1;     // [1] unused
    uint64_t* a_full = acc2_init + 1;        // [1] the A tiles have landed (PROJ: `att`; else the pre-norm input)
    uint64_t* a_ready = a_full + 1;          // [1] PROJ: LayerNorm(x1) written into the A tiles by the epilogue warps
    uint64_t* res_bar = a_ready + 1;         // [kEpiWarps][kCh]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kCh * kEpiWarps);
    static_assert((15 + kCh * kEpiWarps) * 8 + 4 <= kBarBytes, "barrier block too small");
    if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();   // the swizzled tiles need the declared alignment

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int row_tile = blockIdx.x * 128;
    unsigned long long* dbg = (p.dbg && blockIdx.x == 0) ? p.dbg : nullptr;
#define PD_FSTAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
    if (threadIdx.x == 0) PD_FSTAMP(0);
    if (threadIdx.x == 32) prefetch_l2_share(p.pf, blockIdx.x, gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&w_full[s], 1);
            ptx::mbar_init(&w_empty[s], 1);
        }
        ptx::mbar_init(acc1_full, 1);
        ptx::mbar_init(acc1_empty, kEpiWarps);
        for (int h = 0; h < 2; ++h) {
            ptx::mbar_init(&mid_full[h], kEpiWarps);
            ptx::mbar_init(&mid_empty[h], 1);
        }
        ptx::mbar_init(acc2_full, 1);
        ptx::mbar_init(acc2_init, kEpiWarps);
        ptx::mbar_init(a_full, 1);
        ptx::mbar_init(a_ready, kEpiWarps);
        for (int i = 0; i < kCh * kEpiWarps; ++i) ptx::mbar_init(&res_bar[i], 1);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_w1);
        ptx::prefetch_tmap(&tmap_w2);
        ptx::prefetch_tmap(&tmap_x);
        if (p.ln_gamma) ptx::prefetch_tmap(&tmap_ln);
        if (PROJ) {
            ptx::prefetch_tmap(&tmap_att);
            ptx::prefetch_tmap(&tmap_wp);
        }
    }
    // The first kStages weight tiles (Wp when PROJ, else W1 chunk 0) do not depend on the preceding kernel: requested
    // before the dependency wait (common.cuh); the activation tiles follow after it.
    if (threadIdx.x == 0) {
        for (int kb = 0; kb < kStages; ++kb) {
            ptx::mbar_arrive_expect_tx(&w_full[kb], kTileW);
            ptx::tma_load_2d(sRing + kb * kTileW, PROJ ? &tmap_wp : &tmap_w1, &w_full[kb], kb * 64, 0);
        }
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    if (warp >= 2) {
        // the bias / LayerNorm vectors (10 KB of parameters, independent of the preceding kernel) into L1 now: the epilogues
        // read them through the read-only path in the middle of dependent chains (the normalising pass of E0 took 3 000
        // cycles for 64 values per thread, most of it L2 latency on these vectors)
        const int t = threadIdx.x - 64;   // one 128-byte line per thread
        if (t < 32) ptx::prefetch_l1(p.b1 + t * 32);
        else if (t < 40) ptx::prefetch_l1(p.b2 + (t - 32) * 32);
        else if (t < 48) { if (p.ln_gamma) ptx::prefetch_l1(p.ln_gamma + (t - 40) * 32); }
        else if (t < 56) { if (p.ln_gamma) ptx::prefetch_l1(p.ln_beta + (t - 48) * 32); }
        else if (PROJ && t < 64) ptx::prefetch_l1(p.bp + (t - 56) * 32);
        else if (PROJ && t < 72) ptx::prefetch_l1(p.ln1_gamma + (t - 64) * 32);
        else if (PROJ && t < 80) ptx::prefetch_l1(p.ln1_beta + (t - 72) * 32);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // dependents are released only now: a dependent CTA that became co-resident before this CTA owned its TMEM columns
    // could take them and then sit in its own dependency wait forever
    grid_dep_launch();
    grid_dep_wait();

    if (warp == 0) {
        if (lane == 0) {
            // the resident A operand: `att` for GEMM-0 (PROJ; the epilogue warps overwrite it with LayerNorm(x1) afterwards)
            // or the pre-norm input of GEMM-1
            ptx::mbar_arrive_expect_tx(a_full, kABytes);
            for (int kb = 0; kb < 4; ++kb)
                ptx::tma_load_3d(sA + kb * kTileA, PROJ ? &tmap_att : &tmap_a, a_full, kb * 64, row_tile, 0);
            int it = 0;
            auto wload = [&](const CUtensorMap* m, int c0, int c1) {   // the next weight tile of the ring
                const int s = it % kStages;
                if (it >= kStages) {   // the first kStages tiles were requested in the prologue
                    ptx::mbar_wait(&w_empty[s], ((it / kStages) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&w_full[s], kTileW);
                    ptx::tma_load_2d(sRing + s * kTileW, m, &w_full[s], c0, c1);
                }
                ++it;
            };
            if (PROJ)
                for (int kb = 0; kb < 4; ++kb) wload(&tmap_wp, kb * 64, 0);
            for (int kb = 0; kb < 4; ++kb) wload(&tmap_w1, kb * 64, 0);
            for (int c = 0; c < kNumChunks; ++c) {
                if (c + 1 < kNumChunks)
                    for (int kb = 0; kb < 4; ++kb) wload(&tmap_w1, kb * 64, (c + 1) * kChunk);
                for (int kb = 0; kb < 2; ++kb) wload(&tmap_w2, c * kChunk + kb * 64, 0);
                for (int kb = 0; kb < 2; ++kb) wload(&tmap_w2, c * kChunk + 128 + kb * 64, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 256);
            int it = 0;
            auto g1 = [&](int c) {
                ptx::mbar_wait(acc1_empty, (c & 1) ^ 1);              // E1(c - 1) has read the accumulator out
                ptx::tc_fence_after();
                for (int kb = 0; kb < 4; ++kb, ++it) {
                    const int s = it % kStages;
                    ptx::mbar_wait(&w_full[s], (it / kStages) & 1);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(sA + kb * kTileA);
                    const uint32_t b_addr = ptx::smem_u32(sRing + s * kTileW);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16(tmem_base, ptx::make_smem_desc_sw128(a_addr + k * 32),
                                      ptx::make_smem_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&w_empty[s]);
                }
                ptx::umma_commit(acc1_full);
            };
            auto g2h = [&](int c, int h) {
                ptx::mbar_wait(&mid_full[h], c & 1);                  // E1(c) phase h has written its two k-blocks
                ptx::tc_fence_after();
                for (int kb = 0; kb < 2; ++kb, ++it) {
                    const int s = it % kStages;
                    ptx::mbar_wait(&w_full[s], (it / kStages) & 1);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(sMid + (h * 2 + kb) * kTileA);
                    const uint32_t b_addr = ptx::smem_u32(sRing + s * kTileW);
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // PROJ: acc2 already holds x1, always accumulate
                        ptx::umma_f16(tmem_base + 256, ptx::make_smem_desc_sw128(a_addr + k * 32),
                                      ptx::make_smem_desc_sw128(b_addr + k * 32), idesc,
                                      (PROJ || (c | h | kb | k) != 0) ? 1u : 0u);
                    ptx::umma_commit(&w_empty[s]);
                }
                ptx::umma_commit(&mid_empty[h]);
            };
            PD_FSTAMP(1);
            ptx::mbar_wait(a_full, 0);
            ptx::tc_fence_after();
            PD_FSTAMP(11);                // A tiles landed
            if (PROJ) {   // GEMM-0: acc2 = att . Wp^T (the epilogue warps park x + bp in the idle acc1 columns meanwhile)
                for (int kb = 0; kb < 4; ++kb, ++it) {
                    const int s = it % kStages;
                    ptx::mbar_wait(&w_full[s], (it / kStages) & 1);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(sA + kb * kTileA);
                    const uint32_t b_addr = ptx::smem_u32(sRing + s * kTileW);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16(tmem_base + 256, ptx::make_smem_desc_sw128(a_addr + k * 32),
                                      ptx::make_smem_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&w_empty[s]);
                }
                ptx::umma_commit(acc2_full);   // phase 0: att . Wp^T complete
                PD_FSTAMP(20);            // GEMM-0 issued
                // the epilogue warps have put x1 into acc2, replaced `att` by LayerNorm(x1) in the A tiles and read acc1 out
                ptx::mbar_wait(a_ready, 0);
                ptx::tc_fence_after();
            }
            g1(0);
            PD_FSTAMP(2);
            for (int c = 0; c < kNumChunks; ++c) {
                // GEMM-1 of the next chunk first: E1(c) hands acc1 back as soon as it has copied its columns to registers,
                // so G1(c + 1) runs under phase 0 of E1(c), the two K halves of GEMM-2 under phase 1 / phase 0 of E1(c + 1)
                if (c + 1 < kNumChunks) g1(c + 1);
                g2h(c, 0);
                g2h(c, 1);
                PD_FSTAMP(3 + c);          // G2(c) issued
            }
            ptx::umma_commit(acc2_full);
        }
    } else {
        const int e = warp - 2;
        const int q = warp & 3;          // TMEM lane quarter (fixed by the hardware: warp index modulo 4)
        const int part = e >> 2;         // which of the kParts column ranges of the quarter's rows this warp owns
        const int et = threadIdx.x - 64;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        const int row0 = row_tile + q * 32;
        const int c_begin = part * kCh;            // this warp's kCh 32-column chunks of the 256-wide row
        uint64_t* my_bar = res_bar + kCh * e;
        uint8_t* islab = sMid + e * (kIs * 4096);
        const int ln_slot = (q * kParts) * 32 + lane;   // + part * 32: the row's partial sums, one per column range
        // row statistics of the 256-wide row from the kParts threads that share it (fixed order: identical in all of them)
        auto row_stats = [&](float2* ln_x, float s1, float s2, float* mean, float* rstd) {
            ln_x[ln_slot + part * 32] = make_float2(s1, s2);
            asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(32 * kParts) : "memory");
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int pp = 0; pp < kParts; ++pp) {
                const float2 o = ln_x[ln_slot + pp * 32];
                t1 = pp == 0 ? o.x : t1 + o.x;
                t2 = pp == 0 ? o.y : t2 + o.y;
            }
            *mean = t1 * (1.0f / kC);
            *rstd = rsqrtf(fmaxf(t2 * (1.0f / kC) - *mean * *mean, 0.f) + p.ln_eps);
        };
        auto load_x_round = [&](int r) {           // PROJ: kIs residual slabs of this warp's rows -> the idle mid region
            for (int j = 0; j < kIs; ++j) {
                ptx::mbar_arrive_expect_tx(&my_bar[kIs * r + j], 4096);
                ptx::tma_load_3d(islab + j * 4096, &tmap_x, &my_bar[kIs * r + j], (c_begin + kIs * r + j) * 32, row0, 0);
            }
        };
        if (PROJ && lane == 0) load_x_round(0);
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        if (PROJ) {
            // ---- x + bp -> the idle acc1 columns (rounds of kIs 4 KB slabs per warp through the idle mid region), while
            //      GEMM-0 runs: nothing of this is on its critical path (round 2 preset acc2 itself, GEMM-0 waited for it) ----
#pragma unroll 1
            for (int r = 0; r < kCh / kIs; ++r) {
                if (lane == 0 && r > 0) load_x_round(r);
#pragma unroll 1
                for (int j = 0; j < kIs; ++j) {
                    const int c = c_begin + kIs * r + j;
                    ptx::mbar_wait(&my_bar[kIs * r + j], 0);
                    const uint8_t* my_row = islab + j * 4096 + lane * 128;
                    uint32_t v[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 xv = *reinterpret_cast<const float4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4));
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bp + c * 32 + 4 * i));
                        v[4 * i] = __float_as_uint(xv.x + bb.x);
                        v[4 * i + 1] = __float_as_uint(xv.y + bb.y);
                        v[4 * i + 2] = __float_as_uint(xv.z + bb.z);
                        v[4 * i + 3] = __float_as_uint(xv.w + bb.w);
                    }
                    ptx::tmem_st_32x32(t_lane + c * 32, v);
                }
                __syncwarp();   // the slabs are consumed before the next round overwrites them
            }
            ptx::tmem_st_wait();
            // the row-statistics exchange below crosses in the first bytes of mid: every warp must be done with its slabs
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            if (et == 0) PD_FSTAMP(7);    // x + bp parked
            // ---- E0: x1 = (x + bp) + att . Wp^T -> acc2 (GEMM-2 accumulates on top); LayerNorm(x1) -> bf16 -> the resident
            //      A tiles (the A operand of GEMM-1) ----
            ptx::mbar_wait(acc2_full, 0);
            ptx::tc_fence_after();
            if (et == 0) PD_FSTAMP(8);    // GEMM-0 complete
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int idx = 0; idx < kCh; ++idx) {
                uint32_t v[32], w[32];
                ptx::tmem_ld_32x32(t_lane + 256 + (c_begin + idx) * 32, v);
                ptx::tmem_ld_32x32(t_lane + (c_begin + idx) * 32, w);
                ptx::tmem_ld_wait();
                float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};   // four independent chains, not one of 32
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float a = __uint_as_float(w[i]) + __uint_as_float(v[i]);
                    v[i] = __float_as_uint(a);
                    p1[i & 3] += a;
                    p2[i & 3] = fmaf(a, a, p2[i & 3]);
                }
                ptx::tmem_st_32x32(t_lane + 256 + (c_begin + idx) * 32, v);
                s1 += (p1[0] + p1[1]) + (p1[2] + p1[3]);
                s2 += (p2[0] + p2[1]) + (p2[2] + p2[3]);
            }
            ptx::tmem_st_wait();   // the normalising pass below reads x1 back; GEMM-2 accumulates onto it after a_ready
            float mean, rstd;
            row_stats(reinterpret_cast<float2*>(sMid), s1, s2, &mean, &rstd);
            if (et == 0) PD_FSTAMP(9);    // x1 in acc2, row statistics known
#pragma unroll 1
            for (int j = 0; j < kCh / 2; ++j) {    // A tile (k-block) c_begin / 2 + j = my chunks 2j, 2j+1 (64 columns)
                const uint32_t arow = ptx::smem_u32(sA + (c_begin / 2 + j) * kTileA + (q * 32 + lane) * 128);
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int colbase = (c_begin + 2 * j + cc) * 32;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_lane + 256 + colbase, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln1_gamma + colbase + 8 * k));
                        const float4 g1v = __ldg(reinterpret_cast<const float4*>(p.ln1_gamma + colbase + 8 * k + 4));
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln1_beta + colbase + 8 * k));
                        const float4 b1v = __ldg(reinterpret_cast<const float4*>(p.ln1_beta + colbase + 8 * k + 4));
                        float a[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) a[t] = __uint_as_float(v[8 * k + t]);
                        ptx::st_shared_v4(arow + ((static_cast<uint32_t>(cc * 4 + k) ^ sw) << 4),
                                          pack_bf16x2(fmaf((a[0] - mean) * rstd, g0.x, b0.x), fmaf((a[1] - mean) * rstd, g0.y, b0.y)),
                                          pack_bf16x2(fmaf((a[2] - mean) * rstd, g0.z, b0.z), fmaf((a[3] - mean) * rstd, g0.w, b0.w)),
                                          pack_bf16x2(fmaf((a[4] - mean) * rstd, g1v.x, b1v.x), fmaf((a[5] - mean) * rstd, g1v.y, b1v.y)),
                                          pack_bf16x2(fmaf((a[6] - mean) * rstd, g1v.z, b1v.z), fmaf((a[7] - mean) * rstd, g1v.w, b1v.w)));
                    }
                }
            }
            // generic-proxy writes -> visible to the tensor core's (async-proxy) reads of the A tiles. The partial sums in
            // mid were read by the quarter's threads before any of them got here (row_stats' barrier precedes its reads,
            // and E1(0) of a warp starts after a_ready, i.e. after every warp's arrival below)
            ptx::tc_fence_before();
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(a_ready);
            if (et == 0) PD_FSTAMP(10);   // LayerNorm(x1) in the A tiles
        }
        // ---- E1: GELU chunks -> swizzled A tiles of GEMM-2 ----
        constexpr int kColsPh = 128 / kParts;      // hidden columns per thread and phase (64 or 32)
        constexpr int kLdPh = kColsPh / 32;
        const int mid_kb = (part * kColsPh) / 64;  // k-block (of the phase's two) and first 16-byte cell the columns fall into
        const int cell0 = ((part * kColsPh) % 64) / 8;
#pragma unroll 1
        for (int c = 0; c < kNumChunks; ++c) {
            ptx::mbar_wait(acc1_full, c & 1);
            ptx::tc_fence_after();
            if (et == 0) PD_FSTAMP(12 + 2 * c);   // E1(c) begins
            // the thread's accumulator columns of both phases leave TMEM at once, so that GEMM-1 of the next chunk can
            // start under phase 0 already (with acc1 handed back only after phase 1's loads, E1(c) phase 0 -> G2h0(c) ->
            // G1(c + 1) -> E1(c + 1) was the critical path: 6 150 cycles per chunk with a 5 000-cycle epilogue)
            uint32_t va[2][kLdPh][32];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int l = 0; l < kLdPh; ++l) ptx::tmem_ld_32x32(t_lane + h * 128 + part * kColsPh + l * 32, va[h][l]);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acc1_empty);
#pragma unroll
            for (int h = 0; h < 2; ++h) {          // phase h: hidden columns [128 h, 128 h + 128) of the chunk
                ptx::mbar_wait(&mid_empty[h], (c & 1) ^ 1);           // G2h(c - 1) no longer reads these k-blocks
                const uint32_t my_row = ptx::smem_u32(sMid + (h * 2 + mid_kb) * kTileA + (q * 32 + lane) * 128);
                const float* bias = p.b1 + c * kChunk + h * 128 + part * kColsPh;   // read-only path: hoistable loads
#pragma unroll
                for (int cell = 0; cell < kColsPh / 8; ++cell) {
                    const float4 bv0 = __ldg(reinterpret_cast<const float4*>(bias + cell * 8));
                    const float4 bv1 = __ldg(reinterpret_cast<const float4*>(bias + cell * 8 + 4));
                    const float bb[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
                    uint32_t pk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t r0 = va[h][cell / 4][(cell % 4) * 8 + 2 * k];
                        const uint32_t r1 = va[h][cell / 4][(cell % 4) * 8 + 2 * k + 1];
                        pk[k] = gelu_pair_bf16(__uint_as_float(r0) + bb[2 * k], __uint_as_float(r1) + bb[2 * k + 1]);
                    }
                    ptx::st_shared_v4(my_row + ((static_cast<uint32_t>(cell0 + cell) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&mid_full[h]);
            }
            if (et == 0) PD_FSTAMP(13 + 2 * c);   // E1(c) done
        }
        // ---- final epilogue: acc2 + b2 + residual -> x (fp32, TMA store), fused LayerNorm -> bf16 ----
        ptx::mbar_wait(acc2_full, PROJ ? 1 : 0);
        ptx::tc_fence_after();
        if (et == 0) PD_FSTAMP(28);               // accumulator 2 complete
        uint8_t* slabs = smem + e * (kCh * 4096);                  // aliases A + the ring (all MMAs have completed)
        if (!PROJ && lane == 0) {                                  // PROJ: the residual is already inside acc2
            for (int idx = 0; idx < kCh; ++idx) {
                ptx::mbar_arrive_expect_tx(&my_bar[idx], 4096);
                ptx::tma_load_3d(slabs + idx * 4096, &tmap_x, &my_bar[idx], (c_begin + idx) * 32, row0, 0);
            }
        }
        float ln_s1 = 0.f, ln_s2 = 0.f;
        const int gsample = GN ? row0 / p.gn_rows : 0;   // GroupNorm sample of my rows
#pragma unroll 1
        for (int idx = 0; idx < kCh; ++idx) {
            const int c = c_begin + idx;
            uint8_t* slab = slabs + idx * 4096;
            uint32_t v[32];
            ptx::tmem_ld_32x32(t_lane + 256 + c * 32, v);
            if (!PROJ) ptx::mbar_wait(&my_bar[idx], 0);
            ptx::tmem_ld_wait();
            uint8_t* my_row = slab + lane * 128;
            // the residual cells first: through generic addresses a store between two loads serialises them
            float4 resv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                resv[i] = PROJ ? make_float4(0.f, 0.f, 0.f, 0.f)
                               : *reinterpret_cast<const float4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4));
            float gs[8], gq[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + c * 32 + 4 * i));
                float4* cell = reinterpret_cast<float4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4));
                const float4 r = resv[i];
                float4 a = make_float4(__uint_as_float(v[4 * i]) + bb.x + r.x, __uint_as_float(v[4 * i + 1]) + bb.y + r.y,
                                       __uint_as_float(v[4 * i + 2]) + bb.z + r.z, __uint_as_float(v[4 * i + 3]) + bb.w + r.w);
                *cell = a;
                gs[i] = (a.x + a.y) + (a.z + a.w);
                gq[i] = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
                ln_s1 += gs[i];
                ln_s2 += gq[i];
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                ptx::tma_store_3d(&tmap_x, slab, c * 32, row0, 0);
                ptx::bulk_commit();
            }
            if constexpr (GN) if (row0 < p.M) {
                const int shift = p.gn_cpg == 8 ? 3 : (p.gn_cpg == 16 ? 4 : 5);
                gn_chunk_accumulate(gs, gq, row0 + lane < p.M, p.gn_cpg,
                                    p.gn_sums + ((size_t)gsample * p.gn_groups + ((c * 32) >> shift)) * 2, lane);
            }
        }
        if (p.ln_gamma) {
            float mean, rstd;   // partial sums cross in the dead tail of the ring (the fp32 slabs end at kEpiWarps * kCh * 4 KB)
            row_stats(reinterpret_cast<float2*>(smem + kEpiWarps * kCh * 4096), ln_s1, ln_s2, &mean, &rstd);
            uint8_t* bslabs = sMid + e * ((kCh / 2) * 4096);       // aliases mid
#pragma unroll
            for (int j = 0; j < kCh / 2; ++j) {
                uint8_t* brow = bslabs + j * 4096 + lane * 128;
                float4 av[2][8];   // loads before the stores (generic addresses)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const uint8_t* frow = slabs + (2 * j + cc) * 4096 + lane * 128;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        av[cc][i] = *reinterpret_cast<const float4*>(frow + ((static_cast<uint32_t>(i) ^ sw) << 4));
                }
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int colbase = (c_begin + 2 * j + cc) * 32;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 a0 = av[cc][2 * k];
                        const float4 a1 = av[cc][2 * k + 1];
                        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + colbase + 8 * k));
                        const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + colbase + 8 * k + 4));
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + colbase + 8 * k));
                        const float4 b1v = __ldg(reinterpret_cast<const float4*>(p.ln_beta + colbase + 8 * k + 4));
                        uint4 pk;
                        pk.x = pack_bf16x2(fmaf((a0.x - mean) * rstd, g0.x, b0.x), fmaf((a0.y - mean) * rstd, g0.y, b0.y));
                        pk.y = pack_bf16x2(fmaf((a0.z - mean) * rstd, g0.z, b0.z), fmaf((a0.w - mean) * rstd, g0.w, b0.w));
                        pk.z = pack_bf16x2(fmaf((a1.x - mean) * rstd, g1.x, b1v.x), fmaf((a1.y - mean) * rstd, g1.y, b1v.y));
                        pk.w = pack_bf16x2(fmaf((a1.z - mean) * rstd, g1.z, b1v.z), fmaf((a1.w - mean) * rstd, g1.w, b1v.w));
                        *reinterpret_cast<uint4*>(brow + ((static_cast<uint32_t>(cc * 4 + k) ^ sw) << 4)) = pk;
                    }
                }
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < kCh / 2; ++j) ptx::tma_store_3d(&tmap_ln, bslabs + j * 4096, (c_begin + 2 * j) * 32, row0, 0);
                ptx::bulk_commit();
            }
        }
        if (lane == 0) ptx::bulk_wait_read<0>();   // smem must stay valid until the last bulk store has read it
        __syncwarp();
        if (et == 0) PD_FSTAMP(29);               // final epilogue done
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) PD_FSTAMP(30);
#undef PD_FSTAMP
}

}  // namespace

struct FfnFusedOpImpl {
    CUtensorMap tmap_a, tmap_w1, tmap_w2, tmap_x, tmap_ln, tmap_att, tmap_wp;
    FfnParams p;
    WRange own_w;
    int tiles;
    int proj;
};
static_assert(sizeof(FfnFusedOpImpl) <= sizeof(FfnFusedOp), "FfnFusedOp storage too small");

int ffn_fused_make(FfnFusedOp* op_, const bf16* ln_in, int M, const bf16* w1, const float* b1, const bf16* w2,
                   const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta, bf16* ln_out,
                   float ln_eps, unsigned long long* dbg, const FfnProjArgs* proj) {
    FfnFusedOpImpl* op = reinterpret_cast<FfnFusedOpImpl*>(op_);
    PD_TRY(gemm_init());
    PD_CHECK(ln_in && w1 && b1 && w2 && b2 && x_inout && M > 0, PD_ERR_ARG, "ffn_fused: null argument");
    PD_CHECK((ln_gamma != nullptr) == (ln_out != nullptr), PD_ERR_ARG, "ffn_fused: ln_gamma and ln_out go together");
    PD_CHECK(!proj || (proj->att && proj->wp && proj->bp && proj->ln1_gamma && proj->ln1_beta), PD_ERR_ARG,
             "ffn_fused: incomplete projection arguments");
    static bool attr_set = false;
    if (!attr_set) {
        PD_CUDA(cudaFuncSetAttribute((ffn_fused_kernel<false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        PD_CUDA(cudaFuncSetAttribute((ffn_fused_kernel<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        PD_CUDA(cudaFuncSetAttribute((ffn_fused_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        PD_CUDA(cudaFuncSetAttribute((ffn_fused_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set = true;
    }
    const uint64_t dims_b[3] = {kC, (uint64_t)M, 1}, st_b[2] = {kC * 2, (uint64_t)kC * 2 * M};   // bf16 [M][256]
    const uint32_t box_ld[3] = {64, 128, 1}, box_st[3] = {64, 32, 1};
    // A of GEMM-1: ln [M][256] bf16, boxes of 128 rows x 64 columns (PROJ: written by the kernel itself first)
    PD_TRY(tmap_encode_sw128(&op->tmap_a, true, 3, ln_in, dims_b, st_b, box_ld));
    {   // W1 [1024][256] (K-major), W2 [256][1024], Wp [256][256]: boxes of 256 rows x 64 K
        const uint64_t d1[2] = {kC, kHid}, s1[1] = {kC * 2};
        const uint64_t d2[2] = {kHid, kC}, s2[1] = {kHid * 2};
        const uint64_t dp[2] = {kC, kC};
        const uint32_t box[2] = {64, 256};
        PD_TRY(tmap_encode_sw128(&op->tmap_w1, true, 2, w1, d1, s1, box));
        PD_TRY(tmap_encode_sw128(&op->tmap_w2, true, 2, w2, d2, s2, box));
        op->tmap_wp = op->tmap_w1;
        op->tmap_att = op->tmap_a;
        if (proj) {
            PD_TRY(tmap_encode_sw128(&op->tmap_wp, true, 2, proj->wp, dp, s1, box));
            PD_TRY(tmap_encode_sw128(&op->tmap_att, true, 3, proj->att, dims_b, st_b, box_ld));
        }
    }
    {   // x [M][256] fp32 (residual in, result out): 32-row x 32-column boxes
        const uint64_t dims[3] = {kC, (uint64_t)M, 1}, st[2] = {kC * 4, (uint64_t)kC * 4 * M};
        const uint32_t box[3] = {32, 32, 1};
        PD_TRY(tmap_encode_sw128(&op->tmap_x, false, 3, x_inout, dims, st, box));
    }
    op->tmap_ln = op->tmap_a;
    if (ln_out) PD_TRY(tmap_encode_sw128(&op->tmap_ln, true, 3, ln_out, dims_b, st_b, box_st));
    op->p.b1 = b1;
    op->p.b2 = b2;
    op->p.ln_gamma = ln_gamma;
    op->p.ln_beta = ln_beta;
    op->p.ln_eps = ln_eps;
    op->p.M = M;
    op->p.dbg = dbg;
    op->p.bp = proj ? proj->bp : nullptr;
    op->p.ln1_gamma = proj ? proj->ln1_gamma : nullptr;
    op->p.ln1_beta = proj ? proj->ln1_beta : nullptr;
    op->p.pf = WRange{};
    op->own_w = WRange{};
    op->own_w.p[0] = reinterpret_cast<const uint8_t*>(w1); op->own_w.n[0] = kHid * kC * 2;
    op->own_w.p[1] = reinterpret_cast<const uint8_t*>(w2); op->own_w.n[1] = kHid * kC * 2;
    if (proj) { op->own_w.p[2] = reinterpret_cast<const uint8_t*>(proj->wp); op->own_w.n[2] = kC * kC * 2; }
    op->p.gn_sums = nullptr;
    op->p.gn_cpg = op->p.gn_groups = op->p.gn_rows = 0;
    op->proj = proj ? 1 : 0;
    op->tiles = ceil_div(M, 128);
    return PD_OK;
}

void ffn_fused_set_prefetch(FfnFusedOp* op_, const WRange& next) { reinterpret_cast<FfnFusedOpImpl*>(op_)->p.pf = next; }
WRange ffn_fused_weights(const FfnFusedOp& op_) { return reinterpret_cast<const FfnFusedOpImpl&>(op_).own_w; }

int ffn_fused_set_gn(FfnFusedOp* op_, double* gn_sums, int groups, int rows) {
    FfnFusedOpImpl* op = reinterpret_cast<FfnFusedOpImpl*>(op_);
    PD_CHECK(gn_sums && gemm_gn_fusable(kC, groups, rows), PD_ERR_SHAPE, "ffn_fused: GroupNorm statistics not fusable");
    op->p.gn_sums = gn_sums;
    op->p.gn_groups = groups;
    op->p.gn_rows = rows;
    op->p.gn_cpg = kC / groups;
    return PD_OK;
}

int ffn_fused_launch(const FfnFusedOp& op_, cudaStream_t st) {
    const FfnFusedOpImpl& op = reinterpret_cast<const FfnFusedOpImpl&>(op_);
#define PD_FFN_LAUNCH(PROJ, GN)                                                                                          \
    PD_LAUNCH((ffn_fused_kernel<PROJ, GN>), op.tiles, kThreads, kSmem, st, op.tmap_a, op.tmap_w1, op.tmap_w2, op.tmap_x,     \
              op.tmap_ln, op.tmap_att, op.tmap_wp, op.p)
    if (op.proj) {
        if (op.p.gn_sums) PD_FFN_LAUNCH(true, true);
        else PD_FFN_LAUNCH(true, false);
    } else {
        if (op.p.gn_sums) PD_FFN_LAUNCH(false, true);
        else PD_FFN_LAUNCH(false, false);
    }
#undef PD_FFN_LAUNCH
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
