// tcgen05 + TMA implicit-GEMM kernel (see gemm.cuh). One 128 x BN output tile per CTA, warp-specialised:
//   warp 0    : TMA producer (one elected lane) - A tile through a rank-5 tensor map with per-tap shifts,
//               B (weights) through a rank-2 map; STAGES-deep mbarrier ring
//   warp 1    : allocates TMEM, one lane issues tcgen05.mma (128 x BN x 16, bf16 -> fp32 in TMEM)
//   warps 2-5 : epilogue - tcgen05.ld 32x32 chunks, transpose through smem, fused bias / time-embedding /
//               GELU / residual, coalesced 128-bit stores of fp32 and/or bf16
#include "gemm.cuh"
#include "ptx.cuh"
#include <cudaTypedefs.h>
#include <cstdlib>
#include <type_traits>

namespace pd {

namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;  // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;  // 16 KB

template <int BN, int STAGES>
struct Cfg {
    static constexpr int kBBytes = BN * kGemmBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kPipeBytes = STAGES * kStageBytes;
    static constexpr int kBarBytes = 512;
    // fused-LN gamma/beta + row-sum exchange inside the CTA + [128] row sums and one mbarrier written by the peer CTA
    static constexpr int kLnBytes = 2 * BN * 4 + kEpiWarps * 32 * 8 + 4 * kGemmBlockM * 8 + 16;   // peers: up to 4 ranks
    static_assert(kPipeBytes >= kEpiWarps * 2 * 4096, "epilogue slabs alias the pipeline stages");
    // barriers + tmem slot (512 B) then the per-CTA epilogue vector (bias + time-embedding row), BN floats
    static constexpr int kSmem = kPipeBytes + 1024 /*align slack*/ + kBarBytes + BN * 4 + kLnBytes;
    static constexpr int kTmemCols = BN < 32 ? 32 : BN;
    // <= ~110 KB of smem lets two CTAs share an SM, so one CTA's epilogue overlaps the other's mainloop
    static constexpr int kMinBlocks = (kSmem <= 112 * 1024) ? 2 : 1;
};

template <int BN, int STAGES, bool GN>   // GN: also accumulate GroupNorm statistics of the output (GemmEpilogue::gn_sums)
__global__ void __launch_bounds__(kThreads, Cfg<BN, STAGES>::kMinBlocks)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
               const __grid_constant__ CUtensorMap tmap_ln, const __grid_constant__ GemmKernelParams p) {
    using C = Cfg<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kPipeBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    uint64_t* res_bar = reinterpret_cast<uint64_t*>(smem + C::kPipeBytes + 128);  // [kEpiWarps][4]
    float* vec_s = reinterpret_cast<float*>(smem + C::kPipeBytes + C::kBarBytes);
    float* ln_g = vec_s + BN;                                   // fused LayerNorm: gamma, beta, per-row partial sums
    float* ln_b = ln_g + BN;
    float2* ln_x = reinterpret_cast<float2*>(ln_b + BN);
    float2* ln_peer = ln_x + kEpiWarps * 32;                    // [4][128]: (sum, sumsq) of my rows, by sender rank
    uint64_t* ln_bar = reinterpret_cast<uint64_t*>(ln_peer + 4 * kGemmBlockM);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // dbg_block = x + 1000 * z selects the stamped CTA (y = 0)
    unsigned long long* dbg = (p.dbg && (int)blockIdx.x == p.dbg_block % 1000 && blockIdx.y == 0 &&
                               (int)blockIdx.z == p.dbg_block / 1000) ? p.dbg : nullptr;
#define PD_STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
    if (threadIdx.x == 32)   // a lane of the MMA warp, idle during set-up
        prefetch_l2_share(p.pf, (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x, gridDim.x * gridDim.y * gridDim.z);
    if (threadIdx.x == 0) {
        PD_STAMP(0);
        if (dbg) {   // wall-clock (ns) stamps: effective SM frequency + launch skew between CTAs
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            dbg[9] = gt;
        }
    }

    const int tile = blockIdx.x;
    const int n0 = blockIdx.y * BN;
    const int sample = tile / p.tiles_per_sample;
    const int p0 = (tile - sample * p.tiles_per_sample) * kGemmBlockM;
    const int num_k_total = p.ntaps * p.cblks;
    // split-K: CTA z handles k-blocks [k_begin, k_end); z = 0 owns bias / residual / plain store, z = 1 adds its
    // partial with a TMA reduce-add once z = 0 has published the tile (fixed order -> deterministic sums)
    const int split = blockIdx.z;
    const int k_begin = (int)(((long long)num_k_total * split) / gridDim.z);
    const int k_end = (int)(((long long)num_k_total * (split + 1)) / gridDim.z);
    const int num_k = k_end - k_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(tmem_full_bar, 1);
        for (int i = 0; i < 4 * kEpiWarps; ++i) ptx::mbar_init(&res_bar[i], 1);
        // cluster LayerNorm: the peers' row sums arrive as asynchronous DSMEM stores counted in bytes on this barrier
        ptx::mbar_init(ln_bar, 1);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
        if (p.ln_gamma && p.ln_cluster > 1)   // 8 bytes per row from each of the other CTAs of the cluster
            ptx::mbar_arrive_expect_tx(ln_bar, (uint32_t)(p.ln_cluster - 1) * kGemmBlockM * 8u);
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        ptx::prefetch_tmap(&tmap_out);
        if (p.has_res) ptx::prefetch_tmap(&tmap_res);
        if (p.ln_gamma) ptx::prefetch_tmap(&tmap_ln);
    }
    // Weight tiles of the first stages do not depend on the preceding kernel: request them before the dependency wait
    // (thread 0 initialised the barriers above). The activation halves of those stages follow after the wait.
    int b_pre = 0;
    if (threadIdx.x == 0 && !p.b_batched) {
        b_pre = num_k < STAGES ? num_k : STAGES;
        for (int it = 0; it < b_pre; ++it) {
            ptx::mbar_arrive_expect_tx(&full_bar[it], C::kStageBytes);
            ptx::tma_load_3d(smem + it * C::kStageBytes + kABytes, &tmap_b, &full_bar[it], (k_begin + it) * p.kblk,
                             n0, 0);
        }
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, C::kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    // cluster LayerNorm: the peer's mbarrier must be initialised before anything is sent to it
    // (split phase: arrive here, wait right before the first store into a peer - nobody stalls at start-up)
    if (p.ln_gamma && p.ln_cluster > 1) ptx::cluster_arrive();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // dependents are released only now: a dependent CTA that became co-resident before this CTA owned its TMEM columns
    // could take them and then sit in its own dependency wait forever
    grid_dep_launch();
    grid_dep_wait();     // everything below reads or writes tensors other kernels of the stream touch
    if (threadIdx.x == 0) PD_STAMP(1);

    if (warp == 0) {
        if (lane == 0) {
            const int z0 = p0 / p.HW;
            const int rem = p0 - z0 * p.HW;
            const int y0 = rem / p.W;
            const int x0 = rem - y0 * p.W;
            const int bz = p.b_batched ? sample : 0;
            int tap = k_begin / p.cblks, cb = k_begin - tap * p.cblks;
            for (int it = 0; it < num_k; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                uint8_t* sa = smem + s * C::kStageBytes;
                if (it >= b_pre) {
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    ptx::mbar_arrive_expect_tx(&full_bar[s], C::kStageBytes);
                }
                ptx::tma_load_5d(sa, &tmap_a, &full_bar[s], cb * p.kblk, x0 + p.dx[tap], y0 + p.dy[tap],
                                 z0 + p.dz[tap], sample);
                if (it >= b_pre) ptx::tma_load_3d(sa + kABytes, &tmap_b, &full_bar[s], (k_begin + it) * p.kblk, n0, bz);
                if (++cb == p.cblks) { cb = 0; ++tap; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // 4 MMAs per 128-byte k-block either way: 16 bf16 or 8 tf32 per instruction
            auto issue = [&](auto tf32_tag) {
                constexpr bool TF32 = decltype(tf32_tag)::value;
                constexpr uint32_t idesc = TF32 ? ptx::make_idesc_tf32(kGemmBlockM, BN) : ptx::make_idesc_bf16(kGemmBlockM, BN);
                for (int it = 0; it < num_k; ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    if (it == 0) PD_STAMP(2);
                    const uint32_t a_addr = ptx::smem_u32(smem + s * C::kStageBytes);
                    const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 16; ++k) {
                        const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * 32);
                        const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * 32);
                        ptx::umma_ss<TF32>(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
                }
            };
            if (p.tf32) issue(std::true_type{});
            else issue(std::false_type{});
            ptx::umma_commit(tmem_full_bar);      // accumulator complete
            PD_STAMP(3);
        }
    } else {
        // ---- epilogue: warps 2..9. Warp w may only touch TMEM lanes [32*(w%4), +32); two warps share a lane
        // quarter and split the BN columns. Each thread owns one output row. All global traffic of the epilogue is
        // TMA: the residual sub-tile (32 rows x 128 B) is bulk-loaded into a swizzled smem slab, combined in place
        // with the accumulator (+ bias / time-embedding from smem, activation) and bulk-stored, so the L1/LSU path
        // never sees the row-per-thread (uncoalesced) pattern and ragged rows are clipped by the tensor map.
        const int e = warp - 2;
        const int q = warp & 3;
        const int half = e >> 2;
        const int et = threadIdx.x - 64;
        for (int i = et; i < BN; i += 32 * kEpiWarps) {
            float v = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
            if (p.rowvec) v += __ldg(p.rowvec + (size_t)sample * p.rowvec_ld + n0 + i);
            vec_s[i] = v;
            if (p.ln_gamma) {
                ln_g[i] = __ldg(p.ln_gamma + n0 + i);
                ln_b[i] = __ldg(p.ln_beta + n0 + i);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        uint8_t* slab0 = smem + e * 8192;               // two 4 KB slabs per warp, aliasing the (finished) pipeline
        uint64_t* my_bar = res_bar + 4 * e;
        const int row0 = p0 + q * 32;                   // first row of this warp inside the sample
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        const int act = p.act;
        ptx::mbar_wait(tmem_full_bar, 0);
        ptx::tc_fence_after();
        if (e == 0 && lane == 0) PD_STAMP(4);
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        if (!p.out_is_bf16) {
            constexpr int kChunks = BN / 32;
            constexpr int kPerHalf = (kChunks + 1) / 2;
            constexpr int kMaxSlabs = C::kPipeBytes / (kEpiWarps * 4096);
            constexpr int kSlabs = kPerHalf < kMaxSlabs ? kPerHalf : kMaxSlabs;   // 4 KB slabs per warp
            constexpr bool kOwnSlab = kSlabs >= kPerHalf;                         // one slab per chunk: no reuse waits
            uint8_t* slabs = smem + e * (kSlabs * 4096);   // aliases the (finished) pipeline stages
            const int c_begin = half * kPerHalf;
            const int c_end = (c_begin + kPerHalf) < kChunks ? (c_begin + kPerHalf) : kChunks;
            const bool has_res = p.has_res != 0 && split == 0;
            if (split != 0) {
                // wait until split 0 has stored this tile, then add the partial sums on top (TMA reduce-add)
                if (et == 0) {
                    const int* flag = p.split_flags + blockIdx.y * gridDim.x + blockIdx.x;
                    while (ptx::ld_acquire_gpu(flag) == 0) {}
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                ptx::fence_proxy_async_all();
#pragma unroll 1
                for (int c = c_begin, idx = 0; c < c_end; ++c, ++idx) {
                    uint8_t* slab = slabs + (kOwnSlab ? idx : (idx % kSlabs)) * 4096;
                    if (!kOwnSlab) {
                        if (lane == 0) ptx::bulk_wait_read<kSlabs - 1>();
                        __syncwarp();
                    }
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_lane + c * 32, v);
                    ptx::tmem_ld_wait();
                    uint8_t* my_row = slab + lane * 128;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<uint4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4)) =
                            make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_reduce_add_3d(&tmap_out, slab, n0 + c * 32, row0, sample);
                        ptx::bulk_commit();
                    }
                }
                if (lane == 0) ptx::bulk_wait_all<0>();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                if (et == 0) p.split_flags[blockIdx.y * gridDim.x + blockIdx.x] = 0;   // re-arm for the next launch
            } else {
                if (has_res && lane == 0) {
                    if (kOwnSlab) {
                        for (int c = c_begin, idx = 0; c < c_end; ++c, ++idx) {
                            ptx::mbar_arrive_expect_tx(&my_bar[idx], 4096);
                            ptx::tma_load_3d(slabs + idx * 4096, &tmap_res, &my_bar[idx], n0 + c * 32, row0, sample);
                        }
                    } else if (c_begin < c_end) {
                        ptx::mbar_arrive_expect_tx(&my_bar[0], 4096);
                        ptx::tma_load_3d(slabs, &tmap_res, &my_bar[0], n0 + c_begin * 32, row0, sample);
                    }
                }
                float ln_s1 = 0.f, ln_s2 = 0.f;
                const int gsample = GN ? (sample * p.rows_per_sample + row0) / p.gn_rows : 0;   // GroupNorm sample of my rows
#pragma unroll 1
                for (int c = c_begin, idx = 0; c < c_end; ++c, ++idx) {
                    const int s = kOwnSlab ? idx : (idx % kSlabs);
                    uint8_t* slab = slabs + s * 4096;
                    if (!kOwnSlab) {
                        if (lane == 0) {
                            if (has_res) {
                                ptx::bulk_wait_read<0>();   // the store that used the next slab has drained
                                if (c + 1 < c_end) {
                                    const int s1 = (idx + 1) % kSlabs;
                                    ptx::mbar_arrive_expect_tx(&my_bar[s1], 4096);
                                    ptx::tma_load_3d(slabs + s1 * 4096, &tmap_res, &my_bar[s1], n0 + (c + 1) * 32, row0,
                                                     sample);
                                }
                            } else {
                                ptx::bulk_wait_read<kSlabs - 1>();
                            }
                        }
                        __syncwarp();
                    }
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_lane + c * 32, v);
                    if (has_res) ptx::mbar_wait(&my_bar[s], kOwnSlab ? 0u : static_cast<uint32_t>((idx / kSlabs) & 1));
                    ptx::tmem_ld_wait();
                    if (e == 0 && lane == 0 && idx == 0) PD_STAMP(5);
                    uint8_t* my_row = slab + lane * 128;
                    // the residual cells first: through generic addresses a store between two loads serialises them
                    float4 resv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        resv[i] = has_res ? *reinterpret_cast<const float4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                    float gs[8], gq[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = *reinterpret_cast<const float4*>(vec_s + c * 32 + 4 * i);
                        float4 a = make_float4(__uint_as_float(v[4 * i]) + b.x, __uint_as_float(v[4 * i + 1]) + b.y,
                                               __uint_as_float(v[4 * i + 2]) + b.z, __uint_as_float(v[4 * i + 3]) + b.w);
                        if (act == ACT_GELU) {
                            a.x = gelu_fast(a.x); a.y = gelu_fast(a.y); a.z = gelu_fast(a.z); a.w = gelu_fast(a.w);
                        } else if (act == ACT_SILU) {
                            a.x = silu_f(a.x); a.y = silu_f(a.y); a.z = silu_f(a.z); a.w = silu_f(a.w);
                        }
                        if (p.round_out) {   // this output is the next tf32 GEMM's operand
                            a.x = tf32_rna(a.x); a.y = tf32_rna(a.y); a.z = tf32_rna(a.z); a.w = tf32_rna(a.w);
                        }
                        float4* cell = reinterpret_cast<float4*>(my_row + ((static_cast<uint32_t>(i) ^ sw) << 4));
                        if (has_res) {
                            const float4 r = resv[i];
                            a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
                        }
                        *cell = a;
                        gs[i] = (a.x + a.y) + (a.z + a.w);
                        gq[i] = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
                        ln_s1 += gs[i];
                        ln_s2 += gq[i];
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_3d(&tmap_out, slab, n0 + c * 32, row0, sample);
                        ptx::bulk_commit();
                    }
                    if constexpr (GN) if (row0 < p.rows_per_sample) {   // statistics for the GroupNorm that reads this output
                        const int shift = p.gn_cpg == 8 ? 3 : (p.gn_cpg == 16 ? 4 : 5);
                        gn_chunk_accumulate(gs, gq, row0 + lane < p.rows_per_sample, p.gn_cpg,
                                            p.gn_sums + ((size_t)gsample * p.gn_groups + ((n0 + c * 32) >> shift)) * 2, lane);
                    }
                }
                if constexpr (kOwnSlab && ((BN == 256 && kPerHalf == 4) || (BN == 128 && kPerHalf == 2))) {
                    constexpr int kBf = kPerHalf / 2;   // bf16 slabs (64 columns each) per warp
                    if (p.ln_gamma) {
                        // ---- fused LayerNorm of the finished rows (this CTA / cluster owns all N columns) ----
                        // each row lives in two warps (column halves): exchange partial sums through smem
                        ln_x[(q * 2 + half) * 32 + lane] = make_float2(ln_s1, ln_s2);
                        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
                        const float2 o = ln_x[(q * 2 + (half ^ 1)) * 32 + lane];
                        float tot1 = ln_s1 + o.x, tot2 = ln_s2 + o.y, inv_n = 1.0f / BN;
                        if (p.ln_cluster > 1) {
                            // N = nc x BN: send this CTA's row sums to every peer (DSMEM store + remote mbarrier
                            // arrive), wait for theirs; all CTAs add the nc partials in rank order, so the statistics are
                            // bit-identical on every side
                            const uint32_t nc = (uint32_t)p.ln_cluster, me = ptx::cluster_ctarank();
                            ptx::cluster_wait();   // every peer has initialised and armed its ln_bar (arrive: kernel start)
                            if (half == 0) {
                                for (uint32_t d = 1; d < nc; ++d) {
                                    const uint32_t peer = (me + d) & (nc - 1);
                                    ptx::st_async_cluster_f32x2(
                                        ptx::mapa(ptx::smem_u32(&ln_peer[me * kGemmBlockM + q * 32 + lane]), peer), tot1, tot2,
                                        ptx::mapa(ptx::smem_u32(ln_bar), peer));
                                }
                            }
                            ptx::mbar_wait_cluster(ln_bar, 0);
                            float s1 = 0.f, s2 = 0.f;
                            for (uint32_t r = 0; r < nc; ++r) {
                                const float2 v = r == me ? make_float2(tot1, tot2) : ln_peer[r * kGemmBlockM + q * 32 + lane];
                                s1 = r == 0 ? v.x : s1 + v.x;
                                s2 = r == 0 ? v.y : s2 + v.y;
                            }
                            tot1 = s1;
                            tot2 = s2;
                            inv_n = 1.0f / (float)(nc * BN);
                        }
                        const float mean = tot1 * inv_n;
                        const float var = fmaxf(tot2 * inv_n - mean * mean, 0.f);
                        const float rstd = rsqrtf(var + p.ln_eps);
                        if (p.tf32) {
                            // tf32 operands: the normalised rows are the next GEMM's fp32 operand - formed in place in the fp32
                            // slabs once the bulk stores of x have read them, rounded to tf32, stored through the fp32 map
                            if (lane == 0) ptx::bulk_wait_read<0>();
                            __syncwarp();
#pragma unroll
                            for (int idx = 0; idx < kPerHalf; ++idx) {
                                uint8_t* frow = slabs + idx * 4096 + lane * 128;
                                const int colbase = (c_begin + idx) * 32;
                                float4 av[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    av[i] = *reinterpret_cast<const float4*>(frow + ((static_cast<uint32_t>(i) ^ sw) << 4));
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 g0 = *reinterpret_cast<const float4*>(ln_g + colbase + 4 * i);
                                    const float4 b0 = *reinterpret_cast<const float4*>(ln_b + colbase + 4 * i);
                                    float4 o;
                                    o.x = tf32_rna(fmaf((av[i].x - mean) * rstd, g0.x, b0.x));
                                    o.y = tf32_rna(fmaf((av[i].y - mean) * rstd, g0.y, b0.y));
                                    o.z = tf32_rna(fmaf((av[i].z - mean) * rstd, g0.z, b0.z));
                                    o.w = tf32_rna(fmaf((av[i].w - mean) * rstd, g0.w, b0.w));
                                    *reinterpret_cast<float4*>(frow + ((static_cast<uint32_t>(i) ^ sw) << 4)) = o;
                                }
                            }
                            ptx::fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) {
#pragma unroll
                                for (int idx = 0; idx < kPerHalf; ++idx)
                                    ptx::tma_store_3d(&tmap_ln, slabs + idx * 4096, n0 + (c_begin + idx) * 32, row0, sample);
                                ptx::bulk_commit();
                            }
                        } else {
                        uint8_t* bslabs = smem + kEpiWarps * kPerHalf * 4096 + e * (kBf * 4096);   // kBf bf16 slabs per warp
#pragma unroll
                        for (int j = 0; j < kBf; ++j) {        // bf16 slab j = my fp32 chunks 2j, 2j+1 (64 columns)
                            uint8_t* brow = bslabs + j * 4096 + lane * 128;
#pragma unroll
                            for (int cc = 0; cc < 2; ++cc) {
                                const uint8_t* frow = slabs + (2 * j + cc) * 4096 + lane * 128;
                                const int colbase = (c_begin + 2 * j + cc) * 32;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {  // two fp32 cells -> one 16-byte bf16 cell
                                    const float4 a0 = *reinterpret_cast<const float4*>(frow + ((static_cast<uint32_t>(2 * k) ^ sw) << 4));
                                    const float4 a1 = *reinterpret_cast<const float4*>(frow + ((static_cast<uint32_t>(2 * k + 1) ^ sw) << 4));
                                    const float4 g0 = *reinterpret_cast<const float4*>(ln_g + colbase + 8 * k);
                                    const float4 g1 = *reinterpret_cast<const float4*>(ln_g + colbase + 8 * k + 4);
                                    const float4 b0 = *reinterpret_cast<const float4*>(ln_b + colbase + 8 * k);
                                    const float4 b1 = *reinterpret_cast<const float4*>(ln_b + colbase + 8 * k + 4);
                                    uint4 pk;
                                    pk.x = pack_bf16x2(fmaf((a0.x - mean) * rstd, g0.x, b0.x), fmaf((a0.y - mean) * rstd, g0.y, b0.y));
                                    pk.y = pack_bf16x2(fmaf((a0.z - mean) * rstd, g0.z, b0.z), fmaf((a0.w - mean) * rstd, g0.w, b0.w));
                                    pk.z = pack_bf16x2(fmaf((a1.x - mean) * rstd, g1.x, b1.x), fmaf((a1.y - mean) * rstd, g1.y, b1.y));
                                    pk.w = pack_bf16x2(fmaf((a1.z - mean) * rstd, g1.z, b1.z), fmaf((a1.w - mean) * rstd, g1.w, b1.w));
                                    *reinterpret_cast<uint4*>(brow + ((static_cast<uint32_t>(cc * 4 + k) ^ sw) << 4)) = pk;
                                }
                            }
                        }
                        ptx::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
#pragma unroll
                            for (int j = 0; j < kBf; ++j)
                                ptx::tma_store_3d(&tmap_ln, bslabs + j * 4096, n0 + (c_begin + 2 * j) * 32, row0, sample);
                            ptx::bulk_commit();
                        }
                        }
                    }
                }
                if (gridDim.z > 1) {
                    // publish the tile for the split-1 CTA: all bulk stores complete -> fence -> flag
                    if (lane == 0) ptx::bulk_wait_all<0>();
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                    if (et == 0) {
                        ptx::fence_proxy_async_all();
                        __threadfence();
                        ptx::st_release_gpu(p.split_flags + blockIdx.y * gridDim.x + blockIdx.x, 1);
                    }
                }
            }
        } else {
            constexpr int kChunks = BN / 64;  // 64 bf16 columns = one 128-byte slab row
            constexpr int kPerHalf = (kChunks + 1) / 2;
            const int c_begin = half * kPerHalf;
            const int c_end = (c_begin + kPerHalf) < kChunks ? (c_begin + kPerHalf) : kChunks;
#pragma unroll 1
            static_assert(kPerHalf <= 2, "bf16 epilogue: one slab per chunk");
            for (int c = c_begin, idx = 0; c < c_end; ++c, ++idx) {
                uint8_t* slab = slab0 + idx * 4096;
                uint8_t* my_row = slab + lane * 128;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_lane + c * 64 + hh * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {   // 8 columns -> one 16-byte cell
                        float a[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            a[k] = __uint_as_float(v[8 * i + k]) + vec_s[c * 64 + hh * 32 + 8 * i + k];
                            if (act == ACT_GELU) a[k] = gelu_fast(a[k]);
                            else if (act == ACT_SILU) a[k] = silu_f(a[k]);
                        }
                        const uint32_t cellidx = static_cast<uint32_t>(hh * 4 + i) ^ sw;
                        *reinterpret_cast<uint4*>(my_row + (cellidx << 4)) =
                            make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]),
                                       pack_bf16x2(a[6], a[7]));
                    }
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_3d(&tmap_out, slab, n0 + c * 64, row0, sample);
                    ptx::bulk_commit();
                }
            }
        }
        if (e == 0 && lane == 0) PD_STAMP(6);
        if (lane == 0) ptx::bulk_wait_read<0>();   // smem must stay valid until the last bulk store has read it
        __syncwarp();
        if (e == 0 && lane == 0) PD_STAMP(7);
    }
    if (warp < 2 && p.ln_gamma && p.ln_cluster > 1) ptx::cluster_wait();   // pairs with the arrive of these two warps
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, C::kTmemCols);
    if (threadIdx.x == 0) {
        PD_STAMP(8);
        if (dbg) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            dbg[10] = gt;
        }
    }
#undef PD_STAMP
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant for the multi-wave, bf16-output GEMMs (QKV, FFN-1: K <= 512, several tiles per SM): one CTA per
// SM walks tiles m-major; the accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i
// (8 warps: TMEM -> +bias -> GELU -> bf16 -> swizzled slab -> TMA store) runs under the TMA loads + MMAs of tile i+1,
// and barrier setup / TMEM allocation / first-tile latency are paid once per SM instead of once per tile.
constexpr int kPersBN = 256;
constexpr int kPersStages = 3;
struct PersCfg {
    static constexpr int kStageBytes = kABytes + kPersBN * kGemmBlockK * 2;   // 48 KB
    static constexpr int kPipeBytes = kPersStages * kStageBytes;              // 144 KB
    static constexpr int kSlabBytes = kEpiWarps * 2 * 4096;                   // 64 KB, NOT aliased (pipeline stays live)
    static constexpr int kBarBytes = 256;
    static constexpr int kSmem = kPipeBytes + kSlabBytes + 1024 + kBarBytes + 2 * kPersBN * 4;
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                          const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ GemmKernelParams p,
                          int n_tiles_n, int total_tiles) {
    constexpr int BN = kPersBN, STAGES = kPersStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* slab_base = smem + PersCfg::kPipeBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(slab_base + PersCfg::kSlabBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;     // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* vec_s = reinterpret_cast<float*>(slab_base + PersCfg::kSlabBytes + PersCfg::kBarBytes);  // [2][BN]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_k = p.ntaps * p.cblks;
    unsigned long long* dbg = (p.dbg && (int)blockIdx.x == p.dbg_block) ? p.dbg : nullptr;
    if (threadIdx.x == 32) prefetch_l2_share(p.pf, blockIdx.x, gridDim.x);
    if (dbg && threadIdx.x == 0) {
        dbg[0] = clock64();
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg[9] = gt;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&acc_full[a], 1);
            ptx::mbar_init(&acc_empty[a], kEpiWarps);
        }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        ptx::prefetch_tmap(&tmap_out);
    }
    int b_pre = 0;   // weight tiles of the first tile's first stages, requested before the dependency wait
    if (threadIdx.x == 0 && !p.b_batched) {
        b_pre = num_k < STAGES ? num_k : STAGES;
        const int n0 = ((int)blockIdx.x % n_tiles_n) * BN;
        for (int it = 0; it < b_pre; ++it) {
            ptx::mbar_arrive_expect_tx(&full_bar[it], PersCfg::kStageBytes);
            ptx::tma_load_3d(smem + it * PersCfg::kStageBytes + kABytes, &tmap_b, &full_bar[it], it * kGemmBlockK, n0, 0);
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
    // dependents are released only now: a dependent CTA that became co-resident before this CTA owned its TMEM columns
    // could take them and then sit in its own dependency wait forever
    grid_dep_launch();
    grid_dep_wait();
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;  // running k-block counter across tiles: stage = it % STAGES
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_tile = tile / n_tiles_n;
                const int n0 = (tile - m_tile * n_tiles_n) * BN;
                const int sample = m_tile / p.tiles_per_sample;
                const int p0 = (m_tile - sample * p.tiles_per_sample) * kGemmBlockM;
                const int z0 = p0 / p.HW;
                const int rem = p0 - z0 * p.HW;
                const int y0 = rem / p.W;
                const int x0 = rem - y0 * p.W;
                const int bz = p.b_batched ? sample : 0;
                int tap = 0, cb = 0;
                for (int k = 0; k < num_k; ++k, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    uint8_t* sa = smem + s * PersCfg::kStageBytes;
                    if (it >= b_pre) {
                        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                        ptx::mbar_arrive_expect_tx(&full_bar[s], PersCfg::kStageBytes);
                    }
                    ptx::tma_load_5d(sa, &tmap_a, &full_bar[s], cb * kGemmBlockK, x0 + p.dx[tap], y0 + p.dy[tap],
                                     z0 + p.dz[tap], sample);
                    if (it >= b_pre) ptx::tma_load_3d(sa + kABytes, &tmap_b, &full_bar[s], k * kGemmBlockK, n0, bz);
                    if (++cb == p.cblks) { cb = 0; ++tap; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kGemmBlockM, BN);
            int it = 0, li = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
                const int acc = li & 1;
                ptx::mbar_wait(&acc_empty[acc], ((li >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int k = 0; k < num_k; ++k, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(smem + s * PersCfg::kStageBytes);
                    const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                    for (int kk = 0; kk < kGemmBlockK / 16; ++kk) {
                        const uint64_t da = ptx::make_smem_desc_sw128(a_addr + kk * 32);
                        const uint64_t db = ptx::make_smem_desc_sw128(b_addr + kk * 32);
                        ptx::umma_f16(tmem_d, da, db, idesc, (k | kk) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[s]);
                }
                ptx::umma_commit(&acc_full[acc]);
            }
        }
    } else {
        const int e = warp - 2;
        const int q = warp & 3;
        const int half = e >> 2;
        const int et = threadIdx.x - 64;
        uint8_t* slab0 = slab_base + e * 8192;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        const int act = p.act;
        constexpr int kChunks = BN / 64, kPerHalf = kChunks / 2;   // 2 chunks of 64 bf16 columns per warp
        int li = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
            const int acc = li & 1;
            const int m_tile = tile / n_tiles_n;
            const int n0 = (tile - m_tile * n_tiles_n) * BN;
            const int sample = m_tile / p.tiles_per_sample;
            const int p0 = (m_tile - sample * p.tiles_per_sample) * kGemmBlockM;
            const int row0 = p0 + q * 32;
            float* vs = vec_s + acc * BN;
            // bias (+ per-sample row vector) for this tile's columns; the other buffer may still be in use by slower warps
            {
                float v = p.bias ? __ldg(p.bias + n0 + et) : 0.f;
                if (p.rowvec) v += __ldg(p.rowvec + (size_t)sample * p.rowvec_ld + n0 + et);
                vs[et] = v;   // 256 epilogue threads == BN columns
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            if (lane == 0) ptx::bulk_wait_read<0>();   // my slabs: the previous tile's stores have been read out
            __syncwarp();
            ptx::mbar_wait(&acc_full[acc], (li >> 1) & 1);
            ptx::tc_fence_after();
            if (dbg && et == 0 && li < 3) dbg[2 + 2 * li] = clock64();
            const uint32_t t_lane = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int c = half * kPerHalf, idx = 0; c < (half + 1) * kPerHalf; ++c, ++idx) {
                uint8_t* my_row = slab0 + idx * 4096 + lane * 128;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_lane + c * 64 + hh * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float a[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            a[k] = __uint_as_float(v[8 * i + k]) + vs[c * 64 + hh * 32 + 8 * i + k];
                            if (act == ACT_GELU) a[k] = gelu_fast(a[k]);
                            else if (act == ACT_SILU) a[k] = silu_f(a[k]);
                        }
                        const uint32_t cellidx = static_cast<uint32_t>(hh * 4 + i) ^ sw;
                        *reinterpret_cast<uint4*>(my_row + (cellidx << 4)) =
                            make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]),
                                       pack_bf16x2(a[6], a[7]));
                    }
                }
                if (idx == kPerHalf - 1) {   // all TMEM reads of this tile are done: hand the accumulator back
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_3d(&tmap_out, slab0 + idx * 4096, n0 + c * 64, row0, sample);
                    ptx::bulk_commit();
                }
            }
            if (dbg && et == 0 && li < 3) dbg[3 + 2 * li] = clock64();
        }
        if (lane == 0) ptx::bulk_wait_read<0>();
        __syncwarp();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
    if (dbg && threadIdx.x == 0) {
        dbg[8] = clock64();
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg[10] = gt;
    }
}

PFN_cuTensorMapEncodeTiled g_encode = nullptr;
bool g_inited = false;

template <int BN, int STAGES>
int set_smem_attr() {
    PD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg<BN, STAGES>::kSmem));
    if (BN == 256 && STAGES == 4)
        PD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<256, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg<256, 4>::kSmem));
    return PD_OK;
}

int launch_persistent(const GemmOp& op, cudaStream_t stream) {
    const int total = (int)(op.grid_x * op.grid_y);
    const int grid = total < kNumSMs ? total : kNumSMs;
    PD_LAUNCH(gemm_tc_persistent_kernel, grid, kThreads, PersCfg::kSmem, stream, op.tmap_a, op.tmap_b, op.tmap_out, op.p,
              (int)op.grid_y, total);
    return PD_OK;
}

template <int BN, int STAGES>
int launch_cfg(const GemmOp& op, cudaStream_t stream) {
    // cluster_y > 1: the CTAs of one row tile (along N) form a thread-block cluster
    if (op.p.gn_sums) {   // fused GroupNorm statistics exist for the 256 x 4-stage configuration only (gemm_make checks)
        PD_CUDA(launch_pdl(gemm_tc_kernel<256, 4, true>, dim3(op.grid_x, op.grid_y, op.split_k), dim3(kThreads),
                           (size_t)Cfg<256, 4>::kSmem, stream, dim3(1, (unsigned)op.cluster_y, 1), op.tmap_a, op.tmap_b,
                           op.tmap_out, op.tmap_res, op.tmap_ln, op.p));
        return PD_OK;
    }
    PD_CUDA(launch_pdl(gemm_tc_kernel<BN, STAGES, false>, dim3(op.grid_x, op.grid_y, op.split_k), dim3(kThreads),
                       (size_t)Cfg<BN, STAGES>::kSmem, stream, dim3(1, (unsigned)op.cluster_y, 1), op.tmap_a, op.tmap_b,
                       op.tmap_out, op.tmap_res, op.tmap_ln, op.p));
    return PD_OK;
}

}  // namespace

int gemm_init() {
    if (g_inited) return PD_OK;
    int dev = 0;
    PD_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    PD_CUDA(cudaGetDeviceProperties(&prop, dev));
    PD_CHECK(prop.major == 10, PD_ERR_ARCH,
             "prediff_b200 requires an sm_100 (Blackwell B200) device; found sm_%d%d (%s). There is no fallback path.",
             prop.major, prop.minor, prop.name);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PD_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, PD_ERR_CUDA,
             "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    PD_TRY((set_smem_attr<32, 4>()));
    PD_TRY((set_smem_attr<64, 4>()));
    PD_TRY((set_smem_attr<128, 3>()));
    PD_TRY((set_smem_attr<256, 2>()));
    PD_TRY((set_smem_attr<256, 4>()));
    PD_CUDA(cudaFuncSetAttribute(gemm_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PersCfg::kSmem));
    g_inited = true;
    return PD_OK;
}

// cuTensorMapEncodeTiled with the settings every kernel here uses (128-byte swizzle, no OOB fill); for other
// translation units (ffn_fused.cu) that build their own tensor maps.
int tmap_encode_sw128(CUtensorMap* m, bool is_bf16, int rank, const void* ptr, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box) {
    PD_TRY(gemm_init());
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], es[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
    PD_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, PD_ERR_ARG, "tensor map: pointer must be 16-byte aligned");
    CUresult r = g_encode(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                          const_cast<void*>(ptr), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PD_CHECK(r == CUDA_SUCCESS, PD_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
    return PD_OK;
}

// Tile-shape preference of gemm_make for the ops built from now on (set around the construction of a plan): 0 = fill the
// machine (narrow tiles when a wide one would leave more than half of the SMs idle), 1 = widest tile that divides N - for a
// network that runs BESIDE another one (the knowledge-alignment net next to the UNet): a 128 x 32 tcgen05.mma costs the same
// ~128 cycles as 128 x 256, so the wide tile has the same latency on 8 x fewer SMs.
static int g_tile_pref = 0;
void gemm_set_tile_preference(int wide) { g_tile_pref = wide; }

int gemm_split_flags_needed(const GemmGeom& g, int N, int force_split) {
    const int out_D = g.out_D ? g.out_D : g.D;
    const int num_k = g.ntaps * (g.C / g.kblk());
    if ((num_k < 128 && force_split != 2) || N % 256 != 0) return 0;
    return ceil_div(out_D * g.H * g.W, kGemmBlockM) * g.samples * (N / 256);
}

int gemm_make(GemmOp* op, const void* A, const GemmGeom& g, const void* Wt, int N, const GemmEpilogue& e,
              int force_block_n) {
    PD_TRY(gemm_init());
    PD_CHECK(A && Wt, PD_ERR_ARG, "gemm_make: null operand");
    const int kblk = g.kblk();
    const int esz = g.tf32 ? 4 : 2;   // operand element size
    const CUtensorMapDataType op_dtype = g.tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    PD_CHECK(g.C > 0 && g.C % kblk == 0, PD_ERR_SHAPE, "gemm: channels per tap (%d) must be a multiple of %d", g.C, kblk);
    PD_CHECK(!g.tf32 || !e.out_bf16, PD_ERR_ARG,
             "gemm: tf32 operands go with fp32 outputs only (the fused LayerNorm output is then fp32 rounded to tf32)");
    PD_CHECK(N > 0 && N % 32 == 0, PD_ERR_SHAPE, "gemm: N (%d) must be a multiple of 32", N);
    PD_CHECK(g.ntaps >= 1 && g.ntaps <= kMaxTaps, PD_ERR_SHAPE, "gemm: ntaps %d out of range", g.ntaps);
    PD_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wt) & 15) == 0, PD_ERR_ARG,
             "gemm: operands must be 16-byte aligned");
    const int ldo = e.ldo ? e.ldo : N;
    PD_CHECK(ldo % 4 == 0, PD_ERR_SHAPE, "gemm: output row stride must be a multiple of 4");
    PD_CHECK((e.out_f32 != nullptr) != (e.out_bf16 != nullptr), PD_ERR_ARG,
             "gemm: exactly one of out_f32 / out_bf16 must be given");
    PD_CHECK(!e.residual || e.out_f32, PD_ERR_ARG, "gemm: a residual needs the fp32 output");

    // ---- A box: 128 rows = bw x bh x bd positions -------------------------------------------------------------
    const int bw = g.W < kGemmBlockM ? g.W : kGemmBlockM;
    PD_CHECK(kGemmBlockM % bw == 0, PD_ERR_SHAPE, "gemm: W=%d must divide 128 (or exceed it for a plain linear)", g.W);
    PD_CHECK(g.W <= kGemmBlockM || (g.H == 1 && g.D == 1), PD_ERR_SHAPE, "gemm: W=%d > 128 only for plain linears", g.W);
    int rows_left = kGemmBlockM / bw;
    const int bh = g.H < rows_left ? g.H : rows_left;
    PD_CHECK(rows_left % bh == 0 && g.H % bh == 0, PD_ERR_SHAPE, "gemm: H=%d incompatible with the 128-row tile", g.H);
    const int bd = rows_left / bh;
    PD_CHECK(bd == 1 || bh == g.H, PD_ERR_SHAPE, "gemm: tile geometry");
    const int out_D = g.out_D ? g.out_D : g.D;
    PD_CHECK(out_D == g.D || bd == 1, PD_ERR_SHAPE, "gemm: plane-indexed A needs H*W >= 128");

    const int64_t sW = g.sW ? g.sW : g.C;
    const int64_t sH = g.sH ? g.sH : sW * g.W;
    const int64_t sD = g.sD ? g.sD : sH * g.H;
    const int64_t sN = g.sN ? g.sN : sD * g.D;
    {
        cuuint64_t dims[5] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.D, (cuuint64_t)g.samples};
        cuuint64_t strides[4] = {(cuuint64_t)sW * esz, (cuuint64_t)sH * esz, (cuuint64_t)sD * esz, (cuuint64_t)sN * esz};
        cuuint32_t box[5] = {(cuuint32_t)kblk, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = g_encode(&op->tmap_a, op_dtype, 5, const_cast<void*>(A), dims, strides, box,
                              es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PD_CHECK(r == CUDA_SUCCESS, PD_ERR_CUDA,
                 "cuTensorMapEncodeTiled(A) failed: %d (dims %d,%d,%d,%d,%d box %d,%d,%d,%d)", (int)r, g.C, g.W, g.H, g.D,
                 g.samples, kblk, bw, bh, bd);
    }

    // ---- tile shape -------------------------------------------------------------------------------------------
    const int rows_per_sample = out_D * g.H * g.W;
    const int tiles_per_sample = ceil_div(rows_per_sample, kGemmBlockM);
    const int m_tiles = tiles_per_sample * g.samples;
    int bn = force_block_n;
    const bool want_ln = e.ln_out != nullptr;
    if (want_ln) {
        PD_CHECK((N == 256 || N == 512) && e.out_f32 && e.ln_gamma && e.ln_beta, PD_ERR_SHAPE,
                 "gemm: the fused output LayerNorm needs N == 256 or 512 and an fp32 output (got N=%d)", N);
        // N = 512: four 128-wide CTAs per row tile in a cluster (half the epilogue per CTA, twice the CTAs; a
        // 128 x 128 MMA costs as much as 128 x 256, but these GEMMs are epilogue-bound) - PD_LN_BN256=1: two 256-wide
        bn = (N == 512 && getenv("PD_LN_BN256") == nullptr) ? 128 : 256;
    }
    if (!bn) {
        if (const char* s = getenv("PD_GEMM_BN")) bn = atoi(s);
        if (bn && N % bn != 0) bn = 0;
    }
    const int num_k = g.ntaps * (g.C / kblk);
    if (!bn && e.split_flags && gemm_split_flags_needed(g, N, e.force_split) > 0 && e.out_f32 && e.act == ACT_NONE)
        bn = 256;
    if (!bn && e.gn_sums && N % 256 == 0) bn = 256;   // the statistics epilogue exists for the 256-wide, 4-stage kernel
    // A 128 x 128 tcgen05.mma takes the same ~128 cycles as 128 x 256 (measured, tools/gemm_phases.py), so once the
    // mainloop matters (>= 16 k-blocks) BN = 256 halves it even if fewer CTAs run.
    if (!bn && N % 256 == 0 && num_k >= 16 && (int64_t)m_tiles * (N / 256) >= kNumSMs / 4) bn = 256;
    if (!bn) {
        // Wide tiles cut L2->SM operand traffic (the binding resource of the implicit GEMM); shrink only when the
        // grid would leave more than half of the SMs idle.
        const int cands[4] = {256, 128, 64, 32};
        int smallest = 0;
        for (int c : cands) {
            if (N % c != 0 || (e.out_bf16 && c < 64)) continue;
            if (c >= 64 || !smallest) smallest = c;
            if ((int64_t)m_tiles * (N / c) >= kNumSMs / 2 || g_tile_pref == 1) { bn = c; break; }
        }
        if (!bn) bn = smallest;
    }
    PD_CHECK(bn == 32 || bn == 64 || bn == 128 || bn == 256, PD_ERR_SHAPE, "gemm: bad BLOCK_N %d", bn);
    PD_CHECK(N % bn == 0, PD_ERR_SHAPE, "gemm: N %d not a multiple of BLOCK_N %d", N, bn);
    {
        const int64_t ktot = (int64_t)g.ntaps * g.C;
        const int64_t ldb = g.ldb ? g.ldb : ktot;
        const int nb = g.b_sample_stride ? g.samples : 1;
        const int64_t bs = g.b_sample_stride ? g.b_sample_stride : ldb * N;
        cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)N, (cuuint64_t)nb};
        cuuint64_t strides[2] = {(cuuint64_t)ldb * esz, (cuuint64_t)bs * esz};
        cuuint32_t box[3] = {(cuuint32_t)kblk, (cuuint32_t)bn, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = g_encode(&op->tmap_b, op_dtype, 3, const_cast<void*>(Wt), dims, strides, box,
                              es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PD_CHECK(r == CUDA_SUCCESS, PD_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }

    PD_CHECK(!e.out_bf16 || bn >= 64, PD_ERR_SHAPE, "gemm: bf16 output needs N to be a multiple of 64 (got %d)", N);
    op->ldo = ldo;
    op->out_rows = rows_per_sample;
    op->out_samples = g.samples;
    op->out_N = N;
    PD_TRY(gemm_bind_output(op, e.out_f32, e.out_bf16, e.residual));

    GemmKernelParams& p = op->p;
    memset(&p, 0, sizeof(p));
    p.rows_per_sample = rows_per_sample;
    p.tiles_per_sample = tiles_per_sample;
    p.samples = g.samples;
    p.N = N;
    p.H = g.H;
    p.W = g.W;
    p.HW = g.H * g.W;
    p.ntaps = g.ntaps;
    p.cblks = g.C / kblk;
    p.tf32 = g.tf32 ? 1 : 0;
    p.kblk = kblk;
    p.round_out = e.round_tf32;
    p.b_batched = g.b_sample_stride ? 1 : 0;
    memcpy(p.dz, g.dz, sizeof(p.dz));
    memcpy(p.dy, g.dy, sizeof(p.dy));
    memcpy(p.dx, g.dx, sizeof(p.dx));
    p.bias = e.bias;
    p.rowvec = e.rowvec;
    p.has_res = e.residual ? 1 : 0;
    p.out_is_bf16 = e.out_bf16 ? 1 : 0;
    p.act = e.act;
    p.dbg = e.dbg;
    p.dbg_block = e.dbg_block;
    p.rowvec_ld = e.rowvec_ld ? e.rowvec_ld : N;
    op->block_n = bn;
    // short K: the kernel is epilogue/memory-bound -> 2 stages so that two CTAs fit on an SM
    // ... unless the whole grid is a single wave anyway, where deeper prefetch wins
    const bool multi_wave = (int64_t)m_tiles * (N / bn) > kNumSMs;
    op->stages = bn == 256 ? ((num_k <= 8 && multi_wave) ? 2 : 4) : (bn == 128 ? 3 : 4);
    // split-K = 2 for very long reductions (the level-1 Conv3d, 216 k-blocks): depends on the layer shape only, never
    // on the batch, so results stay batch-invariant; the two partial sums are combined in a fixed order.
    op->split_k = (e.split_flags && (num_k >= 128 || (e.force_split == 2 && num_k >= 8)) && bn == 256 && e.out_f32 &&
                   e.act == ACT_NONE && !want_ln) ? 2 : 1;
    if (want_ln && bn == 256) op->stages = 4;   // the LN pass needs every chunk resident in its own slab (192 KB of stages)
    p.ln_gamma = want_ln ? e.ln_gamma : nullptr;
    p.ln_cluster = want_ln ? N / bn : 1;
    op->cluster_y = p.ln_cluster;
    p.ln_beta = e.ln_beta;
    p.ln_eps = e.ln_eps;
    op->tmap_ln = op->tmap_out;
    if (want_ln) {
        cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)rows_per_sample, (cuuint64_t)g.samples};
        // bf16 operands: ln_out is bf16 [M][N]; tf32 operands: the same pointer holds fp32 (rounded to tf32), boxes of 32 columns
        const cuuint64_t lsz = g.tf32 ? 4 : 2;
        cuuint64_t strides[2] = {(cuuint64_t)N * lsz, (cuuint64_t)N * lsz * rows_per_sample};
        cuuint32_t box[3] = {g.tf32 ? 32u : 64u, 32, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = g_encode(&op->tmap_ln, g.tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                              e.ln_out, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PD_CHECK(r == CUDA_SUCCESS, PD_ERR_CUDA, "cuTensorMapEncodeTiled(ln) failed: %d", (int)r);
    }
    p.split_flags = e.split_flags;
    if (!g.b_sample_stride && (g.ldb == 0 || g.ldb == (int64_t)g.ntaps * g.C)) {   // a plain weight matrix [N][K]
        op->own_w.p[0] = reinterpret_cast<const uint8_t*>(Wt);
        op->own_w.n[0] = (uint32_t)((size_t)N * g.ntaps * g.C * esz);
    }
    p.gn_sums = e.gn_sums;
    p.gn_groups = e.gn_groups;
    p.gn_rows = e.gn_rows;
    p.gn_cpg = e.gn_sums ? N / e.gn_groups : 0;
    if (e.gn_sums) {
        PD_CHECK(e.out_f32 && gemm_gn_fusable(N, e.gn_groups, e.gn_rows) && bn % 32 == 0, PD_ERR_SHAPE,
                 "gemm: fused GroupNorm statistics need an fp32 output, 8/16/32 channels per group and rows %% 32 == 0");
        PD_CHECK(op->split_k == 1, PD_ERR_SHAPE, "gemm: fused GroupNorm statistics are not available with split-K");
        PD_CHECK(bn == 256, PD_ERR_SHAPE, "gemm: fused GroupNorm statistics need N %% 256 == 0 (BLOCK_N = 256)");
        op->stages = 4;
    }
    // multi-wave bf16-output GEMMs (QKV, FFN-1) take the persistent kernel: epilogue under the next tile's mainloop
    op->persistent = (e.out_bf16 && bn == 256 && (int64_t)m_tiles * (N / bn) > kNumSMs && getenv("PD_NO_PERSISTENT") == nullptr) ? 1 : 0;
    op->grid_x = (unsigned)m_tiles;
    op->grid_y = (unsigned)(N / bn);
    op->flops = 2.0 * (double)rows_per_sample * g.samples * (double)N * (double)g.ntaps * g.C;
    return PD_OK;
}

// (Re)encodes the output / residual tensor maps: [samples][rows][N] with row stride ldo, 32-row x 128-byte boxes.
int gemm_bind_output(GemmOp* op, float* out_f32, bf16* out_bf16, const float* residual) {
    PD_TRY(gemm_init());
    auto enc = [&](CUtensorMap* m, void* ptr, bool is_bf16) -> int {
        const int es_bytes = is_bf16 ? 2 : 4;
        PD_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, PD_ERR_ARG, "gemm: output must be 16-byte aligned");
        cuuint64_t dims[3] = {(cuuint64_t)op->out_N, (cuuint64_t)op->out_rows, (cuuint64_t)op->out_samples};
        cuuint64_t strides[2] = {(cuuint64_t)op->ldo * es_bytes, (cuuint64_t)op->ldo * es_bytes * op->out_rows};
        cuuint32_t box[3] = {(cuuint32_t)(128 / es_bytes), 32, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = g_encode(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr,
                              dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PD_CHECK(r == CUDA_SUCCESS, PD_ERR_CUDA, "cuTensorMapEncodeTiled(out) failed: %d", (int)r);
        return PD_OK;
    };
    if (out_bf16) PD_TRY(enc(&op->tmap_out, out_bf16, true));
    else PD_TRY(enc(&op->tmap_out, out_f32, false));
    if (residual) PD_TRY(enc(&op->tmap_res, const_cast<float*>(residual), false));
    else op->tmap_res = op->tmap_out;
    return PD_OK;
}

int gemm_launch(const GemmOp& op, cudaStream_t stream) {
    if (op.sk_segs) return gemm_streamk_launch(op, stream);
    if (op.persistent) return launch_persistent(op, stream);
    switch (op.block_n) {
        case 32: return launch_cfg<32, 4>(op, stream);
        case 64: return launch_cfg<64, 4>(op, stream);
        case 128: return launch_cfg<128, 3>(op, stream);
        case 256: return op.stages == 2 ? launch_cfg<256, 2>(op, stream) : launch_cfg<256, 4>(op, stream);
    }
    set_error("gemm_launch: op not initialised");
    return PD_ERR_STATE;
}

}  // namespace pd
