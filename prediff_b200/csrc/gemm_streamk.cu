// Stream-K schedule for the long-K implicit-GEMM convolutions (Conv3d 3x3x3: 108 / 216 k-blocks per 128 x 256 tile).
// The plain kernel gives every tile to one CTA (104 tiles on 148 SMs at level 0) or to two with serialised epilogues
// (level 1); the mainloop already runs at the tcgen05 floor, so the only thing left is SM fill. Here the (tile, k-block)
// units of a sample are cut into equal contiguous ranges, one per CTA (36 CTAs per sample -> 144 at batch 4). A range
// spans at most two tiles: the tail of one tile, then the head of the next.
//   tail / middle part of a tile ("dump")  : raw fp32 accumulator -> global workspace with coalesced 16-byte stores
//                                            straight from the TMEM registers (no shared memory: it runs under the
//                                            CTA's next mainloop, which accumulates into the second TMEM buffer), then
//                                            a release increment of the tile's flag;
//   head part of a tile ("final")          : always the CTA's last segment, finishing when the dumpers are long done:
//                                            accumulator + partials (fixed order) + bias / time embedding + residual,
//                                            TMA store, optional fused LayerNorm - the usual epilogue.
// Results are deterministic (fixed summation order) and batch-invariant (the cut depends on the layer shape only).
#include "gemm.cuh"
#include "ptx.cuh"
#include <type_traits>

namespace pd {
namespace {

constexpr int BN = 256, STAGES = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;
constexpr int kBBytes = BN * kGemmBlockK * 2;
constexpr int kStageBytes = kABytes + kBBytes;          // 48 KB
constexpr int kPipeBytes = STAGES * kStageBytes;        // 192 KB
constexpr int kBarBytes = 512;
constexpr int kSmem = kPipeBytes + 1024 + kBarBytes + BN * 4 + 2 * BN * 4 + kEpiWarps * 32 * 8;

template <bool GN>   // GN: also accumulate GroupNorm statistics of the output (GemmEpilogue::gn_sums)
__global__ void __launch_bounds__(kThreads, 1)
conv_streamk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                    const __grid_constant__ CUtensorMap tmap_ln, const __grid_constant__ GemmKernelParams p,
                    const SkSeg* __restrict__ segs, float* __restrict__ partials, int* __restrict__ flags) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kPipeBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;        // [2]: one per segment / accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 2);
    uint64_t* res_bar = reinterpret_cast<uint64_t*>(smem + kPipeBytes + 128);   // [kEpiWarps][4]
    float* vec_s = reinterpret_cast<float*>(smem + kPipeBytes + kBarBytes);
    float* ln_g = vec_s + BN;
    float* ln_b = ln_g + BN;
    float2* ln_x = reinterpret_cast<float2*>(ln_b + BN);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // A tile's finishing CTA waits for CTAs holding later k ranges (higher schedule index): give those the lower block
    // indices, which are dispatched first, so a launch larger than one wave cannot park its waiters on every SM.
    const int cta = gridDim.x - 1 - blockIdx.x;
    if (threadIdx.x == 32) prefetch_l2_share(p.pf, blockIdx.x, gridDim.x);   // next GEMM's weights -> L2 (common.cuh)
    const SkSeg seg0 = segs[2 * cta], seg1 = segs[2 * cta + 1];
    // phase stamps of schedule CTA p.dbg_block (tools/streamk_phases.py): [0] entry, [1] setup done, [2] first operand
    // stage landed, [3]/[6] last MMA of segment 0 / 1 issued, [4]/[7] accumulator 0 / 1 complete, [5] partial dumped,
    // [8] partials of the tile arrived, [9] epilogue done, [10] exit, [11]/[12] %globaltimer at entry / exit
    unsigned long long* dbg = (p.dbg && cta == p.dbg_block) ? p.dbg : nullptr;
#define SK_STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
    if (dbg && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg[11] = gt;
        dbg[0] = clock64();
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(&tmem_full[0], 1);
        ptx::mbar_init(&tmem_full[1], 1);
        for (int i = 0; i < 4 * kEpiWarps; ++i) ptx::mbar_init(&res_bar[i], 1);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        ptx::prefetch_tmap(&tmap_out);
        if (p.has_res) ptx::prefetch_tmap(&tmap_res);
        if (p.ln_gamma) ptx::prefetch_tmap(&tmap_ln);
    }
    int b_pre = 0;   // weight tiles of the first stages, requested before the dependency wait (common.cuh)
    if (threadIdx.x == 0 && seg0.role != SK_NONE) {
        b_pre = seg0.k_end - seg0.k_begin < STAGES ? seg0.k_end - seg0.k_begin : STAGES;
        for (int it = 0; it < b_pre; ++it) {
            ptx::mbar_arrive_expect_tx(&full_bar[it], kStageBytes);
            ptx::tma_load_3d(smem + it * kStageBytes + kABytes, &tmap_b, &full_bar[it], (seg0.k_begin + it) * p.kblk,
                             seg0.n_tile * BN, 0);
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
    if (threadIdx.x == 0) SK_STAMP(1);

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
#pragma unroll 1
            for (int sg = 0; sg < 2; ++sg) {
                const SkSeg sgm = sg ? seg1 : seg0;
                if (sgm.role == SK_NONE) continue;
                const int sample = sgm.m_tile / p.tiles_per_sample;
                const int p0 = (sgm.m_tile - sample * p.tiles_per_sample) * kGemmBlockM;
                const int z0 = p0 / p.HW;
                const int rem = p0 - z0 * p.HW;
                const int y0 = rem / p.W;
                const int x0 = rem - y0 * p.W;
                const int n0 = sgm.n_tile * BN;
                int tap = sgm.k_begin / p.cblks, cb = sgm.k_begin - tap * p.cblks;
                for (int k = sgm.k_begin; k < sgm.k_end; ++k, ++it) {
                    const int s = it % STAGES;
                    uint8_t* sa = smem + s * kStageBytes;
                    if (it >= b_pre) {
                        ptx::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                        ptx::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                    }
                    ptx::tma_load_5d(sa, &tmap_a, &full_bar[s], cb * p.kblk, x0 + p.dx[tap], y0 + p.dy[tap],
                                     z0 + p.dz[tap], sample);
                    if (it >= b_pre) ptx::tma_load_3d(sa + kABytes, &tmap_b, &full_bar[s], k * p.kblk, n0, 0);
                    if (++cb == p.cblks) { cb = 0; ++tap; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            auto issue = [&](auto tf32_tag) {   // the precision picks the loop, not the instruction (ptx.cuh umma_ss)
                constexpr bool TF32 = decltype(tf32_tag)::value;
                constexpr uint32_t idesc = TF32 ? ptx::make_idesc_tf32(kGemmBlockM, BN) : ptx::make_idesc_bf16(kGemmBlockM, BN);
                int it = 0;
#pragma unroll 1
                for (int sg = 0; sg < 2; ++sg) {
                    const SkSeg sgm = sg ? seg1 : seg0;
                    if (sgm.role == SK_NONE) continue;
                    const uint32_t tmem_d = tmem_base + sg * BN;   // the two segments use different accumulators
                    for (int k = sgm.k_begin; k < sgm.k_end; ++k, ++it) {
                        const int s = it % STAGES;
                        ptx::mbar_wait(&full_bar[s], (it / STAGES) & 1);
                        ptx::tc_fence_after();
                        if (it == 0) SK_STAMP(2);
                        const uint32_t a_addr = ptx::smem_u32(smem + s * kStageBytes);
                        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                        for (int kk = 0; kk < kGemmBlockK / 16; ++kk)
                            ptx::umma_ss<TF32>(tmem_d, ptx::make_smem_desc_sw128(a_addr + kk * 32),
                                               ptx::make_smem_desc_sw128(b_addr + kk * 32), idesc,
                                               (k != sgm.k_begin || kk != 0) ? 1u : 0u);
                        ptx::umma_commit(&empty_bar[s]);
                    }
                    ptx::umma_commit(&tmem_full[sg]);
                    SK_STAMP(sg ? 6 : 3);
                }
            };
            if (p.tf32) issue(std::true_type{});
            else issue(std::false_type{});
        }
    } else {
        const int e = warp - 2;
        const int q = warp & 3;
        const int half = e >> 2;
        const int et = threadIdx.x - 64;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        const int c_begin = half * 4;                  // this warp's four 32-column chunks
        const int row = q * 32 + lane;                 // row inside the tile
#pragma unroll 1
        for (int sg = 0; sg < 2; ++sg) {
            const SkSeg sgm = sg ? seg1 : seg0;
            if (sgm.role == SK_NONE) continue;
            const uint32_t t_lane = tmem_base + sg * BN + (static_cast<uint32_t>(q * 32) << 16);
            if (sgm.role == SK_DUMP) {
                // ---- partial tile -> workspace [chunk][cell][row][4 floats]: consecutive lanes, consecutive 16 B ----
                ptx::mbar_wait(&tmem_full[sg], 0);
                ptx::tc_fence_after();
                if (et == 0) SK_STAMP(sg ? 7 : 4);
                float4* dst = reinterpret_cast<float4*>(partials + (size_t)sgm.slot * (kGemmBlockM * BN));
#pragma unroll 1
                for (int idx = 0; idx < 4; ++idx) {
                    const int c = c_begin + idx;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_lane + c * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[(size_t)(c * 8 + i) * kGemmBlockM + row] =
                            make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                        __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                }
                // the CTA barrier orders every warp's stores before thread 0, whose gpu-scope fence then publishes them
                // (cumulativity) ahead of the flag increment - one fence per CTA instead of one per thread
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                if (et == 0) {
                    __threadfence();
                    atomicAdd(flags + sgm.flag, 1);
                    SK_STAMP(5);
                }
                continue;
            }
            // ---- final segment: the usual epilogue, plus the partial sums of the tile's other parts ----
            const int sample = sgm.m_tile / p.tiles_per_sample;
            const int p0 = (sgm.m_tile - sample * p.tiles_per_sample) * kGemmBlockM;
            const int n0 = sgm.n_tile * BN;
            const int row0 = p0 + q * 32;
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
            ptx::mbar_wait(&tmem_full[sg], 0);   // last segment: every mainloop of this CTA is finished, the ring is idle
            ptx::tc_fence_after();
            if (et == 0) SK_STAMP(sg ? 7 : 4);
            uint8_t* slabs = smem + e * (4 * 4096);      // aliases the (finished) pipeline stages
            uint64_t* my_bar = res_bar + 4 * e;
            const bool has_res = p.has_res != 0;
            if (has_res && lane == 0) {
                for (int idx = 0; idx < 4; ++idx) {
                    ptx::mbar_arrive_expect_tx(&my_bar[idx], 4096);
                    ptx::tma_load_3d(slabs + idx * 4096, &tmap_res, &my_bar[idx], n0 + (c_begin + idx) * 32, row0, sample);
                }
            }
            if (sgm.n_part > 0) {   // the other parts of this tile were dumped long ago (they are first segments)
                if (et == 0)
                    while (ptx::ld_acquire_gpu(flags + sgm.flag) < sgm.n_part) {}
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            }
            if (et == 0) SK_STAMP(8);
            float ln_s1 = 0.f, ln_s2 = 0.f;
            const int gsample = GN ? (sample * p.rows_per_sample + row0) / p.gn_rows : 0;   // GroupNorm sample of my rows
#pragma unroll 1
            for (int idx = 0; idx < 4; ++idx) {
                const int c = c_begin + idx;
                uint8_t* slab = slabs + idx * 4096;
                uint32_t v[32];
                ptx::tmem_ld_32x32(t_lane + c * 32, v);
                float4 part[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) part[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int pp = 0; pp < sgm.n_part; ++pp) {   // fixed order: increasing k range
                    const float4* src = reinterpret_cast<const float4*>(partials + (size_t)(sgm.slot + pp) * (kGemmBlockM * BN));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 t = __ldcg(src + (size_t)(c * 8 + i) * kGemmBlockM + row);
                        part[i].x += t.x; part[i].y += t.y; part[i].z += t.z; part[i].w += t.w;
                    }
                }
                if (has_res) ptx::mbar_wait(&my_bar[idx], 0);
                ptx::tmem_ld_wait();
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
                    float4 a = make_float4((__uint_as_float(v[4 * i]) + part[i].x) + b.x,
                                           (__uint_as_float(v[4 * i + 1]) + part[i].y) + b.y,
                                           (__uint_as_float(v[4 * i + 2]) + part[i].z) + b.z,
                                           (__uint_as_float(v[4 * i + 3]) + part[i].w) + b.w);
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
            if (p.ln_gamma) {   // fused LayerNorm of the finished rows (N == 256: this CTA owns whole rows)
                ln_x[(q * 2 + half) * 32 + lane] = make_float2(ln_s1, ln_s2);
                asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
                const float2 o = ln_x[(q * 2 + (half ^ 1)) * 32 + lane];
                const float mean = (ln_s1 + o.x) * (1.0f / BN);
                const float var = fmaxf((ln_s2 + o.y) * (1.0f / BN) - mean * mean, 0.f);
                const float rstd = rsqrtf(var + p.ln_eps);
                if (p.tf32) {   // fp32 (tf32-rounded) operand for the next GEMM, formed in place in the fp32 slabs (see gemm.cu)
                    if (lane == 0) ptx::bulk_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int idx = 0; idx < 4; ++idx) {
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
                        for (int idx = 0; idx < 4; ++idx)
                            ptx::tma_store_3d(&tmap_ln, slabs + idx * 4096, n0 + (c_begin + idx) * 32, row0, sample);
                        ptx::bulk_commit();
                    }
                } else {
                uint8_t* bslabs = smem + kEpiWarps * 4 * 4096 + e * (2 * 4096);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint8_t* brow = bslabs + j * 4096 + lane * 128;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const uint8_t* frow = slabs + (2 * j + cc) * 4096 + lane * 128;
                        const int colbase = (c_begin + 2 * j + cc) * 32;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
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
                    ptx::tma_store_3d(&tmap_ln, bslabs, n0 + (c_begin + 0) * 32, row0, sample);
                    ptx::tma_store_3d(&tmap_ln, bslabs + 4096, n0 + (c_begin + 2) * 32, row0, sample);
                    ptx::bulk_commit();
                }
                }
            }
            if (sgm.n_part > 0) {   // re-arm the tile's flag for the next launch
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                if (et == 0) flags[sgm.flag] = 0;
            }
            if (lane == 0) ptx::bulk_wait_read<0>();
            __syncwarp();
            if (et == 0) SK_STAMP(9);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (dbg && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg[12] = gt;
        dbg[10] = clock64();
    }
#undef SK_STAMP
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Cuts the (tile, k-block) units of each sample into `ctas_per_sample` equal contiguous ranges. Returns PD_ERR_SHAPE if
// the op is not eligible (needs BN = 256, fp32 output, no activation, no split-K, ranges of at most two tiles).
int gemm_streamk_schedule(const GemmOp& op, int ctas_per_sample, std::vector<SkSeg>* segs, int* n_slots, int* n_flags) {
    const GemmKernelParams& p = op.p;
    const int num_k = p.ntaps * p.cblks;
    const int n_tiles = (int)op.grid_y;
    PD_CHECK(op.block_n == 256 && !p.out_is_bf16 && p.act == ACT_NONE && !op.persistent && op.cluster_y == 1 && !p.b_batched,
             PD_ERR_SHAPE, "stream-K: op not eligible");
    const int tiles_ps = p.tiles_per_sample * n_tiles;
    const long long U = (long long)tiles_ps * num_k;
    PD_CHECK(ctas_per_sample >= 1 && U / ctas_per_sample <= num_k && U / ctas_per_sample >= 8, PD_ERR_SHAPE,
             "stream-K: %d CTAs per sample do not fit %lld units of %d k-blocks", ctas_per_sample, U, num_k);
    const int total_ctas = ctas_per_sample * p.samples;
    segs->assign((size_t)2 * total_ctas, SkSeg{0, 0, 0, 0, SK_NONE, 0, 0, 0});
    struct Part { int cta, sg, k0; };
    std::vector<std::vector<Part>> per_tile((size_t)tiles_ps * p.samples);
    for (int s = 0; s < p.samples; ++s)
        for (int j = 0; j < ctas_per_sample; ++j) {
            const long long a = U * j / ctas_per_sample, b = U * (j + 1) / ctas_per_sample;
            const int cta = s * ctas_per_sample + j;
            int sg = 0;
            for (long long u = a; u < b;) {
                const int tl = (int)(u / num_k);
                const int k0 = (int)(u - (long long)tl * num_k);
                const long long tile_end = (long long)(tl + 1) * num_k;
                const long long e = b < tile_end ? b : tile_end;
                PD_CHECK(sg < 2, PD_ERR_SHAPE, "stream-K: a range spans more than two tiles");
                SkSeg& g = (*segs)[(size_t)2 * cta + sg];
                g.m_tile = s * p.tiles_per_sample + tl / n_tiles;
                g.n_tile = tl % n_tiles;
                g.k_begin = k0;
                g.k_end = (int)(e - (long long)tl * num_k);
                g.role = k0 == 0 ? SK_FINAL : SK_DUMP;
                g.flag = s * tiles_ps + tl;
                per_tile[(size_t)s * tiles_ps + tl].push_back(Part{cta, sg, k0});
                ++sg;
                u = e;
            }
        }
    int slots = 0;
    for (auto& parts : per_tile) {
        // parts were appended in increasing k order; the first one starts at k = 0 and finalises
        PD_CHECK(!parts.empty() && parts[0].k0 == 0, PD_ERR_STATE, "stream-K: tile without a head segment");
        SkSeg& fin = (*segs)[(size_t)2 * parts[0].cta + parts[0].sg];
        fin.slot = slots;
        fin.n_part = (int)parts.size() - 1;
        for (size_t i = 1; i < parts.size(); ++i) (*segs)[(size_t)2 * parts[i].cta + parts[i].sg].slot = slots++;
    }
    // a final segment must be the last segment of its CTA (its epilogue reuses the pipeline's shared memory)
    for (int c = 0; c < total_ctas; ++c) {
        const SkSeg& a = (*segs)[(size_t)2 * c];
        const SkSeg& b = (*segs)[(size_t)2 * c + 1];
        PD_CHECK(!(a.role == SK_FINAL && b.role != SK_NONE), PD_ERR_SHAPE, "stream-K: final segment followed by another");
    }
    *n_slots = slots;
    *n_flags = tiles_ps * p.samples;
    return PD_OK;
}

int gemm_streamk_attach(GemmOp* op, const SkSeg* segs_dev, int n_ctas, float* partials, int* flags) {
    static bool attr_set = false;
    if (!attr_set) {
        PD_CUDA(cudaFuncSetAttribute(conv_streamk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        PD_CUDA(cudaFuncSetAttribute(conv_streamk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set = true;
    }
    op->sk_segs = segs_dev;
    op->sk_ctas = n_ctas;
    op->sk_partials = partials;
    op->sk_flags = flags;
    op->split_k = 1;
    return PD_OK;
}

int gemm_streamk_launch(const GemmOp& op, cudaStream_t stream) {
    if (op.p.gn_sums)
        PD_LAUNCH(conv_streamk_kernel<true>, op.sk_ctas, kThreads, kSmem, stream, op.tmap_a, op.tmap_b, op.tmap_out,
                  op.tmap_res, op.tmap_ln, op.p, op.sk_segs, op.sk_partials, op.sk_flags);
    else
        PD_LAUNCH(conv_streamk_kernel<false>, op.sk_ctas, kThreads, kSmem, stream, op.tmap_a, op.tmap_b, op.tmap_out,
                  op.tmap_res, op.tmap_ln, op.p, op.sk_segs, op.sk_partials, op.sk_flags);
    return PD_OK;
}

}  // namespace pd
