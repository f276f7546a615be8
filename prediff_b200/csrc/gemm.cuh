// tcgen05 / TMA implicit-GEMM: the one dense-contraction kernel of the sampling path.
//
//   out[row, n] = epilogue( sum_{tap, c} A[row shifted by tap, c] * Wt[n, tap, c] )
//
// A is a channels-last bf16 activation tensor viewed as [samples][D][H][W][C]; a "tap" is a (dz, dy, dx) shift
// (one tap with zero shift = a plain Linear; 27 taps = Conv3d 3x3x3; 9 taps = Conv2d 3x3). Shifted tiles are
// fetched with *tiled* TMA whose out-of-bounds zero fill implements the zero padding, so no im2col buffer exists.
// Replaces the cuDNN / cuBLAS call sites listed in SURVEY.md section 2.1 (reference:
// src/prediff/models/time_embed.py:93,120; cuboid_transformer.py:735,767,157,164; taming/resnet.py:405,421).
#pragma once
#include "common.cuh"
#include <vector>

namespace pd {

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row (tf32 operands: 32 fp32 = the same 128 bytes)
constexpr int kGemmBlockKTf32 = 32;
constexpr int kMaxTaps = 27;

enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2 };

struct GemmGeom {
    int samples = 1;  // independent samples (tiles never straddle a sample)
    int D = 1, H = 1, W = 1;
    int out_D = 0;    // D extent of the OUTPUT grid (0 -> D); differs from D only for the parity-plane stride-2 conv
    int C = 0;        // channels of A per tap (multiple of 64)
    int ntaps = 1;
    int8_t dz[kMaxTaps] = {0}, dy[kMaxTaps] = {0}, dx[kMaxTaps] = {0};
    // strides of A in elements, in case the activation buffer is a view (default: dense channels-last)
    int64_t sW = 0, sH = 0, sD = 0, sN = 0;
    // B operand: row stride in elements (0 -> ntaps*C) and, for a per-sample B (attention: K or V^T of that
    // sample instead of shared weights), the element stride between samples (0 -> shared)
    int64_t ldb = 0, b_sample_stride = 0;
    // Operand precision: 0 = bf16 operands (A and Wt are bf16), 1 = tf32 (A and Wt are fp32 holding tf32-rounded values;
    // tcgen05.mma kind::tf32 at half the bf16 rate, K block = 32 elements). C must be a multiple of 64 / 32.
    int tf32 = 0;
    int kblk() const { return tf32 ? kGemmBlockKTf32 : kGemmBlockK; }

    static GemmGeom linear(int M, int K) {
        GemmGeom g;
        g.W = M;
        g.C = K;
        return g;
    }
    // Conv with a (kt, kh, kw) window, stride 1, zero padding k/2 on every axis.
    static GemmGeom conv(int samples, int D, int H, int W, int C, int kt, int kh, int kw) {
        GemmGeom g;
        g.samples = samples; g.D = D; g.H = H; g.W = W; g.C = C;
        g.ntaps = kt * kh * kw;
        int i = 0;
        for (int a = 0; a < kt; ++a)
            for (int b = 0; b < kh; ++b)
                for (int c = 0; c < kw; ++c, ++i) {
                    g.dz[i] = (int8_t)(a - kt / 2);
                    g.dy[i] = (int8_t)(b - kh / 2);
                    g.dx[i] = (int8_t)(c - kw / 2);
                }
        return g;
    }
    // Stride-2 3x3 conv with right/bottom zero pad (taming/resnet.py:183-188) over parity planes
    // [F][4][Ho][Wo][C] (plane = (y%2)*2 + x%2): tap (kh, kw) reads plane (kh%2, kw%2) shifted by (kh/2, kw/2).
    static GemmGeom conv_s2_planes(int F, int Ho, int Wo, int C) {
        GemmGeom g;
        g.samples = F; g.D = 4; g.out_D = 1; g.H = Ho; g.W = Wo; g.C = C;
        g.ntaps = 9;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                const int i = kh * 3 + kw;
                g.dz[i] = (int8_t)((kh & 1) * 2 + (kw & 1));
                g.dy[i] = (int8_t)(kh >> 1);
                g.dx[i] = (int8_t)(kw >> 1);
            }
        return g;
    }
};

struct GemmEpilogue {
    const float* bias = nullptr;      // [N]
    const float* rowvec = nullptr;    // [samples][rowvec_ld], added to every row of the sample (time embedding)
    int rowvec_ld = 0;                // row stride of rowvec in elements (0 -> N)
    const float* residual = nullptr;  // [M][ldo] fp32, may alias out_f32
    float* out_f32 = nullptr;         // [M][ldo]   (exactly one of out_f32 / out_bf16)
    bf16* out_bf16 = nullptr;         // [M][ldo]   (needs N % 64 == 0, no residual)
    int ldo = 0;                      // row stride of residual/out in elements (0 -> N)
    int act = ACT_NONE;               // applied after bias/rowvec, before the residual add
    int round_tf32 = 0;               // fp32 output that feeds a tf32 GEMM: store round-to-nearest tf32 values
    // Optional fused LayerNorm of the OUTPUT rows (needs out_f32; N == 256: one CTA owns whole rows; N == 512: the two
    // CTAs of a row form a thread-block cluster and exchange their row sums through distributed shared memory):
    // ln_out[M][N] bf16 = LN(out row) * ln_gamma + ln_beta - the next layer's pre-norm, saving its kernel + a pass.
    const float* ln_gamma = nullptr;
    const float* ln_beta = nullptr;
    bf16* ln_out = nullptr;
    float ln_eps = 1e-5f;
    // Optional GroupNorm statistics of the OUTPUT (fp32 output, no split-K): (sum, sum of squares) per (sample, group)
    // added to gn_sums[S][gn_groups][2] (double, zeroed by the caller) - the table gn_stats() would fill. gn_rows = rows
    // per GroupNorm sample (multiple of 32); N / gn_groups must be 8, 16 or 32.
    double* gn_sums = nullptr;
    int gn_groups = 0, gn_rows = 0;
    int* split_flags = nullptr;       // optional zeroed [m_tiles * n_tiles] ints: enables split-K (see gemm_make)
    int force_split = 0;              // 2: split K in two even below the automatic threshold (few tiles, long K);
                                      // must depend on the layer shape only so results stay batch-invariant
    unsigned long long* dbg = nullptr;  // optional phase-timestamp buffer (9 x u64), see GemmKernelParams::dbg
    int dbg_block = 0;
};

struct GemmKernelParams {
    int rows_per_sample, tiles_per_sample, samples;
    int N, H, W, HW;
    int ntaps, cblks, b_batched;
    int tf32, kblk;            // operand precision (GemmGeom::tf32) and elements per 128-byte k-block (64 bf16 / 32 tf32)
    int round_out;             // GemmEpilogue::round_tf32
    int8_t dz[kMaxTaps], dy[kMaxTaps], dx[kMaxTaps];
    const float* bias;
    const float* rowvec;
    int act, rowvec_ld;
    int has_res, out_is_bf16;  // residual / output live in the tmap_res / tmap_out tensor maps
    const float* ln_gamma;     // fused output LayerNorm (null = off); result goes through tmap_ln
    const float* ln_beta;
    float ln_eps;
    int ln_cluster;            // CTAs along N that share a row (cluster size): 1, or 2 when the fused LN spans N = 512
    double* gn_sums;           // fused GroupNorm statistics of the output (null = off), see GemmEpilogue::gn_sums
    int gn_cpg, gn_groups, gn_rows;
    int* split_flags;          // per-tile handshake between the two split-K CTAs (self re-arming)
    unsigned long long* dbg;   // optional: 9 clock64() phase stamps of CTA (dbg_block, 0) - tools/gemm_phases.py
    int dbg_block;
    WRange pf;                 // weights of the next GEMM of the plan, requested into L2 at kernel start (common.cuh)
};

// Stream-K schedule entry (gemm_streamk.cu): one contiguous k-block range of one output tile.
enum SkRole : int { SK_NONE = 0, SK_DUMP = 1, SK_FINAL = 2 };
struct SkSeg {
    int m_tile, n_tile;   // output tile (m_tile counts over all samples)
    int k_begin, k_end;   // k-block range
    int role;             // SK_DUMP: write the raw partial to workspace slot `slot`, bump flag; SK_FINAL: epilogue
    int slot;             // DUMP: slot written; FINAL: first of the n_part consecutive slots to add
    int n_part;           // FINAL: number of partials of this tile
    int flag;             // index of the tile's arrival counter
};

struct GemmOp {
    CUtensorMap tmap_a, tmap_b, tmap_out, tmap_res, tmap_ln;
    GemmKernelParams p;
    int ldo = 0, out_rows = 0, out_samples = 0, out_N = 0;
    int block_n = 0, stages = 0, split_k = 1, persistent = 0, cluster_y = 1;
    const SkSeg* sk_segs = nullptr;   // non-null: launch through the stream-K kernel with sk_ctas CTAs
    int sk_ctas = 0;
    float* sk_partials = nullptr;
    int* sk_flags = nullptr;
    unsigned grid_x = 0, grid_y = 0;
    size_t smem = 0;
    double flops = 0;
    WRange own_w = {};   // this op's weight bytes (what the preceding GEMM should prefetch)
};

// Builds tensor maps + launch geometry. `A` is bf16 (fp32 if g.tf32) [samples][D][H][W][C]; `Wt` is bf16 (fp32 if g.tf32)
// [N][ntaps*C] (K-major).
int gemm_make(GemmOp* op, const void* A, const GemmGeom& g, const void* Wt, int N, const GemmEpilogue& e,
              int force_block_n = 0);
int gemm_launch(const GemmOp& op, cudaStream_t stream);
void gemm_set_tile_preference(int wide);   // see gemm.cu
// Points an already-built op at new output / residual buffers of the same shape (per-call user pointers).
int gemm_bind_output(GemmOp* op, float* out_f32, bf16* out_bf16, const float* residual);
// Number of ints gemm_make may need in GemmEpilogue::split_flags for this geometry (0 if it will not split).
int gemm_split_flags_needed(const GemmGeom& g, int N, int force_split = 0);
// True if GemmEpilogue::gn_sums can be honoured for this shape (channels per group 8 / 16 / 32, rows per sample % 32 == 0).
inline bool gemm_gn_fusable(int N, int groups, int rows) {
    if (groups <= 0 || N % groups != 0 || rows % 32 != 0 || N % 256 != 0) return false;
    const int cpg = N / groups;
    return cpg == 8 || cpg == 16 || cpg == 32;
}
// Stream-K for long-K fp32-output convolutions (gemm_streamk.cu). schedule(): host-side cut of an op built by gemm_make
// into 2 * ctas_per_sample * samples segments (+ the number of 128 KB partial slots / flags it needs); attach(): points
// the op at the device copy of the schedule and at the workspace, after which gemm_launch uses the stream-K kernel.
int gemm_streamk_schedule(const GemmOp& op, int ctas_per_sample, std::vector<SkSeg>* segs, int* n_slots, int* n_flags);
int gemm_streamk_attach(GemmOp* op, const SkSeg* segs_dev, int n_ctas, float* partials, int* flags);
int gemm_streamk_launch(const GemmOp& op, cudaStream_t stream);
int tmap_encode_sw128(CUtensorMap* m, bool is_bf16, int rank, const void* ptr, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box);
int gemm_init();  // resolves the driver entry point + raises the dynamic smem limits (idempotent)

}  // namespace pd
