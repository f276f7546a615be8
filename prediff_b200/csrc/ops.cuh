// Launchers for the non-GEMM kernels of the sampling path. All tensors are channels-last; fp32 is the residual
// stream, bf16 is what feeds the tensor-core GEMM. Every launcher returns a pd error code.
#pragma once
#include "common.cuh"
#include <vector>

namespace pd {

// ---- normalisation (norm.cu) -------------------------------------------------------------------------------
// GroupNorm statistics: x fp32 [S][R][C], groups of cpg contiguous channels. Accumulates (sum, sumsq) as doubles
// into sums[S][G][2]; the caller zeroes `sums` beforehand (one memset per forward covers every GN of the model).
// Reference: torch.nn.GroupNorm in time_embed.py:90-92,116-118 and taming/resnet.py:403,419.
int gn_stats(const float* x, double* sums, int S, int R, int C, int G, cudaStream_t st);
// y = act((x - mean) * rstd * gamma + beta) -> bf16 [S][R][C]; mean/rstd derived from sums (biased variance).
// y_f32 (here and below): the operand is written as tf32-rounded fp32 instead of bf16 (PD_PRECISION_TF32, common.cuh)
int gn_apply(const float* x, const double* sums, const float* gamma, const float* beta, void* y, int S, int R, int C,
             int G, float eps, int silu, cudaStream_t st, int y_f32 = 0);
// LayerNorm over the last dim (eps 1e-5): fp32 [P][C] -> bf16 [P][C]. C in {64,128,256,512,1024,2048}.
// Reference: models/utils.py:218 (nn.LayerNorm) as used by cuboid_transformer.py:813,195.
int layer_norm(const float* x, const float* gamma, const float* beta, void* y, int P, int C, float eps,
               cudaStream_t st, int y_f32 = 0);
// PatchMerging3D gather + LayerNorm(4C): x fp32 [B*T][H][W][C] -> bf16 [B*T][H/2][W/2][4C], merged channel order
// (dh, dw, c). Reference: cuboid_transformer.py:286-294.
int patch_merge_ln(const float* x, const float* gamma, const float* beta, void* y, int BT, int H, int W, int C,
                   float eps, cudaStream_t st, int y_f32 = 0);

// ---- attention (attention.cu) ------------------------------------------------------------------------------
// Axial cuboid self-attention core: qkv bf16 [B][T][H][W][3C] (q|k|v, head-major), one softmax per line along
// `axis` (0=T, 1=H, 2=W) and head, relative-position bias table fp32 [2L-1][heads]; out bf16 [B][T][H][W][C].
// Reference: cuboid_transformer.py:849-861,949 with cuboids (T,1,1)/(1,H,1)/(1,1,W) (patterns.py:34-36).
// f32 = 1: qkv and out are fp32 (out tf32-rounded) and the whole core runs in fp32 on the CUDA cores.
// The global vectors as extra keys of the token grid's queries (cuboid_transformer.py:902-913: unmasked, no position bias).
// Shared global_qkv net: k / v point into the rows of one [B][n][3C] tensor (ld = 3C, offsets C and 2C). separate_global_qkv
// (:866-891): k / v rows of l2g_global_kv_net and q2 = the tokens' l2g_q_net rows, which meet these keys instead of q.
struct GvKeys {
    const bf16* k = nullptr;    // [B][n][ld]: keys (head-major channels)
    const bf16* v = nullptr;    // [B][n][ld]: values
    int ld = 0;                 // elements between consecutive global rows
    int n = 0;                  // number of global vectors (0 = none)
    const bf16* q2 = nullptr;   // [B][N][q2_ld]: the tokens' queries for the global keys (null: the layer's own q)
    int q2_ld = 0;
};
// gk (bf16 mode only, gk->n <= 16): every query of a line also attends to the sample's global keys.
int axial_attention(const void* qkv, const float* bias_table, void* out, int B, int T, int H, int W, int C, int heads,
                    int axis, cudaStream_t st, int f32 = 0, const GvKeys* gk = nullptr);
// General cuboid self-attention core (any cuboid size, 'l' / 'd' strategy, shifted windows, end padding with
// padding_type 'zeros' (0) or 'ignore' (1)): cuboid_transformer.py:812-966 without global vectors.
struct CuboidLayerSpec {   // constructor arguments of one CuboidSelfAttentionLayer
    int size[3];           // cuboid_size (bT, bH, bW)
    int strategy[3];       // 0 = 'l' (local), 1 = 'd' (dilated)
    int shift[3];          // shift_size
};
struct CuboidTables {      // host tables of one layer on a (T, H, W) grid
    int size[3], shift[3], pad[3];                 // effective size / shift (:563-592), end padding per axis
    int num_cuboids = 0, volume = 0, rel_off = 0, n_rel = 0;
    int axial_axis = -1;                           // >= 0: the layer is exactly axial_attention() along that axis
    std::vector<int> tok;  // [num_cuboids * volume] token row inside a sample's [T*H*W] block, -1 = padding slot
    std::vector<int> lab;  // [num_cuboids * volume] shifted-window region label, -1 = masked out ('ignore' padding)
    std::vector<int> rel;  // [volume] relative_position_index[i][j] == rel[i] - rel[j] + rel_off
    // padding_type 'nearest' only (empty otherwise): the slot's result is written to token dst (-1 = to nobody); `tok` is
    // then the token the slot's q|k|v rows are COPIES of (the reference resamples the grid with F.interpolate before the
    // attention and back after it, models/utils.py:228-270)
    std::vector<int> dst;
    // padding_type 'ignore' only (empty otherwise): [num_cuboids * volume] 1 = the slot is visible to the global vectors'
    // queries (cuboid_transformer.py:915-924; raster-order validity of the padded, rolled frame, see build_cuboid_tables)
    std::vector<int> gmask;
};
int build_cuboid_tables(int T, int H, int W, const CuboidLayerSpec& spec, int padding_type, CuboidTables* out);
struct CuboidDev {         // device copies
    const int *tok = nullptr, *lab = nullptr, *rel = nullptr;
    const int* dst = nullptr;                                  // null: results go to `tok` (every padding type but 'nearest')
    const int* gmask = nullptr;                                // null: every slot is visible to the global queries
    int num_cuboids = 0, volume = 0, rel_off = 0, n_rel = 0;   // n_rel: rows of the bias table
};
struct CuboidTablesDev {   // owner of the device copies
    void* mem = nullptr;
    CuboidDev dev;
    CuboidTablesDev() = default;
    CuboidTablesDev(const CuboidTablesDev&) = delete;
    CuboidTablesDev& operator=(const CuboidTablesDev&) = delete;
    ~CuboidTablesDev();
    int upload(const CuboidTables& t);
};
// qkv bf16 [B][N][3C] (N = T*H*W tokens per sample, q|k|v head-major), bias_table fp32 [n_rel][heads] -> out bf16 [B][N][C]
// impl: 0 = choose (tcgen05 tile kernel when eligible, see below), 1 = warp-level mma.sync kernel, 2 = tcgen05 tile kernel
// gk (gk->n <= 64): every query also attends to the sample's global keys (GvKeys above); mma.sync kernel only
int cuboid_attention(const bf16* qkv, const float* bias_table, bf16* out, int B, int N, int C, int heads,
                     const CuboidDev& g, cudaStream_t st, int impl = 0, const GvKeys* gk = nullptr);
// The same contract on tcgen05 tensor-core tiles (attention_tc.cu): 128-query tile per (cuboid, head, sample), S and O in
// TMEM, K / V chunks of 128 keys in a swizzled shared-memory ring. Eligible for head dims 64 / 128 and volumes >= 128
// (cuboid_attention() dispatches to it; PD_CUBOID_NO_TC=1 keeps the mma.sync kernel for A/B runs).
bool cuboid_attention_tc_eligible(int hd, int volume);
int cuboid_attention_tc(const bf16* qkv, const float* bias_table, bf16* out, int B, int N, int C, int heads,
                        const CuboidDev& g, cudaStream_t st);
// ---- global vectors (global_vec.cu; cuboid_transformer.py:864-945, 1130-1145) ------------------------------------------
// out[r][n] = (res ? res[r][n] : 0) + act( sum_k f(in[r][k]) W[n][k] + bias[n] ) on M = B * K rows, fp32 on the fp32 weights
// as loaded; f = LayerNorm(ln_gamma, ln_beta, eps 1e-5) when ln_gamma is given (K <= 512), act 1 = GELU (erf). res may alias
// out_f32. out_f32 / out_bf16: either or both.
int gv_linear(const float* in, const float* ln_gamma, const float* ln_beta, const float* W, const float* bias, const float* res,
              float* out_f32, bf16* out_bf16, int M, int K, int N, int act, cudaStream_t st);
// g[b] = init for every sample (init_global_vectors.expand, cuboid_transformer_unet.py:432-434)
int gv_broadcast(const float* init, float* g, int B, int K, int C, cudaStream_t st);
// The global queries' attention (:928-945): gqkv fp32 [B][K][3C] (q part used, scaled by hd^-0.5 inside), local q|k|v bf16
// [B][N][3C] gathered through the layer's slot table (padding slots: zero rows, hidden under 'ignore' by g.gmask), and with
// self_attn the global keys / values gkv bf16 [B][K][3C] appended -> out fp32 [B][K][C] (before global_proj).
// workspace: global_attention_workspace_floats(...) floats. K <= 32.
size_t global_attention_workspace_floats(int B, int heads, int K, int hd, int n_keys);
// Operands of the global queries' attention. Shared net: q = sq = the fp32 q|k|v rows (ld 3C), tok_kv = the layer's q|k|v,
// sk / sv = the bf16 copy at offsets C / 2C. separate_global_qkv: q = g2l_global_q rows, tok_kv = the tokens' l2g_q | g2l_k |
// g2l_v rows (k at +C, v at +2C like q|k|v), sq / sk / sv = the g2g_global_qkv rows.
struct GvQuery {
    const float* q = nullptr;       // [B][K][q_ld] fp32: queries for the token keys (scaled by hd^-0.5 inside)
    int q_ld = 0;
    const bf16* tok_kv = nullptr;   // [B][N][3C] bf16: the tokens' rows, keys at +C, values at +2C
    const float* sq = nullptr;      // [B][K][sq_ld] fp32: queries for the global keys (self-attention; null: none)
    int sq_ld = 0;
    const bf16* sk = nullptr;       // [B][K][s_ld] bf16: global keys / values of the self-attention
    const bf16* sv = nullptr;
    int s_ld = 0;
};
int global_attention(const GvQuery& a, float* out, float* workspace, int B, int N, int C, int heads, int K, const CuboidDev& g,
                     cudaStream_t st);
// Row layouts of the global vectors' projections as the models produce them (one gv_linear over stacked weights):
//   shared net     [q | k | v]                                          ld = 3C
//   separate nets  [l2g_k | l2g_v | g2l_q | g2g_q | g2g_k | g2g_v]      ld = 3C (no global self-attention) or 6C
// g32 / g16: the fp32 rows and their bf16 copy; tok_qkv: the layer's q|k|v; tok2: the tokens' l2g_q | g2l_k | g2l_v rows (separate).
inline GvKeys gv_keys(const bf16* g16, int C, int K, bool separate, int ld, const bf16* tok2) {
    GvKeys k;
    k.k = separate ? g16 : g16 + C;
    k.v = separate ? g16 + C : g16 + 2 * C;
    k.ld = ld;
    k.n = K;
    k.q2 = separate ? tok2 : nullptr;
    k.q2_ld = 3 * C;
    return k;
}
inline GvQuery gv_query(const float* g32, const bf16* g16, const bf16* tok_qkv, const bf16* tok2, int C, bool separate,
                        bool self_attn, int ld) {
    GvQuery q;
    q.q = separate ? g32 + 2 * C : g32;
    q.q_ld = ld;
    q.tok_kv = separate ? tok2 : tok_qkv;
    if (self_attn) {
        q.sq = separate ? g32 + 3 * C : g32;
        q.sq_ld = ld;
        q.sk = separate ? g16 + 4 * C : g16 + C;
        q.sv = separate ? g16 + 5 * C : g16 + 2 * C;
        q.s_ld = ld;
    }
    return q;
}
// Row softmax for the VAE AttentionBlock: s fp32 [rows][L] -> p bf16 [rows][L], p = softmax(scale * s).
int softmax_rows(const float* s, bf16* p, int rows, int L, float scale, cudaStream_t st);
// Batched transpose bf16: in [S][R][ld_in] (first C columns used) -> out [S][C][R].
int transpose_bf16(const bf16* in, bf16* out, int S, int R, int C, int ld_in, cudaStream_t st);

// ---- elementwise / data movement (elementwise.cu) ----------------------------------------------------------
// UNet input assembly (cuboid_transformer_unet.py:425-428): cat(cond, x) on T, append the observed-indicator
// channel, zero-pad channels to Cpad. Writes fp32 and bf16 copies [B][Tc+Tx][HW][Cpad].
int unet_assemble(const float* x, const float* cond, float* out_f32, void* out_op, int B, int Tx, int Tc, int HW,
                  int C, int Cpad, cudaStream_t st, int op_f32 = 0);
// x[b,t,h,w,:] += Temb[t] + Hemb[h] + Wemb[w]   (cuboid_transformer.py:78-85)
int pos_embed_add(float* x, const float* Te, const float* He, const float* We, int B, int T, int H, int W, int C,
                  cudaStream_t st);
// nearest 2x spatial upsample + cast: fp32 [F][H][W][C] -> bf16 [F][2H][2W][C]  (cuboid_transformer.py:340,373;
// taming/resnet.py:128)
int upsample2x_cast(const float* x, void* y, int F, int H, int W, int C, cudaStream_t st, int y_f32 = 0);
// fp32 -> bf16 cast of a [S][R][C] slice taken from x with sample stride `in_sample_stride` elements.
int cast_bf16(const float* x, void* y, int S, int64_t RC, int64_t in_sample_stride, cudaStream_t st, int y_f32 = 0);
// Downsample2D prep (taming/resnet.py:183-188): fp32 [F][H][W][C] -> bf16 parity planes [F][4][H/2][W/2][C],
// plane = (y%2)*2 + (x%2), so the stride-2 3x3 conv becomes 9 unit-stride shifted loads.
int parity_split_cast(const float* x, bf16* y, int F, int H, int W, int C, cudaStream_t st);
// sinusoidal timestep embedding (models/utils.py:77-83): out[b] = [cos(t f_i) | sin(t f_i)], dim even.
// If `step` (device int) is non-null, t is a table [n_steps][B] and row *step is used (device-resident loop).
int timestep_embedding(const int64_t* t, const int* step, int t_stride, float* out, int B, int dim, cudaStream_t st);
// out[b][n] = out_act( sum_k in_act(in[b][k]) * W[n][k] + bias[n] ), fp32, tiny-M linear (time-embedding MLP).
int small_linear(const float* in, const float* W, const float* bias, float* out, int B, int K, int N, int in_silu,
                 int out_silu, cudaStream_t st);
// VAE conv_in (1 -> Cout, 3x3, pad 1): x fp32 [F][H][W] -> fp32 [F][H][W][Cout]. w fp32 [Cout][9].
int conv3x3_c1_in(const float* x, const float* w, const float* bias, float* y, int F, int H, int W, int Cout,
                  cudaStream_t st);
// VAE decoder conv_out (Cin -> 1, 3x3, pad 1): x bf16 [F][H][W][Cin] -> fp32 [F][H][W]. w fp32 [9][Cin].
int conv3x3_c1_out(const bf16* x, const float* w, float bias, float* y, int F, int H, int W, int Cin, cudaStream_t st);

// ---- weight repacking (elementwise.cu) ---------------------------------------------------------------------
// fp32 [N][K] -> bf16 [N][Kpad] (zero padded)
int pack_linear(const float* w, void* out, int N, int K, int Kpad, cudaStream_t st, int out_f32 = 0);
// fp32 [Co][Ci][taps] -> bf16 [Co][taps][Cipad]
int pack_conv(const float* w, void* out, int Co, int Ci, int taps, int Cipad, cudaStream_t st, int out_f32 = 0);

// ---- input-gradient kernels (backward.cu) - knowledge-alignment guidance only ---------------------------------
// GroupNorm(+SiLU) backward: x, dy fp32 [S][R][C]; sums = forward (sum, sumsq); bsums = zeroed scratch [S][G][2].
// dx_io (fp32, optional) = (accumulate ? dx_io : 0) + dx; dxb (bf16, optional) = the same value.
int gn_bwd(const float* x, const float* dy, const double* sums, double* bsums, const float* gamma, const float* beta,
           float* dx_io, bf16* dxb, int S, int R, int C, int G, float eps, int silu, int accumulate, cudaStream_t st);
// LayerNorm backward: x (forward input), dy fp32 [P][C]; dx_io = (accumulate ? dx_io : 0) + dx; dxb = bf16(dx_io).
int layer_norm_bwd(const float* x, const float* gamma, const float* dy, float* dx_io, bf16* dxb, int P, int C, float eps,
                   int accumulate, cudaStream_t st);
// Backward of patch_merge_ln: dy fp32 [BT*(H/2)*(W/2)][4C] -> dx fp32 / dxb bf16 [BT][H][W][C] (plain store).
int patch_merge_ln_bwd(const float* x, const float* gamma, const float* dy, float* dx, bf16* dxb, int BT, int H, int W,
                       int C, float eps, cudaStream_t st);
// y = GELU_erf(pre) -> bf16; dpre = dmid * GELU'(pre).
int gelu_fwd(const float* pre, bf16* y, int64_t n, cudaStream_t st);
int gelu_bwd(const float* pre, const bf16* dmid, bf16* dpre, int64_t n, cudaStream_t st);
// Backward of axial_attention: qkv (forward input), dout bf16 [B][T][H][W][C] -> dqkv bf16 [B][T][H][W][3C].
int axial_attention_bwd(const bf16* qkv, const float* bias_table, const bf16* dout, bf16* dqkv, int B, int T, int H,
                        int W, int C, int heads, int axis, cudaStream_t st);
// dgrad operands: fp32 [N][K] -> bf16 [K][N];  fp32 [Co][Ci][taps] -> bf16 [Ci][taps reversed][Co].
int pack_linear_t(const float* w, bf16* out, int N, int K, cudaStream_t st);
int pack_conv_dgrad(const float* w, bf16* out, int Co, int Ci, int taps, cudaStream_t st);

// ---- knowledge-alignment read-out head (ka_head.cu; reference knowledge_alignment/models.py:19-104,500-528) ----
int ka_tokens(const float* x, const double* sums, const float* gamma, const float* beta, const float* pos_tc, bf16* tok,
              int F, int R, int C, int G, float eps, cudaStream_t st);
int ka_pool(const float* qkv, const float* cw, float cb, float* wsave, float* out, int F, int L, int C, int heads,
            cudaStream_t st);
int ka_loss_grad(const float* out, const float* target, float* dout, float* loss, int B, int T, float guide_scale,
                 cudaStream_t st);
int ka_pool_bwd(const float* qkv, const float* cw, const float* wsave, const float* dout, bf16* dqkv, int F, int L, int C,
                int heads, cudaStream_t st);
int ka_tokens_bwd(const float* dtok, float* dout, int F, int R, int C, cudaStream_t st);
int transpose_f32(const float* in, float* out, int R, int C, cudaStream_t st);

// ---- fused FFN (ffn_fused.cu) ---------------------------------------------------------------------------------
// x[M][256] += W2 GELU(W1 ln + b1) + b2 in one kernel per 128-row tile (hidden width 1024 stays on the SM), optionally
// followed by the fused LayerNorm of the new rows -> ln_out (bf16). ln_in bf16 [M][256]; W1 bf16 [1024][256]; W2 bf16
// [256][1024]. Reference: PositionwiseFFN.forward, cuboid_transformer.py:182-208.
struct FfnFusedOp {
    alignas(64) unsigned char storage[1536];
};
// Optional front-end: the attention output projection x1 = x + att Wp^T + bp and the FFN's pre-norm LayerNorm(x1)
// (gamma / beta below) are computed inside the kernel as well; `ln_in` is then only the bf16 scratch tensor the
// normalised tile makes its L2 round trip through, x1 itself never leaves the SM (it seeds the GEMM-2 accumulator).
struct FfnProjArgs {
    const bf16* att;        // [M][256] attention output (A operand of the projection)
    const bf16* wp;         // [256][256]
    const float* bp;        // [256]
    const float* ln1_gamma; // [256]
    const float* ln1_beta;
};
int ffn_fused_make(FfnFusedOp* op, const bf16* ln_in, int M, const bf16* w1, const float* b1, const bf16* w2,
                   const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta, bf16* ln_out,
                   float ln_eps, unsigned long long* dbg = nullptr, const FfnProjArgs* proj = nullptr);
// Also accumulate the GroupNorm statistics of the new x rows into gn_sums[S][groups][2] (see GemmEpilogue::gn_sums).
int ffn_fused_set_gn(FfnFusedOp* op, double* gn_sums, int groups, int rows_per_sample);
// L2 weight prefetch chain (common.cuh WRange): this op's weight ranges / the ranges it should request for its successor.
void ffn_fused_set_prefetch(FfnFusedOp* op, const WRange& next);
WRange ffn_fused_weights(const FfnFusedOp& op);
int ffn_fused_launch(const FfnFusedOp& op, cudaStream_t st);

// ---- fused FFN for the width-512 level (ffn_cluster.cu) ------------------------------------------------------------
// x[M][512] += W2 GELU(W1 ln + b1) + b2 in one kernel, hidden dimension (2048) split over a 4-CTA cluster with a
// distributed-shared-memory reduce-scatter, optionally followed by the fused LayerNorm of the new rows -> ln_out (bf16)
// and the GroupNorm statistics of the new rows. ln_in bf16 [M][512]; W1 bf16 [2048][512]; W2 bf16 [512][2048].
struct FfnClusterOp {
    alignas(64) unsigned char storage[1536];
};
// Optional front-end (FfnProjArgs, width 512 here): x1 = x + att Wp^T + bp and the FFN's pre-norm LayerNorm(x1) are computed
// by the kernel as well - each CTA of the cluster its 128 output columns, row sums exchanged through distributed shared
// memory, the normalised slices published to `ln_in` (then a scratch tensor) through L2 and loaded back by all four CTAs.
// `workspace`: ffn_cluster_workspace_bytes(M) bytes of device memory the partial slices cross L2 in (may be shared by ops
// that run one at a time on a stream).
size_t ffn_cluster_workspace_bytes(int M);
int ffn_cluster_make(FfnClusterOp* op, const bf16* ln_in, int M, const bf16* w1, const float* b1, const bf16* w2,
                     const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta, bf16* ln_out, float ln_eps,
                     float* workspace, const FfnProjArgs* proj = nullptr);
int ffn_cluster_set_gn(FfnClusterOp* op, double* gn_sums, int groups, int rows_per_sample);
void ffn_cluster_set_dbg(FfnClusterOp* op, unsigned long long* stamps32);   // clock64() phase stamps of CTA 0
WRange ffn_cluster_weights(const FfnClusterOp& op);
int ffn_cluster_launch(const FfnClusterOp& op, cudaStream_t st);

// ---- fused QKV projection + axial attention core (qkv_attn.cu) -----------------------------------------------------
// att[B][T][H][W][C] (bf16) = axial attention along `axis` (0 = T, 1 = H, 2 = W; line length <= 16) of q|k|v = ln Wqkv^T
// (no bias), heads of C / heads channels, relative-position bias table [2 L - 1][heads]: the QKV GEMM and
// axial_attention() in one kernel, bit-identical to that pair (q|k|v are rounded to bf16 before the attention core there
// as here). ln bf16 [B][T][H][W][C]; Wqkv bf16 [3 C][C]. Reference: cuboid_transformer.py:812-861, 949.
struct QkvAttnOp {
    alignas(64) unsigned char storage[768];
};
bool qkv_attn_supported(int T, int H, int W, int C, int heads, int axis);
int qkv_attn_make(QkvAttnOp* op, const bf16* ln, const bf16* wqkv, const float* bias_table, bf16* out, int B, int T, int H,
                  int W, int C, int heads, int axis);
void qkv_attn_set_prefetch(QkvAttnOp* op, const WRange& next);
void qkv_attn_set_dbg(QkvAttnOp* op, unsigned long long* stamps32, int cta_x = 0, int cta_y = 0);   // phase stamps of one CTA
WRange qkv_attn_weights(const QkvAttnOp& op);
double qkv_attn_flops(const QkvAttnOp& op);
int qkv_attn_launch(const QkvAttnOp& op, cudaStream_t st);

// ---- evaluation (eval.cu) ----------------------------------------------------------------------------------
// SEVIR skill-score contingency counts + error sums, accumulated on the device (evaluation.py:197-245).
// pred / target fp32 [N][T][H][W] in [0,1]; counts int64 [n_thr][T][3] (hits, misses, false alarms) and sums double [2]
// (sum sq err, sum abs err) are ADDED to; thresholds: host pointer, VIL units 0..255.
int sevir_eval_update(const float* pred, const float* target, long long* counts, double* sums, int N, int T, int H, int W,
                      int pool, const float* thresholds, int n_thr, cudaStream_t st);

// SSIM of torchmetrics.image.StructuralSimilarityIndexMeasure() (defaults) for single-channel frames: pred / target fp32
// [N][H][W]; data_range <= 0 = None (taken from the batch); state double [2] += {sum of per-image SSIM, N}.
int ssim_update(const float* pred, const float* target, int N, int H, int W, float data_range, double* state,
                cudaStream_t st);

// ---- input side (io.cu) ------------------------------------------------------------------------------------
// Raw uint8 VIL events [n_events][H][W][T_raw] (events event_base .. event_base + n_events - 1) -> sequent windows
// first_seq .. first_seq + batch - 1 as fp32 [batch][seq_len][H][W] = scale * (x + offset) (sevir_dataloader.py:834-877,
// :610-650).
int sevir_windows(const uint8_t* events, int event_base, int n_events, int H, int W, int T_raw, long long first_seq, int batch,
                  int seq_len, int stride, float scale, float offset, float* out, cudaStream_t st);

// ---- sampler (sampler.cu) ----------------------------------------------------------------------------------
// One fused update of the latent (latent_diffusion.py:553-566,620-631 for DDPM; SURVEY section 8 S6 for DDIM):
//   z0 = c[0] z - c[1] eps ;  z <- c[2] z0 + c[3] z + c[4] eps + c[5] noise - c[6] guide
// coef points at 8 floats in device memory (the row of the resident schedule table for this step).
// If `step` (device int) is non-null, coef is a table [n_steps][8] and noise a stack [n_steps][n]; row *step is used.
// noise_step_stride: elements between consecutive steps of the noise stack (0 -> n; larger when z is a sub-batch).
int sampler_update(float* z, const float* eps, const float* noise, const float* guide, const float* coef,
                   const int* step, int64_t n, int64_t noise_step_stride, cudaStream_t st);
// Forward diffusion (q_sample, latent_diffusion.py:489-492): out = sqrt_ac[t_b] x0 + sqrt_1mac[t_b] noise; tables and t
// on the device, n elements per sample.
int q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac, const float* sqrt_1mac,
             float* out, int B, int64_t n, cudaStream_t st);
// p_losses reductions (latent_diffusion.py:534-549): per_sample[b] = mean |target - pred|^p over the sample, then
// out4 = {mean loss_simple, loss_vlb, loss, loss_gamma}.
int diffusion_loss_reduce(const float* pred, const float* target, const int64_t* t, const float* lvlb, float logvar,
                          float w_simple, float w_elbo, int l1, float* per_sample, float* out4, int B, int64_t n,
                          cudaStream_t st);
// *step += 1 (one thread): closes one iteration of the device-resident sampling loop.
int advance_step(int* step, cudaStream_t st);

}  // namespace pd
