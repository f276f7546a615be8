// CuboidTransformerUNet on the GPU: weights (reference key names), repacked bf16 operands and per-batch plans.
#pragma once
#include "model_common.cuh"
#include <cstdlib>

namespace pd {

// Operand pointers (`bf16*` members below, activation buffers in the plans) hold bf16 data, or fp32 data rounded to tf32
// when the model's precision is PD_PRECISION_TF32 (the buffers are then sized 4 bytes per element).
struct ResW {   // TimeEmbedResBlock
    const float *gn1_w, *gn1_b, *conv1_b, *gn2_w, *gn2_b, *conv2_b;
    bf16 *conv1_w, *conv2_w;
};
struct AttnW {  // CuboidSelfAttentionLayer
    const float *ln_w, *ln_b, *table, *proj_b;
    bf16 *qkv_w, *proj_w;
    // global vectors (fp32 weights in the reference layout, used as loaded): global_vec_norm, global_proj and the rows'
    // projection - global_qkv [3C][C], or with separate_global_qkv the stack l2g_global_kv | g2l_global_q [| g2g_global_qkv]
    // ([3C or 6C][C], concatenated at finalize) - plus, separate only, the tokens' second projection l2g_q | g2l_k | g2l_v
    const float *g_ln_w = nullptr, *g_ln_b = nullptr, *g_qkv_w = nullptr, *g_proj_w = nullptr, *g_proj_b = nullptr;
    bf16* tok2_w = nullptr;   // packed [3C][C]
};
struct FfnW {   // PositionwiseFFN
    const float *ln_w, *ln_b, *b1, *b2;
    bf16 *w1, *w2;
};
struct GFfnW {  // PositionwiseFFN of the global vectors (global_ffn_l), fp32 weights as loaded
    const float *ln_w, *ln_b, *w1, *b1, *w2, *b2;
};
struct StackW {  // StackCuboidSelfAttentionBlock: n x (attn, ffn), n = number of cuboid layers of the level's pattern
    std::vector<AttnW> a;
    std::vector<FfnW> f;
    std::vector<GFfnW> gf;   // empty without global vectors / use_global_vector_ffn
};

class UNet {
public:
    struct Bufs;
    struct BatchPlan;

    // n_global > 0: global vectors (cuboid_transformer_unet.py:55-60 with separate_global_qkv=False, global_dim_ratio=1)
    UNet(const pd_unet_config& c, const pd_unet_pattern* pattern, int n_global = 0, int global_ffn = 1, int global_self_attn = 0,
         int global_separate = 0);
    ~UNet();
    int validate() const;
    int finalize();
    // Operand precision of every tensor-core GEMM of the model (PD_PRECISION_*): bf16 (default) or tf32 - the
    // reference's own GPU arithmetic (cfg.yaml:32; train_sevirlr_prediff.py:1143). Takes effect at the next finalize().
    int set_precision(int prec);
    int set_streamk_ctas(int n);
    int precision = 0;
    int streamk_cps = 0;   // CTAs per sample of the stream-K cut (0 = default, see set_streamk_ctas)
    // t: device int64; if `step` (device int) is given, t is a table and row *step is used (sampler loop)
    // replica: independent workspace index (sub-batches of one call running concurrently on different streams);
    // t_stride: row length of the t table when `step` is given (0 -> B)
    int forward(const float* x, const int64_t* t, const int* step, const float* cond, float* out, int B, cudaStream_t st,
                PlanProfile* prof = nullptr, int replica = 0, int t_stride = 0, unsigned long long* trace_ns = nullptr);
    int plan_labels(int B, std::vector<std::string>* out);   // one label per plan step (trace_ns has size + 1 slots)
    int plan_flops(int B, std::vector<double>* out);          // algorithmic FLOPs per plan step (0 for non-GEMM steps)
    int kernels_per_forward(int B, int* n);
    int get_plan(int B, BatchPlan** out, int replica = 0);

    pd_unet_config cfg;
    int C0, C1, T, TE;
    // attention layers of each level's stack block (block_attn_patterns) and their geometry tables (built in finalize)
    std::vector<CuboidLayerSpec> layers[2];
    int padding_type = 0;   // 0 = 'zeros', 1 = 'ignore', 2 = 'nearest'
    int n_global = 0;       // num_global_vectors
    bool global_ffn = true, global_self_attn = false;
    bool global_separate = false;   // separate_global_qkv
    int g_row_ld(int C) const { return (global_separate && global_self_attn ? 6 : 3) * C; }   // width of the global rows' projection
    bool all_axial = true;
    WeightStore ws;
    bool finalized = false;
    // bumped by every finalize(): packed weights and plans (activation arenas) of older generations are gone, so anything
    // that cached device addresses of this model (the sampler's CUDA graphs) must be rebuilt
    unsigned long long generation = 0;

private:
    void declare_weights();
    void declare_resblock(const std::string& p, int cin, int cout, bool emb);
    void declare_stack(const std::string& p, int dim, int lvl);
    int pack_conv_w(const std::string& name, int co, int ci, int taps, int cipad, bf16** out);
    int pack_linear_w(const std::string& name, int n, int k, bf16** out);
    int finalize_resblock(const std::string& p, int cin, int cinpad, int cout, ResW* r);
    int finalize_stack(const std::string& p, int dim, int lvl, StackW* s);
    int build_plan(int B, BatchPlan* bp);
    // `next`: the stack block that follows - when the level width is 256 its first LayerNorm is fused into the
    // resblock's conv2 epilogue (and add_stack skips that LayerNorm launch)
    // x_stats_ready: the producer of x already accumulated the first GroupNorm's statistics (GemmEpilogue::gn_sums)
    int add_resblock(Plan& pl, const Bufs& b, int B, int lvl, const ResW& r, int emb_index, int* gn_slot,
                     const StackW* next, bool x_stats_ready);
    // gn_next: statistics table of the resblock that follows the stack (filled by the stack's last kernel), or null
    int add_stack(Plan& pl, const Bufs& b, int B, int lvl, const StackW& s, double* gn_next);
    int make_conv(GemmOp* op, const void* a, const GemmGeom& g, const void* w, int N, GemmEpilogue e, const Bufs& b,
                  bool* gn_fused = nullptr);
    // The LayerNorm that follows a GEMM is computed in that GEMM's epilogue when a CTA (width 256) or a 2-CTA cluster
    // (width 512) owns whole rows; the resblock's conv2 only when it is not split-K (27 * C / 64 < 128 k-blocks).
    bool ln_fusable(int lvl) const {
        static const bool no_tf32_ln = getenv("PD_NO_TF32_LN_FUSION") != nullptr;   // A/B: layer_norm launches in tf32 mode
        if (precision && no_tf32_ln) return false;
        const int c = lvl ? C1 : C0;
        static const bool no_cluster = getenv("PD_NO_LN_CLUSTER") != nullptr;   // A/B: separate LayerNorm launches at 512
        return c == 256 || (c == 512 && !no_cluster);
    }
    bool ln_fusable_conv(int lvl) const { const int c = lvl ? C1 : C0; return c == 256 && (!precision || ln_fusable(lvl)); }
    // GEMM geometry / operand-output helpers for the model's precision
    GemmGeom geom(GemmGeom g) const { g.tf32 = precision; return g; }
    void operand_out(GemmEpilogue& e, bf16* buf) const {
        if (precision) { e.out_f32 = reinterpret_cast<float*>(buf); e.round_tf32 = 1; }
        else e.out_bf16 = buf;
    }
    size_t op_bytes() const { return precision ? 4 : 2; }
    int num_gn_slots() const;
    template <class A>
    void carve(A& ar, int B, Bufs* b) const;

    BatchPlan* building_ = nullptr;   // plan under construction (make_conv attaches stream-K workspaces to it)
    // plan under construction, global vectors: marks (Plan::mark) of the last cuboid-attention kernel (reader of the global
    // k | v buffer) and of the last global attention (reader of the grid's q|k|v buffer)
    int gv_mark_attn_ = -1, gv_mark_gvattn_ = -1;
    std::vector<std::unique_ptr<DevMem>> packed;
    std::vector<std::unique_ptr<DevMem>> gv_stacked;   // concatenated fp32 weights of the separate global nets
    std::vector<std::unique_ptr<CuboidTablesDev>> cub_dev[2];   // per level, per layer (null for axial fast-path layers)
    std::vector<int> cub_axis[2];                               // axial fast-path axis or -1
    std::map<std::pair<int, int>, std::unique_ptr<BatchPlan>> plans;  // (batch, replica)
    ResW first{}, down_res[2]{}, up_res[2]{};
    std::vector<StackW> down_stack[2], up_stack[2];
    bf16 *first_skip_w = nullptr, *pm_w = nullptr, *up_w = nullptr, *final_w = nullptr;
    const float *first_skip_b = nullptr, *pos_T = nullptr, *pos_H = nullptr, *pos_W = nullptr;
    const float *te_w0 = nullptr, *te_b0 = nullptr, *te_w2 = nullptr, *te_b2 = nullptr;
    const float *pm_ln_w = nullptr, *pm_ln_b = nullptr, *up_b = nullptr, *final_b = nullptr;
    const float *gv_init = nullptr, *gv_down_w = nullptr, *gv_down_b = nullptr, *gv_up_w = nullptr, *gv_up_b = nullptr;
    DevMem first_gn_pad, emb_cat;
    int emb_total = 0, emb_off[4] = {0, 0, 0, 0};
};

}  // namespace pd
