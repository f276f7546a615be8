// Knowledge-alignment network U(z_t, t) (NoisyCuboidTransformerEncoder, reference
// src/prediff/diffusion/knowledge_alignment/models.py:459-528) and its guidance gradient
//   g = guide_scale * d || mean_T U(z_t, t) - avg_x_gt ||_2 / d z_t     (sevir.py:76-104, alignment_pl.py:441-445)
// as two static launch plans: a forward that keeps the inputs of every non-linearity, and a hand-written
// input-gradient backward (no weight gradients, so no GEMM input is ever stored).
#pragma once
#include "model_common.cuh"
#include "unet.cuh"   // ResW / AttnW / FfnW / StackW

namespace pd {

struct ResBwdW {    // dgrad operands of a TimeEmbedResBlock
    bf16 *conv1_d, *conv2_d;
};
struct StackBwdW {  // transposed Linear weights of a StackCuboidSelfAttentionBlock
    bf16 *qkv_t[3], *proj_t[3], *w1_t[3], *w2_t[3];
};

class KANet {
public:
    struct Bufs;
    struct BatchPlan;

    explicit KANet(const pd_ka_config& c);
    ~KANet();
    int validate() const;
    int finalize();
    // pred [B][T] = U(zt, t). t: device int64 [B] (or a table indexed by the device counter `step`, row length t_stride).
    int forward(const float* zt, const int64_t* t, const int* step, int t_stride, float* pred, int B, cudaStream_t st);
    // grad_out [B][T][H][W][C] (may be null: result stays in guide_buffer(B)); avg_x_gt: device fp32 [B].
    int mean_shift(const float* zt, const int64_t* t, const int* step, int t_stride, const float* avg_x_gt,
                   float guide_scale, float* grad_out, int B, cudaStream_t st);
    // Device buffer the last mean_shift(.., B) left the guidance in; *loss_dev = the alignment value (1 float).
    int guide_buffer(int B, float** g, float** loss_dev);
    int kernels(int B, int* n_fwd, int* n_bwd);

    pd_ka_config cfg;
    int C0, C1, T, TE;
    WeightStore ws;
    bool finalized = false;
    unsigned long long generation = 0;   // see UNet::generation

private:
    void declare_weights();
    void declare_resblock(const std::string& p, int cin, int cout, bool emb);
    void declare_stack(const std::string& p, int dim, int lvl);
    int pack_conv_w(const std::string& name, int co, int ci, int taps, bf16** fwd, bf16** dgrad);
    int pack_linear_w(const std::string& name, int n, int k, bf16** fwd, bf16** tr);
    int finalize_resblock(const std::string& p, int cin, int cout, ResW* r, ResBwdW* rb);
    int finalize_stack(const std::string& p, int dim, StackW* s, StackBwdW* sb);
    int get_plan(int B, BatchPlan** out);
    int build_plan(int B, BatchPlan* bp);
    int bind(BatchPlan* bp, const float* zt, const int64_t* t, const int* step, int t_stride, int B);
    int num_gn_slots() const { return 2 + 2 * (cfg.depth[0] + cfg.depth[1]) + 1; }
    int num_blocks() const { return cfg.depth[0] + cfg.depth[1]; }
    template <class A>
    void carve(A& ar, int B, Bufs* b) const;

    std::vector<std::unique_ptr<DevMem>> packed;
    std::map<int, std::unique_ptr<BatchPlan>> plans;
    ResW first{}, res[2]{};
    ResBwdW first_b{}, res_b[2]{};
    std::vector<StackW> stack[2];
    std::vector<StackBwdW> stack_b[2];
    bf16 *skip_w = nullptr, *skip_t = nullptr, *pm_w = nullptr, *pm_t = nullptr, *hq_w = nullptr, *hq_t = nullptr;
    const float *skip_b = nullptr, *pos_T = nullptr, *pos_H = nullptr, *pos_W = nullptr;
    const float *te_w0 = nullptr, *te_b0 = nullptr, *te_w2 = nullptr, *te_b2 = nullptr;
    const float *pm_ln_w = nullptr, *pm_ln_b = nullptr, *out_gn_w = nullptr, *out_gn_b = nullptr, *hq_b = nullptr;
    const float* cproj_w = nullptr;
    float cproj_b = 0.f;
    DevMem emb_cat, head_pos;
    int emb_total = 0, emb_off[2] = {0, 0};
};

}  // namespace pd
