// AutoencoderKL on the GPU: weights under the reference key names, repacked operands, per-frame-count plans.
#pragma once
#include "model_common.cuh"

namespace pd {

struct Res2W {  // ResnetBlock2D
    int cin, cout;
    const float *gn1_w, *gn1_b, *conv1_b, *gn2_w, *gn2_b, *conv2_b, *sc_b;
    bf16 *conv1_w, *conv2_w, *sc_w;
};
struct MidW {   // UNetMidBlock2D: resnet, single-head attention, resnet
    Res2W r0, r1;
    const float *gn_w, *gn_b, *qkv_b, *proj_b;
    bf16 *qkv_w, *proj_w;
};

class VAE {
public:
    struct Ctx;
    struct DirPlan;
    explicit VAE(const pd_vae_config& c);
    ~VAE();
    int validate() const;
    int finalize();
    int encode(const float* x, float* moments, int N, cudaStream_t st);
    int decode(const float* z, float* out, int N, cudaStream_t st);

    pd_vae_config cfg;
    WeightStore ws;
    bool finalized = false;

private:
    void declare_weights();
    void declare_resnet(const std::string& p, int cin, int cout);
    void declare_mid(const std::string& p, int c);
    int pack_conv_w(const std::string& name, int co, int ci, int taps, bf16** out);
    int finalize_resnet(const std::string& p, int cin, int cout, Res2W* r);
    int finalize_mid(const std::string& p, int c, MidW* m);
    int add_gn(Ctx& c, const float* x, const float* w, const float* b, bf16* y, int R, int C, int silu);
    int add_resnet(Ctx& c, const Res2W& r, int h, int w);
    int add_mid(Ctx& c, const MidW& m, int h, int w, int ch);
    int64_t max_elems(int N) const;
    template <class A>
    void carve(A& ar, int N, Ctx* c, int n_gn) const;
    int build(int N, bool encode, DirPlan* dp);
    int get_plan(int N, bool encode, DirPlan** out);

    std::vector<std::unique_ptr<DevMem>> packed;
    std::map<int, std::unique_ptr<DirPlan>> enc_plans, dec_plans;
    std::vector<Res2W> enc_res, dec_res;
    MidW enc_mid{}, dec_mid{};
    const float *enc_in_w = nullptr, *enc_in_b = nullptr, *enc_no_w = nullptr, *enc_no_b = nullptr, *enc_out_b = nullptr;
    const float *quant_b = nullptr, *pquant_b = nullptr, *dec_in_b = nullptr, *dec_no_w = nullptr, *dec_no_b = nullptr;
    const float *down_b[3] = {nullptr, nullptr, nullptr}, *up_b[3] = {nullptr, nullptr, nullptr};
    bf16 *down_w[3] = {nullptr, nullptr, nullptr}, *up_w[3] = {nullptr, nullptr, nullptr};
    bf16 *enc_out_w = nullptr, *quant_w = nullptr, *pquant_w = nullptr, *dec_in_w = nullptr;
    DevMem dec_out_w;
    float dec_out_b = 0.f;
};

}  // namespace pd
