// extern "C" model-level entry points (include/prediff_b200.h): UNet, sampler. (VAE: api_vae.cu)
#include "sampler_host.cuh"
#include "unet.cuh"
#include "ka_api.cuh"

using namespace pd;

struct pd_unet {
    UNet impl;
    pd_unet(const pd_unet_config& c, const pd_unet_pattern* p, int n_global = 0, int gffn = 1, int gsa = 0, int gsep = 0)
        : impl(c, p, n_global, gffn, gsa, gsep) {}
};
struct pd_sampler {
    Sampler impl;
    pd_sampler(int n, double a, double b) : impl(n, a, b) {}
};

namespace {
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
}

extern "C" {

int pd_unet_create(const pd_unet_config* cfg, pd_unet** out) { return pd_unet_create_ex(cfg, nullptr, out); }
int pd_unet_create_ex(const pd_unet_config* cfg, const pd_unet_pattern* pattern, pd_unet** out) {
    return pd_unet_create_gv(cfg, pattern, 0, 1, 0, out);
}
int pd_unet_create_gv(const pd_unet_config* cfg, const pd_unet_pattern* pattern, int num_global_vectors,
                      int use_global_vector_ffn, int use_global_self_attn, pd_unet** out) {
    return pd_unet_create_gv_ex(cfg, pattern, num_global_vectors, use_global_vector_ffn, use_global_self_attn, 0, out);
}
int pd_unet_create_gv_ex(const pd_unet_config* cfg, const pd_unet_pattern* pattern, int num_global_vectors,
                         int use_global_vector_ffn, int use_global_self_attn, int separate_global_qkv, pd_unet** out) {
    PD_CHECK(cfg && out, PD_ERR_ARG, "pd_unet_create: null argument");
    PD_CHECK(num_global_vectors >= 0 && num_global_vectors <= 32, PD_ERR_ARG,
             "pd_unet_create_gv: num_global_vectors %d (0..32 are built)", num_global_vectors);
    if (pattern)
        for (int l = 0; l < 2; ++l)
            PD_CHECK(pattern->n_layers[l] >= 1 && pattern->n_layers[l] <= PD_MAX_ATTN_LAYERS, PD_ERR_ARG,
                     "pd_unet_create_ex: level %d has %d attention layers (1..%d)", l, pattern->n_layers[l], PD_MAX_ATTN_LAYERS);
    pd_unet* m = new (std::nothrow) pd_unet(*cfg, pattern, num_global_vectors, use_global_vector_ffn, use_global_self_attn,
                                            separate_global_qkv);
    PD_CHECK(m, PD_ERR_CUDA, "pd_unet_create: out of host memory");
    const int rc = m->impl.validate();
    if (rc != PD_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return PD_OK;
}
void pd_unet_destroy(pd_unet* m) { delete m; }
int pd_unet_num_weights(const pd_unet* m) { return m ? m->impl.ws.size() : 0; }
int pd_unet_weight_info(const pd_unet* m, int i, const char** name, int64_t shape[5]) {
    PD_CHECK(m && name && shape && i >= 0 && i < m->impl.ws.size(), PD_ERR_ARG, "pd_unet_weight_info: bad argument");
    const WeightEntry& e = m->impl.ws.at(i);
    *name = e.name.c_str();
    for (size_t d = 0; d < 5; ++d) shape[d] = d < e.shape.size() ? e.shape[d] : 0;
    return (int)e.shape.size();
}
int pd_unet_load_weight(pd_unet* m, const char* name, const float* data, const int64_t* shape, int ndim) {
    PD_CHECK(m, PD_ERR_ARG, "pd_unet_load_weight: null model");
    m->impl.finalized = false;
    return m->impl.ws.load(name, data, shape, ndim);
}
int pd_unet_set_precision(pd_unet* m, int precision) {
    PD_CHECK(m, PD_ERR_ARG, "pd_unet_set_precision: null model");
    return m->impl.set_precision(precision);
}
int pd_unet_set_streamk_ctas(pd_unet* m, int ctas_per_sample) {
    PD_CHECK(m, PD_ERR_ARG, "pd_unet_set_streamk_ctas: null model");
    return m->impl.set_streamk_ctas(ctas_per_sample);
}
int pd_unet_finalize(pd_unet* m) {
    PD_CHECK(m, PD_ERR_ARG, "pd_unet_finalize: null model");
    return m->impl.finalize();
}
int pd_unet_forward(pd_unet* m, const float* x, const int64_t* t, const float* cond, float* out, int batch,
                    void* stream) {
    PD_CHECK(m, PD_ERR_ARG, "pd_unet_forward: null model");
    return m->impl.forward(x, t, nullptr, cond, out, batch, S(stream));
}

int pd_unet_profile_forward(pd_unet* m, const float* x, const int64_t* t, const float* cond, float* out, int batch,
                            void* stream, double stats[5]) {
    PD_CHECK(m && stats, PD_ERR_ARG, "pd_unet_profile_forward: null argument");
    PlanProfile p;
    PD_TRY(m->impl.forward(x, t, nullptr, cond, out, batch, S(stream), &p));
    stats[0] = p.gemm_ms; stats[1] = p.n_gemm; stats[2] = p.other_ms; stats[3] = p.n_other; stats[4] = p.gemm_flops;
    return PD_OK;
}
int pd_unet_trace_forward(pd_unet* m, const float* x, const int64_t* t, const float* cond, float* out, int batch,
                          void* stream, unsigned long long* ns_dev, int max_slots, char* labels, int labels_bytes) {
    PD_CHECK(m && ns_dev, PD_ERR_ARG, "pd_unet_trace_forward: null argument");
    std::vector<std::string> lab;
    PD_TRY(m->impl.plan_labels(batch, &lab));
    PD_CHECK((int)lab.size() + 1 <= max_slots, PD_ERR_ARG, "pd_unet_trace_forward: need %d slots", (int)lab.size() + 1);
    if (labels) {
        std::string all;
        for (const auto& l : lab) { all += l; all += '\n'; }
        PD_CHECK((int)all.size() + 1 <= labels_bytes, PD_ERR_ARG, "pd_unet_trace_forward: label buffer too small");
        memcpy(labels, all.c_str(), all.size() + 1);
    }
    PD_TRY(m->impl.forward(x, t, nullptr, cond, out, batch, S(stream), nullptr, 0, 0, ns_dev));
    return (int)lab.size();
}
int pd_unet_step_flops(pd_unet* m, int batch, double* flops, int max_slots) {
    PD_CHECK(m && flops, PD_ERR_ARG, "pd_unet_step_flops: null argument");
    std::vector<double> f;
    PD_TRY(m->impl.plan_flops(batch, &f));
    PD_CHECK((int)f.size() <= max_slots, PD_ERR_ARG, "pd_unet_step_flops: need %d slots", (int)f.size());
    for (size_t i = 0; i < f.size(); ++i) flops[i] = f[i];
    return (int)f.size();
}
int pd_unet_kernels_per_forward(pd_unet* m, int batch, int* n) {
    PD_CHECK(m && n, PD_ERR_ARG, "pd_unet_kernels_per_forward: null argument");
    return m->impl.kernels_per_forward(batch, n);
}

int pd_sampler_create(int num_timesteps, double linear_start, double linear_end, pd_sampler** out) {
    PD_CHECK(out && num_timesteps >= 2 && linear_start > 0 && linear_end > linear_start, PD_ERR_ARG,
             "pd_sampler_create: bad argument");
    *out = new (std::nothrow) pd_sampler(num_timesteps, linear_start, linear_end);
    PD_CHECK(*out, PD_ERR_CUDA, "pd_sampler_create: out of host memory");
    return PD_OK;
}
void pd_sampler_destroy(pd_sampler* s) { delete s; }
int pd_sampler_get_buffer(const pd_sampler* s, const char* name, float* out) {
    PD_CHECK(s, PD_ERR_ARG, "pd_sampler_get_buffer: null sampler");
    return s->impl.get_buffer(name, out);
}
int pd_sample_loop(pd_sampler* s, pd_unet* unet, float* z, const float* cond, const float* noise, int batch, int mode,
                   int n_steps, float eta, void* stream) {
    PD_CHECK(s && unet, PD_ERR_ARG, "pd_sample_loop: null handle");
    return s->impl.loop(&unet->impl, z, cond, noise, batch, mode, n_steps, eta, 0, n_steps, S(stream));
}
int pd_sample_loop_range(pd_sampler* s, pd_unet* unet, float* z, const float* cond, const float* noise, int batch,
                         int mode, int n_steps, float eta, int k_begin, int k_end, void* stream) {
    PD_CHECK(s && unet, PD_ERR_ARG, "pd_sample_loop_range: null handle");
    return s->impl.loop(&unet->impl, z, cond, noise, batch, mode, n_steps, eta, k_begin, k_end, S(stream));
}
int pd_sample_loop_aligned(pd_sampler* s, pd_unet* unet, pd_ka* ka, float* z, const float* cond, const float* noise,
                           const float* avg_x_gt, float guide_scale, int batch, int mode, int n_steps, float eta,
                           int k_begin, int k_end, void* stream) {
    PD_CHECK(s && unet && ka && avg_x_gt, PD_ERR_ARG, "pd_sample_loop_aligned: null argument");
    Sampler::Align al;
    al.ka = &ka->impl;
    al.avg_x_gt = avg_x_gt;
    al.guide_scale = guide_scale;
    return s->impl.loop(&unet->impl, z, cond, noise, batch, mode, n_steps, eta, k_begin, k_end, S(stream), al);
}
int pd_diffusion_losses(pd_sampler* s, pd_unet* unet, const float* x_start, const float* cond, const int64_t* t,
                        const float* noise, int batch, int loss_l1, float logvar, float l_simple_weight,
                        float original_elbo_weight, float* per_sample, float* out4, void* stream) {
    PD_CHECK(s && unet, PD_ERR_ARG, "pd_diffusion_losses: null handle");
    PD_CHECK(unet->impl.finalized, PD_ERR_WEIGHT, "pd_diffusion_losses: pd_unet_finalize has not been called");
    return s->impl.losses(&unet->impl, x_start, cond, t, noise, batch, loss_l1, logvar, l_simple_weight,
                          original_elbo_weight, per_sample, out4, S(stream));
}
int pd_sampler_set_clip_denoised(pd_sampler* s, int on) {
    PD_CHECK(s, PD_ERR_ARG, "pd_sampler_set_clip_denoised: null sampler");
    s->impl.set_clip_denoised(on != 0);
    return PD_OK;
}
int pd_sampler_sub_batches(const pd_sampler* s, int batch) { return s ? s->impl.n_sub_for(batch) : 0; }
int pd_sample_step_ddpm(pd_sampler* s, pd_unet* unet, float* z, const float* cond, const float* noise, int batch, int t,
                        void* stream) {
    PD_CHECK(s && unet, PD_ERR_ARG, "pd_sample_step_ddpm: null handle");
    return s->impl.step_ddpm(&unet->impl, z, cond, noise, batch, t, S(stream));
}

}  // extern "C"
