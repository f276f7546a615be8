// Noise schedule + device-resident sampling loop.
// Reference: LatentDiffusion.register_schedule (latent_diffusion.py:228-278), p_sample / p_sample_loop
// (:598-684), make_ddim_timesteps / make_ddim_sampling_parameters (diffusion/utils.py:42-70).
#pragma once
#include "model_common.cuh"
#include "unet.cuh"
#include "ka.cuh"

namespace pd {

// Optional knowledge alignment: ka != null adds g = ka->mean_shift(z_t, t, avg_x_gt, guide_scale) to every step
// (computed on its own stream, concurrently with the UNet) and the update subtracts coef[6] * g.
struct SamplerAlign {
    KANet* ka = nullptr;
    const float* avg_x_gt = nullptr;   // device fp32 [B]
    float guide_scale = 0.f;
};

class Sampler {
public:
    using Align = SamplerAlign;
    Sampler(int num_timesteps, double linear_start, double linear_end);
    ~Sampler();
    int get_buffer(const char* name, float* out) const;
    // Fills rows[k][8] + timesteps[k] for the k-th executed step.
    int coefficients(int mode, int n_steps, float eta, std::vector<float>* rows, std::vector<int64_t>* ts) const;
    // Runs executed steps [k_begin, k_end) of the n_total-step schedule; noise[0] belongs to step k_begin.
    int loop(UNet* unet, float* z, const float* cond, const float* noise, int B, int mode, int n_total, float eta,
             int k_begin, int k_end, cudaStream_t st, const Align& al = Align());
    int step_ddpm(UNet* unet, float* z, const float* cond, const float* noise, int B, int t, cudaStream_t st);
    // Forward-only LatentDiffusion.p_losses (latent_diffusion.py:517-551; eps-parameterization, fixed logvar): q_sample
    // -> UNet -> per-sample loss -> {loss_simple, loss_vlb, loss, loss_gamma}. All pointers on the device; t int64 [B];
    // per_sample [B] and out4 [4] are outputs. This is the validation-loss path (validation_step -> self(batch)); no
    // gradients are produced.
    int losses(UNet* unet, const float* x_start, const float* cond, const int64_t* t, const float* noise, int B,
               int loss_l1, float logvar, float l_simple_weight, float elbo_weight, float* per_sample, float* out4,
               cudaStream_t st);

    int T;
    // clip_denoised of p_mean_variance (latent_diffusion.py:580-581): z_0 estimate clamped to [-1, 1] inside the update
    void set_clip_denoised(bool on) { if (on != clip_) { clip_ = on; table_key_.valid = false; } }
    bool clip_denoised() const { return clip_; }
    // number of concurrent sub-batches a batch of B is cut into (env PD_SUB_BATCHES, default 2, must divide B)
    int n_sub_for(int B) const;

private:
    int enter(cudaStream_t user);
    int leave(cudaStream_t user);
    int loop_on(cudaStream_t st, UNet* unet, float* z, const float* cond, const float* noise, int B, int mode,
                int n_total, float eta, int k_begin, int k_end, const Align& al);
    int step_on(cudaStream_t st, UNet* unet, float* z, const float* cond, const float* noise, int B, int t);
    int upload_tables(const std::vector<float>& rows, const std::vector<int64_t>& ts, int B, cudaStream_t st);
    int capture(cudaStream_t st, UNet* unet, const float* noise, int B, const Align& al, int iterations,
                cudaGraphExec_t* out);
    int one_iteration(UNet* unet, float* z, const float* cond, const float* noise, int B, cudaStream_t st,
                      const Align& al = Align());
    void drop_graph();

    std::map<std::string, std::vector<float>> buf_;  // the reference's registered fp32 buffers
    DevMem coef_dev_, t_dev_, step_dev_, eps_dev_;
    // Stable device homes of the loop state: the captured graphs only ever see these addresses (the caller's z / cond /
    // avg_x_gt are copied in before the first replay and z is copied back after the last), so a graph survives any
    // change of the caller's tensor addresses.
    DevMem z_buf_, cond_buf_, target_buf_;
    bool clip_ = false;
    // what the resident coefficient / timestep tables currently hold (re-uploaded only when this changes)
    struct TableKey {
        bool valid = false;
        int mode = 0, n_total = 0, k_begin = 0, k_end = 0, B = 0;
        float eta = 0.f;
        bool operator==(const TableKey& o) const {
            return valid && o.valid && mode == o.mode && n_total == o.n_total && k_begin == o.k_begin && k_end == o.k_end &&
                   B == o.B && eta == o.eta;
        }
    } table_key_;
    bool table_needs_noise_ = false;
    DevMem loss_tab_, loss_ws_;   // {sqrt_ac, sqrt_1mac, lvlb} tables; x_noisy + eps workspace of losses()
    // cached graphs: one iteration (replayed n times: long DDPM stretches) and a whole loop of graph_loop_steps_
    // iterations (one launch per loop: the 50-step DDIM benchmark)
    cudaGraphExec_t graph_exec_ = nullptr, graph_loop_ = nullptr;
    int graph_loop_steps_ = 0;
    static constexpr int kWholeLoopMaxSteps = 64;
    // sub-batch concurrency: the batch is cut into n_sub independent slices (samples never interact) that run the
    // same launch sequence on parallel streams, so one slice's prologue / epilogue bubbles are filled by another's work
    static constexpr int kMaxSub = 4;
    cudaStream_t sub_stream_[kMaxSub] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork_ = nullptr, ev_sub_[kMaxSub] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t ka_stream_ = nullptr;
    cudaEvent_t ev_ka_ = nullptr;
    cudaStream_t loop_stream_ = nullptr;
    cudaEvent_t ev_in_ = nullptr, ev_out_ = nullptr;
    // A graph is valid for one (UNet, its weight generation, KA net, its generation, batch, noise stack, guide scale):
    // finalize() frees the packed weights and activation arenas the captured kernels point at and bumps the generation.
    struct Key {
        const void *unet, *noise, *ka;
        unsigned long long unet_gen, ka_gen;
        int B;
        float gs;
        bool operator==(const Key& o) const {
            return unet == o.unet && noise == o.noise && B == o.B && ka == o.ka && unet_gen == o.unet_gen &&
                   ka_gen == o.ka_gen && gs == o.gs;
        }
    } graph_key_{};
};

}  // namespace pd
