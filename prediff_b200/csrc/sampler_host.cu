#include "sampler_host.cuh"
#include <cmath>
#include <cstdlib>

namespace pd {

Sampler::Sampler(int num_timesteps, double linear_start, double linear_end) : T(num_timesteps) {
    // "linear" beta schedule: linspace(sqrt(start), sqrt(end), T)^2 in float64 (diffusion/utils.py:17-22)
    std::vector<double> betas(T), ac(T), ac_prev(T);
    const double s0 = std::sqrt(linear_start), s1 = std::sqrt(linear_end);
    double cp = 1.0;
    for (int i = 0; i < T; ++i) {
        // torch.linspace(float64): start + i*step for the first half, end - (T-1-i)*step for the second half
        const double step = (s1 - s0) / (double)(T - 1);
        const double v = (i < T / 2) ? s0 + step * i : s1 - step * (T - 1 - i);
        betas[i] = v * v;
        ac_prev[i] = cp;
        cp *= (1.0 - betas[i]);
        ac[i] = cp;
    }
    auto put = [&](const char* name, auto fn) {
        std::vector<float> v(T);
        for (int i = 0; i < T; ++i) v[i] = (float)fn(i);
        buf_[name] = std::move(v);
    };
    auto pv = [&](int i) { return betas[i] * (1.0 - ac_prev[i]) / (1.0 - ac[i]); };
    put("betas", [&](int i) { return betas[i]; });
    put("alphas_cumprod", [&](int i) { return ac[i]; });
    put("alphas_cumprod_prev", [&](int i) { return ac_prev[i]; });
    put("sqrt_alphas_cumprod", [&](int i) { return std::sqrt(ac[i]); });
    put("sqrt_one_minus_alphas_cumprod", [&](int i) { return std::sqrt(1.0 - ac[i]); });
    put("log_one_minus_alphas_cumprod", [&](int i) { return std::log(1.0 - ac[i]); });
    put("sqrt_recip_alphas_cumprod", [&](int i) { return std::sqrt(1.0 / ac[i]); });
    put("sqrt_recipm1_alphas_cumprod", [&](int i) { return std::sqrt(1.0 / ac[i] - 1.0); });
    put("posterior_variance", pv);
    put("posterior_log_variance_clipped", [&](int i) { return std::log(std::max(pv(i), 1e-20)); });
    put("posterior_mean_coef1", [&](int i) { return betas[i] * std::sqrt(ac_prev[i]) / (1.0 - ac[i]); });
    put("posterior_mean_coef2", [&](int i) { return (1.0 - ac_prev[i]) * std::sqrt(1.0 - betas[i]) / (1.0 - ac[i]); });
    // lvlb_weights (eps-parameterization, latent_diffusion.py:270-277): fp32 tensor arithmetic on the registered buffers,
    // betas^2 / (2 * posterior_variance * alphas * (1 - alphas_cumprod)), entry 0 replaced by entry 1
    {
        std::vector<float> w(T);
        const std::vector<float>&b = buf_["betas"], &pvar = buf_["posterior_variance"], &acf = buf_["alphas_cumprod"];
        for (int i = 0; i < T; ++i) {
            const float alpha = (float)(1.0 - betas[i]);
            float den = 2.f * pvar[i];
            den = den * alpha;
            den = den * (1.f - acf[i]);
            w[i] = (b[i] * b[i]) / den;
        }
        if (T > 1) w[0] = w[1];
        buf_["lvlb_weights"] = std::move(w);
    }
}

int Sampler::losses(UNet* unet, const float* x_start, const float* cond, const int64_t* t, const float* noise, int B,
                    int loss_l1, float logvar, float l_simple_weight, float elbo_weight, float* per_sample, float* out4,
                    cudaStream_t st) {
    PD_CHECK(unet && x_start && cond && t && noise && per_sample && out4, PD_ERR_ARG, "losses: null argument");
    const int64_t n = (int64_t)unet->cfg.t_out * unet->cfg.h * unet->cfg.w * unet->cfg.c;
    if (!loss_tab_.p) {
        PD_TRY(loss_tab_.alloc((size_t)3 * T * sizeof(float)));
        const char* names[3] = {"sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "lvlb_weights"};
        for (int k = 0; k < 3; ++k)
            PD_CUDA(cudaMemcpy(loss_tab_.as<float>() + (size_t)k * T, buf_[names[k]].data(), T * sizeof(float),
                               cudaMemcpyHostToDevice));
    }
    const size_t need = (size_t)2 * B * n * sizeof(float);
    if (loss_ws_.bytes < need) {
        PD_CUDA(cudaStreamSynchronize(st));   // a previous call on this stream may still use the old workspace
        PD_TRY(loss_ws_.alloc(need));
    }
    float* x_noisy = loss_ws_.as<float>();
    float* eps = x_noisy + (size_t)B * n;
    const float* tab = loss_tab_.as<float>();
    PD_TRY(q_sample(x_start, noise, t, tab, tab + T, x_noisy, B, n, st));
    PD_TRY(unet->forward(x_noisy, t, nullptr, cond, eps, B, st));
    // eps-parameterization: target = noise (:527-530)
    return diffusion_loss_reduce(eps, noise, t, tab + 2 * (size_t)T, logvar, l_simple_weight, elbo_weight, loss_l1,
                                 per_sample, out4, B, n, st);
}

Sampler::~Sampler() {
    drop_graph();
    for (int i = 0; i < kMaxSub; ++i) {
        if (sub_stream_[i]) cudaStreamDestroy(sub_stream_[i]);
        if (ev_sub_[i]) cudaEventDestroy(ev_sub_[i]);
    }
    if (ev_fork_) cudaEventDestroy(ev_fork_);
    if (ev_ka_) cudaEventDestroy(ev_ka_);
    if (ka_stream_) cudaStreamDestroy(ka_stream_);
    if (ev_in_) cudaEventDestroy(ev_in_);
    if (ev_out_) cudaEventDestroy(ev_out_);
    if (loop_stream_) cudaStreamDestroy(loop_stream_);
}

// The loop runs on a private non-blocking stream (stream capture is not allowed on the legacy default stream,
// which is what torch hands us by default); events order it after / before the caller's stream.
int Sampler::enter(cudaStream_t user) {
    if (!loop_stream_) {
        PD_CUDA(cudaStreamCreateWithFlags(&loop_stream_, cudaStreamNonBlocking));
        PD_CUDA(cudaEventCreateWithFlags(&ev_in_, cudaEventDisableTiming));
        PD_CUDA(cudaEventCreateWithFlags(&ev_out_, cudaEventDisableTiming));
    }
    PD_CUDA(cudaEventRecord(ev_in_, user));
    PD_CUDA(cudaStreamWaitEvent(loop_stream_, ev_in_, 0));
    return PD_OK;
}
int Sampler::leave(cudaStream_t user) {
    PD_CUDA(cudaEventRecord(ev_out_, loop_stream_));
    PD_CUDA(cudaStreamWaitEvent(user, ev_out_, 0));
    return PD_OK;
}

void Sampler::drop_graph() {
    if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
    if (graph_loop_) cudaGraphExecDestroy(graph_loop_);
    graph_exec_ = nullptr;
    graph_loop_ = nullptr;
    graph_loop_steps_ = 0;
}

int Sampler::get_buffer(const char* name, float* out) const {
    PD_CHECK(name && out, PD_ERR_ARG, "sampler_get_buffer: null argument");
    auto it = buf_.find(name);
    PD_CHECK(it != buf_.end(), PD_ERR_ARG, "sampler_get_buffer: unknown buffer '%s'", name);
    memcpy(out, it->second.data(), (size_t)T * sizeof(float));
    return PD_OK;
}

int Sampler::coefficients(int mode, int n_steps, float eta, std::vector<float>* rows, std::vector<int64_t>* ts) const {
    PD_CHECK(n_steps >= 1 && n_steps <= T, PD_ERR_ARG, "sampler: n_steps %d outside [1, %d]", n_steps, T);
    rows->assign((size_t)n_steps * 8, 0.f);
    ts->assign(n_steps, 0);
    if (mode == PD_MODE_DDPM) {
        // p_sample_loop: for i in reversed(range(timesteps)) (latent_diffusion.py:663)
        const auto& r = buf_.at("sqrt_recip_alphas_cumprod");
        const auto& rm1 = buf_.at("sqrt_recipm1_alphas_cumprod");
        const auto& c1 = buf_.at("posterior_mean_coef1");
        const auto& c2 = buf_.at("posterior_mean_coef2");
        const auto& lv = buf_.at("posterior_log_variance_clipped");
        for (int k = 0; k < n_steps; ++k) {
            const int t = n_steps - 1 - k;
            float* c = rows->data() + (size_t)k * 8;
            const float sigma = std::exp(0.5f * lv[t]);
            c[0] = r[t]; c[1] = rm1[t]; c[2] = c1[t]; c[3] = c2[t]; c[4] = 0.f;
            c[5] = t == 0 ? 0.f : sigma;   // no noise when t == 0 (latent_diffusion.py:624-626)
            c[6] = sigma;                  // aligned_mean: mean - exp(0.5 logvar) * grad (:594-595)
            c[7] = clip_ ? 1.f : 0.f;      // clip_denoised: z_recon.clamp_(-1, 1) (:580-581)
            (*ts)[k] = t;
        }
    } else if (mode == PD_MODE_DDIM) {
        PD_CHECK(T % n_steps == 0 || T / n_steps >= 1, PD_ERR_ARG, "sampler: bad DDIM step count");
        const auto& ac = buf_.at("alphas_cumprod");
        const int c = T / n_steps;
        std::vector<int> steps;
        for (int i = 0; i < T; i += c) steps.push_back(i + 1);  // make_ddim_timesteps('uniform') (+1)
        PD_CHECK((int)steps.size() == n_steps && steps.back() < T, PD_ERR_ARG,
                 "sampler: DDIM with %d steps is not a uniform sub-sequence of %d", n_steps, T);
        for (int k = 0; k < n_steps; ++k) {
            const int i = n_steps - 1 - k;
            const double a = ac[steps[i]];
            const double a_prev = i == 0 ? ac[0] : ac[steps[i - 1]];
            const double sigma = (double)eta * std::sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev));
            float* cf = rows->data() + (size_t)k * 8;
            const double c0 = 1.0 / std::sqrt(a), c1 = std::sqrt(1 - a) / std::sqrt(a), c2 = std::sqrt(a_prev);
            const double c4 = std::sqrt(std::max(1 - a_prev - sigma * sigma, 0.0));
            cf[0] = (float)c0; cf[1] = (float)c1; cf[2] = (float)c2; cf[3] = 0.f; cf[4] = (float)c4;
            cf[5] = (float)sigma;
            cf[6] = (float)((c2 * c1 - c4) * std::sqrt(1 - a));  // eps_hat = eps + sqrt(1-a_t) * guide (S6)
            (*ts)[k] = steps[i];
        }
    } else {
        set_error("sampler: unknown mode %d", mode);
        return PD_ERR_ARG;
    }
    return PD_OK;
}

int Sampler::upload_tables(const std::vector<float>& rows, const std::vector<int64_t>& ts, int B, cudaStream_t st) {
    const int n = (int)ts.size();
    std::vector<int64_t> tt((size_t)n * B);
    for (int k = 0; k < n; ++k)
        for (int b = 0; b < B; ++b) tt[(size_t)k * B + b] = ts[k];
    if (coef_dev_.bytes < rows.size() * sizeof(float)) {
        PD_TRY(coef_dev_.alloc(rows.size() * sizeof(float)));
        drop_graph();  // the captured iteration holds the old table address
    }
    if (t_dev_.bytes < tt.size() * sizeof(int64_t)) {
        PD_TRY(t_dev_.alloc(tt.size() * sizeof(int64_t)));
        drop_graph();
    }
    if (!step_dev_.p) PD_TRY(step_dev_.alloc(sizeof(int)));
    // synchronous small copies: the host vectors die at return
    PD_CUDA(cudaStreamSynchronize(st));
    PD_CUDA(cudaMemcpy(coef_dev_.p, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice));
    PD_CUDA(cudaMemcpy(t_dev_.p, tt.data(), tt.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    PD_CUDA(cudaMemset(step_dev_.p, 0, sizeof(int)));
    return PD_OK;
}

// Concurrent sub-batch slices of one loop step (PD_SUB_BATCHES). Default 1: with the stream-K convolutions a batch-4
// step already puts 144 CTAs on the 148 SMs and every other kernel is latency-bound, so slices only add contention
// (measured: 775 sample-steps/s with 1 slice, 751 with 2).
int Sampler::n_sub_for(int B) const {
    int n = 1;
    if (const char* e = getenv("PD_SUB_BATCHES")) n = atoi(e);
    if (n < 1) n = 1;
    if (n > kMaxSub) n = kMaxSub;
    while (n > 1 && B % n != 0) --n;
    return n;
}

int Sampler::one_iteration(UNet* unet, float* z, const float* cond, const float* noise, int B, cudaStream_t st,
                           const Align& al) {
    const int64_t per = (int64_t)unet->cfg.t_out * unet->cfg.h * unet->cfg.w * unet->cfg.c;
    const int64_t per_c = (int64_t)unet->cfg.t_in * unet->cfg.h * unet->cfg.w * unet->cfg.c;
    const int* step = step_dev_.as<int>();
    const int ns = n_sub_for(B);
    const bool fork = ns > 1 || al.ka != nullptr;
    const float* guide = nullptr;
    if (fork) {
        if (!ev_fork_) PD_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
        PD_CUDA(cudaEventRecord(ev_fork_, st));
    }
    if (al.ka) {
        // the guidance depends on z_t only, not on eps: it runs beside the UNet and fills its idle SMs
        if (!ka_stream_) {
            PD_CUDA(cudaStreamCreateWithFlags(&ka_stream_, cudaStreamNonBlocking));
            PD_CUDA(cudaEventCreateWithFlags(&ev_ka_, cudaEventDisableTiming));
        }
        PD_CUDA(cudaStreamWaitEvent(ka_stream_, ev_fork_, 0));
        PD_TRY(al.ka->mean_shift(z, t_dev_.as<int64_t>(), step, B, al.avg_x_gt, al.guide_scale, nullptr, B, ka_stream_));
        PD_CUDA(cudaEventRecord(ev_ka_, ka_stream_));
        float* g = nullptr;
        PD_TRY(al.ka->guide_buffer(B, &g, nullptr));
        guide = g;
    }
    if (ns == 1) {
        PD_TRY(unet->forward(z, t_dev_.as<int64_t>(), step, cond, eps_dev_.as<float>(), B, st));
        if (al.ka) PD_CUDA(cudaStreamWaitEvent(st, ev_ka_, 0));
        PD_TRY(sampler_update(z, eps_dev_.as<float>(), noise, guide, coef_dev_.as<float>(), step, B * per, 0, st));
        return advance_step(step_dev_.as<int>(), st);
    }
    const int Bs = B / ns;
    for (int i = 0; i < ns; ++i) {
        if (!sub_stream_[i]) {
            PD_CUDA(cudaStreamCreateWithFlags(&sub_stream_[i], cudaStreamNonBlocking));
            PD_CUDA(cudaEventCreateWithFlags(&ev_sub_[i], cudaEventDisableTiming));
        }
        cudaStream_t ss = sub_stream_[i];
        PD_CUDA(cudaStreamWaitEvent(ss, ev_fork_, 0));
        float* zi = z + (int64_t)i * Bs * per;
        float* ei = eps_dev_.as<float>() + (int64_t)i * Bs * per;
        PD_TRY(unet->forward(zi, t_dev_.as<int64_t>() + (int64_t)i * Bs, step, cond + (int64_t)i * Bs * per_c, ei, Bs, ss,
                             nullptr, i, B));
        if (!al.ka)   // with guidance z_t must stay intact until the KA stream has read it: update after the join
            PD_TRY(sampler_update(zi, ei, noise ? noise + (int64_t)i * Bs * per : nullptr, nullptr, coef_dev_.as<float>(),
                                  step, Bs * per, B * per, ss));
        PD_CUDA(cudaEventRecord(ev_sub_[i], ss));
        PD_CUDA(cudaStreamWaitEvent(st, ev_sub_[i], 0));
    }
    if (al.ka) {
        PD_CUDA(cudaStreamWaitEvent(st, ev_ka_, 0));
        PD_TRY(sampler_update(z, eps_dev_.as<float>(), noise, guide, coef_dev_.as<float>(), step, B * per, 0, st));
    }
    return advance_step(step_dev_.as<int>(), st);
}

int Sampler::loop(UNet* unet, float* z, const float* cond, const float* noise, int B, int mode, int n_total, float eta,
                  int k_begin, int k_end, cudaStream_t user, const Align& al) {
    PD_CHECK(unet && z && cond, PD_ERR_ARG, "sample_loop: null pointer");
    PD_CHECK(!al.ka || al.avg_x_gt, PD_ERR_ARG, "sample_loop: alignment needs avg_x_gt");
    if (al.ka)
        PD_CHECK(al.ka->cfg.t == unet->cfg.t_out && al.ka->cfg.h == unet->cfg.h && al.ka->cfg.w == unet->cfg.w &&
                     al.ka->cfg.c == unet->cfg.c,
                 PD_ERR_SHAPE, "sample_loop: the alignment network's input shape differs from the UNet's target shape");
    PD_TRY(enter(user));
    const int rc = loop_on(loop_stream_, unet, z, cond, noise, B, mode, n_total, eta, k_begin, k_end, al);
    const int rl = leave(user);
    return rc != PD_OK ? rc : rl;
}

// Captures `iterations` consecutive loop iterations on the internal state buffers into one executable graph.
int Sampler::capture(cudaStream_t st, UNet* unet, const float* noise, int B, const Align& al, int iterations,
                     cudaGraphExec_t* out) {
    cudaGraph_t graph = nullptr;
    PD_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = PD_OK;
    for (int i = 0; i < iterations && rc == PD_OK; ++i)
        rc = one_iteration(unet, z_buf_.as<float>(), cond_buf_.as<float>(), noise, B, st, al);
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != PD_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    PD_CUDA(ce);
    const cudaError_t ie = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    PD_CUDA(ie);
    return PD_OK;
}

int Sampler::loop_on(cudaStream_t st, UNet* unet, float* z, const float* cond, const float* noise, int B, int mode,
                     int n_total, float eta, int k_begin, int k_end, const Align& al_user) {
    PD_CHECK(unet->finalized, PD_ERR_STATE, "sample_loop: pd_unet_finalize() has not been called since the last weight load");
    PD_CHECK(!al_user.ka || al_user.ka->finalized, PD_ERR_STATE, "sample_loop: pd_ka_finalize() has not been called");
    PD_CHECK(0 <= k_begin && k_begin <= k_end && k_end <= n_total, PD_ERR_ARG, "sample_loop: bad step range [%d, %d)",
             k_begin, k_end);
    if (k_begin == k_end) return PD_OK;
    const int n_steps = k_end - k_begin;
    const size_t per = (size_t)unet->cfg.t_out * unet->cfg.h * unet->cfg.w * unet->cfg.c;
    const size_t per_c = (size_t)unet->cfg.t_in * unet->cfg.h * unet->cfg.w * unet->cfg.c;
    const size_t n = (size_t)B * per;
    if (eps_dev_.bytes < n * sizeof(float)) {
        PD_TRY(eps_dev_.alloc(n * sizeof(float)));
        drop_graph();
    }
    // coefficient rows + timesteps of the executed range: resident on the device, re-uploaded only when the request changes
    TableKey tk;
    tk.valid = true; tk.mode = mode; tk.n_total = n_total; tk.k_begin = k_begin; tk.k_end = k_end; tk.B = B; tk.eta = eta;
    if (!(tk == table_key_)) {
        std::vector<float> rows;
        std::vector<int64_t> ts;
        PD_TRY(coefficients(mode, n_total, eta, &rows, &ts));
        rows = std::vector<float>(rows.begin() + (size_t)k_begin * 8, rows.begin() + (size_t)k_end * 8);
        ts = std::vector<int64_t>(ts.begin() + k_begin, ts.begin() + k_end);
        table_needs_noise_ = false;
        for (int k = 0; k < n_steps; ++k) table_needs_noise_ |= rows[(size_t)k * 8 + 5] != 0.f;
        table_key_.valid = false;
        PD_TRY(upload_tables(rows, ts, B, st));
        table_key_ = tk;
    }
    PD_CHECK(!table_needs_noise_ || noise, PD_ERR_ARG, "sample_loop: this sampler is stochastic; pass the noise stack");
    PD_CUDA(cudaMemsetAsync(step_dev_.p, 0, sizeof(int), st));

    const bool use_graph = getenv("PD_NO_GRAPH") == nullptr && n_steps > 2;
    if (!use_graph) {
        for (int k = 0; k < n_steps; ++k) PD_TRY(one_iteration(unet, z, cond, noise, B, st, al_user));
        return PD_OK;
    }
    // ---- graph path: the loop state lives in the sampler's own buffers --------------------------------------------
    if (z_buf_.bytes < n * sizeof(float) || cond_buf_.bytes < (size_t)B * per_c * sizeof(float) ||
        target_buf_.bytes < (size_t)B * sizeof(float)) {
        PD_CUDA(cudaStreamSynchronize(st));
        drop_graph();
        if (z_buf_.bytes < n * sizeof(float)) PD_TRY(z_buf_.alloc(n * sizeof(float)));
        if (cond_buf_.bytes < (size_t)B * per_c * sizeof(float)) PD_TRY(cond_buf_.alloc((size_t)B * per_c * sizeof(float)));
        if (target_buf_.bytes < (size_t)B * sizeof(float)) PD_TRY(target_buf_.alloc((size_t)B * sizeof(float)));
    }
    Align al = al_user;
    if (al.ka) {
        PD_CUDA(cudaMemcpyAsync(target_buf_.p, al_user.avg_x_gt, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
        al.avg_x_gt = target_buf_.as<float>();
    }
    PD_CUDA(cudaMemcpyAsync(cond_buf_.p, cond, (size_t)B * per_c * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const Key key{unet, noise, al.ka, unet->generation, al.ka ? al.ka->generation : 0ull, B, al.guide_scale};
    if (!(key == graph_key_) || (!graph_exec_ && !graph_loop_)) {
        drop_graph();
        // one throw-away eager iteration on the internal buffers: performs every lazy one-time setup (plans, arenas,
        // kernel attributes) outside of stream capture; z_buf_ is (re)loaded below
        PD_CUDA(cudaMemcpyAsync(z_buf_.p, z, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
        PD_TRY(one_iteration(unet, z_buf_.as<float>(), cond_buf_.as<float>(), noise, B, st, al));
        PD_CUDA(cudaMemsetAsync(step_dev_.p, 0, sizeof(int), st));
        graph_key_ = key;
    }
    PD_CUDA(cudaMemcpyAsync(z_buf_.p, z, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (n_steps <= kWholeLoopMaxSteps && getenv("PD_NO_LOOP_GRAPH") == nullptr) {
        // the whole loop is ONE graph launch (50-step DDIM: 50 x 286 kernel nodes)
        if (!graph_loop_ || graph_loop_steps_ != n_steps) {
            if (graph_loop_) cudaGraphExecDestroy(graph_loop_);
            graph_loop_ = nullptr;
            PD_TRY(capture(st, unet, noise, B, al, n_steps, &graph_loop_));
            graph_loop_steps_ = n_steps;
        }
        PD_CUDA(cudaGraphLaunch(graph_loop_, st));
    } else {
        if (!graph_exec_) PD_TRY(capture(st, unet, noise, B, al, 1, &graph_exec_));
        for (int k = 0; k < n_steps; ++k) PD_CUDA(cudaGraphLaunch(graph_exec_, st));
    }
    PD_CUDA(cudaMemcpyAsync(z, z_buf_.p, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return PD_OK;
}

int Sampler::step_ddpm(UNet* unet, float* z, const float* cond, const float* noise, int B, int t, cudaStream_t user) {
    PD_CHECK(unet && z && cond, PD_ERR_ARG, "sample_step: null pointer");
    PD_TRY(enter(user));
    const int rc = step_on(loop_stream_, unet, z, cond, noise, B, t);
    const int rl = leave(user);
    return rc != PD_OK ? rc : rl;
}

int Sampler::step_on(cudaStream_t st, UNet* unet, float* z, const float* cond, const float* noise, int B, int t) {
    PD_CHECK(t >= 0 && t < T, PD_ERR_ARG, "sample_step: t=%d outside the schedule", t);
    std::vector<float> rows;
    std::vector<int64_t> ts;
    PD_TRY(coefficients(PD_MODE_DDPM, t + 1, 0.f, &rows, &ts));  // row 0 is timestep t
    rows.resize(8);
    ts.resize(1);
    PD_CHECK(rows[5] == 0.f || noise, PD_ERR_ARG, "sample_step: noise required for t > 0");
    const size_t n = (size_t)B * unet->cfg.t_out * unet->cfg.h * unet->cfg.w * unet->cfg.c;
    if (eps_dev_.bytes < n * sizeof(float)) {
        PD_TRY(eps_dev_.alloc(n * sizeof(float)));
        drop_graph();
    }
    table_key_.valid = false;
    PD_TRY(upload_tables(rows, ts, B, st));
    return one_iteration(unet, z, cond, noise, B, st);
}

}  // namespace pd
