// Fused sampler update: the ~25 elementwise ATen launches of LatentDiffusion.p_sample
// (reference latent_diffusion.py:553-566,598-631) collapse into one kernel that reads its coefficients from
// the device-resident schedule table, so the step loop never returns to the host.
#include "ops.cuh"
#include "model_common.cuh"

namespace pd {
namespace {

__global__ void __launch_bounds__(256) sampler_update_kernel(float* __restrict__ z, const float* __restrict__ eps,
                                                             const float* __restrict__ noise,
                                                             const float* __restrict__ guide,
                                                             const float* __restrict__ coef,
                                                             const int* __restrict__ step, int64_t n4,
                                                             int64_t noise_step_stride) {
    grid_dep_launch();
    grid_dep_wait();
    if (step) {
        const int k = *step;
        coef += (size_t)k * 8;
        if (noise) noise += (size_t)k * noise_step_stride;
    }
    const float c0 = coef[0], c1 = coef[1], c2 = coef[2], c3 = coef[3], c4 = coef[4], c5 = coef[5], c6 = coef[6];
    const bool clip = coef[7] != 0.f;   // clip_denoised (latent_diffusion.py:580-581): z_0 estimate clamped to [-1, 1]
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 zv = reinterpret_cast<const float4*>(z)[i];
        const float4 ev = __ldg(reinterpret_cast<const float4*>(eps) + i);
        float4 nv = make_float4(0.f, 0.f, 0.f, 0.f), gv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (noise && c5 != 0.f) nv = __ldg(reinterpret_cast<const float4*>(noise) + i);
        if (guide && c6 != 0.f) gv = __ldg(reinterpret_cast<const float4*>(guide) + i);
        float zi[4] = {zv.x, zv.y, zv.z, zv.w};
        const float ei[4] = {ev.x, ev.y, ev.z, ev.w};
        const float ni[4] = {nv.x, nv.y, nv.z, nv.w};
        const float gi[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float z0 = c0 * zi[k] - c1 * ei[k];
            if (clip) z0 = fminf(fmaxf(z0, -1.f), 1.f);
            float m = c2 * z0 + c3 * zi[k] + c4 * ei[k];
            m -= c6 * gi[k];
            zi[k] = m + c5 * ni[k];
        }
        reinterpret_cast<float4*>(z)[i] = make_float4(zi[0], zi[1], zi[2], zi[3]);
    }
}

// ---- forward diffusion loss (LatentDiffusion.q_sample / p_losses, latent_diffusion.py:489-551) -----------------
// x_noisy = sqrt_ac[t_b] * x_start + sqrt_1mac[t_b] * noise; one grid row per sample, float4 lanes.
__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                       const int64_t* __restrict__ t, const float* __restrict__ sqrt_ac,
                                                       const float* __restrict__ sqrt_1mac, float* __restrict__ out,
                                                       int64_t n4) {
    grid_dep_launch();
    grid_dep_wait();
    const int b = blockIdx.y;
    const int64_t tb = t[b];
    const float ca = __ldg(sqrt_ac + tb), cn = __ldg(sqrt_1mac + tb);
    const float4* xa = reinterpret_cast<const float4*>(x0) + (size_t)b * n4;
    const float4* na = reinterpret_cast<const float4*>(noise) + (size_t)b * n4;
    float4* oa = reinterpret_cast<float4*>(out) + (size_t)b * n4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = __ldg(xa + i), n = __ldg(na + i);
        // mul, mul, add with separate roundings = the reference's tensor expression (bit-exact, no fma contraction)
        oa[i] = make_float4(__fadd_rn(__fmul_rn(ca, x.x), __fmul_rn(cn, n.x)), __fadd_rn(__fmul_rn(ca, x.y), __fmul_rn(cn, n.y)),
                            __fadd_rn(__fmul_rn(ca, x.z), __fmul_rn(cn, n.z)), __fadd_rn(__fmul_rn(ca, x.w), __fmul_rn(cn, n.w)));
    }
}

// per_sample[b] = mean_i |pred - target|^p (p = 2: 'l2' / mse, p = 1: 'l1'); one block per sample, double
// accumulation in a fixed order (deterministic, independent of the batch size).
__global__ void __launch_bounds__(1024) loss_per_sample_kernel(const float* __restrict__ pred,
                                                               const float* __restrict__ target,
                                                               float* __restrict__ per_sample, int64_t n4, int l1) {
    grid_dep_launch();
    grid_dep_wait();
    const int b = blockIdx.x;
    const float4* pa = reinterpret_cast<const float4*>(pred) + (size_t)b * n4;
    const float4* ta = reinterpret_cast<const float4*>(target) + (size_t)b * n4;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 p = __ldg(pa + i), t = __ldg(ta + i);
        const float d[4] = {t.x - p.x, t.y - p.y, t.z - p.z, t.w - p.w};
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += l1 ? fabsf(d[k]) : d[k] * d[k];
        acc += (double)s;
    }
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) per_sample[b] = (float)(v / (double)(n4 * 4));
    }
}

// out = {mean(loss_simple), loss_vlb = mean(lvlb[t] * loss_simple), loss = w_simple * mean(loss_simple / exp(logvar)
// + logvar) + w_elbo * loss_vlb, loss_gamma = mean(loss_simple / exp(logvar) + logvar)}   (p_losses :534-549)
__global__ void loss_finish_kernel(const float* __restrict__ per_sample, const int64_t* __restrict__ t,
                                   const float* __restrict__ lvlb, float logvar, float w_simple, float w_elbo, int B,
                                   float* __restrict__ out) {
    grid_dep_launch();
    grid_dep_wait();
    if (threadIdx.x != 0) return;
    float s = 0.f, g = 0.f, v = 0.f;
    for (int b = 0; b < B; ++b) {
        const float ls = per_sample[b];
        s += ls;
        g += ls / expf(logvar) + logvar;
        v += lvlb[t[b]] * ls;
    }
    s /= (float)B;
    g /= (float)B;
    v /= (float)B;
    out[0] = s;
    out[1] = v;
    out[2] = w_simple * g + w_elbo * v;
    out[3] = g;
}

__global__ void advance_step_kernel(int* step) {
    grid_dep_launch();
    grid_dep_wait(); *step += 1; }
__global__ void stamp_globaltimer_kernel(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}

}  // namespace

int stamp_globaltimer(unsigned long long* slot, cudaStream_t st) {
    stamp_globaltimer_kernel<<<1, 1, 0, st>>>(slot);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac, const float* sqrt_1mac,
             float* out, int B, int64_t n, cudaStream_t st) {
    PD_CHECK(n % 4 == 0 && B >= 1 && B <= 65535, PD_ERR_SHAPE, "q_sample: B=%d n=%lld", B, (long long)n);
    const int64_t n4 = n / 4;
    int bx = (int)((n4 + 255) / 256);
    if (bx > kNumSMs * 4) bx = kNumSMs * 4;
    PD_LAUNCH(q_sample_kernel, dim3(bx, B), 256, 0, st, x0, noise, t, sqrt_ac, sqrt_1mac, out, n4);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int diffusion_loss_reduce(const float* pred, const float* target, const int64_t* t, const float* lvlb, float logvar,
                          float w_simple, float w_elbo, int l1, float* per_sample, float* out4, int B, int64_t n,
                          cudaStream_t st) {
    PD_CHECK(n % 4 == 0 && B >= 1, PD_ERR_SHAPE, "diffusion_loss_reduce: B=%d n=%lld", B, (long long)n);
    PD_LAUNCH(loss_per_sample_kernel, B, 1024, 0, st, pred, target, per_sample, n / 4, l1);
    PD_LAUNCH_CHECK();
    PD_LAUNCH(loss_finish_kernel, 1, 32, 0, st, (const float*)per_sample, t, lvlb, logvar, w_simple, w_elbo, B, out4);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int advance_step(int* step, cudaStream_t st) {
    PD_LAUNCH(advance_step_kernel, 1, 1, 0, st, step);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

int sampler_update(float* z, const float* eps, const float* noise, const float* guide, const float* coef,
                   const int* step, int64_t n, int64_t noise_step_stride, cudaStream_t st) {
    PD_CHECK(n % 4 == 0, PD_ERR_SHAPE, "sampler_update: n must be a multiple of 4");
    const int64_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    PD_LAUNCH(sampler_update_kernel, blocks, 256, 0, st, z, eps, noise, guide, coef, step, n4,
                                                  noise_step_stride ? noise_step_stride : n);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
